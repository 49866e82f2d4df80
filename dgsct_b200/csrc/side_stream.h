// Auxiliary CUDA streams forked from the caller's stream inside one library call (event record / wait: graph-capturable, so
// inside a captured step they become parallel graph branches) and joined before the call returns.
#pragma once
#include <cuda_runtime.h>

#include <deque>

namespace dgsct {

// ---- side stream for the parameter-gradient products ------------------------------------------------------------------
// Every weight gradient (sum_b A_b . B_b^T, accumulated atomically into G.*) is off the critical path of the backward: the
// activation gradients never read it.  They are therefore issued on an auxiliary stream forked from the caller's stream
// (event record / wait: graph-capturable, so inside a captured train step they become parallel branches) and joined before
// the call returns.  One auxiliary stream per (thread, device, caller stream), created on first use.
// A second auxiliary stream (which = 1) carries the per-frame-vector chain that ends in d obar (aq1 / aq2 / temporal-gate
// backward, ~10 small launches): its result is not needed before the dV products of the token-attention backward.
struct SideStream { cudaStream_t s; cudaEvent_t ev; };
inline SideStream* side_stream_for(cudaStream_t main, int which = 0) {
  struct Key { int dev; cudaStream_t st; int which; SideStream side; };
  static thread_local std::deque<Key> cache;      // (deque: references handed out stay valid when it grows)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (Key& k : cache)
    if (k.dev == dev && k.st == main && k.which == which) return &k.side;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(main, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;   // never create inside a capture
  SideStream sd{};
  if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&sd.ev, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(sd.s); return nullptr; }
  cache.push_back(Key{dev, main, which, sd});
  return &cache.back().side;
}


}  // namespace dgsct

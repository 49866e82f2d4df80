#include "profile.h"

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/dgsct.h"

namespace dgsct {

static const char* kNames[ST_COUNT] = {
    "resample_conv", "resample_fc", "tok_logits", "col_softmax", "tok_update", "x_logits", "row_softmax", "x_update",
    "small_linear", "channel_query", "spatial_query", "spatial_map", "gate_ln", "down_sampler", "bn_finalize",
    "up_sampler", "out_ln", "misc",
    "bwd_out", "bwd_updown", "bwd_gate_ln", "bwd_spatial", "bwd_channel", "bwd_small", "bwd_token", "bwd_fc", "bwd_dwres",
    "bwd_dother", "bwd_dw", "adam"};

const char* stage_name(int s) { return (s >= 0 && s < ST_COUNT) ? kNames[s] : "?"; }

struct Rec { int stage; cudaEvent_t a, b; };
static std::atomic<int> g_on{0};
static std::atomic<long long> g_launches{0};
static std::mutex g_mu;
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static thread_local cudaEvent_t t_open = nullptr;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool prof_enabled() { return g_on.load(std::memory_order_relaxed) != 0; }

void prof_begin(int, cudaStream_t st) {
  if (!g_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  t_open = get_event();
  cudaEventRecord(t_open, st);
}

void prof_end(int stage, cudaStream_t st) {
  if (!g_on.load(std::memory_order_relaxed) || !t_open) return;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaEvent_t b = get_event();
  cudaEventRecord(b, st);
  g_recs.push_back(Rec{stage, t_open, b});
  t_open = nullptr;
}

}  // namespace dgsct

using namespace dgsct;

extern "C" {

int dgsct_num_stages(void) { return ST_COUNT; }
const char* dgsct_stage_name(int s) { return stage_name(s); }
long long dgsct_launch_count(void) { return g_launches.load(); }

int dgsct_profile_enable(int on) {
  g_on.store(on ? 1 : 0);
  return DGSCT_OK;
}

int dgsct_profile_collect(double* ms, long long* calls, int n) {
  if (!ms || !calls || n < ST_COUNT) return DGSCT_ERR_NULL;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; calls[i] = 0; }
  for (Rec& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return DGSCT_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return DGSCT_ERR_CUDA;
    ms[r.stage] += t; calls[r.stage] += 1;
    g_pool.push_back(r.a); g_pool.push_back(r.b);
  }
  g_recs.clear();
  return DGSCT_OK;
}

}  // extern "C"

// Built-in stage profiler: when enabled, every pipeline stage is bracketed by CUDA events on the launching
// stream; bench.py reads the per-stage device time (dgsct_profile_collect) for the roofline of the dominant
// kernel.  The launch counter is always on (bench.py's `gpu_launches`).
#pragma once
#include <cuda_runtime.h>

namespace dgsct {

enum Stage {
  ST_RESAMPLE_CONV = 0, ST_RESAMPLE_FC, ST_TOK_LOGITS, ST_COL_SOFTMAX, ST_TOK_UPDATE, ST_X_LOGITS, ST_ROW_SOFTMAX,
  ST_X_UPDATE, ST_SMALL_LINEAR, ST_CHANNEL_QUERY, ST_SPATIAL_QUERY, ST_SPATIAL_MAP, ST_GATE_LN, ST_DOWN, ST_BN_FINALIZE,
  ST_UP, ST_OUT_LN, ST_MISC,
  ST_BWD_OUT, ST_BWD_UPDOWN, ST_BWD_GATE_LN, ST_BWD_SPATIAL, ST_BWD_CHANNEL, ST_BWD_SMALL, ST_BWD_TOKEN, ST_BWD_FC, ST_BWD_DWRES,
  ST_BWD_DOTHER, ST_BWD_DW, ST_ADAM, ST_COUNT
};

const char* stage_name(int s);
void prof_begin(int stage, cudaStream_t st);
void prof_end(int stage, cudaStream_t st);
void count_launch(int n = 1);
bool prof_enabled();

struct StageScope {
  int stage; cudaStream_t st;
  StageScope(int s, cudaStream_t stream) : stage(s), st(stream) { prof_begin(stage, st); }
  ~StageScope() { prof_end(stage, st); }
};

}  // namespace dgsct

// Declarations shared between the kernels' translation units and the C-ABI layer.
#pragma once
#include "device.cuh"
#include "llh_math.cuh"

namespace krepp {

// Per-stage CUDA events on the slot's stream (measurement only: bench.py reads them through krepp_batch_stage_times).
struct StageClock {
  static constexpr int kMax = 24;
  cudaEvent_t ev[kMax] = {};
  const char* name[kMax] = {};
  int n = 0;
  void tick(const char* nm, cudaStream_t s) { if (n < kMax && ev[n]) { cudaEventRecord(ev[n], s); name[n++] = nm; } }
};

int match_resident_warps(int device, uint32_t k, bool staged);
cudaError_t launch_match(const DevIndex& ix, const MatchArgs& a, int resident_warps, bool staged, bool tap, cudaStream_t stream);
// sorted.cu: the bucket-sorted form of the match stage (same outputs as launch_match)
int sorted_resolve_warps(int sms);
int sorted_resolve_warps_per_cta();
// begin[i] = sum of in[0..i), begin[n] = the total; cursor (optional) = copy of begin[0..n); partials: n / 4096 + 2 words of scratch
cudaError_t exclusive_scan(const uint32_t* in, uint32_t n, uint32_t* partials, uint32_t* begin, uint32_t* cursor, cudaStream_t stream);
cudaError_t launch_match_sorted(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream, uint32_t* launches, StageClock* clk = nullptr);
// mode B: the chain of launch_match_sorted cut at its two exchange points
cudaError_t launch_shard_lookup(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream, StageClock* clk = nullptr);
cudaError_t launch_shard_join(const DevIndex& ix, const SortArgs& s, uint32_t th, uint32_t* counters, unsigned long long* stats, int sms, cudaStream_t stream);
cudaError_t launch_shard_finish(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, cudaStream_t stream, StageClock* clk = nullptr);
cudaError_t launch_solve(const SolveArgs& a, const LlhTables& tab, int sms, cudaStream_t stream, StageClock* clk = nullptr);
cudaError_t launch_segment_combine(const SegArgs& g, int ctas, cudaStream_t stream); // ctas * 8 warps, each with its own slice of g.scratch
cudaError_t launch_seek(const SolveArgs& a, const LlhTables& tab, double* out, int sms, cudaStream_t stream); // after launch_solve, sketch handles only
constexpr int kPlaceWarpsPerCta = 4;
cudaError_t launch_place(const PlaceArgs& a, const LlhTables& tab, int grid, int sms, cudaStream_t stream, StageClock* clk = nullptr);

} // namespace krepp

// sorted.cu -- the bucket-sorted form of K1..K3 (sm_100a): the same lookups, bucket scans, colour expansion and Hamming
// histograms as the fused kernel of match.cu, organised so that every bucket of the index is read from HBM ONCE per
// batch instead of once per lookup.
//
//   L1  lookup_kernel<COUNT>    warp per read: bases -> eligible lookups of both strands (tile_lookups, the A0/A1 stages
//                               of match.cu; ref src/query.cpp:40-94, src/lshf.cpp:62-69, src/index.hpp:27); counts the
//                               lookups of every bucket row; per-read onmers / wnmers
//   S1  scan                    row counts -> row ranges of the batch's lookup list
//   L2  lookup_kernel<SCATTER>  the same lookups again (cheaper than storing them), written into their row's range as
//                               {q, read, local lookup index | strand}: a counting sort by LSH bucket
//   J   join_kernel             warp per bucket row: the row's index entries sit in registers (one coalesced read of
//                               cmer), the row's queries stream past them from shared memory; XOR / OR / popc
//                               (ref src/common.hpp:175, src/query.cpp:361-368); hit entries {read, lookup, colour, hd}
//                               are queued per warp and appended to a batch-wide list
//   S2  scan + hit_scatter      hit entries grouped by read (counting sort by read), colour id -> its flattened leaf list
//   R   resolve_kernel          warp per read: hit entries -> (strand, leaf, lookup, hd) keys -> bitonic sort (in registers
//                               with shuffles up to 256 keys, in shared memory up to 1,024, in HBM scratch beyond) ->
//                               per (strand, leaf): one count per lookup at its minimum distance
//                               (Minfo::update_match, ref src/query.hpp:153-176), per-strand hdist_filt and its gate
//                               (ref src/query.cpp:101-106,116-119), records in (strand, leaf) order
//
// Everything is integer work and bit-exact with the fused kernel (same records in the same order).  Batches or reads
// that do not fit this pipeline's buffers raise a flag; the host then grows the buffer or redoes the batch with the
// fused kernel (api.cu).  For an index split by bucket range over several GPUs (SURVEY.md 8e mode B) the same chain is cut
// after L2 and after J, where lookups and hit entries change owner (launch_shard_lookup / _join / _finish at the end).
#include "device.cuh"
#include "match_common.cuh"
#include "solve.cuh"

#include <algorithm>
#include <cstdlib>

namespace krepp {

constexpr int kLkWarps = 16;           // lookup kernel: warps per CTA (two CTAs per SM beside the 28-32 kB LUT)
constexpr uint32_t kLkClaim = 8;       // reads claimed per atomic
constexpr int kJoinWarps = 8;
constexpr uint32_t kRowClaim = 32;     // bucket rows claimed per atomic: one per lane
constexpr int kJoinChunk = 128;        // index entries held in registers per pass: four per lane
constexpr int kHitQ = 128;             // per-warp queue of hit entries; flushed once fewer than 32 slots (one ballot's worth) are free
constexpr int kResWarps = 8;
constexpr int kResKeys = 1024;         // 32-bit sort keys per warp in shared memory (half as many 64-bit keys)
constexpr uint32_t kResClaim = 4;
constexpr uint32_t kMaxLoc = 1u << 26; // local lookup indices must fit bits 5..30 of the hit word

// ---------------------------------------------------------------------------------------------------- L1 / L2

template <bool SCATTER, bool TAP>
__global__ void __launch_bounds__(kLkWarps * 32, 2) lookup_kernel(const DevIndex ix, const MatchArgs a, const SortArgs s)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nchunks = lut_chunks(ix.k);
  uint4* lut = reinterpret_cast<uint4*>(smem_raw);
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_raw + nchunks * 256 * sizeof(uint4));
  if (SCATTER && s.row_begin[s.nrows] > s.cap_lookups) { // the lookup list does not fit: the host grows it and runs the batch again
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.counters + 2, kErrLookupOverflow);
    return;
  }
  for (uint32_t i = threadIdx.x; i < nchunks * 256; i += blockDim.x) lut[i] = ix.lut[i];
  __syncthreads();
  const bool wide = nchunks > 7;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, k = ix.k;
  WarpSmem& sm = smem[warp];
  uint32_t* claimctr = s.sc + (SCATTER ? 3 : 2);
  unsigned long long st_bytes = 0, st_lookups = 0;
  uint32_t claim = 0, claim_end = 0;
  for (;;) {
    if (claim == claim_end) {
      if (lane == 0) claim = atomicAdd(claimctr, kLkClaim);
      claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
      claim_end = min(claim + kLkClaim, a.n_reads);
      if (claim >= a.n_reads) break;
    }
    const uint32_t read = claim++;
    uint64_t off, len;
    read_span(a, read, off, len);
    uint32_t onmers = 0, wn0 = 0, wn1 = 0, loc = 0;
    for (uint64_t t0 = 0; t0 + k <= len; t0 += kTileWindows) {
      const uint32_t nl = tile_lookups<TAP && !SCATTER>(ix, a, sm, lut, wide, read, off, len, t0, onmers, wn0, wn1);
      if (loc + nl >= kMaxLoc) { if (lane == 0) atomicOr(a.counters + 2, kErrSortFallback); break; }
      if (!SCATTER) {
        for (uint32_t i = lane; i < nl; i += 32) atomicAdd(&s.row_count[sm.lk_a[i] & 0x7FFFFFFFu], 1u);
      } else {
        // the cursor atomics return positions: four are kept in flight per lane before the stores that need them
        for (uint32_t i0 = lane; i0 < nl; i0 += 128) {
          uint32_t ob[4], pos[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + 32 * u;
            ob[u] = 0; pos[u] = 0;
            if (i < nl) { ob[u] = sm.lk_a[i]; pos[u] = atomicAdd(&s.row_cursor[ob[u] & 0x7FFFFFFFu], 1u); }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + 32 * u;
            if (i < nl) s.tuples[pos[u]] = make_uint4(sm.lk_q[i], read, (loc + i) | (ob[u] & 0x80000000u), 0u);
          }
        }
      }
      loc += nl;
      __syncwarp();
    }
    if (!SCATTER && lane == 0) {
      a.onmers[read] = onmers; a.wn[2 * read] = wn0; a.wn[2 * read + 1] = wn1;
      st_bytes += len; st_lookups += wn0 + wn1;
    }
  }
  if (!SCATTER && lane == 0 && (st_bytes | st_lookups)) { atomicAdd(a.stats, st_bytes + 16ull * st_lookups); atomicAdd(a.stats + 1, st_lookups); }
}

// ---------------------------------------------------------------------------------------------------- S1 / S2: exclusive scan

constexpr int kScanThreads = 256, kScanItems = 16, kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& block_total)
{
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t before = 0, total = 0;
  for (uint32_t w = 0; w < nwarps; ++w) { const uint32_t x = warp_sums[w]; if (w < warp) before += x; total += x; }
  __syncthreads();
  block_total = total;
  return before + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ partials)
{
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const uint32_t base = blockIdx.x * kScanTile;
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { const uint32_t j = base + i * kScanThreads + threadIdx.x; if (j < n) v += in[j]; }
  uint32_t total;
  block_exclusive_scan(v, warp_sums, total);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_top_kernel(uint32_t* partials, uint32_t nb)
{
  __shared__ uint32_t warp_sums[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t j = base + threadIdx.x;
    const uint32_t v = j < nb ? partials[j] : 0u;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, warp_sums, total);
    if (j < nb) partials[j] = carry + ex;
    carry += total;
  }
}

// begin[i] = sum of in[0..i), begin[n] = the total; cursor (optional) starts as a copy of begin[0..n)
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ partials,
                                                                  uint32_t* __restrict__ begin, uint32_t* __restrict__ cursor)
{
  __shared__ uint32_t tile[kScanTile + kScanTile / 32];
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const uint32_t base = blockIdx.x * kScanTile;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const uint32_t t = i * kScanThreads + threadIdx.x, j = base + t;
    tile[t + (t >> 5)] = j < n ? in[j] : 0u;
  }
  __syncthreads();
  uint32_t v[kScanItems], sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { const uint32_t t = threadIdx.x * kScanItems + i; v[i] = tile[t + (t >> 5)]; sum += v[i]; }
  uint32_t total;
  uint32_t run = partials[blockIdx.x] + block_exclusive_scan(sum, warp_sums, total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) { const uint32_t t = threadIdx.x * kScanItems + i; tile[t + (t >> 5)] = run; run += v[i]; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const uint32_t t = i * kScanThreads + threadIdx.x, j = base + t;
    if (j < n) { const uint32_t x = tile[t + (t >> 5)]; begin[j] = x; if (cursor) cursor[j] = x; }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) begin[n] = partials[blockIdx.x] + total;
}

cudaError_t exclusive_scan(const uint32_t* in, uint32_t n, uint32_t* partials, uint32_t* begin, uint32_t* cursor, cudaStream_t stream)
{
  const uint32_t nb = (n + kScanTile - 1) / kScanTile;
  scan_sums_kernel<<<nb, kScanThreads, 0, stream>>>(in, n, partials);
  scan_top_kernel<<<1, 1024, 0, stream>>>(partials, nb);
  scan_apply_kernel<<<nb, kScanThreads, 0, stream>>>(in, n, partials, begin, cursor);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------- L1' / L2': two-level sort
//
// The two passes above sort the lookups by row with one global atomic per lookup per pass; the second pass needs the value the
// atomic returns, and an SM can keep only so many of those in flight (it runs at a third of the speed of the first pass, which
// does the same work with fire-and-forget reductions).  The two-level form computes every lookup ONCE and needs no per-lookup
// global atomic:
//
//   L1' lookup_partition_kernel  two CTAs of 16 warps per SM.  A round = every warp turns one tile of a read into lookups (in its
//                                own shared-memory segment), then the CTA appends the round's lookups to at most 512 coarse bins
//                                (2^bin_shift rows each): arrivals are counted per bin with shared-memory atomics, joined to the few
//                                tuples the bin still holds from earlier rounds, and leave for the bin's region in HBM in whole
//                                64-byte lines at a position taken with ONE global atomic per bin per round; the remainder (fewer
//                                than a line) waits in shared memory for the next round.
//   L2' bin_sort_kernel          CTA per bin: counting sort of the bin's tuples by row through shared-memory counters (histogram,
//                                scan, scatter with shared-memory cursors) into the dense row-grouped list, and the rows' ranges --
//                                exactly what the two-pass form leaves for the join (and, in mode B, for the exchange).
//
// Bins have room for a quarter more than an even share of the batch's lookups; reads that pile their lookups on few rows
// (low-complexity input) can overflow one, which is flagged -- the host then runs the batch through the two-pass form.

constexpr int kLpWarps = 16;           // two CTAs per SM: one computes tiles while the other sits in the barriers of its append
constexpr int kLpBinsMax = 512;
constexpr int kLpLine = 4;             // tuples per line sent to a bin (64 bytes)
constexpr int kBsThreads = 1024;       // one CTA per SM, bins one after the other: few bins are open at a time, so the partly
                                       // written sectors at the tail of every row stay in L2 until their second half arrives
constexpr uint32_t kBsRowsMax = 8192;  // rows per bin the bin sort's shared-memory counters hold

struct LpShared {
  WarpSmemLite w[kLpWarps];
  uint4 stage[kLpBinsMax][kLpLine];    // per bin: tuples waiting for a full line
  uint32_t cnt[kLpBinsMax];            // arrivals of the round (counted by the warps as they find their lookups)
  uint32_t rank[kLpBinsMax];           // the cursor that ranks the arrivals
  uint32_t left[kLpBinsMax];           // tuples waiting in stage[]
  uint32_t gbase[kLpBinsMax];          // where the bin's lines of this round start in its region
  uint32_t lim[kLpBinsMax];            // tuples of this round's lines (a multiple of kLpLine)
  uint32_t nl[kLpWarps], rd[kLpWarps], locb[kLpWarps];
};

template <bool TAP>
__global__ void __launch_bounds__(kLpWarps * 32, 2) lookup_partition_kernel(const DevIndex ix, const MatchArgs a, const SortArgs s)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nchunks = lut_chunks(ix.k);
  uint4* lut = reinterpret_cast<uint4*>(smem_raw);
  LpShared& sh = *reinterpret_cast<LpShared*>(smem_raw + nchunks * 256 * sizeof(uint4));
  for (uint32_t i = threadIdx.x; i < nchunks * 256; i += blockDim.x) lut[i] = ix.lut[i];
  for (uint32_t b = threadIdx.x; b < (uint32_t)kLpBinsMax; b += blockDim.x) { sh.cnt[b] = 0; sh.left[b] = 0; }
  __syncthreads();
  const bool wide = nchunks > 7;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, k = ix.k;
  WarpSmemLite& sm = sh.w[warp];
  unsigned long long st_bytes = 0, st_lookups = 0;
  uint32_t claim = 0, claim_end = 0;
  bool have = false, exhausted = false;
  uint32_t read = 0, onmers = 0, wn0 = 0, wn1 = 0, loc = 0;
  uint64_t off = 0, len = 0, t0 = 0;
  for (;;) {
    // ---- every warp: the next tile of its read (a read shorter than k has none and is finished at once)
    while (!have && !exhausted) {
      if (claim == claim_end) {
        if (lane == 0) claim = atomicAdd(s.sc + 2, kLkClaim);
        claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
        claim_end = min(claim + kLkClaim, a.n_reads);
        if (claim >= a.n_reads) { exhausted = true; break; }
      }
      read = claim++;
      read_span(a, read, off, len);
      onmers = wn0 = wn1 = loc = 0; t0 = 0;
      if (len >= k) have = true;
      else if (lane == 0) { a.onmers[read] = 0; a.wn[2 * read] = 0; a.wn[2 * read + 1] = 0; st_bytes += len; }
    }
    uint32_t nlk = 0;
    const bool did = have;
    if (have) {
      nlk = tile_lookups<TAP>(ix, a, sm, lut, wide, read, off, len, t0, onmers, wn0, wn1, sh.cnt, s.bin_shift);
      if (loc + nlk >= kMaxLoc) { if (lane == 0) atomicOr(a.counters + 2, kErrSortFallback); t0 = len; } // the read cannot be indexed: the host redoes the batch
      if (lane == 0) { sh.nl[warp] = nlk; sh.rd[warp] = read; sh.locb[warp] = loc; }
      loc += nlk;
      t0 += kTileWindows;
      if (t0 + k > len) { // the read is finished
        if (lane == 0) { a.onmers[read] = onmers; a.wn[2 * read] = wn0; a.wn[2 * read + 1] = wn1; st_bytes += len; st_lookups += wn0 + wn1; }
        have = false;
      }
    } else if (lane == 0) sh.nl[warp] = 0;
    if (!__syncthreads_or(did ? 1 : 0)) break; // no warp had a tile: all reads are done
    // ---- b. per bin: whole lines leave; their place in the bin's region; the tuples that waited go first
    for (uint32_t b = threadIdx.x; b < s.nbins; b += blockDim.x) {
      const uint32_t l = sh.left[b], t = l + sh.cnt[b], out = t - t % kLpLine;
      uint32_t g = 0;
      if (out) {
        g = atomicAdd(&s.bin_cursor[b], out);
        if ((uint64_t)g + out > s.bin_cap) atomicOr(a.counters + 2, kErrBinOverflow);
        uint4* dst = s.binned + (size_t)b * s.bin_cap;
        for (uint32_t i = 0; i < l; ++i) if (g + i < s.bin_cap) dst[g + i] = sh.stage[b][i];
      }
      sh.gbase[b] = g; sh.lim[b] = out; sh.rank[b] = l; sh.left[b] = t - out; sh.cnt[b] = 0;
    }
    __syncthreads();
    // ---- c. every arrival takes its rank in its bin: into one of the bin's lines, or into the waiting slots
    for (uint32_t slot = threadIdx.x; slot < (uint32_t)(kLpWarps * kMaxLookups); slot += blockDim.x) {
      const uint32_t w = slot / kMaxLookups, i = slot % kMaxLookups;
      if (i < sh.nl[w]) {
        const uint32_t ob = sh.w[w].lk_a[i], row = ob & 0x7FFFFFFFu, b = row >> s.bin_shift;
        const uint4 tup = make_uint4(sh.w[w].lk_q[i], sh.rd[w], (sh.locb[w] + i) | (ob & 0x80000000u), row);
        const uint32_t c = atomicAdd(&sh.rank[b], 1u), lim = sh.lim[b];
        if (c < lim) { const uint32_t g = sh.gbase[b] + c; if (g < s.bin_cap) s.binned[(size_t)b * s.bin_cap + g] = tup; }
        else sh.stage[b][c - lim] = tup;
      }
    }
    __syncthreads(); // the next round's tiles count into cnt[] (zeroed in b) and overwrite the warps' lookup lists
  }
  // ---- the tuples still waiting leave as short lines
  for (uint32_t b = threadIdx.x; b < s.nbins; b += blockDim.x) {
    const uint32_t l = sh.left[b];
    if (!l) continue;
    const uint32_t g = atomicAdd(&s.bin_cursor[b], l);
    if ((uint64_t)g + l > s.bin_cap) atomicOr(a.counters + 2, kErrBinOverflow);
    uint4* dst = s.binned + (size_t)b * s.bin_cap;
    for (uint32_t i = 0; i < l; ++i) if (g + i < s.bin_cap) dst[g + i] = sh.stage[b][i];
  }
  if (lane == 0 && (st_bytes | st_lookups)) { atomicAdd(a.stats, st_bytes + 16ull * st_lookups); atomicAdd(a.stats + 1, st_lookups); }
}

__global__ void __launch_bounds__(kBsThreads) bin_sort_kernel(const SortArgs s, uint32_t* counters)
{
  extern __shared__ uint32_t bs_smem[];
  uint32_t* hist = bs_smem;                              // [rows per bin] counts, then cursors
  uint32_t* bin_begin = bs_smem + (1u << s.bin_shift);   // [nbins + 1] exclusive prefix of the bin sizes
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t claimed;
  const uint32_t tid = threadIdx.x, rows_per_bin = 1u << s.bin_shift;
  if (counters[2] & kErrRedo) return;
  { // every CTA scans the (at most 1,024) bin sizes itself
    const uint32_t per = (s.nbins + kBsThreads - 1) / kBsThreads, lo = min(tid * per, s.nbins), hi = min(lo + per, s.nbins);
    uint32_t sum = 0;
    for (uint32_t b = lo; b < hi; ++b) sum += s.bin_cursor[b];
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, warp_sums, total);
    for (uint32_t b = lo; b < hi; ++b) { bin_begin[b] = run; run += s.bin_cursor[b]; }
    if (tid == 0) bin_begin[s.nbins] = total;
    __syncthreads();
    if (total > s.cap_lookups) { // the dense list does not fit: the host grows it (s.row_begin[nrows] is the demand) and runs the batch again
      if (blockIdx.x == 0 && tid == 0) { s.row_begin[s.nrows] = total; atomicOr(counters + 2, kErrLookupOverflow); }
      return;
    }
  }
  for (;;) {
    __syncthreads();
    if (tid == 0) claimed = atomicAdd(s.sc + 6, 1u);
    __syncthreads();
    const uint32_t b = claimed;
    if (b >= s.nbins) break;
    const uint32_t n = s.bin_cursor[b], dst0 = bin_begin[b], row0 = b << s.bin_shift, nr = min(rows_per_bin, s.nrows - row0);
    const uint4* src = s.binned + (size_t)b * s.bin_cap;
    for (uint32_t r = tid; r < nr; r += kBsThreads) hist[r] = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 8 * kBsThreads) { // eight loads in flight per thread: the pass is bound by memory latency
      uint32_t rw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + u * kBsThreads + tid; rw[u] = i < n ? __ldg(reinterpret_cast<const uint32_t*>(src + i) + 3) : 0xFFFFFFFFu; }
#pragma unroll
      for (int u = 0; u < 8; ++u) if (rw[u] != 0xFFFFFFFFu) atomicAdd(&hist[rw[u] - row0], 1u);
    }
    __syncthreads();
    { // exclusive scan of the row counts; hist[] becomes the rows' cursors into the dense list
      const uint32_t per = (nr + kBsThreads - 1) / kBsThreads, lo = min(tid * per, nr), hi = min(lo + per, nr);
      uint32_t sum = 0;
      for (uint32_t r = lo; r < hi; ++r) sum += hist[r];
      uint32_t total;
      uint32_t run = dst0 + block_exclusive_scan(sum, warp_sums, total);
      for (uint32_t r = lo; r < hi; ++r) { const uint32_t c = hist[r]; hist[r] = run; s.row_begin[row0 + r] = run; run += c; }
      if (b == s.nbins - 1 && tid == 0) s.row_begin[s.nrows] = dst0 + n;
    }
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 4 * kBsThreads) {
      uint4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + u * kBsThreads + tid; t[u] = i < n ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0xFFFFFFFFu); }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t[u].w != 0xFFFFFFFFu) { const uint32_t pos = atomicAdd(&hist[t[u].w - row0], 1u); s.tuples[pos] = make_uint4(t[u].x, t[u].y, t[u].z, 0u); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- J: join

struct __align__(16) JoinWarpSmem {
  uint2 q2[32];       // the queries being compared, split into their bit-planes: {q low half, q high half}; read two at a time
  uint2 meta[32];     // {read, strand << 31 | local lookup index << 5}
  uint4 hq[kHitQ];    // queued hit entries: {read, strand << 31 | local lookup index << 5 | hd, colour id, 0}
};

// appends the warp's n queued hit entries to the batch-wide list
__device__ __noinline__ void join_flush(JoinWarpSmem* w, const SortArgs s, uint32_t* counters, uint32_t n)
{
  const uint32_t lane = threadIdx.x & 31;
  __syncwarp();
  if (n) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(s.sc, n);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if ((uint64_t)base + n > s.cap_hits && lane == 0) atomicOr(counters + 2, (uint64_t)base + n > 0xFFFFFFF0ull ? (kErrHitOverflow | kErrHitWrap) : kErrHitOverflow); // s.sc[0] ends as the demand
    for (uint32_t i = lane; i < n; i += 32) if (base + i < s.cap_hits) s.hits_tmp[base + i] = w->hq[i];
  }
  __syncwarp();
}

__device__ __forceinline__ int min_of(const int (&p)[1]) { return p[0]; }
__device__ __forceinline__ int min_of(const int (&p)[2]) { return min(p[0], p[1]); }
__device__ __forceinline__ int min_of(const int (&p)[3]) { return __vimin3_s32(p[0], p[1], p[2]); }
__device__ __forceinline__ int min_of(const int (&p)[4]) { return min(__vimin3_s32(p[0], p[1], p[2]), p[3]); }

// One block of up to 32 queries (in w->q2 / w->meta) against the E entries each lane holds (nv of them real; the others are
// all-ones words, whose distance to any query is at least 16).  The residual encoding keeps bit 0 of the 16 kept positions in
// its low half and bit 1 in its high half (ref src/lshf.cpp:64-69), so the mismatch mask is (e.lo ^ q.lo) | (e.hi ^ q.hi): with
// both halves of entries and queries split once, a comparison is two LOP3 and one POPC.  Two queries per trip; per query the
// minimum over the lane's entries is what is tested, so the loop carries one ISETP per query and one vote per trip.  Hits are
// rare per comparison (a few per thousand) but not per trip (96 entries x 2 queries), so what follows a vote is kept short:
// one more vote per query, and only for a query with a hit the E comparisons again, each compacted into the warp's queue by
// ballot (the queue length lives in a register; no shared-memory atomics).
template <int E, bool COUNT>
__device__ __forceinline__ void join_block(JoinWarpSmem* w, const SortArgs& s, uint32_t* counters, const uint32_t (&elo)[4], const uint32_t (&ehi)[4],
                                           const uint32_t (&se)[4], uint32_t nv, uint32_t cnt, int th, uint32_t& hq_n)
{
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
  for (uint32_t j = 0; j < cnt; j += 2) {
    const uint4 qq = *reinterpret_cast<const uint4*>(&w->q2[j]); // {q[j].lo, q[j].hi, q[j+1].lo, q[j+1].hi}
    int p0[E], p1[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      p0[e] = __popc((elo[e] ^ qq.x) | (ehi[e] ^ qq.y));
      p1[e] = __popc((elo[e] ^ qq.z) | (ehi[e] ^ qq.w));
    }
    const bool h0 = min_of(p0) <= th, h1 = min_of(p1) <= th;
    if (__any_sync(0xFFFFFFFFu, h0 | h1)) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (j + q < cnt && __any_sync(0xFFFFFFFFu, q ? h1 : h0)) {
          const uint2 mt = w->meta[j + q];
          const uint32_t qlo = q ? qq.z : qq.x, qhi = q ? qq.w : qq.y;
          uint32_t total = 0;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int hd = __popc((elo[e] ^ qlo) | (ehi[e] ^ qhi));
            const bool hit = hd <= th && (uint32_t)e < nv;
            const uint32_t bm = __ballot_sync(0xFFFFFFFFu, hit);
            if (bm) {
              if (hit) w->hq[hq_n + __popc(bm & lt_mask)] = make_uint4(mt.x, mt.y | (uint32_t)hd, se[e], 0u);
              hq_n += __popc(bm); total += __popc(bm);
              if (hq_n > (uint32_t)(kHitQ - 32)) { join_flush(w, s, counters, hq_n); hq_n = 0; }
            }
          }
          if (COUNT && total && lane == 0) atomicAdd(&s.hit_count[mt.x], total);
        }
      }
    }
  }
}

// COUNT: also count the hit entries of every read (the batch's own reads).  Without it the rows are those of a bucket-range
// shard and the queries another rank's (SURVEY.md 8e mode B): s.row_begin is then the sender's slice for these rows, still
// holding positions in the sender's list, so everything is taken relative to its first element.
template <bool COUNT>
__global__ void __launch_bounds__(kJoinWarps * 32, 5) join_kernel(const DevIndex ix, const SortArgs s, uint32_t th, uint32_t* counters, unsigned long long* stats)
{
  __shared__ JoinWarpSmem jsm[kJoinWarps];
  if (counters[2] & (kErrBinOverflow | kErrLookupOverflow | kErrSortFallback)) return; // the lookup list is incomplete: the host runs the batch again
  const uint32_t rb0 = s.row_begin[0];
  if (s.row_begin[s.nrows] - rb0 > s.cap_lookups) return; // flagged by the lookup kernel
  const uint32_t lane = threadIdx.x & 31;
  JoinWarpSmem* w = &jsm[threadIdx.x >> 5];
  uint32_t hq_n = 0; // entries in the warp's hit queue (warp-uniform)
  unsigned long long st_entries = 0;
  for (;;) {
    uint32_t r0 = 0;
    if (lane == 0) r0 = atomicAdd(s.sc + 1, kRowClaim);
    r0 = __shfl_sync(0xFFFFFFFFu, r0, 0);
    if (r0 >= s.nrows) break;
    const uint32_t row = r0 + lane; // every lane describes one row of the claim
    uint32_t qb = 0, qe = 0, eb = 0, ee = 0;
    if (row < s.nrows) {
      qb = s.row_begin[row] - rb0; qe = s.row_begin[row + 1] - rb0;
      if (qe > qb) { eb = row ? __ldg(&ix.inc32[row - 1]) : 0u; ee = __ldg(&ix.inc32[row]); } // ref src/index.cpp:160-168, src/table.hpp:121-136
    }
    st_entries += (unsigned long long)(qe - qb) * (ee - eb);
    uint32_t active = __ballot_sync(0xFFFFFFFFu, qe > qb && ee > eb);
    while (active) {
      const int src = __ffs(active) - 1;
      active &= active - 1;
      const uint32_t rqb = __shfl_sync(0xFFFFFFFFu, qb, src), nq = __shfl_sync(0xFFFFFFFFu, qe, src) - rqb;
      const uint32_t reb = __shfl_sync(0xFFFFFFFFu, eb, src), ne = __shfl_sync(0xFFFFFFFFu, ee, src) - reb;
      for (uint32_t c = 0; c < ne; c += kJoinChunk) {
        uint32_t elo[4], ehi[4], se[4], nv = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t i = c + 32 * e + lane;
          elo[e] = 0xFFFFFFFFu; ehi[e] = 0xFFFFFFFFu; se[e] = 0;
          if (i < ne) { const uint2 v = __ldg(&ix.cmer[(size_t)reb + i]); elo[e] = v.x & 0xFFFFu; ehi[e] = v.x >> 16; se[e] = v.y; nv = e + 1; }
        }
        const uint32_t E = min(4u, (ne - c + 31u) >> 5);
        for (uint32_t q0 = 0; q0 < nq; q0 += 32) {
          const uint32_t cnt = min(32u, nq - q0);
          __syncwarp();
          if (lane < cnt) {
            const uint4 t = s.tuples[rqb + q0 + lane]; // {q, read, local lookup index | strand << 31, 0}
            w->q2[lane] = make_uint2(t.x & 0xFFFFu, t.x >> 16);
            w->meta[lane] = make_uint2(t.y, (t.z & 0x80000000u) | ((t.z & (kMaxLoc - 1u)) << 5));
          } else if (lane == cnt) w->q2[lane] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); // the odd query's partner: at distance >= 16 of everything
          __syncwarp();
          switch (E) {
            case 1: join_block<1, COUNT>(w, s, counters, elo, ehi, se, nv, cnt, (int)th, hq_n); break;
            case 2: join_block<2, COUNT>(w, s, counters, elo, ehi, se, nv, cnt, (int)th, hq_n); break;
            case 3: join_block<3, COUNT>(w, s, counters, elo, ehi, se, nv, cnt, (int)th, hq_n); break;
            default: join_block<4, COUNT>(w, s, counters, elo, ehi, se, nv, cnt, (int)th, hq_n); break;
          }
        }
      }
    }
  }
  join_flush(w, s, counters, hq_n);
  for (int o = 16; o; o >>= 1) st_entries += __shfl_xor_sync(0xFFFFFFFFu, st_entries, o);
  if (lane == 0 && st_entries) { atomicAdd(stats, 8ull * st_entries); atomicAdd(stats + 2, st_entries); }
}

// ---------------------------------------------------------------------------------------------------- S2: hits by read

__global__ void __launch_bounds__(256) hit_scatter_kernel(const DevIndex ix, const SortArgs s, const uint32_t* counters)
{
  if (counters[2] & kErrRedo) return;
  const uint32_t n = s.sc[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 h = s.hits_tmp[i];
    const uint32_t pos = atomicAdd(&s.hit_cursor[h.x], 1u);
    const uint32_t cs = __ldg(&ix.cbeg[h.z]), ce = __ldg(&ix.cbeg[h.z + 1]);
    s.hits[pos] = make_uint4(cs, ce - cs, h.y, 0u);
  }
}

// mode B, home side: the hit entries came back from the shard owners in no particular order; count them per read
__global__ void __launch_bounds__(256) hit_count_kernel(const SortArgs s, uint32_t n_reads, uint32_t* counters)
{
  const uint32_t n = s.sc[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t read = s.hits_tmp[i].x;
    if (read < n_reads) atomicAdd(&s.hit_count[read], 1u);
    else atomicOr(counters + 2, kErrShardData); // not a hit entry of this batch
  }
}

// ---------------------------------------------------------------------------------------------------- R: resolve

template <typename K>
__device__ __forceinline__ void bitonic_sort(K* keys, uint32_t n) // n: power of two >= 32; ascending
{
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t k = 2; k <= n; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = lane; t < (n >> 1); t += 32) {
        const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u)), l = i | j;
        const K x = keys[i], y = keys[l];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { keys[i] = y; keys[l] = x; }
      }
      __syncwarp();
    }
  }
}

// The same network with the keys in registers, E per lane (key i sits in lane i / E, slot i % E): exchanges between lanes
// are shuffles, exchanges inside a lane are register swaps.  n = 32 * E.
template <typename K, int E>
__device__ __forceinline__ void bitonic_sort_regs(K* keys)
{
  const uint32_t lane = threadIdx.x & 31;
  K v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = keys[lane * E + e];
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= E) { // partner in lane ^ (j / E), same slot
        const bool lower = (lane & (uint32_t)(j / E)) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const bool up = ((lane * E + e) & (uint32_t)k) == 0;
          const K o = __shfl_xor_sync(0xFFFFFFFFu, v[e], j / E);
          v[e] = (lower == up) ? (v[e] < o ? v[e] : o) : (v[e] < o ? o : v[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & j) == 0) {
            const bool up = ((lane * E + e) & (uint32_t)k) == 0;
            const K x = v[e], y = v[e | j];
            if ((x > y) == up) { v[e] = y; v[e | j] = x; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) keys[lane * E + e] = v[e];
}

struct ResolveOut { uint32_t total, rbegin; bool fits; };

// One read's leaf hits, as sorted keys strand | leaf rank | lookup | hd, to its records (see the header).  seg_shift =
// bits of lookup + hd; every lane returns the same ResolveOut.
template <typename K>
__device__ __forceinline__ ResolveOut emit_sorted(const DevIndex& ix, const MatchArgs& a, const K* keys, uint32_t T, uint32_t seg_shift, uint32_t rank_bits,
                                                  uint32_t read, uint32_t g0, uint32_t g1, bool small_counts)
{
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u, stride = a.th + 1;
  // pass A: (strand, leaf) segments that pass the hdist_filt gate
  uint32_t total = 0;
  for (uint32_t c = 0; c < T; c += 32) {
    const uint32_t i = c + lane;
    bool pass = false;
    if (i < T) {
      const K kk = keys[i];
      const K sg = kk >> seg_shift;
      if (i == 0 || (keys[i - 1] >> seg_shift) != sg) {
        uint32_t hdmin = (uint32_t)kk & 31u;
        for (uint32_t j = i + 1; j < T; ++j) { const K kj = keys[j]; if ((kj >> seg_shift) != sg) break; hdmin = min(hdmin, (uint32_t)kj & 31u); }
        const uint32_t strand = (uint32_t)(sg >> rank_bits);
        pass = a.keep_all || !(hdmin > (strand ? g1 : g0));
      }
    }
    total += __popc(__ballot_sync(0xFFFFFFFFu, pass));
  }
  ResolveOut out;
  out.total = total; out.rbegin = 0; out.fits = true;
  if (!total) return out;
  uint32_t rbegin = 0;
  if (lane == 0) rbegin = atomicAdd(a.counters, total);
  rbegin = __shfl_sync(0xFFFFFFFFu, rbegin, 0);
  const bool fits = (uint64_t)rbegin + total <= a.rec_cap;
  if (!fits && lane == 0) atomicOr(a.counters + 2, kErrRecOverflow);
  out.rbegin = rbegin; out.fits = fits;
  // pass B: histograms of the passing segments, one count per lookup at its smallest distance (a lookup's keys are
  // adjacent and ascending in hd, so its first key carries the minimum)
  uint32_t done = 0;
  for (uint32_t c = 0; c < T; c += 32) {
    const uint32_t i = c + lane;
    bool pass = false;
    uint32_t hv[kMaxTh + 1], strand = 0, rank = 0;
#pragma unroll
    for (int x = 0; x <= kMaxTh; ++x) hv[x] = 0;
    if (i < T) {
      const K kk = keys[i];
      const K sg = kk >> seg_shift;
      if (i == 0 || (keys[i - 1] >> seg_shift) != sg) {
        uint32_t hdmin = 0xFFFFFFFFu;
        if (small_counts) { // every count < 256 and th < 8: eight 8-bit counters in one word
          unsigned long long packed = 0;
          K prev_lk = ~(K)0;
          for (uint32_t j = i; j < T; ++j) {
            const K kj = keys[j];
            if ((kj >> seg_shift) != sg) break;
            const K lk = kj >> 5;
            if (lk != prev_lk) { const uint32_t hd = (uint32_t)kj & 31u; packed += 1ull << (8u * hd); hdmin = min(hdmin, hd); prev_lk = lk; }
          }
#pragma unroll
          for (int x = 0; x < 8; ++x) hv[x] = (uint32_t)(packed >> (8 * x)) & 0xFFu;
        } else {
          K prev_lk = ~(K)0;
          for (uint32_t j = i; j < T; ++j) {
            const K kj = keys[j];
            if ((kj >> seg_shift) != sg) break;
            const K lk = kj >> 5;
            if (lk != prev_lk) {
              const uint32_t hd = (uint32_t)kj & 31u;
#pragma unroll
              for (int x = 0; x <= kMaxTh; ++x) hv[x] += (hd == (uint32_t)x);
              hdmin = min(hdmin, hd); prev_lk = lk;
            }
          }
        }
        strand = (uint32_t)(sg >> rank_bits);
        rank = (uint32_t)sg & ((1u << rank_bits) - 1u);
        pass = a.keep_all || !(hdmin > (strand ? g1 : g0));
      }
    }
    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pass);
    if (pass && fits) {
      const uint32_t at = rbegin + done + __popc(pm & lt_mask);
      a.rec_read[at] = read;
      a.rec_slot[at] = strand << 31 | __ldg(&ix.leaf_se[rank]);
#pragma unroll
      for (int x = 0; x <= kMaxTh; ++x) if ((uint32_t)x < stride) a.rec_hist[(size_t)at * stride + x] = hv[x];
    }
    done += __popc(pm);
  }
  return out;
}

// Expands one read's hit entries into keys (written to `keys`, padded to a power of two), sorts them and emits the records.
template <typename K>
__device__ __forceinline__ ResolveOut resolve_read(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, K* keys, uint32_t hb, uint32_t nh, uint32_t T,
                                                   uint32_t rank_bits, uint32_t loc_bits, uint32_t read, uint32_t g0, uint32_t g1, bool small_counts)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t seg_shift = loc_bits + 5u, strand_shift = seg_shift + rank_bits;
  uint32_t base = 0;
  for (uint32_t c = 0; c < nh; c += 32) {
    const uint32_t i = c + lane;
    uint4 h = make_uint4(0u, 0u, 0u, 0u);
    if (i < nh) h = s.hits[hb + i];
    const uint32_t cnt = h.y;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
    const uint32_t at = base + incl - cnt;
    const K fixed = ((K)(h.z >> 31) << strand_shift) | (K)(h.z & 0x7FFFFFFFu); // strand | lookup << 5 | hd
    if (cnt <= 8u) { // all leaf loads of the lane in flight before the first one is used
      uint32_t lf[8];
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) lf[j] = j < cnt ? __ldg(&ix.cleaf[h.x + j]) : 0u;
#pragma unroll
      for (uint32_t j = 0; j < 8u; ++j) if (j < cnt) keys[at + j] = fixed | ((K)lf[j] << seg_shift);
    }
    uint32_t big = __ballot_sync(0xFFFFFFFFu, cnt > 8u); // long leaf lists: all lanes together
    while (big) {
      const int src = __ffs(big) - 1;
      big &= big - 1;
      const uint32_t bx = __shfl_sync(0xFFFFFFFFu, h.x, src), bn = __shfl_sync(0xFFFFFFFFu, cnt, src), bat = __shfl_sync(0xFFFFFFFFu, at, src);
      const uint32_t bz = __shfl_sync(0xFFFFFFFFu, h.z, src);
      const K bfixed = ((K)(bz >> 31) << strand_shift) | (K)(bz & 0x7FFFFFFFu);
      for (uint32_t j = lane; j < bn; j += 32) keys[bat + j] = bfixed | ((K)__ldg(&ix.cleaf[bx + j]) << seg_shift);
    }
    base += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
  uint32_t n = 32;
  while (n < T) n <<= 1;
  for (uint32_t i = T + lane; i < n; i += 32) keys[i] = ~(K)0; // never a real key: hd <= 16 < 31
  __syncwarp();
  if (n == 32) bitonic_sort_regs<K, 1>(keys);
  else if (n == 64) bitonic_sort_regs<K, 2>(keys);
  else if (n == 128) bitonic_sort_regs<K, 4>(keys);
  else if (n == 256 && sizeof(K) == 4) bitonic_sort_regs<K, (sizeof(K) == 4 ? 8 : 4)>(keys);
  else bitonic_sort(keys, n);
  __syncwarp();
  return emit_sorted<K>(ix, a, keys, T, seg_shift, rank_bits, read, g0, g1, small_counts);
}

__global__ void __launch_bounds__(kResWarps * 32) resolve_kernel(const DevIndex ix, const MatchArgs a, const SortArgs s, uint32_t rank_bits)
{
  __shared__ __align__(16) uint32_t skeys[kResWarps][kResKeys];
  if (a.counters[2] & kErrRedo) return;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * kResWarps + warp;
  uint64_t* gkeys = s.keys_g + (size_t)gwarp * s.cap_keys_g;
  unsigned long long st_records = 0;
  uint32_t claim = 0, claim_end = 0;
  for (;;) {
    if (claim == claim_end) {
      if (lane == 0) claim = atomicAdd(s.sc + 4, kResClaim);
      claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
      claim_end = min(claim + kResClaim, a.n_reads);
      if (claim >= a.n_reads) break;
    }
    const uint32_t read = claim++;
    const uint32_t hb = s.hit_begin[read], nh = s.hit_begin[read + 1] - hb;
    uint32_t filt0 = 0xFFFFFFFFu, filt1 = 0xFFFFFFFFu;
    ResolveOut out;
    out.total = 0; out.rbegin = 0; out.fits = true;
    if (nh) {
      // leaf hits of the read and the per-strand minimum distance over all hit entries (IMers::hdist_filt, ref src/query.cpp:366)
      unsigned long long T64 = 0;
      for (uint32_t i = lane; i < nh; i += 32) {
        const uint4 h = s.hits[hb + i];
        T64 += h.y;
        const uint32_t hd = h.z & 31u;
        if (h.z >> 31) filt1 = min(filt1, hd); else filt0 = min(filt0, hd);
      }
      for (int o = 16; o; o >>= 1) {
        T64 += __shfl_xor_sync(0xFFFFFFFFu, T64, o);
        filt0 = min(filt0, __shfl_xor_sync(0xFFFFFFFFu, filt0, o)); filt1 = min(filt1, __shfl_xor_sync(0xFFFFFFFFu, filt1, o));
      }
      const uint32_t g0 = 2u * filt0 + 1u, g1 = 2u * filt1 + 1u; // uint32 wrap kept, as in the reference (src/query.cpp:101-102)
      const uint32_t nlk = a.wn[2 * read] + a.wn[2 * read + 1];
      const uint32_t loc_bits = nlk <= 1u ? 0u : 32u - __clz(nlk - 1u);
      const bool narrow = 1u + rank_bits + loc_bits + 5u <= 32u;
      const bool small_counts = nlk < 256u && a.th < 8u;
      if (T64 == 0) { /* only colours without leaves: no records */ }
      else if (narrow && T64 <= (unsigned long long)kResKeys)
        out = resolve_read<uint32_t>(ix, a, s, skeys[warp], hb, nh, (uint32_t)T64, rank_bits, loc_bits, read, g0, g1, small_counts);
      else if (!narrow && T64 <= (unsigned long long)(kResKeys / 2))
        out = resolve_read<uint64_t>(ix, a, s, reinterpret_cast<uint64_t*>(skeys[warp]), hb, nh, (uint32_t)T64, rank_bits, loc_bits, read, g0, g1, small_counts);
      else if (T64 <= (unsigned long long)s.cap_keys_g)
        out = resolve_read<uint64_t>(ix, a, s, gkeys, hb, nh, (uint32_t)T64, rank_bits, loc_bits, read, g0, g1, small_counts);
      else if (lane == 0) { // more leaf hits than the warp's scratch holds: the host grows it to the largest demand and runs the batch again
        atomicMax(s.sc + 5, (uint32_t)min(T64, 0xFFFFFFFFull));
        atomicOr(a.counters + 2, kErrKeysOverflow);
      }
      __syncwarp();
    }
    if (lane == 0) {
      a.hdfilt[2 * read] = filt0; a.hdfilt[2 * read + 1] = filt1;
      a.rec_begin[read] = out.fits ? out.rbegin : 0; a.rec_count[read] = out.fits ? out.total : 0;
      st_records += out.total;
    }
  }
  if (lane == 0 && st_records) atomicAdd(a.stats, 64ull * st_records);
}

// ---------------------------------------------------------------------------------------------------- host launcher

static size_t lookup_smem(uint32_t k) { return lut_chunks(k) * 256 * sizeof(uint4) + kLkWarps * sizeof(WarpSmem); }

int sorted_resolve_warps(int sms) { return sms * 6 * kResWarps; } // grid of the resolve kernel (sizes SortArgs::keys_g)

// CTAs per SM of the scatter pass (KREPP_SCATTER_CTAS = 1 or 2, default 2).  Measured on B200 (r06, config 3): the pass takes
// the same 4.6 ms per 1M reads with one CTA per SM as with two -- it is bound by the SM's limit on outstanding returning
// atomics, not by issue slots or registers -- so at 1 it leaves half the register file to kernels of other batch slots.
// Together with KREPP_JOIN_CTAS / KREPP_RESOLVE_CTAS (CTAs per SM of those grids) this is the knob set for co-scheduling the
// atomic-bound and the issue-bound kernels of neighbouring batches; with join and resolve throttled to make room, pipelined
// slots did overlap (end-to-end above the serial sum) but lost more in the throttled kernels than they hid, so the defaults
// keep every kernel at full occupancy.
static int env_int(const char* name, int dflt)
{
  const char* e = getenv(name);
  const int x = e ? atoi(e) : dflt;
  return x > 0 ? x : dflt;
}
static int scatter_ctas_per_sm()
{
  static const int v = [] { const char* e = getenv("KREPP_SCATTER_CTAS"); const int x = e ? atoi(e) : 2; return x == 1 ? 1 : 2; }();
  return v;
}

static int resolve_grid(const SortArgs& s, int sms)
{ // keys_g holds one region per warp of the grid it was sized for (api.cu grow_keys shrinks the grid when the regions get large)
  const int dflt = std::min(sorted_resolve_warps(sms) / kResWarps, sms * env_int("KREPP_RESOLVE_CTAS", 6));
  return s.res_ctas ? std::min<int>((int)s.res_ctas, dflt) : dflt;
}
int sorted_resolve_warps_per_cta() { return kResWarps; }

template <bool SCATTER>
static cudaError_t launch_lookup(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream)
{
  const size_t sm = lookup_smem(ix.k);
  cudaError_t e;
  if (SCATTER) sms = sms * scatter_ctas_per_sm() / 2; // the launches below use sms * 2 CTAs
  if (tap && !SCATTER) {
    e = cudaFuncSetAttribute(lookup_kernel<SCATTER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    lookup_kernel<SCATTER, true><<<sms * 2, kLkWarps * 32, sm, stream>>>(ix, a, s);
  } else {
    e = cudaFuncSetAttribute(lookup_kernel<SCATTER, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    lookup_kernel<SCATTER, false><<<sms * 2, kLkWarps * 32, sm, stream>>>(ix, a, s);
  }
  return cudaGetLastError();
}

// L1 + S1 + L2, or L1' + L2' when the slot has bins (s.nbins): leaves the lookups grouped by row in s.tuples / s.row_begin
static cudaError_t launch_lookup_sort(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream, StageClock* clk, uint32_t* launches)
{
  cudaError_t e;
  if (s.nbins) {
    if ((e = cudaMemsetAsync(s.bin_cursor, 0, 4ull * s.nbins, stream)) != cudaSuccess) return e;
    const size_t sm1 = lut_chunks(ix.k) * 256 * sizeof(uint4) + sizeof(LpShared);
    if (tap) {
      if ((e = cudaFuncSetAttribute(lookup_partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)) != cudaSuccess) return e;
      lookup_partition_kernel<true><<<sms * 2, kLpWarps * 32, sm1, stream>>>(ix, a, s);
    } else {
      if ((e = cudaFuncSetAttribute(lookup_partition_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)) != cudaSuccess) return e;
      lookup_partition_kernel<false><<<sms * 2, kLpWarps * 32, sm1, stream>>>(ix, a, s);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (clk) clk->tick("lookup_partition_kernel", stream);
    const size_t sm2 = 4ull * ((1ull << s.bin_shift) + s.nbins + 1);
    if ((e = cudaFuncSetAttribute(bin_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2)) != cudaSuccess) return e;
    bin_sort_kernel<<<std::min<uint32_t>(s.nbins, (uint32_t)sms), kBsThreads, sm2, stream>>>(s, a.counters);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (clk) clk->tick("bin_sort_kernel", stream);
    if (launches) *launches += 2;
    return cudaSuccess;
  }
  if ((e = cudaMemsetAsync(s.row_count, 0, 4ull * s.nrows, stream)) != cudaSuccess) return e;
  if ((e = launch_lookup<false>(ix, a, s, sms, tap, stream)) != cudaSuccess) return e;
  if (clk) clk->tick("lookup_kernel<count>", stream);
  if ((e = exclusive_scan(s.row_count, s.nrows, s.partials, s.row_begin, s.row_cursor, stream)) != cudaSuccess) return e;
  if (clk) clk->tick("scan(rows)", stream);
  if ((e = launch_lookup<true>(ix, a, s, sms, false, stream)) != cudaSuccess) return e;
  if (clk) clk->tick("lookup_kernel<scatter>", stream);
  if (launches) *launches += 5;
  return cudaSuccess;
}

// Enqueues L1 .. R for one batch.  The caller has zeroed a.counters / a.stats; this zeroes the pipeline's own counters.
cudaError_t launch_match_sorted(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream, uint32_t* launches, StageClock* clk)
{
  cudaError_t e;
  if (launches) *launches = 0;
  if (!a.n_reads) return cudaSuccess;
  if ((e = cudaMemsetAsync(s.hit_count, 0, 4ull * a.n_reads, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(s.sc, 0, 32, stream)) != cudaSuccess) return e;
  if (clk) clk->tick("memsets", stream);
  if ((e = launch_lookup_sort(ix, a, s, sms, tap, stream, clk, launches)) != cudaSuccess) return e;
  join_kernel<true><<<sms * env_int("KREPP_JOIN_CTAS", 8), kJoinWarps * 32, 0, stream>>>(ix, s, a.th, a.counters, a.stats);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (clk) clk->tick("join_kernel", stream);
  if ((e = exclusive_scan(s.hit_count, a.n_reads, s.partials, s.hit_begin, s.hit_cursor, stream)) != cudaSuccess) return e;
  hit_scatter_kernel<<<sms * 8, 256, 0, stream>>>(ix, s, a.counters);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (clk) clk->tick("scan(hits)+hit_scatter_kernel", stream);
  uint32_t rank_bits = 0;
  while ((1ull << rank_bits) < ix.nleaves) ++rank_bits;
  rank_bits += s.extra_rank_bits;
  resolve_kernel<<<resolve_grid(s, sms), kResWarps * 32, 0, stream>>>(ix, a, s, rank_bits);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (clk) clk->tick("resolve_kernel", stream);
  if (launches) *launches += 6;
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------- mode B
// The same chain cut at its two exchange points.  Home rank: L1, S1, L2 -> tuples grouped by row, so the tuples of every
// shard's row range are one contiguous run.  Owner rank: J over one sender's run against its slice of the table, hit entries
// appended to one list.  Home rank again: hit entries of its reads from all owners -> S2, R.  All entries that can match a
// lookup live in one bucket, hence on one owner, and R orders a read's hits itself, so the records are those of the
// unsharded chain bit for bit.

cudaError_t launch_shard_lookup(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, bool tap, cudaStream_t stream, StageClock* clk)
{
  cudaError_t e;
  if ((e = cudaMemsetAsync(s.sc, 0, 32, stream)) != cudaSuccess) return e;
  if (!a.n_reads) return cudaMemsetAsync(s.row_begin, 0, 4ull * (s.nrows + 1), stream);
  return launch_lookup_sort(ix, a, s, sms, tap, stream, clk, nullptr);
}

// s.row_begin / s.tuples: one sender's slice for this shard's rows; s.nrows = rows of the shard; s.sc[0] keeps counting hits
cudaError_t launch_shard_join(const DevIndex& ix, const SortArgs& s, uint32_t th, uint32_t* counters, unsigned long long* stats, int sms, cudaStream_t stream)
{
  cudaError_t e;
  if (!s.nrows) return cudaSuccess;
  if ((e = cudaMemsetAsync(s.sc + 1, 0, 4, stream)) != cudaSuccess) return e;
  join_kernel<false><<<sms * 8, kJoinWarps * 32, 0, stream>>>(ix, s, th, counters, stats);
  return cudaGetLastError();
}

// s.hits_tmp: the n_hits hit entries of this batch's reads (s.sc[0] already holds n_hits)
cudaError_t launch_shard_finish(const DevIndex& ix, const MatchArgs& a, const SortArgs& s, int sms, cudaStream_t stream, StageClock* clk)
{
  cudaError_t e;
  if (!a.n_reads) return cudaSuccess;
  if ((e = cudaMemsetAsync(s.hit_count, 0, 4ull * a.n_reads, stream)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(s.sc + 4, 0, 8, stream)) != cudaSuccess) return e;
  hit_count_kernel<<<sms * 8, 256, 0, stream>>>(s, a.n_reads, a.counters);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = exclusive_scan(s.hit_count, a.n_reads, s.partials, s.hit_begin, s.hit_cursor, stream)) != cudaSuccess) return e;
  hit_scatter_kernel<<<sms * 8, 256, 0, stream>>>(ix, s, a.counters);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (clk) clk->tick("hit_count+scan(hits)+hit_scatter_kernel", stream);
  uint32_t rank_bits = 0;
  while ((1ull << rank_bits) < ix.nleaves) ++rank_bits;
  rank_bits += s.extra_rank_bits;
  resolve_kernel<<<resolve_grid(s, sms), kResWarps * 32, 0, stream>>>(ix, a, s, rank_bits);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (clk) clk->tick("resolve_kernel", stream);
  return cudaSuccess;
}

} // namespace krepp

// solve.cu -- K4 of the query path: maximum-likelihood distance per (read, strand, leaf) record (identical problems of a
// batch solved once: the solve memo of gate_kernel / alias_kernel), then the per-read strand merge / closest selection and
// the likelihood-ratio statistic; and K5, placement, as a chain of four kernels (further down).
//
// THIS FILE IS COMPILED WITH --fmad=false: the reference runs on x86-64 without FMA contraction, and Brent's
// parabolic-vs-golden decisions (tolerance only 2^-15) must follow the same iteration
// sequence, so every product and sum below rounds separately, in the reference's literal evaluation order.
//
//   objective  : optimize::HDistHistLLH::operator()          ref src/hdhistllh.hpp:71-89
//   tables     : HDistHistLLH ctor (host, passed by value)   ref src/hdhistllh.hpp:51-69
//   minimiser  : boost::math::tools::brent_find_minima(f, 1e-10, 0.5, 16)
//                ref external/boost/libs/math/include/boost/math/tools/minima.hpp:23-138, called at src/query.cpp:430
//   gate/merge : IBatch::summarize_matches                   ref src/query.cpp:96-139
//   chisq      : Minfo::likelihood_ratio                     ref src/query.cpp:420-424
#include "../../include/krepp_b200.h"
#include "device.cuh"
#include "solve.cuh"
#include "llh_math.cuh"

#include <cfloat>

namespace krepp {

// Objective with every histogram bin the library supports (placement and chisq: weighted / arbitrary th).
using ObjectiveAny = Objective<kMaxTh + 1>;

template <int N>
struct PlainEval { // brent_minimum functor: the full objective at every abscissa
  const Objective<N>* o;
  const LlhTables* t;
  __device__ double operator()(double u, int) const { return o->eval(*t, u); }
};

template <int N>
struct MemoEval { // the same, with the d-only terms of the first three abscissae taken from the CTA's table (looked up by value)
  const Objective<N>* o;
  const LlhTables* t;
  const double* su;   // [4] brent_first_points
  const DTerms* st;   // [4] d_terms at those points
  __device__ double operator()(double u, int it) const
  {
    if (it < 3) {
      const int slot = it < 2 ? it : (u == su[2] ? 2 : 3);
      if (u == su[slot]) return o->finish(st[slot]);
    }
    return o->eval(*t, u);
  }
};

// One thread per record: match_count / hdist_min from the histogram and the hdist_filt gate of summarize_matches (ref
// src/query.cpp:101-106,116-119).  Records that pass are appended to a work list, so that the Brent kernel below runs
// with full warps (on the 1,000-genome index only about a quarter of the records pass).
__global__ void __launch_bounds__(256) gate_kernel(const SolveArgs a)
{
  if (a.counters[2] & kErrRedo) return; // the host grows the buffers and runs the batch again
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  const uint32_t stride = a.th + 1, lane = threadIdx.x & 31;
  const uint32_t nround = (n + 31) & ~31u; // whole warps stay together for the ballot
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
    bool pass = false;
    if (i < n) {
      const uint32_t read = a.rec_read[i], strand = a.rec_slot[i] >> 31;
      uint32_t match = 0, hdmin = 0xFFFFFFFFu;
      for (uint32_t x = 0; x < stride; ++x) {
        const uint32_t c = a.rec_hist[(size_t)i * stride + x];
        match += c;
        if (c && hdmin == 0xFFFFFFFFu) hdmin = x;
      }
      a.rec_match[i] = match; a.rec_hdmin[i] = hdmin;
      const uint32_t filt = 2u * a.hdfilt[2 * read + strand] + 1u; // uint32 wrap kept (ref src/query.cpp:101-102)
      pass = !(hdmin > filt);
      a.rec_d[i] = DBL_MAX; a.rec_v[i] = nan(""); a.rec_flags[i] = 0; a.rec_chisq[i] = nan(""); // Minfo defaults (ref src/query.hpp:225-226)
      uint32_t alias = 0xFFFFFFFFu;
      if (pass && a.memo_mask) { // same (histogram, onmers, leaf) as an earlier record of the batch?
        const uint32_t se = a.rec_slot[i] & 0x7FFFFFFFu, on = a.onmers[read], lim = (1u << a.memo_bits) - 1u;
        unsigned long long key = 0;
        bool fits = se < (1u << 21) && on < 256u;
        for (uint32_t x = 0; x < stride; ++x) { const uint32_t c = a.rec_hist[(size_t)i * stride + x]; fits = fits && c <= lim; key = key << a.memo_bits | c; }
        if (fits) {
          key = (key << 8 | on) << 21 | se; // never 0: a record has at least one match
          unsigned long long hsh = key * 0x9E3779B97F4A7C15ull;
          hsh ^= hsh >> 29;
          uint32_t slot = (uint32_t)hsh & a.memo_mask;
          for (int probe = 0; probe < 32; ++probe, slot = (slot + 1) & a.memo_mask) {
            unsigned long long prev = __ldcg(a.memo_key + slot); // a slot never changes once set: most records find their key with a plain load
            if (prev == 0ull) prev = atomicCAS(a.memo_key + slot, 0ull, key);
            if (prev == 0ull) { a.memo_owner[slot] = i; break; }          // first of its kind: solved below
            if (prev == key) { alias = slot; pass = false; break; }       // takes the owner's result (alias_kernel)
          }                                                               // 32 occupied slots in a row: solved on its own
        }
      }
      a.rec_alias[i] = alias;
    }
    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pass);
    uint32_t base = 0;
    if (lane == 0 && pm) base = atomicAdd(a.counters + 4, __popc(pm));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (pass) a.work[base + __popc(pm & ((1u << lane) - 1))] = i;
  }
}

// Records that share their problem with an earlier one take its result.
__global__ void __launch_bounds__(256) alias_kernel(const SolveArgs a)
{
  if (a.counters[2] & kErrRedo) return;
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = a.rec_alias[i];
    if (slot == 0xFFFFFFFFu) continue;
    const uint32_t o = a.memo_owner[slot];
    a.rec_d[i] = a.rec_d[o]; a.rec_v[i] = a.rec_v[o]; a.rec_flags[i] = 1u; // KREPP_REC_SOLVED
  }
}

// One thread per work item: Brent on the record's histogram.  N = histogram bins kept in registers (th + 1 <= N).
template <int N>
__global__ void __launch_bounds__(128, 8) solve_kernel(const SolveArgs a, const LlhTables tab)
{
  __shared__ double su[4];
  __shared__ DTerms st[4];
  if (a.counters[2] & kErrRedo) return;
  if (threadIdx.x < 4) {
    double u[4];
    brent_first_points(u);
    const double ut = threadIdx.x == 0 ? u[0] : threadIdx.x == 1 ? u[1] : threadIdx.x == 2 ? u[2] : u[3];
    su[threadIdx.x] = ut;
    st[threadIdx.x] = d_terms(tab, ut, a.k);
  }
  __syncthreads();
  const uint32_t n = a.counters[4];
  const uint32_t stride = a.th + 1;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t i = a.work[j];
    const uint32_t read = a.rec_read[i], se = a.rec_slot[i] & 0x7FFFFFFFu;
    Objective<N> f;
#pragma unroll
    for (int x = 0; x < N; ++x) f.mc[x] = (uint32_t)x < stride ? (double)a.rec_hist[(size_t)i * stride + x] : 0.0;
    f.uc = (double)a.onmers[read] - (double)a.rec_match[i]; f.rho = a.rho[se]; f.k = a.k; f.th = a.th;
    double d, v;
    brent_minimum(MemoEval<N>{&f, &tab, su, st}, d, v);
    a.rec_d[i] = d; a.rec_v[i] = v; a.rec_flags[i] = 1u; // KREPP_REC_SOLVED
  }
}

// One thread per read: closest + per-leaf strand choice, in the fixed visiting order (forward leaves by ascending se,
// then reverse leaves by ascending se; `<=` kept so the last tied entry wins -- SURVEY.md section 0 fact 6).
__global__ void __launch_bounds__(128) merge_kernel(const SolveArgs a)
{
  if (a.counters[2] & kErrRedo) return; // the host grows the buffers and runs the batch again
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
    const uint32_t b = a.rec_begin[r], n = a.rec_count[r];
    int32_t cl = -1;
    double best = DBL_MAX;
    uint32_t nf = 0;
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t g = b + i;
      if (!(a.rec_slot[g] >> 31)) ++nf;
      if ((a.rec_flags[g] & 1u) && a.rec_d[g] <= best) { best = a.rec_d[g]; cl = (int32_t)g; }
    }
    // two-pointer walk over forward [b, b+nf) and reverse [b+nf, b+n), both ascending in se
    uint32_t i = b, j = b + nf, nsel = 0;
    const uint32_t ie = b + nf, je = b + n;
    while (i < ie || j < je) {
      const uint32_t si = i < ie ? (a.rec_slot[i] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
      const uint32_t sj = j < je ? (a.rec_slot[j] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
      if (si < sj) { if (a.rec_flags[i] & 1u) { a.rec_flags[i] |= 2u; ++nsel; } ++i; }
      else if (sj < si) { if (a.rec_flags[j] & 1u) { a.rec_flags[j] |= 2u; ++nsel; } ++j; }
      else {
        const bool fs = a.rec_flags[i] & 1u, rs = a.rec_flags[j] & 1u;
        if (rs) {
          const double dr = a.rec_d[j], df = a.rec_d[i];
          const bool fwd_wins = (dr > df) || ((dr == df) && (a.rec_match[j] < a.rec_match[i]));
          if (fwd_wins) a.rec_flags[i] |= 2u; else a.rec_flags[j] |= 2u;
          ++nsel;
        } else if (fs) { a.rec_flags[i] |= 2u; ++nsel; }
        ++i; ++j;
      }
    }
    if (cl >= 0) { // node_to_minfo[nd_closest] = mi_closest (ref src/query.cpp:136-138): the leaf keeps one entry
      const uint32_t se = a.rec_slot[cl] & 0x7FFFFFFFu;
      for (uint32_t q = b; q < b + n; ++q)
        if ((a.rec_slot[q] & 0x7FFFFFFFu) == se) a.rec_flags[q] &= ~2u;
      a.rec_flags[cl] |= 2u | 4u;
    }
    if (a.nsel) a.nsel[r] = nsel;
    a.closest[r] = cl;
  }
}

// One thread per selected record: chisq = 2 * (f_closest(d_record) - v_closest).
__global__ void __launch_bounds__(128) chisq_kernel(const SolveArgs a, const LlhTables tab)
{
  if (a.counters[2] & kErrRedo) return; // the host grows the buffers and runs the batch again
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  const uint32_t stride = a.th + 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!(a.rec_flags[i] & 2u)) continue;
    const uint32_t read = a.rec_read[i];
    const int32_t cl = a.closest[read];
    if (cl < 0) continue;
    ObjectiveAny f;
    for (uint32_t x = 0; x <= (uint32_t)kMaxTh; ++x) f.mc[x] = x < stride ? (double)a.rec_hist[(size_t)cl * stride + x] : 0.0;
    f.uc = (double)a.onmers[read] - (double)a.rec_match[cl]; f.rho = a.rho[a.rec_slot[cl] & 0x7FFFFFFFu]; f.k = a.k; f.th = a.th;
    a.rec_chisq[i] = 2 * (f.eval(tab, a.rec_d[i]) - a.rec_v[cl]);
  }
}

// ------------------------------------------------------------------------------------------------ K5: placement
//
// One warp per read; restates IBatch::report_placement (ref src/query.cpp:218-333, multi mode) over the flattened tree:
//   * skip unless the closest reference has more than one match at Hamming distance <= tau (when filtering, :220)
//   * one selected reference  -> that leaf, lwr = 1, chisq = 0 (:231-241)
//   * otherwise every selected leaf is pushed to all its ancestors with weight prod 1/eff_nchildren (Minfo::add,
//     ref src/query.hpp:139-152; denominators formed by successive division as at src/query.cpp:250-259), in ascending
//     leaf order; internal nodes are re-solved (:273-275); a node is a candidate when its chisq against the closest
//     is below the threshold and it is not the root (:276-279); lwr = exp(-chisq/2) / sum (:284-296).
// Visiting order is fixed to ascending se.  In post-order numbering the subtree of node g
// is the contiguous range (g - subtree[g], g], which makes "is leaf l below g" two comparisons.
__device__ __forceinline__ double jukes_cantor(double d) { return -0.75 * log(1 - 4.0 / 3.0 * d); }

// The work is cut so that every Brent minimisation of the batch runs in one thread-per-item kernel with full warps (an
// internal node costs as much as a record of K4, and a read with 20 selected leaves touches about a hundred nodes):
// (The per-warp scratch lists in HBM are written and read by the same warp with __syncwarp() in between, i.e. on one SM,
// so ordinary cached loads see the stores and the lists are served from L1.)
//   place_collect_kernel  warp per read: gates, the single-reference shortcut, marked nodes in ascending se, one lane per
//                         node accumulating its weighted histogram -> node entries + work list
//   place_solve_kernel    thread per internal node entry: the same minimiser as solve_kernel
//   place_chisq_kernel    thread per node entry: likelihood-ratio test against the closest reference
//   place_emit_kernel     warp per read: candidates in ascending se, lwr, placement rows
template <int N> // N = histogram bins kept in registers (th + 1 <= N)
__global__ void __launch_bounds__(128, 8) place_collect_kernel(const PlaceArgs a)
{
  const SolveArgs& s = a.s;
  if (s.counters[2] & kErrRedo) return;
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t stride = s.th + 1, nbm = (a.nnodes + 32) >> 5;
  // marked-node bitmap of the warp: in shared memory when the tree has at most 8,192 nodes (marking is a chain of dependent
  // atomics up the tree, a microsecond each in L2), else in the warp's HBM scratch (zeroed at allocation, left clean by step 2)
  constexpr uint32_t kSmemBm = 256;
  __shared__ uint32_t sbm[kPlaceWarpsPerCta][kSmemBm];
  // exact mode (see 3a'): prefix sums over the read's selected references of their scaled histograms (+ match count)
  constexpr uint32_t kPfxCap = 64;
  __shared__ double spfx[kPlaceWarpsPerCta][kPfxCap + 1][N + 1];
  __shared__ double srho[kPlaceWarpsPerCta][kPfxCap];  // exact mode: rho of the selected references, by rank
  double (*pfx)[N + 1] = spfx[threadIdx.x >> 5];
  double* sel_rho = srho[threadIdx.x >> 5];
  uint32_t* bm = a.node_bitmap + (size_t)gwarp * nbm;
  if (nbm <= kSmemBm) {
    bm = sbm[threadIdx.x >> 5];
    for (uint32_t i = lane; i < nbm; i += 32) bm[i] = 0;
    __syncwarp();
  }
  uint32_t* list = a.node_list + (size_t)gwarp * a.nnodes;
  krepp_placement_t* out = static_cast<krepp_placement_t*>(a.placements);

  for (uint32_t r = gwarp; r < s.n_reads; r += nwarps) {
    const uint32_t b = s.rec_begin[r], n = s.rec_count[r];
    const int32_t cl = s.closest[r];
    uint32_t pbegin = 0, pcount = 0, nbegin = 0, ncount = 0;
    if (cl >= 0) {
      // number of selected references; records are forward leaves by ascending se, then reverse leaves by ascending se
      uint32_t nsel = 0, nf = 0;
      for (uint32_t i = lane; i < n; i += 32) { nsel += (s.rec_flags[b + i] >> 1) & 1u; nf += !(s.rec_slot[b + i] >> 31); }
      for (int o = 16; o; o >>= 1) { nsel += __shfl_xor_sync(0xFFFFFFFFu, nsel, o); nf += __shfl_xor_sync(0xFFFFFFFFu, nf, o); }
      double leq_cl = 0; // Minfo::get_leq_tau of the closest (ref src/query.hpp:189-196)
      for (uint32_t x = 0; x <= a.tau && x < stride; ++x) leq_cl += (double)s.rec_hist[(size_t)cl * stride + x];
      const uint32_t cl_se = s.rec_slot[cl] & 0x7FFFFFFFu;
      if (a.no_filter || leq_cl > 1.0) {
        if (nsel == 1) {
          if (lane == 0) {
            pbegin = atomicAdd(a.counters + 3, 1u); pcount = 1;
            if (pbegin < a.place_cap) {
              const double bl = a.blen[cl_se], mid = isnan(bl) ? 0.0 : bl / 2.0, d = s.rec_d[cl];
              krepp_placement_t p; p.read = r; p.se = cl_se; p.pendant = jukes_cantor(d) - mid; p.distal = mid; p.loglik = -s.rec_v[cl];
              p.lwr = 1; p.d_llh = d; p.chisq = 0;
              out[pbegin] = p;
            } else atomicOr(a.counters + 2, kErrPlaceOverflow);
          }
        } else {
          // 1. mark every selected leaf and all its ancestors
          for (uint32_t i = lane; i < n; i += 32) {
            if (!(s.rec_flags[b + i] & 2u)) continue;
            uint32_t node = s.rec_slot[b + i] & 0x7FFFFFFFu;
            while (node) {
              const uint32_t bit = 1u << (node & 31);
              if (atomicOr(&bm[node >> 5], bit) & bit) break;
              node = a.parent[node];
            }
          }
          __syncwarp();
          // 2. ascending list of marked nodes
          uint32_t cnt = 0;
          for (uint32_t wb = 0; wb < nbm; wb += 32) {
            uint32_t bits = (wb + lane < nbm) ? *reinterpret_cast<volatile uint32_t*>(&bm[wb + lane]) : 0u; // set by atomics
            const uint32_t c = __popc(bits);
            uint32_t incl = c;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            uint32_t at = cnt + incl - c;
            cnt += __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (wb + lane < nbm && bits) bm[wb + lane] = 0;
            while (bits) { const uint32_t q = __ffs(bits) - 1; bits &= bits - 1; list[at++] = (wb + lane) * 32 + q; }
          }
          __syncwarp();
          if (lane == 0) nbegin = atomicAdd(a.counters + 5, cnt);
          nbegin = __shfl_sync(0xFFFFFFFFu, nbegin, 0);
          ncount = cnt;
          if ((uint64_t)nbegin + cnt > a.node_cap) { // counters[5] ends as the demand; the host grows the node arrays and runs the batch again
            if (lane == 0) atomicOr(a.counters + 2, kErrNodeOverflow);
            ncount = 0; cnt = 0;
          }
          // 3a. the read's selected references by ascending se (at most one per leaf), and for each the weights it carries to
          //     its ancestors: level l = its (l+1)-th ancestor, weight = 1 / nch(1st) / ... / nch((l+1)-th), formed by the
          //     same successive divisions as the reference (ref src/query.cpp:250-259), once per leaf instead of once per
          //     (leaf, ancestor) pair
          uint32_t* sel_rec = a.sel + (size_t)gwarp * 3 * a.nleaves;
          uint32_t* sel_se = sel_rec + a.nleaves;
          uint32_t* sel_off = sel_se + a.nleaves;
          double* chain = a.chain + (size_t)gwarp * a.chain_cap;
          if (n <= 32) { // one record per lane: a selected record's rank = selected records with a smaller se, counted over shuffles
            const bool have = lane < n;
            const uint32_t se = have ? (s.rec_slot[b + lane] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
            const bool picked = have && (s.rec_flags[b + lane] & 2u);
            const uint32_t pm = __ballot_sync(0xFFFFFFFFu, picked);
            uint32_t rank = 0;
            for (uint32_t q = 0; q < n; ++q) { const uint32_t sq = __shfl_sync(0xFFFFFFFFu, se, q); rank += ((pm >> q) & 1u) && sq < se; }
            if (picked) { sel_rec[rank] = b + lane; sel_se[rank] = se; }
          } else
          for (uint32_t i = lane; i < n; i += 32) {
            if (!(s.rec_flags[b + i] & 2u)) continue;
            const uint32_t se = s.rec_slot[b + i] & 0x7FFFFFFFu;
            uint32_t rank = 0;
            for (uint32_t q = 0; q < n; ++q) rank += (s.rec_flags[b + q] & 2u) && (s.rec_slot[b + q] & 0x7FFFFFFFu) < se;
            sel_rec[rank] = b + i; sel_se[rank] = se;
          }
          __syncwarp();
          const uint32_t enmers = (uint32_t)(a.offsets[r + 1] - a.offsets[r]) - s.k + 1;
          // 3a'. Exact mode.  When every ancestor of the selected references has a power-of-two number of children (any binary
          //      tree), a leaf's weight at an ancestor g is exactly 2^-(logw[leaf] - logw[g]) and every sum Minfo::add forms is
          //      exactly representable (counts below 2^16, at most 64 leaves, weights down to 2^-30: 52 bits), so the order of
          //      the additions cannot matter and a node's histogram is a difference of prefix sums over the leaves below it --
          //      O(1) per node instead of one step per leaf below -- with the same bits as the reference's ordered sum.
          bool exact = nsel <= kPfxCap && enmers < 65536u && a.logw != nullptr;
          if (exact) {
            bool bad = false;
            for (uint32_t i = lane; i < nsel; i += 32) bad = bad || a.logw[sel_se[i]] > 30u;
            exact = !__any_sync(0xFFFFFFFFu, bad);
          }
          if (exact) {
            // every reference's scaled histogram (+ match count) and rho into shared memory, one reference per lane and trip ...
            for (uint32_t i = lane; i < nsel; i += 32) {
              const uint32_t rec = sel_rec[i], se = sel_se[i];
              const double scale = __hiloint2double((int)((1023u - a.logw[se]) << 20), 0); // 2^-logw
              for (uint32_t x = 0; x < stride; ++x) pfx[i + 1][x] = (double)s.rec_hist[(size_t)rec * stride + x] * scale;
              pfx[i + 1][stride] = (double)s.rec_match[rec] * scale;
              sel_rho[i] = s.rho[se];
            }
            __syncwarp();
            // ... then the running sums, lane x = component x (lane `stride`: the match count)
            if (lane <= stride && lane <= (uint32_t)N) {
              double run = 0;
              pfx[0][lane] = 0;
              for (uint32_t i = 0; i < nsel; ++i) { run = run + pfx[i + 1][lane]; pfx[i + 1][lane] = run; }
            }
            __syncwarp();
          }
          uint32_t chain_len = 0;
          if (!exact)
          for (uint32_t i0 = 0; i0 < nsel; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint32_t dep = i < nsel ? a.depth[sel_se[i]] : 0u;
            uint32_t incl = dep;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            if (i < nsel) sel_off[i] = chain_len + incl - dep;
            chain_len += __shfl_sync(0xFFFFFFFFu, incl, 31);
          }
          const bool chained = !exact && chain_len <= a.chain_cap;
          __syncwarp();
          if (chained)
            for (uint32_t i = lane; i < nsel; i += 32) {
              double denom = 1.0;
              uint32_t at = sel_off[i];
              for (uint32_t node = a.parent[sel_se[i]]; node; node = a.parent[node]) { denom /= (double)a.eff[node]; chain[at++] = denom; }
            }
          __syncwarp();
          // 3b. one lane per marked node: a selected leaf brings its own record, an internal node accumulates the leaves below it
          //     A node costs as many steps as selected leaves sit below it, and a warp round lasts as long as its heaviest
          //     lane, so the nodes are visited heaviest first (three classes): the handful of near-root nodes that see every
          //     leaf share their rounds instead of each holding up a round of single-leaf nodes.  Entry j of the read stays
          //     node j of the ascending list whatever the visiting order.
          uint32_t* order = a.node_order + (size_t)gwarp * a.nnodes;
          if (!exact) {
            uint32_t placed = 0;
            for (int cls = 0; cls < 3; ++cls)
              for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool mine = false;
                if (j < cnt) {
                  const uint32_t g = list[j], lo = g - a.subtree[g];
                  uint32_t f0 = 0, f1 = 0; // selected references with se <= lo, with se <= g
                  for (uint32_t hi = nsel; f0 < hi;) { const uint32_t mid = (f0 + hi) >> 1; if (sel_se[mid] > lo) hi = mid; else f0 = mid + 1; }
                  for (uint32_t hi = nsel; f1 < hi;) { const uint32_t mid = (f1 + hi) >> 1; if (sel_se[mid] > g) hi = mid; else f1 = mid + 1; }
                  const uint32_t w = f1 - f0;
                  mine = cls == 0 ? w >= 8u : cls == 1 ? (w >= 3u && w < 8u) : w < 3u;
                }
                const uint32_t mm = __ballot_sync(0xFFFFFFFFu, mine);
                if (mine) order[placed + __popc(mm & lt_mask)] = j;
                placed += __popc(mm);
              }
          }
          __syncwarp();
          for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const bool active = j0 + lane < cnt;
            const uint32_t j = active ? (exact ? j0 + lane : order[j0 + lane]) : 0u;
            bool solve = false;
            if (active) {
              const uint32_t g = list[j], e = nbegin + j;
              double d = DBL_MAX, v = nan(""), leq = 0;
              const uint32_t lo = g - a.subtree[g]; // leaves below g have lo < se <= g
              uint32_t first = 0; // first selected reference with se > lo
              for (uint32_t hi = nsel; first < hi;) { const uint32_t mid = (first + hi) >> 1; if (sel_se[mid] > lo) hi = mid; else first = mid + 1; }
              if (a.leaf_rank[g] != 0xFFFFFFFFu) {
                const uint32_t rec = sel_rec[first]; // lo = g - 1: the leaf itself
                d = s.rec_d[rec]; v = s.rec_v[rec];
                for (uint32_t x = 0; x <= a.tau && x < stride; ++x) leq += (double)s.rec_hist[(size_t)rec * stride + x];
              } else {
                double mc[N];
#pragma unroll
                for (int x = 0; x < N; ++x) mc[x] = 0;
                double nmers = 0, mismatch = 0, match = 0, rho = 0;
                const uint32_t gdep = a.depth[g];
                if (exact) {
                  uint32_t last = first; // first selected reference with se > g
                  for (uint32_t hi = nsel; last < hi;) { const uint32_t mid = (last + hi) >> 1; if (sel_se[mid] > g) hi = mid; else last = mid + 1; }
                  const double up = __hiloint2double((int)((1023u + a.logw[g]) << 20), 0); // 2^logw[g]
#pragma unroll
                  for (int x = 0; x < N; ++x) if ((uint32_t)x < stride) mc[x] = (pfx[last][x] - pfx[first][x]) * up;
                  match = (pfx[last][stride] - pfx[first][stride]) * up;
                  mismatch = (double)enmers - match;
                  for (uint32_t i = first; i < last; ++i) rho = fmax(rho, sel_rho[i]);
                } else
                for (uint32_t i = first; i < nsel; ++i) { // ascending leaf se: the order Minfo::add is applied in
                  const uint32_t se = sel_se[i];
                  if (se > g) break;
                  const uint32_t rec = sel_rec[i];
                  double denom = 1.0;
                  if (chained) denom = chain[sel_off[i] + a.depth[se] - gdep - 1];
                  else for (uint32_t node = a.parent[se];; node = a.parent[node]) { denom /= (double)a.eff[node]; if (node == g) break; }
                  const double m = (double)s.rec_match[rec];
                  mismatch = nmers != 0 ? mismatch : (double)enmers;      // Minfo::add (ref src/query.hpp:139-152)
                  match += m * denom;
                  mismatch -= m * denom;
#pragma unroll
                  for (int x = 0; x < N; ++x) if ((uint32_t)x < stride) mc[x] = mc[x] + (double)s.rec_hist[(size_t)rec * stride + x] * denom;
                  nmers = fmax(nmers, (double)enmers);
                  rho = fmax(rho, s.rho[se]);
                }
#pragma unroll
                for (int x = 0; x < N; ++x) if ((uint32_t)x <= a.tau && (uint32_t)x < stride) leq += mc[x];
                solve = a.no_filter || leq > 1.0;
                if (solve) {
#pragma unroll
                  for (int x = 0; x < N; ++x) if ((uint32_t)x < stride) a.pn_mc[(size_t)e * stride + x] = mc[x];
                  a.pn_uc[e] = mismatch; a.pn_rho[e] = rho;
                }
              }
              const bool eligible = a.nchildren[g] == a.eff[g] && a.nchildren[g] != 1 && (a.no_filter || leq > 1.0); // ref src/query.cpp:269-271
              a.pn_read[e] = r; a.pn_se[e] = g; a.pn_flags[e] = (solve ? kPnSolve : 0u) | (eligible ? kPnEligible : 0u);
              a.pn_d[e] = d; a.pn_v[e] = v; a.pn_chisq[e] = nan("");
            }
            const uint32_t sm = __ballot_sync(0xFFFFFFFFu, solve);
            uint32_t wbase = 0;
            if (lane == 0 && sm) wbase = atomicAdd(a.counters + 6, __popc(sm));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            if (solve) a.pn_work[wbase + __popc(sm & lt_mask)] = nbegin + j;
          }
        }
      }
    }
    if (lane == 0) { a.place_begin[r] = pbegin; a.place_count[r] = pcount; a.pn_begin[r] = nbegin; a.pn_count[r] = ncount; }
  }
}

template <int N>
__global__ void __launch_bounds__(128, 8) place_solve_kernel(const PlaceArgs a, const LlhTables tab)
{
  __shared__ double su[4];
  __shared__ DTerms st[4];
  const SolveArgs& s = a.s;
  if (s.counters[2] & kErrRedo) return;
  if (threadIdx.x < 4) {
    double u[4];
    brent_first_points(u);
    const double ut = threadIdx.x == 0 ? u[0] : threadIdx.x == 1 ? u[1] : threadIdx.x == 2 ? u[2] : u[3];
    su[threadIdx.x] = ut;
    st[threadIdx.x] = d_terms(tab, ut, s.k);
  }
  __syncthreads();
  const uint32_t n = a.counters[6], stride = s.th + 1;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t e = a.pn_work[j];
    Objective<N> f;
#pragma unroll
    for (int x = 0; x < N; ++x) f.mc[x] = (uint32_t)x < stride ? a.pn_mc[(size_t)e * stride + x] : 0.0;
    f.uc = a.pn_uc[e]; f.rho = a.pn_rho[e]; f.k = s.k; f.th = s.th;
    double d, v;
    brent_minimum(MemoEval<N>{&f, &tab, su, st}, d, v);
    a.pn_d[e] = d; a.pn_v[e] = v;
  }
}

// chisq of a node = 2 * (f_closest(d_node) - v_closest) (ref src/query.cpp:262-279); candidate when below the threshold and not the root
template <int N>
__global__ void __launch_bounds__(128) place_chisq_kernel(const PlaceArgs a, const LlhTables tab)
{
  const SolveArgs& s = a.s;
  if (s.counters[2] & kErrRedo) return;
  if (s.counters[2] & kErrNodeOverflow) return; // entries are incomplete: the host grows the node arrays and runs the batch again
  const uint32_t n = a.counters[5], stride = s.th + 1;
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t fl = a.pn_flags[e];
    if (!(fl & kPnEligible)) continue;
    const uint32_t r = a.pn_read[e], g = a.pn_se[e];
    const uint32_t cl = (uint32_t)s.closest[r];
    Objective<N> f;
#pragma unroll
    for (int x = 0; x < N; ++x) f.mc[x] = (uint32_t)x < stride ? (double)s.rec_hist[(size_t)cl * stride + x] : 0.0;
    f.uc = (double)s.onmers[r] - (double)s.rec_match[cl]; f.rho = s.rho[s.rec_slot[cl] & 0x7FFFFFFFu]; f.k = s.k; f.th = s.th;
    const double chisq = 2 * (f.eval(tab, a.pn_d[e]) - s.rec_v[cl]);
    a.pn_chisq[e] = chisq;
    if ((chisq < a.chisq_value) && a.parent[g] != 0) a.pn_flags[e] = fl | kPnCandidate;
  }
}

// candidates in ascending se: lwr = exp(-chisq/2) / total, the total summed in that order (ref src/query.cpp:284-296)
__global__ void __launch_bounds__(128) place_emit_kernel(const PlaceArgs a)
{
  const SolveArgs& s = a.s;
  if (s.counters[2] & kErrRedo) return;
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  krepp_placement_t* out = static_cast<krepp_placement_t*>(a.placements);
  for (uint32_t r = gwarp; r < s.n_reads; r += nwarps) {
    const uint32_t nb = a.pn_begin[r], cnt = a.pn_count[r];
    if (!cnt) continue;
    double total = 0;
    uint32_t nc = 0;
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
      const uint32_t j = j0 + lane;
      const bool cand = j < cnt && (a.pn_flags[nb + j] & kPnCandidate);
      const double w = cand ? exp(-a.pn_chisq[nb + j] / 2) : 0.0;
      uint32_t cm = __ballot_sync(0xFFFFFFFFu, cand);
      nc += __popc(cm);
      while (cm) { const int src = __ffs(cm) - 1; cm &= cm - 1; total = total + __shfl_sync(0xFFFFFFFFu, w, src); }
    }
    if (!nc) continue;
    uint32_t pbegin = 0;
    if (lane == 0) pbegin = atomicAdd(a.counters + 3, nc);
    pbegin = __shfl_sync(0xFFFFFFFFu, pbegin, 0);
    if ((uint64_t)pbegin + nc > a.place_cap) { if (lane == 0) atomicOr(a.counters + 2, kErrPlaceOverflow); continue; }
    uint32_t done = 0;
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
      const uint32_t j = j0 + lane;
      const bool cand = j < cnt && (a.pn_flags[nb + j] & kPnCandidate);
      const uint32_t cm = __ballot_sync(0xFFFFFFFFu, cand);
      if (cand) {
        const uint32_t e = nb + j, g = a.pn_se[e];
        const double bl = a.blen[g], mid = isnan(bl) ? 0.0 : bl / 2.0, d = a.pn_d[e], c = a.pn_chisq[e];
        krepp_placement_t p; p.read = r; p.se = g; p.pendant = jukes_cantor(d) - mid; p.distal = mid; p.loglik = -a.pn_v[e];
        p.lwr = exp(-c / 2) / total; p.d_llh = d; p.chisq = c;
        out[pbegin + done + __popc(cm & lt_mask)] = p;
      }
      done += __popc(cm);
    }
    if (lane == 0) { a.place_begin[r] = pbegin; a.place_count[r] = nc; }
  }
}

cudaError_t launch_place(const PlaceArgs& a, const LlhTables& tab, int grid, int sms, cudaStream_t stream, StageClock* clk)
{
  if (a.s.th + 1 <= 5) place_collect_kernel<5><<<grid, 128, 0, stream>>>(a);
  else place_collect_kernel<kMaxTh + 1><<<grid, 128, 0, stream>>>(a);
  if (clk) clk->tick("place_collect_kernel", stream);
  if (a.s.th + 1 <= 5) place_solve_kernel<5><<<sms * 8, 128, 0, stream>>>(a, tab);
  else place_solve_kernel<kMaxTh + 1><<<sms * 8, 128, 0, stream>>>(a, tab);
  if (clk) clk->tick("place_solve_kernel", stream);
  if (a.s.th + 1 <= 5) place_chisq_kernel<5><<<sms * 8, 128, 0, stream>>>(a, tab);
  else place_chisq_kernel<kMaxTh + 1><<<sms * 8, 128, 0, stream>>>(a, tab);
  place_emit_kernel<<<grid, 128, 0, stream>>>(a);
  if (clk) clk->tick("place_chisq_kernel+place_emit_kernel", stream);
  return cudaGetLastError();
}

// Segments of long reads back to reads (SegArgs, device.cuh): one warp per read.  A read that was not cut takes its one segment's
// rows as they are.  A cut read's segments are summed in a dense per-warp table of 2 * nleaves histograms (atomics: two
// segments' rows may hit one slot in the same step), the read's hdist_filt is the minimum over its segments, and the rows are
// written in the order the match step itself uses -- forward references by ascending se, then reverse -- under the gate of
// summarize_matches (ref src/query.cpp:101-106) taken against the READ's hdist_filt.  The match step runs with keep_all on a cut
// batch: a pair that passes for the read may sit beyond a segment's own gate in that segment (its nearest hit there is farther
// than the pair's nearest over the read), and its lookups there still count.  The table is zeroed again on the way out.
__global__ void __launch_bounds__(256) segment_combine_kernel(const SegArgs g)
{
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt_mask = (1u << lane) - 1u, stride = g.th + 1;
  if (g.v_counters[2] & kErrRedo) { // the match step ran out of room: hand its demand and flags to the host, which grows and reruns
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicMax(g.counters, g.v_counters[0]); atomicOr(g.counters + 2, g.v_counters[2]); }
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && g.v_counters[2]) atomicOr(g.counters + 2, g.v_counters[2]);
  uint32_t* acc = g.scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * 2u * g.nleaves * stride;
  const uint32_t nslots = 2u * g.nleaves;
  for (;;) {
    uint32_t r = 0;
    if (lane == 0) r = atomicAdd(g.claim, 1u);
    r = __shfl_sync(0xFFFFFFFFu, r, 0);
    if (r >= g.n_reads) break;
    const uint32_t v0 = g.vbegin[r], v1 = g.vbegin[r + 1];
    if (v1 - v0 == 1) { // not cut: its rows as they are, minus those the gate drops (the match step kept all of them, see below)
      const uint32_t cnt = g.v_rec_count[v0], vb = g.v_rec_begin[v0];
      const uint32_t f0 = g.v_hdfilt[2 * v0], f1 = g.v_hdfilt[2 * v0 + 1], g0 = 2u * f0 + 1u, g1 = 2u * f1 + 1u;
      auto passes = [&](uint32_t i) {
        if (g.keep_all) return true;
        uint32_t hdmin = 0xFFFFFFFFu;
        for (uint32_t x = 0; x < stride; ++x) if (g.v_rec_hist[(size_t)(vb + i) * stride + x] && hdmin == 0xFFFFFFFFu) hdmin = x;
        return !(hdmin > ((g.v_rec_slot[vb + i] >> 31) ? g1 : g0));
      };
      uint32_t total = 0;
      for (uint32_t c = 0; c < cnt; c += 32) total += __popc(__ballot_sync(0xFFFFFFFFu, c + lane < cnt && passes(c + lane)));
      uint32_t rb = 0;
      if (lane == 0 && total) rb = atomicAdd(g.counters, total);
      rb = __shfl_sync(0xFFFFFFFFu, rb, 0);
      const bool fits = (uint64_t)rb + total <= g.rec_cap;
      if (lane == 0) {
        if (!fits) atomicOr(g.counters + 2, kErrRecOverflow);
        g.onmers[r] = g.v_onmers[v0]; g.wn[2 * r] = g.v_wn[2 * v0]; g.wn[2 * r + 1] = g.v_wn[2 * v0 + 1];
        g.hdfilt[2 * r] = f0; g.hdfilt[2 * r + 1] = f1;
        g.rec_begin[r] = fits ? rb : 0; g.rec_count[r] = fits ? total : 0;
      }
      uint32_t done = 0;
      for (uint32_t c = 0; c < cnt && fits; c += 32) {
        const uint32_t i = c + lane;
        const bool pass = i < cnt && passes(i);
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pass);
        if (pass) {
          const uint32_t at = rb + done + __popc(pm & lt_mask);
          g.rec_read[at] = r; g.rec_slot[at] = g.v_rec_slot[vb + i];
          for (uint32_t x = 0; x < stride; ++x) g.rec_hist[(size_t)at * stride + x] = g.v_rec_hist[(size_t)(vb + i) * stride + x];
        }
        done += __popc(pm);
      }
      continue;
    }
    uint32_t onm = 0, w0 = 0, w1 = 0, f0 = 0xFFFFFFFFu, f1 = 0xFFFFFFFFu;
    for (uint32_t v = v0 + lane; v < v1; v += 32) {
      onm += g.v_onmers[v]; w0 += g.v_wn[2 * v]; w1 += g.v_wn[2 * v + 1];
      f0 = min(f0, g.v_hdfilt[2 * v]); f1 = min(f1, g.v_hdfilt[2 * v + 1]);
    }
    for (int o = 16; o; o >>= 1) {
      onm += __shfl_xor_sync(0xFFFFFFFFu, onm, o); w0 += __shfl_xor_sync(0xFFFFFFFFu, w0, o); w1 += __shfl_xor_sync(0xFFFFFFFFu, w1, o);
      f0 = min(f0, __shfl_xor_sync(0xFFFFFFFFu, f0, o)); f1 = min(f1, __shfl_xor_sync(0xFFFFFFFFu, f1, o));
    }
    const uint32_t g0 = 2u * f0 + 1u, g1 = 2u * f1 + 1u; // uint32 wrap kept, as in the reference
    for (uint32_t v = v0; v < v1; ++v) {
      const uint32_t cnt = g.v_rec_count[v], vb = g.v_rec_begin[v];
      for (uint32_t i = lane; i < cnt; i += 32) {
        const uint32_t slot = g.v_rec_slot[vb + i];
        uint32_t* dst = acc + ((size_t)(slot >> 31) * g.nleaves + g.leaf_rank[slot & 0x7FFFFFFFu]) * stride;
        for (uint32_t x = 0; x < stride; ++x) { const uint32_t c = g.v_rec_hist[(size_t)(vb + i) * stride + x]; if (c) atomicAdd(dst + x, c); }
      }
    }
    __syncwarp();
    uint32_t total = 0;
    for (uint32_t c = 0; c < nslots; c += 32) {
      const uint32_t j = c + lane;
      bool pass = false;
      if (j < nslots) {
        uint32_t hdmin = 0xFFFFFFFFu;
        for (uint32_t x = 0; x < stride; ++x) if (__ldcg(acc + (size_t)j * stride + x) && hdmin == 0xFFFFFFFFu) hdmin = x;
        pass = hdmin != 0xFFFFFFFFu && (g.keep_all || !(hdmin > (j >= g.nleaves ? g1 : g0)));
      }
      total += __popc(__ballot_sync(0xFFFFFFFFu, pass));
    }
    uint32_t rb = 0;
    if (lane == 0 && total) rb = atomicAdd(g.counters, total);
    rb = __shfl_sync(0xFFFFFFFFu, rb, 0);
    const bool fits = (uint64_t)rb + total <= g.rec_cap;
    uint32_t done = 0;
    for (uint32_t c = 0; c < nslots; c += 32) {
      const uint32_t j = c + lane;
      bool pass = false, any = false;
      uint32_t hv[kMaxTh + 1];
      if (j < nslots) {
        uint32_t hdmin = 0xFFFFFFFFu;
        for (uint32_t x = 0; x < stride; ++x) { hv[x] = __ldcg(acc + (size_t)j * stride + x); if (hv[x] && hdmin == 0xFFFFFFFFu) hdmin = x; }
        any = hdmin != 0xFFFFFFFFu;
        pass = any && (g.keep_all || !(hdmin > (j >= g.nleaves ? g1 : g0)));
      }
      const uint32_t pm = __ballot_sync(0xFFFFFFFFu, pass);
      if (pass && fits) {
        const uint32_t at = rb + done + __popc(pm & lt_mask);
        g.rec_read[at] = r;
        g.rec_slot[at] = (j >= g.nleaves ? 0x80000000u : 0u) | g.leaf_se[j >= g.nleaves ? j - g.nleaves : j];
        for (uint32_t x = 0; x < stride; ++x) g.rec_hist[(size_t)at * stride + x] = hv[x];
      }
      if (any) for (uint32_t x = 0; x < stride; ++x) acc[(size_t)j * stride + x] = 0u;
      done += __popc(pm);
    }
    if (lane == 0) {
      if (!fits) atomicOr(g.counters + 2, kErrRecOverflow);
      g.onmers[r] = onm; g.wn[2 * r] = w0; g.wn[2 * r + 1] = w1; g.hdfilt[2 * r] = f0; g.hdfilt[2 * r + 1] = f1;
      g.rec_begin[r] = fits ? rb : 0; g.rec_count[r] = fits ? total : 0;
    }
    __syncwarp();
  }
}

cudaError_t launch_segment_combine(const SegArgs& g, int ctas, cudaStream_t stream)
{
  segment_combine_kernel<<<ctas, 256, 0, stream>>>(g);
  return cudaGetLastError();
}

// `krepp seek` (SBatch::seek_sequences ref src/seek.cpp:22-53) on a sketch handle, whose one reference makes a read's records
// its two per-strand summaries: one thread per read prints the smaller of the two strands' distances.  Both strands are
// solved as soon as either matched anything (:36-41), so a strand without a record gets the histogram of zeros (every k-mer a
// mismatch) solved here; NaN when neither strand has a record.
template <int N>
__global__ void __launch_bounds__(128) seek_kernel(const SolveArgs a, const LlhTables tab, double* out)
{
  __shared__ double su[4];
  __shared__ DTerms st[4];
  if (a.counters[2] & kErrRedo) return;
  if (threadIdx.x < 4) {
    double u[4];
    brent_first_points(u);
    const double ut = threadIdx.x == 0 ? u[0] : threadIdx.x == 1 ? u[1] : threadIdx.x == 2 ? u[2] : u[3];
    su[threadIdx.x] = ut;
    st[threadIdx.x] = d_terms(tab, ut, a.k);
  }
  __syncthreads();
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
    const uint32_t b = a.rec_begin[r], c = a.rec_count[r];
    double res = nan("");
    if (c) {
      double d[2] = {0.0, 0.0};
      bool have[2] = {false, false};
      uint32_t se = 1;
      for (uint32_t j = 0; j < c && j < 2; ++j) { const uint32_t slot = a.rec_slot[b + j], sd = slot >> 31; se = slot & 0x7FFFFFFFu; d[sd] = a.rec_d[b + j]; have[sd] = true; }
      for (int sd = 0; sd < 2; ++sd) {
        if (have[sd]) continue;
        Objective<N> f;
#pragma unroll
        for (int x = 0; x < N; ++x) f.mc[x] = 0.0;
        f.uc = (double)a.onmers[r]; f.rho = a.rho[se]; f.k = a.k; f.th = a.th;
        double v;
        brent_minimum(MemoEval<N>{&f, &tab, su, st}, d[sd], v);
      }
      res = d[0] < d[1] ? d[0] : d[1];
    }
    out[r] = res;
  }
}

cudaError_t launch_seek(const SolveArgs& a, const LlhTables& tab, double* out, int sms, cudaStream_t stream)
{
  if (a.th + 1 <= 5) seek_kernel<5><<<sms * 8, 128, 0, stream>>>(a, tab, out);
  else seek_kernel<kMaxTh + 1><<<sms * 8, 128, 0, stream>>>(a, tab, out);
  return cudaGetLastError();
}

cudaError_t launch_solve(const SolveArgs& a, const LlhTables& tab, int sms, cudaStream_t stream, StageClock* clk)
{
  const int grid = sms * 8;
  if (a.memo_mask) { const cudaError_t e = cudaMemsetAsync(a.memo_key, 0, 8ull * ((size_t)a.memo_mask + 1), stream); if (e != cudaSuccess) return e; }
  gate_kernel<<<sms * 8, 256, 0, stream>>>(a); // latency-bound streaming kernels run at full occupancy (64 warps per SM)
  if (clk) clk->tick("gate_kernel", stream);
  if (a.th + 1 <= 5) solve_kernel<5><<<grid, 128, 0, stream>>>(a, tab);
  else solve_kernel<kMaxTh + 1><<<grid, 128, 0, stream>>>(a, tab);
  if (clk) clk->tick("solve_kernel", stream);
  if (a.memo_mask) alias_kernel<<<sms * 8, 256, 0, stream>>>(a);
  merge_kernel<<<sms * 16, 128, 0, stream>>>(a);
  if (a.want_chisq) chisq_kernel<<<sms * 16, 128, 0, stream>>>(a, tab);
  if (clk) clk->tick(a.want_chisq ? "alias_kernel+merge_kernel+chisq_kernel" : "alias_kernel+merge_kernel", stream);
  return cudaGetLastError();
}

} // namespace krepp

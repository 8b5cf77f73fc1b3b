// match.cu -- K1..K3 of the query path fused into one persistent kernel (sm_100a).
//
// One warp owns one read at a time (reads are claimed from a global counter, so ragged lengths balance):
//   1. 128-bit loads of the read's ASCII bases -> 2-bit codes + validity bits, packed MSB-first in shared memory
//      (ref src/common.cpp:10-18 seq_nt4_table, src/common.hpp:225-243 compute/update_encoding).
//   2. every lane extracts k-mer windows from the packed stream, forms the reverse complement
//      (ref src/common.hpp:177-186), the LSH bucket id rix = pext(bp, mask_hash_bp) (ref src/lshf.cpp:62) and the 32-bit
//      residual q = pext(lr, mask_drop_lr) (ref src/lshf.cpp:64-69), eligibility + row offset (ref src/index.hpp:27,
//      src/index.cpp:160-168) for both strands (ref src/query.cpp:82-91).
//   3. the bucket [inc[off-1], inc[off]) of the flat table is scanned with XOR/OR/popc (ref src/common.hpp:175,
//      src/query.cpp:361-368); hits expand their colour through the se->(se,se) DAG to leaves on the device
//      (ref src/query.cpp:369-387) and bump the per-(strand, leaf) Hamming histogram (ref src/query.hpp:153-176).
//
// The reference's Minfo::update_match keeps, per (strand, leaf, position), the MINIMUM Hamming distance over all
// matching entries.  All entries that can match one (read, strand, position) live in one bucket, so that minimum is
// formed inside a single lookup: a lookup with one hit entry commits directly, a lookup with several hit entries goes
// through a per-warp marker array (atomicMin, then commit-and-clear), which makes the result independent of the order
// in which lanes run.  Histograms live in a per-warp accumulator and are emitted as records
// (read, strand<<31|leaf_se, hist[0..th]) in (strand, leaf) order when the read is finished.
#include "device.cuh"
#include "solve.cuh"

namespace krepp {

constexpr int kWarpsPerCta = 8;
constexpr int kTileWindows = 128;              // windows handled per tile: 4 per lane
constexpr int kTileWords = 12;                 // 16 bases per 32-bit word -> 192 bases >= 128 + 32 - 1
constexpr int kLocalStack = 32;

struct WarpSmem {
  uint32_t code[kTileWords + 1];
  uint32_t valid[kTileWords / 2 + 1];
};

__device__ __forceinline__ void encode4(uint32_t u, uint32_t& code8, uint32_t& valid4)
{
  const uint32_t up = u & 0xDFDFDFDFu; // fold lower case
  const uint32_t vm = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) | __vcmpeq4(up, 0x47474747u) | __vcmpeq4(up, 0x54545454u);
  uint32_t x = (u >> 1) & 0x03030303u; // A0 C1 G3 T2
  x ^= (x >> 1) & 0x01010101u;         // A0 C1 G2 T3  (ref nt4_bp_table)
  x &= vm;
  code8 = (x * 0x40100401u) >> 24;                        // first char -> most significant pair
  valid4 = ((vm & 0x01010101u) * 0x08040201u) >> 24;      // first char -> bit 3
}

__device__ __forceinline__ uint32_t pext_runs(uint64_t x, const DevRun* runs, uint32_t n)
{
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; ++i) r |= ((uint32_t)(x >> runs[i].src) & runs[i].mask) << runs[i].dst;
  return r;
}

__device__ __forceinline__ uint32_t even_bits16(uint32_t t)
{
  t &= 0x55555555u;
  t = (t | (t >> 1)) & 0x33333333u;
  t = (t | (t >> 2)) & 0x0F0F0F0Fu;
  t = (t | (t >> 4)) & 0x00FF00FFu;
  t = (t | (t >> 8)) & 0x0000FFFFu;
  return t;
}

__device__ __forceinline__ uint64_t revcomp(uint64_t bp, uint32_t k)
{
  uint64_t y = __brevll(bp);
  y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
  return (~y) >> (64 - 2 * k);
}

struct WarpCtx {
  uint32_t* acc;      // [2*nleaves*(th+1)]
  uint32_t* bitmap;   // [ceil(2*nleaves/32)]
  uint32_t* marker;   // [nleaves]
  uint32_t* stack;
  uint32_t stack_cap;
  uint32_t stride;    // th+1
  uint32_t nleaves;
  uint32_t* err;
};

__device__ __forceinline__ void commit(const WarpCtx& w, uint32_t strand, uint32_t rank, uint32_t hd)
{
  const uint32_t slot = strand * w.nleaves + rank;
  atomicAdd(&w.acc[slot * w.stride + hd], 1u);
  atomicOr(&w.bitmap[slot >> 5], 1u << (slot & 31));
}

// Lane-local colour expansion for a lookup with exactly one hit entry (no dedupe needed): depth-first with a small
// private stack.  The host only enables it (ix.local_expand) when the deepest colour DAG of the index fits.
__device__ __forceinline__ void expand_local(const DevIndex& ix, const WarpCtx& w, uint32_t se, uint32_t strand, uint32_t hd)
{
  uint32_t st[kLocalStack];
  int sp = 0;
  st[sp++] = se;
  while (sp) {
    const uint32_t s = st[--sp];
    const uint32_t kd = ix.kind[s];
    if (kd == 1) commit(w, strand, ix.leaf_rank[s], hd);
    else if (kd == 2) { const uint2 c = ix.pse[s]; st[sp++] = c.y; st[sp++] = c.x; }
  }
}

// Warp-cooperative colour expansion over the per-warp HBM stack.  mode 0: commit every leaf; mode 1: marker[leaf] =
// min(marker, hd); mode 2: commit marker value once per leaf and reset the marker.
__device__ void expand_coop(const DevIndex& ix, const WarpCtx& w, uint32_t se, uint32_t strand, uint32_t hd, int mode)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t size = 1;
  if (lane == 0) w.stack[0] = se;
  __syncwarp();
  while (size) {
    const uint32_t take = min(size, 32u);
    uint32_t s = 0, kd = 0;
    if (lane < take) { s = w.stack[size - 1 - lane]; kd = ix.kind[s]; }
    __syncwarp();
    size -= take;
    if (kd == 1) {
      const uint32_t rank = ix.leaf_rank[s];
      if (mode == 0) commit(w, strand, rank, hd);
      else if (mode == 1) atomicMin(&w.marker[rank], hd);
      else { const uint32_t old = atomicExch(&w.marker[rank], 0xFFFFFFFFu); if (old != 0xFFFFFFFFu) commit(w, strand, rank, old); }
    }
    const uint32_t ex = __ballot_sync(0xFFFFFFFFu, kd == 2);
    const uint32_t nex = __popc(ex);
    if (size + 2 * nex > w.stack_cap) { if (lane == 0) atomicOr(w.err, kErrStackOverflow); return; }
    if (kd == 2) {
      const uint2 c = ix.pse[s];
      const uint32_t at = size + 2 * __popc(ex & ((1u << lane) - 1));
      w.stack[at] = c.x; w.stack[at + 1] = c.y;
    }
    size += 2 * nex;
    __syncwarp();
  }
}

// Careful path for one lookup with several hit entries (or a colour too deep for the private stack): the whole warp
// rescans the bucket, marks per-leaf minima, then commits them.
__device__ void careful_lookup(const DevIndex& ix, const WarpCtx& w, uint64_t begin, uint32_t len, uint32_t q, uint32_t strand, uint32_t th)
{
  const uint32_t lane = threadIdx.x & 31;
  for (int pass = 1; pass <= 2; ++pass) {
    for (uint32_t base = 0; base < len; base += 32) {
      uint32_t se = 0, hd = 0xFFFFFFFFu;
      if (base + lane < len) {
        const uint2 e = ix.cmer[begin + base + lane];
        const uint32_t z = e.x ^ q;
        hd = __popc((z | (z >> 16)) & 0xFFFFu);
        se = e.y;
      }
      uint32_t hits = __ballot_sync(0xFFFFFFFFu, hd <= th);
      while (hits) {
        const int src = __ffs(hits) - 1;
        hits &= hits - 1;
        const uint32_t hse = __shfl_sync(0xFFFFFFFFu, se, src), hhd = __shfl_sync(0xFFFFFFFFu, hd, src);
        expand_coop(ix, w, hse, strand, hhd, pass);
      }
    }
  }
}

template <bool TAP>
__global__ void __launch_bounds__(kWarpsPerCta * 32) match_kernel(const DevIndex ix, const MatchArgs a)
{
  __shared__ WarpSmem smem[kWarpsPerCta];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * kWarpsPerCta + warp;
  WarpSmem& sm = smem[warp];
  const uint32_t k = ix.k, th = a.th, stride = th + 1, nleaves = ix.nleaves;
  const uint32_t nslots = 2 * nleaves, nbm = (nslots + 31) >> 5;

  WarpCtx w;
  w.acc = a.acc + (size_t)gwarp * nslots * stride;
  w.bitmap = a.bitmap + (size_t)gwarp * nbm;
  w.marker = a.marker + (size_t)gwarp * nleaves;
  w.stack = a.stack + (size_t)gwarp * a.stack_cap;
  w.stack_cap = a.stack_cap; w.stride = stride; w.nleaves = nleaves; w.err = a.counters + 2;

  unsigned long long st_bytes = 0, st_lookups = 0, st_entries = 0;

  for (;;) {
    uint32_t read = 0;
    if (lane == 0) read = atomicAdd(a.counters + 1, 1u);
    read = __shfl_sync(0xFFFFFFFFu, read, 0);
    if (read >= a.n_reads) break;
    const uint64_t off = a.offsets[read];
    const uint64_t len = a.offsets[read + 1] - off;
    uint32_t onmers = 0, wn0 = 0, wn1 = 0, filt0 = 0xFFFFFFFFu, filt1 = 0xFFFFFFFFu;
    st_bytes += (lane == 0) ? len : 0;

    for (uint64_t t0 = 0; t0 + k <= len; t0 += kTileWindows) {
      // ---- 1. load + encode the tile's bases: [t0, t0 + kTileWindows + k - 1) clipped to the read
      const uint64_t rem = len - t0;                                  // bases available from t0
      const uint32_t nb = (uint32_t)min((uint64_t)(kTileWindows + k - 1), rem);
      const char* p0 = a.bases + off + t0;
      const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 15);
      const char* al = p0 - sh;
      uint32_t cw = 0, vw = 0;
      if (lane <= kTileWords) {
        const char* cp = al + 16 * lane;
        if (cp < p0 + nb) {
          uint4 u;
          if (cp + 16 <= a.bases + a.n_bases) u = __ldg(reinterpret_cast<const uint4*>(cp));
          else {
            unsigned char b[16];
            for (int i = 0; i < 16; ++i) b[i] = (cp + i < a.bases + a.n_bases) ? (unsigned char)cp[i] : 0;
            u.x = b[0] | b[1] << 8 | b[2] << 16 | (uint32_t)b[3] << 24; u.y = b[4] | b[5] << 8 | b[6] << 16 | (uint32_t)b[7] << 24;
            u.z = b[8] | b[9] << 8 | b[10] << 16 | (uint32_t)b[11] << 24; u.w = b[12] | b[13] << 8 | b[14] << 16 | (uint32_t)b[15] << 24;
          }
          uint32_t c0, c1, c2, c3, v0, v1, v2, v3;
          encode4(u.x, c0, v0); encode4(u.y, c1, v1); encode4(u.z, c2, v2); encode4(u.w, c3, v3);
          cw = c0 << 24 | c1 << 16 | c2 << 8 | c3;
          vw = v0 << 12 | v1 << 8 | v2 << 4 | v3;
        }
      }
      // align the stream to the tile start: word t covers tile bases 16t .. 16t+15
      const uint32_t cn = __shfl_down_sync(0xFFFFFFFFu, cw, 1), vn = __shfl_down_sync(0xFFFFFFFFu, vw, 1);
      uint32_t cwa = __funnelshift_l(cn, cw, 2 * sh);
      uint32_t vwa = (((vw << 16) | vn) << sh) >> 16;
      // clip validity to the bases that belong to this read
      {
        const int first = 16 * (int)lane;
        const int keep = (int)nb - first;             // number of leading bases of this word inside the read
        if (keep <= 0) vwa = 0; else if (keep < 16) vwa &= 0xFFFFu << (16 - keep);
      }
      const uint32_t vhi = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane) & 31), vlo = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane + 1) & 31);
      __syncwarp();
      if (lane <= kTileWords) sm.code[lane] = cwa;
      if (lane <= kTileWords / 2) sm.valid[lane] = (vhi << 16) | vlo;
      __syncwarp();

      // ---- 2 + 3. windows -> lookups -> bucket scans
      const uint32_t nwin = (uint32_t)min((uint64_t)kTileWindows, rem - k + 1);
      for (uint32_t j = 0; j < kTileWindows / 32; ++j) {
        const uint32_t p = lane + 32 * j;
        bool valid = false;
        uint64_t bp = 0;
        if (p < nwin) {
          const uint32_t vj = p >> 5, vs = p & 31;
          const uint32_t vx = __funnelshift_l(sm.valid[vj + 1], sm.valid[vj], vs);
          valid = (vx >> (32 - k)) == (0xFFFFFFFFu >> (32 - k));
          const uint32_t cj = p >> 4, cs = 2 * (p & 15);
          const uint32_t w0 = sm.code[cj], w1 = sm.code[cj + 1], w2 = sm.code[cj + 2];
          const uint64_t x = ((uint64_t)__funnelshift_l(w1, w0, cs) << 32) | __funnelshift_l(w2, w1, cs);
          bp = x >> (64 - 2 * k);
        }
        onmers += valid;
        // per strand: hash, eligibility, bucket scan
        uint32_t n_hit[2] = {0, 0}, hit_se[2] = {0, 0}, hit_hd[2] = {0, 0}, lk_len[2] = {0, 0}, lk_q[2] = {0, 0};
        uint64_t lk_begin[2] = {0, 0};
#pragma unroll
        for (uint32_t strand = 0; strand < 2; ++strand) {
          if (!valid) continue;
          const uint64_t e = strand ? revcomp(bp, k) : bp;
          const uint32_t rix = pext_runs(e, ix.hash_runs, ix.n_hash_runs);
          uint32_t quo, res;
          if (ix.m_shift != 0xFFFFFFFFu) { quo = rix >> ix.m_shift; res = rix & (ix.m - 1); }
          else { quo = rix / ix.m; res = rix - quo * ix.m; }
          const int32_t numer = ix.res_numer[res];
          if (numer == 0) continue;
          const uint32_t offset = numer > 1 ? quo * (uint32_t)numer + res : quo;
          const uint32_t qbp = pext_runs(e, ix.drop_runs, ix.n_drop_runs);
          const uint32_t q = even_bits16(qbp) | (even_bits16(qbp >> 1) << 16);
          if (strand) ++wn1; else ++wn0;
          const uint64_t begin = offset ? ix.inc[offset - 1] : 0ull;
          const uint64_t end = ix.inc[offset];
          const uint32_t blen = (uint32_t)(end - begin);
          if (TAP) {
            const unsigned long long at = atomicAdd(a.tap_count, 1ull);
            const uint32_t pos = strand ? (uint32_t)(len - (t0 + p) - k) : (uint32_t)(t0 + p);
            if (at < a.tap_cap) a.tap[at] = make_uint4(read, strand << 31 | pos, rix, q);
          }
          st_lookups += 1; st_entries += blen; st_bytes += 16 + 8ull * blen;
          uint32_t cnt = 0, fse = 0, fhd = 0, mn = 0xFFFFFFFFu;
          for (uint32_t i = 0; i < blen; ++i) {
            const uint2 ent = __ldg(&ix.cmer[begin + i]);
            const uint32_t z = ent.x ^ q;
            const uint32_t hd = __popc((z | (z >> 16)) & 0xFFFFu);
            if (hd <= th) { if (!cnt) { fse = ent.y; fhd = hd; } ++cnt; mn = min(mn, hd); }
          }
          if (strand) filt1 = min(filt1, mn); else filt0 = min(filt0, mn);
          n_hit[strand] = cnt; hit_se[strand] = fse; hit_hd[strand] = fhd; lk_len[strand] = blen; lk_q[strand] = q; lk_begin[strand] = begin;
        }
        // commit hits
#pragma unroll
        for (uint32_t strand = 0; strand < 2; ++strand) {
          bool careful = n_hit[strand] > 1;
          if (n_hit[strand] == 1) {
            const uint32_t kd = ix.kind[hit_se[strand]];
            if (kd == 1) commit(w, strand, ix.leaf_rank[hit_se[strand]], hit_hd[strand]);
            else if (kd == 2) { if (ix.local_expand) expand_local(ix, w, hit_se[strand], strand, hit_hd[strand]); else careful = true; }
          }
          uint32_t need = __ballot_sync(0xFFFFFFFFu, careful);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint64_t b = __shfl_sync(0xFFFFFFFFu, lk_begin[strand], src);
            const uint32_t l = __shfl_sync(0xFFFFFFFFu, lk_len[strand], src), qq = __shfl_sync(0xFFFFFFFFu, lk_q[strand], src);
            careful_lookup(ix, w, b, l, qq, strand, th);
          }
        }
      }
      __syncwarp();
    }

    // ---- per-read scalars
    for (int o = 16; o; o >>= 1) {
      onmers += __shfl_xor_sync(0xFFFFFFFFu, onmers, o);
      wn0 += __shfl_xor_sync(0xFFFFFFFFu, wn0, o); wn1 += __shfl_xor_sync(0xFFFFFFFFu, wn1, o);
      filt0 = min(filt0, __shfl_xor_sync(0xFFFFFFFFu, filt0, o)); filt1 = min(filt1, __shfl_xor_sync(0xFFFFFFFFu, filt1, o));
    }
    // ---- emit this read's records in (strand, leaf) order and reset the accumulator
    __syncwarp();
    uint32_t total = 0;
    for (uint32_t wbase = 0; wbase < nbm; wbase += 32) {
      const uint32_t bits = (wbase + lane < nbm) ? __ldcg(&w.bitmap[wbase + lane]) : 0u;
      uint32_t c = __popc(bits);
      for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
      total += c;
    }
    uint32_t rbegin = 0;
    if (lane == 0 && total) rbegin = atomicAdd(a.counters, total);
    rbegin = __shfl_sync(0xFFFFFFFFu, rbegin, 0);
    const bool fits = (uint64_t)rbegin + total <= a.rec_cap;
    if (!fits && lane == 0) atomicOr(a.counters + 2, kErrRecOverflow);
    uint32_t done = 0;
    for (uint32_t wbase = 0; wbase < nbm; wbase += 32) {
      uint32_t bits = (wbase + lane < nbm) ? __ldcg(&w.bitmap[wbase + lane]) : 0u;
      const uint32_t c = __popc(bits);
      uint32_t incl = c;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
      uint32_t at = rbegin + done + incl - c;
      done += __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (wbase + lane < nbm && bits) w.bitmap[wbase + lane] = 0;
      while (bits) {
        const uint32_t b = __ffs(bits) - 1;
        bits &= bits - 1;
        const uint32_t slot = (wbase + lane) * 32 + b;
        const uint32_t strand = slot >= nleaves, rank = slot - strand * nleaves;
        uint32_t* h = w.acc + (size_t)slot * stride;
        if (fits) {
          a.rec_read[at] = read;
          a.rec_slot[at] = strand << 31 | ix.leaf_se[rank];
          for (uint32_t x = 0; x < stride; ++x) a.rec_hist[(size_t)at * stride + x] = __ldcg(&h[x]);
        }
        for (uint32_t x = 0; x < stride; ++x) h[x] = 0;
        ++at;
      }
    }
    if (lane == 0) {
      a.onmers[read] = onmers; a.wn[2 * read] = wn0; a.wn[2 * read + 1] = wn1;
      a.hdfilt[2 * read] = filt0; a.hdfilt[2 * read + 1] = filt1;
      a.rec_begin[read] = fits ? rbegin : 0; a.rec_count[read] = fits ? total : 0;
      st_bytes += 64ull * total;
    }
    __syncwarp();
  }
  // ---- roofline accounting
  for (int o = 16; o; o >>= 1) {
    st_bytes += __shfl_xor_sync(0xFFFFFFFFu, st_bytes, o);
    st_lookups += __shfl_xor_sync(0xFFFFFFFFu, st_lookups, o);
    st_entries += __shfl_xor_sync(0xFFFFFFFFu, st_entries, o);
  }
  if (lane == 0) { atomicAdd(a.stats, st_bytes); atomicAdd(a.stats + 1, st_lookups); atomicAdd(a.stats + 2, st_entries); }
}

// ------------------------------------------------------------------------------------------------ host launchers

int match_resident_warps(int device)
{
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel<false>, kWarpsPerCta * 32, 0);
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm * kWarpsPerCta;
}

cudaError_t launch_match(const DevIndex& ix, const MatchArgs& a, int resident_warps, bool tap, cudaStream_t stream)
{
  const int grid = resident_warps / kWarpsPerCta;
  if (tap) match_kernel<true><<<grid, kWarpsPerCta * 32, 0, stream>>>(ix, a);
  else match_kernel<false><<<grid, kWarpsPerCta * 32, 0, stream>>>(ix, a);
  return cudaGetLastError();
}

} // namespace krepp

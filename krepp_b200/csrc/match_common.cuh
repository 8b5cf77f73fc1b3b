// match_common.cuh -- pieces shared by the fused match kernel (match.cu) and the bucket-sorted pipeline (sorted.cu):
// tile geometry, the per-warp lookup list, ASCII -> 2-bit encoding, the byte-LUT software pext and tile_lookups
// (stages A0 + A1 of the path: read bases -> eligible lookups of both strands).
#pragma once
#include "device.cuh"

namespace krepp {

constexpr int kTileWindows = 128;              // windows handled per tile: 4 per lane
constexpr int kTileWords = 12;                 // 16 bases per 32-bit word -> 192 bases >= 128 + 32 - 1
constexpr int kLocalStack = 32;
constexpr uint32_t kClaim = 4;                // reads claimed per atomic on the global work counter
constexpr int kMaxLookups = 2 * kTileWindows;  // both strands

struct WarpSmem {
  uint32_t lk_a[kMaxLookups];  // A1: row offset | strand<<31; A2: first entry of the bucket
  uint32_t lk_l[kMaxLookups];  // A2: bucket length | strand<<31
  uint32_t lk_q[kMaxLookups];  // residual encoding q of the query k-mer
  uint32_t code[kTileWords + 1];
  uint32_t valid[kTileWords / 2 + 1];
  uint32_t cursor;
};
// The same without the bucket fields: what the bucket-sorted pipeline's lookup kernels need per warp.
struct WarpSmemLite {
  uint32_t lk_a[kMaxLookups];  // row offset | strand<<31
  uint32_t lk_q[kMaxLookups];  // residual encoding q of the query k-mer
  uint32_t code[kTileWords + 1];
  uint32_t valid[kTileWords / 2 + 1];
  uint32_t cursor;
};
// LUT layout: [byte of the k-mer word][byte value] -> {rix fwd, q fwd, rix rc, q rc} parts; 7 bytes cover k <= 28
// lut_pext always reads the first seven byte tables, so at least seven are staged (all-zero past the k-mer's last byte)
__host__ __device__ inline uint32_t lut_chunks(uint32_t k) { const uint32_t n = (2 * k + 7) / 8; return n < 7 ? 7 : n; }

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

// Both software pexts of one k-mer word, for BOTH strands, by table lookup: the 2k-bit word is cut into bytes (4 bases
// each); every byte value maps to its pre-positioned contribution to rix = pext(bp, mask_hash_bp) and to
// q = pext(lr, mask_drop_lr) (already in the bit-plane form the index stores) of the forward k-mer (.x, .y) and of its
// reverse complement (.z, .w): a base at position p lands, complemented, at position k-1-p of the reverse complement,
// so the reverse strand is indexed by the SAME forward bytes and no reverse-complement word is ever formed.  Tables
// are built by the host from the index's ppos/npos (api.cu build_lut) and copied to shared memory once per CTA.
__device__ __forceinline__ uint4 lut_pext(const uint4* lut, uint32_t lo, uint32_t hi, bool wide)
{
  uint4 r = lut[lo & 0xFF];
  uint4 t;
#define KREPP_LUT_OR(expr) t = lut[expr]; r.x |= t.x; r.y |= t.y; r.z |= t.z; r.w |= t.w;
  KREPP_LUT_OR(256 + ((lo >> 8) & 0xFF))
  KREPP_LUT_OR(512 + ((lo >> 16) & 0xFF))
  KREPP_LUT_OR(768 + (lo >> 24))
  KREPP_LUT_OR(1024 + (hi & 0xFF))
  KREPP_LUT_OR(1280 + ((hi >> 8) & 0xFF))
  KREPP_LUT_OR(1536 + ((hi >> 16) & 0xFF))
  if (wide) { KREPP_LUT_OR(1792 + (hi >> 24)) } // only k > 28 reaches the eighth byte
#undef KREPP_LUT_OR
  return r;
}

// A0 + A1 of one tile of a read (see the header): ASCII bases -> 2-bit stream in shared memory -> k-mer windows -> bucket
// ids and residual encodings of both strands; the eligible lookups are compacted into sm.lk_a (row offset | strand << 31)
// and sm.lk_q.  Returns their number; onmers / wn0 / wn1 are the read's running counts (warp-uniform).
// bin_cnt (optional): shared-memory counters of the two-level lookup sort (sorted.cu); every eligible lookup also counts itself
// in the coarse bin of its row, bin_cnt[row >> bin_shift].
template <bool TAP, class WS>
__device__ __forceinline__ uint32_t tile_lookups(const DevIndex& ix, const MatchArgs& a, WS& sm, const uint4* lut, bool wide, uint32_t read,
                                                 uint64_t off, uint64_t len, uint64_t t0, uint32_t& onmers, uint32_t& wn0, uint32_t& wn1,
                                                 uint32_t* bin_cnt = nullptr, uint32_t bin_shift = 0)
{
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1, k = ix.k;
  // ---- A0. load + encode the tile's bases: [t0, t0 + kTileWindows + k - 1) clipped to the read
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
  const uint32_t cwa = __funnelshift_l(cn, cw, 2 * sh);
  uint32_t vwa = (((vw << 16) | vn) << sh) >> 16;
  { // clip validity to the bases that belong to this read
    const int keep = (int)nb - 16 * (int)lane;    // number of leading bases of this word inside the read
    if (keep <= 0) vwa = 0; else if (keep < 16) vwa &= 0xFFFFu << (16 - keep);
  }
  const uint32_t vhi = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane) & 31), vlo = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane + 1) & 31);
  __syncwarp();
  if (lane <= kTileWords) sm.code[lane] = cwa;
  if (lane <= kTileWords / 2) sm.valid[lane] = (vhi << 16) | vlo;
  if (lane == 0) sm.cursor = 0;
  __syncwarp();

  // ---- A1. windows -> k-mer words -> bucket ids; eligible lookups compacted into the list
  const uint32_t nwin = (uint32_t)min((uint64_t)kTileWindows, rem - k + 1);
  uint32_t nl = 0;
#pragma unroll
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
    onmers += __popc(__ballot_sync(0xFFFFFFFFu, valid));   // warp-uniform counts: no reduction at the end of the read
    const uint4 rq = lut_pext(lut, (uint32_t)bp, (uint32_t)(bp >> 32), wide);
#pragma unroll
    for (uint32_t strand = 0; strand < 2; ++strand) {
      const uint32_t rix = strand ? rq.z : rq.x, q = strand ? rq.w : rq.y;
      uint32_t quo, res;
      if (ix.m_shift != 0xFFFFFFFFu) { quo = rix >> ix.m_shift; res = rix & (ix.m - 1); }
      else { quo = rix / ix.m; res = rix - quo * ix.m; }
      const int32_t numer = ix.res_numer[res];
      const bool elig = valid && numer != 0;
      const uint32_t offset = (numer > 1 ? quo * (uint32_t)numer + res : quo) + ix.res_base[res];
      const uint32_t em = __ballot_sync(0xFFFFFFFFu, elig);
      if (elig) {
        const uint32_t idx = nl + __popc(em & lt_mask);
        sm.lk_a[idx] = offset | (strand << 31);
        sm.lk_q[idx] = q;
        if (bin_cnt) atomicAdd(&bin_cnt[offset >> bin_shift], 1u);
        if (TAP) {
          const unsigned long long at = atomicAdd(a.tap_count, 1ull);
          const uint32_t pos = strand ? (uint32_t)(len - (t0 + p) - k) : (uint32_t)(t0 + p);
          if (at < a.tap_cap) a.tap[at] = make_uint4(read, strand << 31 | pos, rix, q);
        }
      }
      nl += __popc(em);
      if (strand) wn1 += __popc(em); else wn0 += __popc(em);
    }
  }
  __syncwarp();
  return nl;
}

} // namespace krepp

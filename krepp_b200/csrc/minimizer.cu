// minimizer.cu -- the index-side k-mer kernel: what `krepp index` does to one reference genome
// before the colour unions -- RSeq::extract_mers (ref src/rqseq.cpp:51-144, sdust off, not canonical) followed by the per-bucket
// sort and unique of DynHT::fill_table (ref src/table.cpp:110-117,157-166,248-260).
//
//   every forward-strand k-mer with k valid bases is hashed with xur64_hash (murmur fmix64, ref src/common.hpp:147-155);
//   at every position whose valid run is at least w long, the k-mer with the smallest hash among the w-k+1 k-mers of the
//   window is the minimizer; it is kept when its LSH bucket id rix = pext(bp, mask_hash_bp) has a residue of this partial
//   index (rix % m <= r, or == r without frac) and lands in row (rix / m)(r + 1) + rix % m with the 32-bit residual encoding
//   pext(lr, mask_drop_lr) (ref src/lshf.cpp:62-69).
//
// Kernel: warp per tile of a sequence.  128 k-mers per tile, four per lane, from 128-bit loads of ASCII through the same
// 2-bit stream and byte-LUT pext as the query side (match_common.cuh); hashes go to shared memory, every lane takes the minimum
// of the w-k+1 hashes that end at its window ends, and surviving (row, encoding) keys are appended to one list with a
// warp-aggregated atomic.  The list is then sorted and made unique (cub::DeviceRadixSort / DeviceSelect: library primitives, as
// cuBLAS would be for a plain GEMM), which is exactly one leaf table: sorted unique encodings per row.
//
// The reference's end-of-sequence quirk (ref src/rqseq.cpp:112-116: at the last base an emit also happens when the valid run
// is shorter than w, from whatever the ring buffer of w-k+1 slots holds -- k-mers from before the last run of N, or
// zero-initialised slots) yields at most one key per sequence and depends on sequential state; the host computes it.
#include "../../include/krepp_b200.h"

#include "device.cuh"
#include "builder.hpp"
#include "handles.hpp"
#include "match_common.cuh"

#include <cub/cub.cuh>

#include <algorithm>
#include <string>
#include <vector>

namespace krepp {

constexpr int kMzWarps = 8;
constexpr int kMzKmers = 128;  // k-mers per tile: four per lane

struct MzWarpSmem {
  unsigned long long z[kMzKmers];  // xur64_hash of the tile's k-mers
  uint2 rq[kMzKmers];              // {rix, residual encoding} of the same k-mers
  uint32_t code[kTileWords + 1];
  uint32_t valid[kTileWords / 2 + 1];
};

struct MzArgs {
  const char* bases; const uint64_t* offsets; const uint64_t* tile_begin; // [n_seqs + 1] tiles before every sequence
  uint64_t n_bases, n_tiles;
  uint32_t n_seqs, k, w, m, r, frac, m_shift;
  unsigned long long* keys; unsigned long long cap; unsigned long long* counters; // [0] next tile, [1] keys appended
  unsigned char* hll;  // optional [n_seqs][2][4096]: HyperLogLog registers of every sequence's valid k-mers and of its minimizers (rho)
};

constexpr uint32_t kHllBits = 12, kHllRegs = 1u << kHllBits;

// hll::HyperLogLog::add (ref src/hyperloglog.hpp:103-110) with b = 12 on the low 32 bits of the k-mer's hash: register = top 12
// bits, rank = leading zeros of the rest (capped at 20) + 1.  The registers are bytes; a byte-wide maximum is a CAS on its word,
// tried only when the register would grow (after the first few thousand k-mers almost never).
__host__ __device__ __forceinline__ uint32_t hll_rank(uint32_t hash, uint32_t& index)
{
  index = hash >> (32 - kHllBits);
  const uint32_t x = hash << kHllBits;
#ifdef __CUDA_ARCH__
  const uint32_t lz = (uint32_t)__clz((int)x);
#else
  const uint32_t lz = x ? (uint32_t)__builtin_clz(x) : 32u; // (the reference's __builtin_clz(0) is undefined; 32 is what lzcnt gives)
#endif
  return (lz < 32 - kHllBits ? lz : 32 - kHllBits) + 1;
}

__device__ __forceinline__ void hll_add(unsigned char* regs, uint32_t hash)
{
  uint32_t index;
  const uint32_t rank = hll_rank(hash, index);
  if (__ldcg(regs + index) >= rank) return;
  uint32_t* word = reinterpret_cast<uint32_t*>(regs + (index & ~3u));
  const uint32_t sh = 8 * (index & 3u);
  uint32_t old = __ldcg(word);
  while (((old >> sh) & 0xFFu) < rank) {
    const uint32_t seen = atomicCAS(word, old, (old & ~(0xFFu << sh)) | rank << sh);
    if (seen == old) break;
    old = seen;
  }
}

__device__ __forceinline__ unsigned long long xur64(unsigned long long h)
{ // ref src/common.hpp:147-155
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}

__global__ void __launch_bounds__(kMzWarps * 32) minimizer_kernel(const uint4* __restrict__ g_lut, const MzArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nchunks = lut_chunks(a.k);
  uint4* lut = reinterpret_cast<uint4*>(smem_raw);
  MzWarpSmem* wsm = reinterpret_cast<MzWarpSmem*>(smem_raw + nchunks * 256 * sizeof(uint4));
  for (uint32_t i = threadIdx.x; i < nchunks * 256; i += blockDim.x) lut[i] = g_lut[i];
  __syncthreads();
  const bool wide = nchunks > 7;
  const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u, k = a.k, ldiff = a.w - a.k + 1, adv = kMzKmers + 1 - ldiff;
  MzWarpSmem& sm = wsm[threadIdx.x >> 5];
  for (;;) {
    unsigned long long tile = 0;
    if (lane == 0) tile = atomicAdd(a.counters, 1ull);
    tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
    if (tile >= a.n_tiles) break;
    // the sequence of this tile: last s with tile_begin[s] <= tile
    uint32_t lo = 0, hi = a.n_seqs;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.tile_begin[mid] <= tile) lo = mid; else hi = mid; }
    const uint64_t off = a.offsets[lo], len = a.offsets[lo + 1] - off;
    const uint64_t b0 = (tile - a.tile_begin[lo]) * adv;   // first base of the tile; k-mer i of the tile covers bases [b0 + i, b0 + i + k)
    // ---- bases -> 2-bit stream + validity (as match_common.cuh tile_lookups A0)
    const uint64_t rem = len - b0;
    const uint32_t nb = (uint32_t)min((uint64_t)(kMzKmers + k - 1), rem);
    const char* p0 = a.bases + off + b0;
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
    const uint32_t cn = __shfl_down_sync(0xFFFFFFFFu, cw, 1), vn = __shfl_down_sync(0xFFFFFFFFu, vw, 1);
    const uint32_t cwa = __funnelshift_l(cn, cw, 2 * sh);
    uint32_t vwa = (((vw << 16) | vn) << sh) >> 16;
    { const int keep = (int)nb - 16 * (int)lane; if (keep <= 0) vwa = 0; else if (keep < 16) vwa &= 0xFFFFu << (16 - keep); }
    const uint32_t vhi = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane) & 31), vlo = __shfl_sync(0xFFFFFFFFu, vwa, (2 * lane + 1) & 31);
    __syncwarp();
    if (lane <= kTileWords) sm.code[lane] = cwa;
    if (lane <= kTileWords / 2) sm.valid[lane] = (vhi << 16) | vlo;
    __syncwarp();
    // ---- k-mers -> hash, bucket id, residual encoding; vk[j] bit l: k-mer 32 j + l has k valid bases
    const uint32_t nk = rem >= k ? (uint32_t)min((uint64_t)kMzKmers, rem - k + 1) : 0u;
    uint32_t vk[kMzKmers / 32 + 1];
    vk[kMzKmers / 32] = 0;
#pragma unroll
    for (uint32_t j = 0; j < kMzKmers / 32; ++j) {
      const uint32_t p = lane + 32 * j;
      unsigned long long z = 0;
      uint2 rq = make_uint2(0u, 0u);
      bool ok = false;
      if (p < nk) {
        const uint32_t vj = p >> 5, vs = p & 31;
        const uint32_t vx = __funnelshift_l(sm.valid[vj + 1], sm.valid[vj], vs);
        ok = (vx >> (32 - k)) == (0xFFFFFFFFu >> (32 - k));
        const uint32_t cj = p >> 4, cs = 2 * (p & 15);
        const uint32_t w0 = sm.code[cj], w1 = sm.code[cj + 1], w2 = sm.code[cj + 2];
        const uint64_t x = ((uint64_t)__funnelshift_l(w1, w0, cs) << 32) | __funnelshift_l(w2, w1, cs);
        const uint64_t bp = x >> (64 - 2 * k);
        z = xur64(bp);
        const uint4 t = lut_pext(lut, (uint32_t)bp, (uint32_t)(bp >> 32), wide);
        rq = make_uint2(t.x, t.y);
      }
      vk[j] = __ballot_sync(0xFFFFFFFFu, ok);
      sm.z[p] = z; sm.rq[p] = rq;
      if (a.hll && ok) hll_add(a.hll + (size_t)lo * 2 * kHllRegs, (uint32_t)z); // c1: every valid k-mer (ref src/rqseq.cpp:107-108); k-mers two tiles share are added twice, which a maximum does not see
    }
    __syncwarp();
    // ---- window ends: the w-k+1 k-mers i .. i + ldiff - 1 of the tile must all be valid (a valid run of >= w bases)
    const uint32_t need = ldiff >= 32 ? 0xFFFFFFFFu : (1u << ldiff) - 1u;
#pragma unroll
    for (uint32_t j = 0; j < kMzKmers / 32; ++j) {
      const uint32_t i = lane + 32 * j; // first k-mer of the window
      bool emit = false;
      unsigned long long key = 0;
      if (i < adv && i + ldiff <= nk && (__funnelshift_r(vk[j], vk[j + 1], lane) & need) == need) {
        unsigned long long best = sm.z[i];
        uint32_t at = i;
        for (uint32_t d = 1; d < ldiff; ++d) { const unsigned long long z = sm.z[i + d]; if (z < best) { best = z; at = i + d; } } // equal hashes are equal k-mers
        const uint2 rq = sm.rq[at];
        if (a.hll) hll_add(a.hll + ((size_t)lo * 2 + 1) * kHllRegs, (uint32_t)best); // c2: every window's minimizer, before the residue test (ref :117)
        uint32_t quo, res;
        if (a.m_shift != 0xFFFFFFFFu) { quo = rq.x >> a.m_shift; res = rq.x & (a.m - 1); } else { quo = rq.x / a.m; res = rq.x - quo * a.m; }
        if (a.frac ? res <= a.r : res == a.r) { // ref src/rqseq.cpp:123-126
          const uint32_t row = a.frac ? quo * (a.r + 1) + res : quo;
          key = (unsigned long long)row << 32 | rq.y;
          emit = true;
        }
      }
      const uint32_t em = __ballot_sync(0xFFFFFFFFu, emit);
      if (em) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.counters + 1, (unsigned long long)__popc(em));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (emit) { const unsigned long long pos = base + __popc(em & lt_mask); if (pos < a.cap) a.keys[pos] = key; }
      }
    }
    __syncwarp();
  }
}

} // namespace krepp

using namespace krepp;

namespace {

#define MZ_CU(expr)                                                                                                      \
  do {                                                                                                                   \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess) { rc = set_error(KREPP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); goto done; } \
  } while (0)

inline uint64_t host_xur64(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33; return h; }
inline int nt4(unsigned char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }
inline uint64_t host_pext(uint64_t x, uint64_t mask) { uint64_t r = 0; int b = 0; for (int i = 0; i < 64; ++i) if ((mask >> i) & 1) { r |= ((x >> i) & 1) << b; ++b; } return r; }

// The end-of-sequence emit of ref src/rqseq.cpp:112-116 when the last valid run has k <= l < w bases: the ring holds the last
// w-k+1 valid k-mers of the sequence (pushed in order, never reset), zero-initialised slots where there were fewer.
bool end_quirk_key(const HostIndex& h, const char* s, uint64_t len, uint64_t* key, bool* kept, uint64_t* zmin)
{
  const uint32_t k = h.k, w = h.w, ldiff = w - k + 1;
  uint64_t l = 0;
  while (l < len && nt4((unsigned char)s[len - 1 - l]) < 4) ++l;  // valid run ending at the last base
  if (l < k || l >= w) return false;                               // no k-mer ends there, or an ordinary window (the kernel's)
  // the last ldiff valid k-mers of the sequence: walk the valid runs backwards; a run [a, b] of at least k bases has k-mers
  // ending at b, b - 1, ..., a + k - 1
  std::vector<uint64_t> ring; // bp words
  uint64_t e = len;           // bases [e, len) have been looked at
  while (e > 0 && ring.size() < ldiff) {
    if (nt4((unsigned char)s[e - 1]) >= 4) { --e; continue; }
    const uint64_t b = e - 1;
    uint64_t a = b;
    while (a > 0 && nt4((unsigned char)s[a - 1]) < 4 && b - a + 1 < (uint64_t)k - 1 + ldiff) --a; // no need to see more of a long run
    for (uint64_t end = b; end + 1 >= a + k && ring.size() < ldiff; --end) {
      uint64_t bp = 0;
      for (uint64_t i = end + 1 - k; i <= end; ++i) bp = (bp << 2) | (uint64_t)nt4((unsigned char)s[i]);
      ring.push_back(bp);
      if (end == 0) break;
    }
    e = a;
  }
  bool have = false;
  uint64_t best_z = 0, best_x = 0;
  if (ring.size() < ldiff) { have = true; best_z = 0; best_x = 0; } // a zero slot: hash 0 beats everything (or ties with poly-A, the same key)
  for (uint64_t x : ring) { const uint64_t z = host_xur64(x); if (!have || z < best_z) { have = true; best_z = z; best_x = x; } }
  *zmin = best_z; // the emit happens (and counts towards rho) whether or not the k-mer belongs to this library
  const uint32_t rix = (uint32_t)host_pext(best_x, h.mask_hash_bp), res = rix % h.m;
  *kept = h.frac ? res <= h.r : res == h.r;
  if (!*kept) return true;
  const uint32_t row = h.frac ? rix / h.m * (h.r + 1) + res : rix / h.m;
  // bp -> lr (ref src/common.hpp:188-197,223): low 32 = bit 0 of every base, high 32 = bit 1; position 0 = last base
  uint64_t lr = 0;
  for (uint32_t p = 0; p < k; ++p) { const uint64_t c = (best_x >> (2 * p)) & 3; lr |= (c & 1) << p | (c >> 1) << (32 + p); }
  *key = (uint64_t)row << 32 | (uint32_t)host_pext(lr, h.mask_drop_lr);
  return true;
}

} // namespace

namespace {

// Shared body of krepp_extract_mers / krepp_sketch_write: the sorted unique keys (row << 32 | encoding) of the sequences, and,
// when `est` is given, the two HyperLogLog estimates of RSeq::extract_mers summed over the sequences in order (n1: distinct
// valid k-mers, n2: distinct minimizers; ref src/rqseq.cpp:63-64,107-108,117,142-143), whose ratio is rho (src/rqseq.hpp:79).
int extract_impl(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, uint64_t* keys, uint64_t cap, uint64_t* n_keys,
                 std::vector<uint64_t>* keys_vec, double* est, unsigned long long** keep_dev = nullptr, MzScratch* scratch = nullptr)
{
  // device buffer `slot` of at least `bytes`: from the caller's scratch (grown when too small, kept between calls) or fresh
  auto buffer = [&](int slot, void** out, size_t bytes) -> cudaError_t {
    if (!scratch) return cudaMalloc(out, bytes);
    if (scratch->cap[slot] < bytes) {
      if (scratch->p[slot]) cudaFree(scratch->p[slot]);
      scratch->p[slot] = nullptr; scratch->cap[slot] = 0;
      const cudaError_t e = cudaMalloc(&scratch->p[slot], bytes + bytes / 4);
      if (e != cudaSuccess) return e;
      scratch->cap[slot] = bytes + bytes / 4;
    }
    *out = scratch->p[slot];
    return cudaSuccess;
  };
  if (ix->device == KREPP_DEVICE_NONE) return set_error(KREPP_ERR_CUDA, "the index-side kernels need a handle opened on a GPU (there is no CPU fallback)");
  const HostIndex& h = ix->host;
  if (h.w < h.k || h.w - h.k + 1 > 32) return set_error(KREPP_ERR_UNSUPPORTED, "window of %u with k = %u: at most 32 k-mers per window are supported", h.w, h.k);
  if (cudaSetDevice(ix->device) != cudaSuccess) return set_error(KREPP_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
  const uint32_t ldiff = h.w - h.k + 1, adv = kMzKmers + 1 - ldiff;
  int rc = KREPP_OK;
  const uint64_t nb = n_seqs ? offsets[n_seqs] - offsets[0] : 0;
  std::vector<uint64_t> rel(n_seqs + 1), tile_begin(n_seqs + 1, 0), quirk;
  std::vector<std::pair<uint32_t, uint64_t>> quirk_z; // (sequence, hash of the end-of-sequence emit)
  uint64_t windows = 0;
  for (uint32_t s = 0; s <= n_seqs; ++s) rel[s] = offsets[s] - offsets[0];
  for (uint32_t s = 0; s < n_seqs; ++s) {
    const uint64_t len = rel[s + 1] - rel[s];
    uint64_t tiles = 0;
    if (len >= h.w) { // ref src/rqseq.hpp:80-86: shorter sequences are skipped altogether
      const uint64_t nwin = len - h.w + 1;
      tiles = (nwin + adv - 1) / adv;
      windows += nwin;
      uint64_t key = 0, z = 0;
      bool kept = false;
      if (end_quirk_key(h, bases + offsets[s], len, &key, &kept, &z)) { quirk_z.emplace_back(s, z); if (kept) quirk.push_back(key); }
    }
    tile_begin[s + 1] = tile_begin[s] + tiles;
  }
  char* d_bases = nullptr; uint64_t *d_off = nullptr, *d_tb = nullptr;
  unsigned long long *d_keys = nullptr, *d_sorted = nullptr, *d_uniq = nullptr, *d_cnt = nullptr, *d_nsel = nullptr;
  unsigned char* d_hll = nullptr;
  void* d_tmp = nullptr;
  size_t tmp1 = 0, tmp2 = 0;
  unsigned long long h_cnt[2] = {0, 0}, nsel = 0;
  const uint64_t kcap = windows + quirk.size() + 1;
  MzArgs a{};
  {
    MZ_CU(buffer(0, (void**)&d_bases, nb + 64)); MZ_CU(buffer(1, (void**)&d_off, 8ull * (n_seqs + 1))); MZ_CU(buffer(2, (void**)&d_tb, 8ull * (n_seqs + 1)));
    MZ_CU(buffer(3, (void**)&d_keys, 8ull * kcap)); MZ_CU(buffer(4, (void**)&d_sorted, 8ull * kcap)); MZ_CU(buffer(5, (void**)&d_uniq, 8ull * kcap));
    MZ_CU(buffer(6, (void**)&d_cnt, 16)); MZ_CU(buffer(7, (void**)&d_nsel, 8));
    if (est) { MZ_CU(buffer(8, (void**)&d_hll, 2ull * kHllRegs * std::max<uint32_t>(n_seqs, 1))); MZ_CU(cudaMemset(d_hll, 0, 2ull * kHllRegs * std::max<uint32_t>(n_seqs, 1))); }
    MZ_CU(cudaMemcpy(d_bases, bases + offsets[0], nb, cudaMemcpyHostToDevice));
    MZ_CU(cudaMemcpy(d_off, rel.data(), 8ull * (n_seqs + 1), cudaMemcpyHostToDevice));
    MZ_CU(cudaMemcpy(d_tb, tile_begin.data(), 8ull * (n_seqs + 1), cudaMemcpyHostToDevice));
    MZ_CU(cudaMemset(d_cnt, 0, 16));
    a.bases = d_bases; a.offsets = d_off; a.tile_begin = d_tb; a.n_bases = nb; a.n_tiles = tile_begin[n_seqs]; a.n_seqs = n_seqs;
    a.k = h.k; a.w = h.w; a.m = h.m; a.r = h.r; a.frac = h.frac; a.m_shift = ix->dev.m_shift;
    a.keys = d_keys; a.cap = kcap; a.counters = d_cnt; a.hll = d_hll;
    if (a.n_tiles) {
      const size_t smem = lut_chunks(h.k) * 256 * sizeof(uint4) + kMzWarps * sizeof(MzWarpSmem);
      MZ_CU(cudaFuncSetAttribute(minimizer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      minimizer_kernel<<<ix->sms * 4, kMzWarps * 32, smem>>>(ix->dev.lut, a);
      MZ_CU(cudaGetLastError());
    }
    MZ_CU(cudaMemcpy(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost));
    uint64_t n = h_cnt[1];
    if (n > windows) { rc = set_error(KREPP_ERR_CUDA, "minimizer kernel emitted more keys than windows"); goto done; }
    if (!quirk.empty()) { MZ_CU(cudaMemcpy(d_keys + n, quirk.data(), 8ull * quirk.size(), cudaMemcpyHostToDevice)); n += quirk.size(); }
    if (n) {
      uint32_t row_bits = 1;
      while ((1ull << row_bits) < h.nrows) ++row_bits;
      MZ_CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp1, d_keys, d_sorted, (int)n, 0, 32 + (int)row_bits));
      MZ_CU(cub::DeviceSelect::Unique(nullptr, tmp2, d_sorted, d_uniq, d_nsel, (int)n));
      MZ_CU(buffer(9, &d_tmp, std::max(tmp1, tmp2)));
      if (n > 0x7FFFFFFFull) { rc = set_error(KREPP_ERR_CAPACITY, "more than 2^31 minimizers in one call: pass fewer sequences"); goto done; }
      MZ_CU(cub::DeviceRadixSort::SortKeys(d_tmp, tmp1, d_keys, d_sorted, (int)n, 0, 32 + (int)row_bits));
      MZ_CU(cub::DeviceSelect::Unique(d_tmp, tmp2, d_sorted, d_uniq, d_nsel, (int)n));
      MZ_CU(cudaMemcpy(&nsel, d_nsel, 8, cudaMemcpyDeviceToHost));
      nsel &= 0xFFFFFFFFull; // DeviceSelect writes an int
    }
    if (est) { // hll::HyperLogLog::estimate per sequence (ref src/hyperloglog.hpp:117-140), summed in sequence order
      std::vector<unsigned char> regs(2ull * kHllRegs * std::max<uint32_t>(n_seqs, 1));
      MZ_CU(cudaMemcpy(regs.data(), d_hll, regs.size(), cudaMemcpyDeviceToHost));
      for (const auto& qz : quirk_z) { // the end-of-sequence emit is the host's: its minimizer joins c2
        uint32_t index;
        const uint32_t rank = hll_rank((uint32_t)qz.second, index);
        unsigned char& reg = regs[((size_t)qz.first * 2 + 1) * kHllRegs + index];
        if (rank > reg) reg = (unsigned char)rank;
      }
      const double mm = (double)kHllRegs, alpha_mm = (0.7213 / (1.0 + 1.079 / mm)) * mm * mm;
      est[0] = est[1] = 0.0;
      for (uint32_t s = 0; s < n_seqs; ++s) {
        if (rel[s + 1] - rel[s] < h.w) continue; // no extract_mers call for it
        for (int c = 0; c < 2; ++c) {
          const unsigned char* M = regs.data() + ((size_t)s * 2 + c) * kHllRegs;
          double sum = 0.0;
          for (uint32_t i = 0; i < kHllRegs; ++i) sum += 1.0 / (double)(1 << M[i]);
          double e = alpha_mm / sum;
          if (e <= 2.5 * mm) {
            uint32_t zeros = 0;
            for (uint32_t i = 0; i < kHllRegs; ++i) zeros += M[i] == 0;
            if (zeros) e = mm * std::log(mm / (double)zeros);
          } else if (e > (1.0 / 30.0) * 4294967296.0) e = -4294967296.0 * std::log(1.0 - e / 4294967296.0);
          est[c] += e;
        }
      }
    }
    *n_keys = nsel;
    if (keep_dev) { // the table stays on the device for the library builder (builder.cu), in an allocation of its own size
      *keep_dev = nullptr;
      if (nsel) { MZ_CU(cudaMalloc(keep_dev, 8ull * nsel)); MZ_CU(cudaMemcpy(*keep_dev, d_uniq, 8ull * nsel, cudaMemcpyDeviceToDevice)); }
      goto done;
    }
    if (keys_vec) { keys_vec->resize(nsel); if (nsel) MZ_CU(cudaMemcpy(keys_vec->data(), d_uniq, 8ull * nsel, cudaMemcpyDeviceToHost)); goto done; }
    if (nsel > cap) { rc = cap ? set_error(KREPP_ERR_CAPACITY, "krepp_extract_mers: %llu keys but room for %llu", (unsigned long long)nsel, (unsigned long long)cap) : KREPP_OK; goto done; }
    if (nsel) MZ_CU(cudaMemcpy(keys, d_uniq, 8ull * nsel, cudaMemcpyDeviceToHost));
  }
done:
  if (!scratch) for (void* p : {(void*)d_bases, (void*)d_off, (void*)d_tb, (void*)d_keys, (void*)d_sorted, (void*)d_uniq, (void*)d_cnt, (void*)d_nsel, (void*)d_hll, d_tmp}) if (p) cudaFree(p);
  return rc;
}

} // namespace

namespace krepp {
int extract_to_device(const krepp_index* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, unsigned long long** d_keys, uint64_t* n_keys, double est[2],
                      MzScratch* scratch)
{
  return extract_impl(ix, bases, offsets, n_seqs, nullptr, 0, n_keys, nullptr, est, d_keys, scratch);
}
void MzScratch::release()
{
  for (int i = 0; i < kSlots; ++i) { if (p[i]) cudaFree(p[i]); p[i] = nullptr; cap[i] = 0; }
}
} // namespace krepp

extern "C" int krepp_extract_mers(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, uint64_t* keys, uint64_t cap,
                                  uint64_t* n_keys)
{
  if (!ix || !bases || !offsets || !n_keys || (cap && !keys)) return set_error(KREPP_ERR_ARG, "krepp_extract_mers: null argument");
  return extract_impl(ix, bases, offsets, n_seqs, keys, cap, n_keys, nullptr, nullptr);
}

extern "C" int krepp_sequence_rho(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, double* n_kmers_est, double* n_minimizers_est)
{
  if (!ix || !bases || !offsets || !n_kmers_est || !n_minimizers_est) return set_error(KREPP_ERR_ARG, "krepp_sequence_rho: null argument");
  double est[2] = {0, 0};
  uint64_t n = 0;
  std::vector<uint64_t> keys;
  if (int rc = extract_impl(ix, bases, offsets, n_seqs, nullptr, 0, &n, &keys, est)) return rc;
  *n_kmers_est = est[0]; *n_minimizers_est = est[1];
  return KREPP_OK;
}

extern "C" int krepp_sketch_write(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, const char* out_path, uint64_t* n_kmers,
                                  double* rho_out)
{
  // SketchSingle::create_sketch / save_sketch (ref src/krepp.cpp:110-128): SDynHT::fill_table (extract_mers of every sequence,
  // sort and unique per row; src/table.cpp:236-246), SFlatHT (src/table.cpp:3-21) and its save (:35-41), save_configuration
  // (src/krepp.cpp:18-29), then rho = n2 / n1
  if (!ix || !bases || !offsets || !out_path) return set_error(KREPP_ERR_ARG, "krepp_sketch_write: null argument");
  const HostIndex& h = ix->host;
  double est[2] = {0, 0};
  uint64_t n = 0;
  std::vector<uint64_t> keys;
  if (int rc = extract_impl(ix, bases, offsets, n_seqs, nullptr, 0, &n, &keys, est)) return rc;
  const double rho = est[1] / est[0];
  std::vector<uint32_t> enc(n);
  std::vector<uint64_t> inc(h.nrows, 0);
  for (uint64_t i = 0; i < n; ++i) {
    enc[i] = (uint32_t)keys[i];
    const uint64_t row = keys[i] >> 32;
    if (row >= h.nrows) return set_error(KREPP_ERR_CUDA, "a minimizer fell outside the table (row %llu of %u)", (unsigned long long)row, h.nrows);
    ++inc[row];
  }
  for (uint32_t r = 1; r < h.nrows; ++r) inc[r] += inc[r - 1];
  FILE* f = fopen(out_path, "wb");
  if (!f) return set_error(KREPP_ERR_IO, "Failed to write the sketch!");
  const uint8_t k8 = (uint8_t)h.k, w8 = (uint8_t)h.w, h8 = (uint8_t)h.h, frac8 = h.frac ? 1 : 0;
  bool ok = fwrite(&n, 8, 1, f) == 1 && (!n || fwrite(enc.data(), 4, n, f) == n) && fwrite(&h.nrows, 4, 1, f) == 1 && (!h.nrows || fwrite(inc.data(), 8, h.nrows, f) == h.nrows);
  ok = ok && fwrite(&k8, 1, 1, f) == 1 && fwrite(&w8, 1, 1, f) == 1 && fwrite(&h8, 1, 1, f) == 1 && fwrite(&h.m, 4, 1, f) == 1 && fwrite(&h.r, 4, 1, f) == 1 &&
       fwrite(&frac8, 1, 1, f) == 1 && fwrite(&h.nrows, 4, 1, f) == 1 && fwrite(h.ppos.data(), 1, h.ppos.size(), f) == h.ppos.size() &&
       fwrite(h.npos.data(), 1, h.npos.size(), f) == h.npos.size() && fwrite(&rho, 8, 1, f) == 1;
  ok = (fclose(f) == 0) && ok;
  if (!ok) return set_error(KREPP_ERR_IO, "Failed to write the sketch!");
  if (n_kmers) *n_kmers = n;
  if (rho_out) *rho_out = rho;
  return KREPP_OK;
}

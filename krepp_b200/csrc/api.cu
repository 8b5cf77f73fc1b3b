// api.cu -- the C ABI declared in include/krepp_b200.h: index image upload (whole, or one bucket-range shard), batch slots
// (one CUDA stream each), submit / wait with selectable result rows, the three phases of a batch on a sharded index,
// parity taps.  No CPU fallback exists: every compute entry point needs a CUDA device.
#include "../../include/krepp_b200.h"

#include "device.cuh"
#include "handles.hpp"
#include "index_image.hpp"
#include "fixed5.h"
#include "solve.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <sys/stat.h>
#include <vector>

using namespace krepp;

namespace {

thread_local std::string g_err;
thread_local std::string g_name;

int fail(int code, const char* fmt, ...)
{
  char b[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof b, fmt, ap);
  va_end(ap);
  g_err = b;
  return code;
}

#define CU(expr)                                                                                                 \
  do {                                                                                                           \
    cudaError_t e__ = (expr);                                                                                    \
    if (e__ != cudaSuccess) return fail(KREPP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));      \
  } while (0)

template <class T, class A>
cudaError_t upload(const std::vector<T, A>& v, const T** out, std::vector<void*>& allocs, uint64_t& bytes, size_t pad = 0)
{
  void* p = nullptr;
  const size_t n = (v.size() + pad) * sizeof(T);
  cudaError_t e = cudaMalloc(&p, n ? n : sizeof(T));
  if (e != cudaSuccess) return e;
  allocs.push_back(p);
  bytes += n;
  if (pad) { e = cudaMemset(p, 0, n); if (e != cudaSuccess) return e; }
  if (!v.empty()) e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  *out = static_cast<const T*>(p);
  return e;
}

// Byte-indexed tables for the two software pexts (see match.cu lut_pext).  Position p (0 = last base of the k-mer) sits at
// bits 2p, 2p+1 of the k-mer word.  Forward strand: a base at a hash position contributes its code to rix at the rank of
// that position among ppos (ascending, pext order; ref src/lshf.cpp:47-50,62); a base at a kept position contributes
// bit0 / bit1 of its code to bits r and 16+r of q, r its rank among npos (ref src/lshf.cpp:39-46,64-69).  Reverse
// strand: the same forward base, complemented, is the base at position k-1-p of the reverse complement
// (ref src/common.hpp:177-186).
static std::vector<uint4> build_lut(const HostIndex& h)
{
  std::vector<int> hrank(32, -1), nrank(32, -1);
  { std::vector<uint8_t> pp = h.ppos, np = h.npos; std::sort(pp.begin(), pp.end()); std::sort(np.begin(), np.end());
    for (size_t i = 0; i < pp.size(); ++i) hrank[pp[i]] = (int)i;
    for (size_t i = 0; i < np.size(); ++i) nrank[np[i]] = (int)i; }
  const uint32_t nch = std::max<uint32_t>(7, (2 * h.k + 7) / 8); // bytes of the k-mer word (match.cu lut_chunks: never fewer than seven tables)
  std::vector<uint4> lut(nch * 256, make_uint4(0, 0, 0, 0));
  for (uint32_t strand = 0; strand < 2; ++strand)
    for (uint32_t c = 0; c < nch; ++c)
      for (uint32_t v = 0; v < 256; ++v) {
        uint32_t rix = 0, q = 0;
        for (uint32_t s = 0; s < 4; ++s) {
          const uint32_t p = 4 * c + s;
          if (p >= h.k) continue;
          uint32_t code = (v >> (2 * s)) & 3, pos = p;
          if (strand) { code = 3 - code; pos = h.k - 1 - p; }
          if (hrank[pos] >= 0) rix |= code << (2 * hrank[pos]);
          if (nrank[pos] >= 0) q |= (code & 1) << nrank[pos] | (code >> 1) << (16 + nrank[pos]);
        }
        uint4& t = lut[c * 256 + v];
        if (strand) { t.z = rix; t.w = q; } else { t.x = rix; t.y = q; }
      }
  return lut;
}

// AoS assembly of the public result structs on the device (one D2H copy each, no host-side gather).
__global__ void __launch_bounds__(128) finalize_kernel(const SolveArgs a, krepp_record_t* out_rec, krepp_read_summary_t* out_read,
                                                        const uint32_t* wn, const uint32_t* place_begin, const uint32_t* place_count,
                                                        krepp_brief_t* out_brief, double chisq_value)
{
  if (a.counters[2] & kErrRedo) return; // incomplete records: the host re-runs the batch
  const uint32_t n = a.counters[0] < a.n_records ? a.counters[0] : a.n_records;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (out_brief)
    for (uint32_t i = tid; i < n; i += nth) {
      const uint32_t slot = a.rec_slot[i];
      krepp_brief_t r;
      r.read = a.rec_read[i];
      r.ref = (slot & 0x07FFFFFFu) | (slot >> 31) << 27 | (a.rec_flags[i] & 7u) << 28 | (a.rec_chisq[i] < chisq_value ? 1u : 0u) << 31;
      r.d_llh = a.rec_d[i];
      out_brief[i] = r;
    }
  for (uint32_t i = tid; out_rec && i < n; i += nth) {
    krepp_record_t r;
    const uint32_t slot = a.rec_slot[i];
    r.read = a.rec_read[i]; r.leaf_se = slot & 0x7FFFFFFFu; r.strand = slot >> 31;
    r.match_count = a.rec_match[i]; r.hdist_min = a.rec_hdmin[i]; r.flags = a.rec_flags[i];
    r.rho = a.rho[r.leaf_se]; r.d_llh = a.rec_d[i]; r.v_llh = a.rec_v[i]; r.chisq = a.rec_chisq[i];
    out_rec[i] = r;
  }
  for (uint32_t i = tid; out_read && i < a.n_reads; i += nth) {
    krepp_read_summary_t s;
    s.onmers = a.onmers[i]; s.wn[0] = wn[2 * i]; s.wn[1] = wn[2 * i + 1];
    s.hdist_filt[0] = a.hdfilt[2 * i]; s.hdist_filt[1] = a.hdfilt[2 * i + 1];
    s.rec_begin = a.rec_begin[i]; s.rec_count = a.rec_count[i];
    s.place_begin = place_begin ? place_begin[i] : 0; s.place_count = place_count ? place_count[i] : 0;
    s.closest = a.closest[i];
    s.n_selected = a.nsel ? a.nsel[i] : 0u;
    out_read[i] = s;
  }
}


// ---- KREPP_OUT_DIST: the rows `krepp dist` prints, chosen on the device (IBatch::report_distances, ref src/query.cpp:158-196)
struct DistOut {
  uint32_t n_reads;
  const uint32_t *rec_begin, *rec_count, *rec_slot, *rec_flags;
  const double *rec_d, *rec_chisq;
  const int32_t* closest;
  const uint32_t* leaf_rank;   // by se
  const uint32_t* counters;
  int summarize, multi, no_filter, has_max;
  double dist_max, chisq_value;
  uint32_t* cnt;               // [n_reads] printed rows per read
  const uint32_t* begin;       // [n_reads + 1] exclusive prefix of cnt
  uint32_t* out_begin;         // [n_reads + 1] begin | NA << 31
  void* rows;
  uint32_t row_bytes;
};

// Calls emit(record index) for every row of read r in print order (references by ascending se: the records are forward leaves
// by ascending se, then reverse leaves by ascending se, and at most one strand of a leaf is selected); returns true when the
// read prints "NA\tNaN" instead (ref src/query.cpp:173-176).
template <class F>
__device__ __forceinline__ bool dist_rows_of(const DistOut& a, uint32_t r, F&& emit)
{
  const uint32_t b = a.rec_begin[r], n = a.rec_count[r];
  const int32_t cl = a.closest[r];
  if (!a.summarize) {
    if (cl < 0 || (a.has_max && a.rec_d[cl] > a.dist_max)) return true;
    if (!a.multi) { emit((uint32_t)cl); return false; }       // ref :193-195
  }
  uint32_t nf = 0;
  while (nf < n && !(a.rec_slot[b + nf] >> 31)) ++nf;
  uint32_t i = b, j = b + nf;
  const uint32_t ie = b + nf, je = b + n;
  while (i < ie || j < je) {
    const uint32_t si = i < ie ? (a.rec_slot[i] & 0x7FFFFFFFu) : 0xFFFFFFFFu, sj = j < je ? (a.rec_slot[j] & 0x7FFFFFFFu) : 0xFFFFFFFFu;
    uint32_t pick;
    if (si < sj) pick = i++;
    else if (sj < si) pick = j++;
    else { pick = (a.rec_flags[i] & KREPP_REC_SELECTED) ? i : j; ++i; ++j; }
    if (!(a.rec_flags[pick] & KREPP_REC_SELECTED)) continue;
    const bool dmax_ok = !a.has_max || a.rec_d[pick] < a.dist_max;
    bool keep = dmax_ok;
    if (a.summarize || !a.no_filter) keep = keep && (a.rec_chisq[pick] < a.chisq_value); // ref :164-166,186-188 (NaN never passes)
    if (keep) emit(pick);
  }
  return false;
}

__global__ void __launch_bounds__(128) dist_count_kernel(const DistOut a)
{
  if (a.counters[2] & kErrRedo) return;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
    uint32_t c = 0;
    dist_rows_of(a, r, [&](uint32_t) { ++c; });
    a.cnt[r] = c;
  }
}

__global__ void __launch_bounds__(128) dist_emit_kernel(const DistOut a)
{
  if (a.counters[2] & kErrRedo) return;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
    uint32_t at = a.begin[r];
    const uint32_t first = at;
    const bool na = dist_rows_of(a, r, [&](uint32_t i) {
      const uint32_t se = a.rec_slot[i] & 0x7FFFFFFFu, units = fixed5_units(a.rec_d[i]);
      if (a.row_bytes == 4) static_cast<uint32_t*>(a.rows)[at] = a.leaf_rank[se] << 16 | (units < 0xFFFFu ? units : 0xFFFFu);
      else static_cast<uint2*>(a.rows)[at] = make_uint2(se, units);
      ++at;
    });
    a.out_begin[r] = first | (na ? 0x80000000u : 0u);
    if (r == a.n_reads - 1) a.out_begin[a.n_reads] = at;
  }
}

} // namespace

namespace krepp {
int set_error(int code, const char* fmt, ...)
{
  char b[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(b, sizeof b, fmt, ap);
  va_end(ap);
  g_err = b;
  return code;
}
} // namespace krepp

struct krepp_batch {
  krepp_index* ix = nullptr;
  int device = 0;                       // the index's device (kept here: krepp_batch_destroy must not touch the index handle)
  krepp_params_t p{};
  LlhTables tab{};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evm0 = nullptr, evm1 = nullptr;
  uint32_t max_reads = 0, rec_cap = 0, n_reads = 0, launches = 0;
  uint64_t max_bases = 0, n_bases = 0;
  bool submitted = false, pending = false, device_input = false, keep_all = false; // pending: submitted and not yet waited for
  uint32_t out_rows = KREPP_OUT_ALL;    // which row arrays krepp_batch_wait copies to the host
  const char* in_bases = nullptr;       // device pointers used by the last submit
  const uint64_t* in_offsets = nullptr;
  // pinned host staging + device inputs
  char* h_bases = nullptr; uint64_t* h_offsets = nullptr;
  char* d_bases = nullptr; uint64_t* d_offsets = nullptr;
  // device state
  uint32_t *d_onmers = nullptr, *d_wn = nullptr, *d_hdfilt = nullptr, *d_rec_begin = nullptr, *d_rec_count = nullptr;
  int32_t* d_closest = nullptr; uint32_t* d_nsel = nullptr;
  uint32_t *d_rec_read = nullptr, *d_rec_slot = nullptr, *d_rec_hist = nullptr, *d_rec_flags = nullptr, *d_rec_match = nullptr, *d_rec_hdmin = nullptr, *d_rec_work = nullptr;
  double *d_rec_d = nullptr, *d_rec_v = nullptr, *d_rec_chisq = nullptr;
  uint32_t* d_rec_alias = nullptr; unsigned long long* d_memo_key = nullptr; uint32_t* d_memo_owner = nullptr; uint32_t memo_mask = 0, memo_bits = 0;
  uint32_t* d_counters = nullptr; unsigned long long* d_stats = nullptr;
  uint32_t *d_acc = nullptr, *d_bitmap = nullptr, *d_marker = nullptr, *d_stack = nullptr, *d_tagctr = nullptr;
  uint32_t stack_cap = 0;
  krepp_record_t* d_out_rec = nullptr; krepp_read_summary_t* d_out_read = nullptr;
  krepp_brief_t *d_out_brief = nullptr, *h_brief = nullptr;
  // KREPP_OUT_DIST: printed rows per read, their exclusive prefix, the rows (4 or 8 bytes each; at most one per record)
  double *d_seek = nullptr, *h_seek = nullptr; // KREPP_OUT_SEEK: one distance per read (sketch handles)
  uint32_t *d_dist_cnt = nullptr, *d_dist_begin = nullptr, *d_dist_out_begin = nullptr, *d_dist_partials = nullptr, *h_dist_begin = nullptr;
  void *d_dist_rows = nullptr, *h_dist_rows = nullptr;
  uint32_t dist_row_bytes = 4;
  bool summaries_copied = false;
  // pinned host results
  krepp_record_t* h_rec = nullptr; krepp_read_summary_t* h_read = nullptr; uint32_t* h_hist = nullptr;
  uint32_t* h_counters = nullptr; unsigned long long* h_stats = nullptr;
  // placement (K5)
  uint32_t place_cap = 0, place_warps = 0;
  uint32_t *d_place_begin = nullptr, *d_place_count = nullptr, *d_node_bitmap = nullptr, *d_node_list = nullptr;
  uint32_t* d_sel = nullptr; double* d_chain = nullptr; uint32_t chain_cap = 0; uint32_t* d_node_order = nullptr;
  uint32_t node_cap = 0;      // tree nodes touched by a batch (place_collect_kernel's entries)
  uint32_t *d_pn_read = nullptr, *d_pn_se = nullptr, *d_pn_flags = nullptr, *d_pn_work = nullptr, *d_pn_begin = nullptr, *d_pn_count = nullptr;
  double *d_pn_mc = nullptr, *d_pn_uc = nullptr, *d_pn_rho = nullptr, *d_pn_d = nullptr, *d_pn_v = nullptr, *d_pn_chisq = nullptr;
  krepp_placement_t *d_place = nullptr, *h_place = nullptr;
  // long reads cut into segments (SegArgs, device.cuh): decided per batch by krepp_batch_submit from the host offsets
  uint32_t seg_windows = 0;   // windows per segment (0: never cut); a read is cut when it has more than twice as many
  uint32_t cap_vreads = 0;    // most segments a batch can have: max_reads + max_bases / seg_windows + 1
  uint32_t seg_nv = 0;        // segments of the pending batch, 0 when none of its reads was cut
  uint64_t* h_voff = nullptr; uint32_t* h_vbegin = nullptr;   // pinned: (begin, end) of every segment; first segment of every read
  uint64_t* d_voff = nullptr; uint32_t* d_vbegin = nullptr;
  uint32_t *v_onmers = nullptr, *v_wn = nullptr, *v_hdfilt = nullptr, *v_rec_begin = nullptr, *v_rec_count = nullptr; // per segment
  uint32_t *v_rec_read = nullptr, *v_rec_slot = nullptr, *v_rec_hist = nullptr; // the segments' records (rec_cap rows)
  uint32_t *d_vcounters = nullptr, *d_seg_scratch = nullptr, *d_seg_claim = nullptr;
  int seg_ctas = 0;
  // bucket-sorted pipeline (sorted.cu)
  bool sorted = false, fused_once = false;
  bool bins_off = false;      // a batch overflowed a coarse bin of the two-level lookup sort: this slot keeps to the two-pass sort
  SortArgs so{};
  uint32_t* h_sc = nullptr;   // [0..7] copy of so.sc, [8] lookups of the batch, [9] hit entries handed to finish, [16..] mode B row/hit boundaries
  const void* shard_hits = nullptr; uint64_t shard_n_hits = 0; // mode B: the batch is in its finish phase (krepp_shard_finish)
  StageClock clk;             // per-stage events of the last enqueue
  // tap
  uint4* d_tap = nullptr; unsigned long long* d_tap_count = nullptr; unsigned long long tap_cap = 0;
};

extern "C" {

const char* krepp_last_error(void) { return g_err.c_str(); }
int krepp_abi_version(void) { return KREPP_ABI_VERSION; }

void krepp_params_default(krepp_params_t* p, int place)
{
  if (!p) return;
  p->hdist_th = 4; p->chisq = 2.706; p->dist_max = std::numeric_limits<double>::quiet_NaN(); p->tau = 2;
  p->no_filter = place ? 0 : 1; p->multi = 1; p->summarize = 0; p->place = place ? 1 : 0;
}

int krepp_index_open(const char* index_dir, int device, krepp_index_t** out) { return krepp_index_open_tree(index_dir, device, 0, 1, nullptr, out); }
int krepp_index_open_shard(const char* index_dir, int device, uint32_t shard, uint32_t nshards, krepp_index_t** out) { return krepp_index_open_tree(index_dir, device, shard, nshards, nullptr, out); }

static int open_index(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* nwk_path, bool lineages, krepp_index_t** out);
int krepp_index_open_tree(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* nwk_path, krepp_index_t** out) { return open_index(index_dir, device, shard, nshards, nwk_path, false, out); }
int krepp_index_open_lineages(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* lineage_path, krepp_index_t** out)
{
  if (!lineage_path) return fail(KREPP_ERR_ARG, "krepp_index_open_lineages: null lineage file");
  return open_index(index_dir, device, shard, nshards, lineage_path, true, out);
}

static int geometry_open(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, int64_t seed, const uint8_t* given_ppos, int device, krepp_index_t** out);

int krepp_geometry_open(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, int64_t seed, int device, krepp_index_t** out)
{
  return geometry_open(k, w, h, m, r, frac, seed, nullptr, device, out);
}

int krepp_geometry_open_positions(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, const uint8_t* ppos, int device, krepp_index_t** out)
{
  if (!ppos) return fail(KREPP_ERR_ARG, "krepp_geometry_open_positions: null argument");
  return geometry_open(k, w, h, m, r, frac, -1, ppos, device, out);
}

static int geometry_open(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, int64_t seed, const uint8_t* given_ppos, int device, krepp_index_t** out)
{
  if (!out) return fail(KREPP_ERR_ARG, "krepp_geometry_open: null argument");
  *out = nullptr;
  auto* ix = new krepp_index;
  { // the configuration is judged before any device is looked for, as the reference judges it before reading anything
    std::vector<uint8_t> ppos, npos;
    if (given_ppos && k >= 19 && k <= 31 && h >= 3 && h <= 15 && h < k) { // the caller's hash positions: distinct, below k; the kept positions are the others, ascending
      std::vector<uint8_t> taken(k, 0);
      for (uint32_t i = 0; i < h; ++i) {
        if (given_ppos[i] >= k || taken[given_ppos[i]]) { delete ix; return fail(KREPP_ERR_ARG, "krepp_geometry_open_positions: the %u hash positions must be distinct and below k", h); }
        taken[given_ppos[i]] = 1;
      }
      for (uint32_t p = k; p-- > 0;) if (taken[p]) ppos.push_back((uint8_t)p);  // descending (ref src/lshf.cpp:143)
      for (uint32_t p = 0; p < k; ++p) if (!taken[p]) npos.push_back((uint8_t)p); // ascending (ref src/lshf.cpp:144)
    } else
    if (k >= 19 && k <= 31 && h >= 3 && h <= 15 && h < k) lsh_positions(k, h, seed >= 0, (uint32_t)seed, ppos, npos);
    std::string err = ix->host.set_geometry(k, w, h, m, r, frac != 0, ppos, npos);
    if (!err.empty()) { delete ix; return fail(KREPP_ERR_ARG, "%s", err.c_str()); }
  }
  int ndev = 0;
  if (device != KREPP_DEVICE_NONE) {
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { delete ix; return fail(KREPP_ERR_CUDA, "no CUDA device is available (the krepp_b200 kernels have no CPU fallback)"); }
    if (device < 0 || device >= ndev) { delete ix; return fail(KREPP_ERR_ARG, "device %d out of range (%d devices)", device, ndev); }
  }
  ix->device = device;
  if (device == KREPP_DEVICE_NONE) { *out = ix; return KREPP_OK; }
  const HostIndex& hh = ix->host;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ix->sms, cudaDevAttrMultiProcessorCount, device);
  DevIndex& d = ix->dev;
  if (e == cudaSuccess) e = upload(build_lut(hh), &d.lut, ix->allocs, ix->device_bytes);
  if (e != cudaSuccess) { for (void* p : ix->allocs) cudaFree(p); delete ix; return fail(KREPP_ERR_CUDA, "uploading the hash tables failed: %s", cudaGetErrorString(e)); }
  d.k = hh.k; d.h = hh.h; d.m = hh.m; d.nrows = hh.nrows;
  d.m_shift = (hh.m & (hh.m - 1)) == 0 ? (uint32_t)__builtin_ctz(hh.m) : 0xFFFFFFFFu;
  *out = ix;
  return KREPP_OK;
}

int krepp_sketch_open(const char* sketch_path, int device, krepp_index_t** out)
{
  if (!sketch_path || !out) return fail(KREPP_ERR_ARG, "krepp_sketch_open: null argument");
  struct stat st;
  if (stat(sketch_path, &st) != 0 || !S_ISREG(st.st_mode)) return fail(KREPP_ERR_IO, "Failed to open %s", sketch_path);
  return open_index(sketch_path, device, 0, 1, nullptr, false, out);
}

static int open_index(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* nwk_path, bool lineages, krepp_index_t** out)
{
  if (!index_dir || !out) return fail(KREPP_ERR_ARG, "krepp_index_open: null argument");
  *out = nullptr;
  if (!nshards || shard >= nshards || nshards > KREPP_MAX_SHARDS) return fail(KREPP_ERR_ARG, "krepp_index_open_shard: shard %u of %u is not valid (at most %d shards)", shard, nshards, KREPP_MAX_SHARDS);
  int ndev = 0;
  if (device != KREPP_DEVICE_NONE) {
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KREPP_ERR_CUDA, "no CUDA device is available (the krepp_b200 query path has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(KREPP_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  }
  auto* ix = new krepp_index;
  std::string err = ix->host.load(index_dir, shard, nshards, true, nwk_path ? nwk_path : "", lineages);
  if (!err.empty()) { delete ix; return fail(KREPP_ERR_IO, "%s", err.c_str()); }
  const HostIndex& h = ix->host;
  if (device == KREPP_DEVICE_NONE) { ix->device = device; *out = ix; return KREPP_OK; } // metadata / tree only, no queries
  if (h.m > (uint32_t)kMaxResidues) { const uint32_t m = h.m; delete ix; return fail(KREPP_ERR_UNSUPPORTED, "m = %u exceeds the %d residues supported on the device", m, kMaxResidues); }
  if (h.inc32.size() != (size_t)(h.row1 - h.row0)) {
    delete ix;
    return fail(KREPP_ERR_UNSUPPORTED, "2^32 or more k-mers on one device are not supported: open the index as bucket-range shards (krepp_index_open_shard)");
  }
  ix->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ix->sms, cudaDevAttrMultiProcessorCount, device);
  DevIndex& d = ix->dev;
  std::vector<uint2> tmp2;
  // the image: cmer and pse are uploaded verbatim (8-byte pairs), inc padded by one entry
  if (e == cudaSuccess) e = upload(h.cmer, reinterpret_cast<const uint64_t**>(&d.cmer), ix->allocs, ix->device_bytes, 4);
  if (e == cudaSuccess) e = upload(h.inc32, &d.inc32, ix->allocs, ix->device_bytes, 1);
  if (e == cudaSuccess) e = upload(h.pse, reinterpret_cast<const uint64_t**>(&d.pse), ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.kind, &d.kind, ix->allocs, ix->device_bytes);
  {
    std::vector<uint2> cnode(h.kind.size(), make_uint2(0u, 0u));
    for (size_t se = 0; se < cnode.size(); ++se) {
      if (h.kind[se] == 1) cnode[se] = make_uint2(0x80000000u | h.col_rank[se], 0u);
      else if (h.kind[se] == 2) cnode[se] = make_uint2(0x40000000u | (uint32_t)h.pse[se], (uint32_t)(h.pse[se] >> 32));
    }
    if (h.nsubsets >= (1u << 30)) e = cudaErrorInvalidValue; // colour ids must leave the two flag bits free
    if (e == cudaSuccess) e = upload(cnode, &d.cnode, ix->allocs, ix->device_bytes);
  }
  if (e == cudaSuccess) e = upload(h.rho, &d.rho, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.leaf_rank, &d.leaf_rank, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.leaf_se, &d.leaf_se, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.parent, &d.parent, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.nchildren, &d.nchildren, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.eff_nchildren, &d.eff_nchildren, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.blen, &d.blen, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(build_lut(h), &d.lut, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.subtree, &d.subtree, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.depth, &d.depth, ix->allocs, ix->device_bytes);
  if (e == cudaSuccess) e = upload(h.tree.logw, &d.logw, ix->allocs, ix->device_bytes);
  d.cbeg = nullptr; d.cleaf = nullptr;
  if (!h.cbeg.empty()) { // flattened colours: what the bucket-sorted pipeline (sorted.cu) expands hits with
    if (e == cudaSuccess) e = upload(h.cbeg, &d.cbeg, ix->allocs, ix->device_bytes);
    if (e == cudaSuccess) e = upload(h.cleaf, &d.cleaf, ix->allocs, ix->device_bytes, 1);
    ix->sorted_ok = true;
  }
  if (e != cudaSuccess) {
    for (void* p : ix->allocs) cudaFree(p);
    delete ix;
    return fail(KREPP_ERR_CUDA, "uploading the index image failed: %s", cudaGetErrorString(e));
  }
  d.nkmers = h.nkmers; d.nrows = h.nrows; d.row0 = h.row0; d.nrows_local = h.row1 - h.row0; d.nsubsets = h.nsubsets; d.nnodes = h.tree.nnodes; d.nleaves = h.tree.nleaves;
  d.k = h.k; d.h = h.h; d.m = h.m;
  d.m_shift = (h.m & (h.m - 1)) == 0 ? (uint32_t)__builtin_ctz(h.m) : 0xFFFFFFFFu;
  for (uint32_t i = 0; i < (uint32_t)kMaxResidues; ++i) { d.res_numer[i] = i < h.m ? h.res_numer[i] : 0; d.res_base[i] = i < h.m ? h.res_base[i] : 0; }
  d.local_expand = h.max_expand_depth + 2 <= 32 ? 1u : 0u;
  { // scan strategy (match.cu phase B).  Small buckets: every lane scans whole buckets on its own with 128-bit loads.
    // Once a typical hit bucket spans several 128-byte lines, buckets are streamed through shared memory by bulk copies.
    const double sb = h.size_biased_bucket; // entries in the bucket an indexed k-mer lands in
    ix->staged = sb > 24.0;
    if (const char* env = getenv("KREPP_SCAN")) { if (!strcmp(env, "staged")) ix->staged = true; else if (!strcmp(env, "lane")) ix->staged = false; }
  }
  // Pipeline (see sorted.cu): indexes with large buckets are matched bucket-sorted, so that a bucket is read from HBM once per
  // batch; small-bucket indexes keep the fused kernel.  KREPP_PIPELINE=sorted|fused overrides (read again per batch slot).
  ix->sorted_default = ix->sorted_ok && (ix->staged || h.nshards > 1);
  ix->resident_warps = match_resident_warps(device, h.k, ix->staged);
  if (ix->staged && ix->resident_warps == 0) { ix->staged = false; ix->resident_warps = match_resident_warps(device, h.k, false); } // k > 28: the ring does not fit beside an 8-table LUT
  *out = ix;
  return KREPP_OK;
}

void krepp_index_close(krepp_index_t* ix)
{
  if (!ix) return;
  if (ix->device != KREPP_DEVICE_NONE) cudaSetDevice(ix->device);
  for (void* p : ix->allocs) cudaFree(p);
  delete ix;
}

int krepp_index_shard_info(const krepp_index_t* ix, krepp_shard_info_t* o, uint32_t* row_splits, uint32_t cap)
{
  if (!ix || !o) return fail(KREPP_ERR_ARG, "krepp_index_shard_info: null argument");
  const HostIndex& h = ix->host;
  o->shard = h.shard; o->nshards = h.nshards; o->row0 = h.row0; o->row1 = h.row1; o->first_entry = h.ent0; o->n_entries = h.cmer.size();
  if (row_splits) for (uint32_t g = 0; g <= h.nshards && g < cap; ++g) row_splits[g] = h.row_splits[g];
  return KREPP_OK;
}

int krepp_index_plan_shards(const char* index_dir, int device, uint64_t budget_bytes, uint32_t* nshards, uint64_t* whole_bytes, uint64_t* shard_bytes)
{
  if (!index_dir || !nshards) return fail(KREPP_ERR_ARG, "krepp_index_plan_shards: null argument");
  if (!budget_bytes) {
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) return fail(KREPP_ERR_CUDA, "no CUDA device %d to take the memory budget from", device);
    budget_bytes = (uint64_t)fr / 4 * 3;
  }
  HostIndex h;
  std::string err = h.load(index_dir, 0, 1, false);
  if (!err.empty()) return fail(KREPP_ERR_IO, "%s", err.c_str());
  const uint64_t repl = h.replicated_device_bytes();
  if (whole_bytes) *whole_bytes = repl + h.shard_table_device_bytes(1);
  for (uint32_t n = 1; n <= KREPP_MAX_SHARDS; ++n) {
    const uint64_t b = repl + h.shard_table_device_bytes(n);
    if (b <= budget_bytes) { *nshards = n; if (shard_bytes) *shard_bytes = b; return KREPP_OK; }
  }
  return fail(KREPP_ERR_CAPACITY, "the index does not fit %llu bytes per GPU even as %d bucket-range shards (%llu bytes are replicated on every shard)",
              (unsigned long long)budget_bytes, KREPP_MAX_SHARDS, (unsigned long long)repl);
}

int krepp_index_info(const krepp_index_t* ix, krepp_index_info_t* o)
{
  if (!ix || !o) return fail(KREPP_ERR_ARG, "krepp_index_info: null argument");
  const HostIndex& h = ix->host;
  o->k = h.k; o->w = h.w; o->h = h.h; o->m = h.m; o->r = h.r; o->frac = h.frac; o->nrows = h.nrows; o->nkmers = h.nkmers;
  o->nnodes = h.tree.nnodes; o->nleaves = h.tree.nleaves; o->nsubsets = h.nsubsets; o->root_se = h.tree.root;
  o->mask_hash_bp = h.mask_hash_bp; o->mask_drop_lr = h.mask_drop_lr; o->device_bytes = ix->device_bytes;
  o->mean_bucket = h.mean_bucket; o->size_biased_bucket = h.size_biased_bucket;
  return KREPP_OK;
}

int krepp_index_host_checksums(const krepp_index_t* ix, uint64_t out[4])
{
  if (!ix || !out) return fail(KREPP_ERR_ARG, "krepp_index_host_checksums: null argument");
  const HostIndex& h = ix->host;
  uint64_t a = 0, b = 0, c = 0, d = 0;
  for (uint64_t e : h.cmer) a += e;
  for (uint32_t v : h.inc32) b += v;
  for (size_t col = 0; col + 1 < h.cbeg.size(); ++col) {
    c += (uint64_t)col * (h.cbeg[col + 1] - h.cbeg[col]);
    for (uint32_t i = h.cbeg[col]; i < h.cbeg[col + 1]; ++i) d += (uint64_t)(col + 1) * ((uint64_t)h.cleaf[i] + 1);
  }
  out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  return KREPP_OK;
}

const char* krepp_index_node_name(const krepp_index_t* ix, uint32_t se, int return_na)
{
  if (!ix) return "";
  g_name = ix->host.tree.node_name(se, return_na != 0);
  return g_name.c_str();
}

int krepp_index_tree(const krepp_index_t* ix, uint32_t* parent, uint32_t* nchildren, uint8_t* is_leaf, double* blen)
{
  if (!ix) return fail(KREPP_ERR_ARG, "krepp_index_tree: null index");
  const HostTree& t = ix->host.tree;
  const size_t n = t.nnodes + 1;
  if (parent) std::memcpy(parent, t.parent.data(), n * 4);
  if (nchildren) std::memcpy(nchildren, t.nchildren.data(), n * 4);
  if (is_leaf) std::memcpy(is_leaf, t.is_leaf.data(), n);
  if (blen) std::memcpy(blen, t.blen.data(), n * 8);
  return KREPP_OK;
}

size_t krepp_index_jplace_tree(const krepp_index_t* ix, char* buf, size_t cap)
{
  if (!ix) return 0;
  const std::string s = ix->host.tree.jplace_newick();
  if (buf && cap) { const size_t n = std::min(cap - 1, s.size()); std::memcpy(buf, s.data(), n); buf[n] = 0; }
  return s.size();
}

// ------------------------------------------------------------------------------------------------ batches

static void free_records(krepp_batch* b)
{
  for (void* p : {(void*)b->d_rec_alias, (void*)b->d_rec_work, (void*)b->d_rec_read, (void*)b->d_rec_slot, (void*)b->d_rec_hist, (void*)b->d_rec_flags, (void*)b->d_rec_match,
                  (void*)b->d_rec_hdmin, (void*)b->d_rec_d, (void*)b->d_rec_v, (void*)b->d_rec_chisq, (void*)b->d_out_rec})
    if (p) cudaFree(p);
  for (void* p : {(void*)b->v_rec_read, (void*)b->v_rec_slot, (void*)b->v_rec_hist}) if (p) cudaFree(p);
  b->v_rec_read = b->v_rec_slot = b->v_rec_hist = nullptr;
  if (b->h_rec) cudaFreeHost(b->h_rec);
  if (b->h_hist) cudaFreeHost(b->h_hist);
  if (b->d_out_brief) cudaFree(b->d_out_brief);
  if (b->h_brief) cudaFreeHost(b->h_brief);
  b->d_out_brief = nullptr; b->h_brief = nullptr;
  if (b->d_dist_rows) cudaFree(b->d_dist_rows);
  if (b->h_dist_rows) cudaFreeHost(b->h_dist_rows);
  b->d_dist_rows = nullptr; b->h_dist_rows = nullptr;
  b->d_rec_read = b->d_rec_slot = b->d_rec_hist = b->d_rec_flags = b->d_rec_match = b->d_rec_hdmin = b->d_rec_work = b->d_rec_alias = nullptr;
  b->d_rec_d = b->d_rec_v = b->d_rec_chisq = nullptr; b->d_out_rec = nullptr; b->h_rec = nullptr; b->h_hist = nullptr;
}

static int alloc_placements(krepp_batch* b, uint32_t cap)
{
  if (b->d_place) cudaFree(b->d_place);
  if (b->h_place) cudaFreeHost(b->h_place);
  b->d_place = nullptr; b->h_place = nullptr;
  b->place_cap = cap;
  CU(cudaMalloc(&b->d_place, sizeof(krepp_placement_t) * (size_t)cap));
  return KREPP_OK;
}

static int alloc_place_nodes(krepp_batch* b, uint64_t cap)
{
  if (cap > 0x7FFFFFFFull) return fail(KREPP_ERR_CAPACITY, "batch touches too many tree nodes; submit fewer reads per batch");
  for (void* p : {(void*)b->d_pn_read, (void*)b->d_pn_se, (void*)b->d_pn_flags, (void*)b->d_pn_work, (void*)b->d_pn_mc, (void*)b->d_pn_uc, (void*)b->d_pn_rho,
                  (void*)b->d_pn_d, (void*)b->d_pn_v, (void*)b->d_pn_chisq})
    if (p) cudaFree(p);
  b->d_pn_read = b->d_pn_se = b->d_pn_flags = b->d_pn_work = nullptr;
  b->d_pn_mc = b->d_pn_uc = b->d_pn_rho = b->d_pn_d = b->d_pn_v = b->d_pn_chisq = nullptr;
  b->node_cap = (uint32_t)cap;
  const size_t stride = b->p.hdist_th + 1;
  CU(cudaMalloc(&b->d_pn_read, 4ull * cap)); CU(cudaMalloc(&b->d_pn_se, 4ull * cap)); CU(cudaMalloc(&b->d_pn_flags, 4ull * cap)); CU(cudaMalloc(&b->d_pn_work, 4ull * cap));
  CU(cudaMalloc(&b->d_pn_mc, 8ull * cap * stride)); CU(cudaMalloc(&b->d_pn_uc, 8ull * cap)); CU(cudaMalloc(&b->d_pn_rho, 8ull * cap));
  CU(cudaMalloc(&b->d_pn_d, 8ull * cap)); CU(cudaMalloc(&b->d_pn_v, 8ull * cap)); CU(cudaMalloc(&b->d_pn_chisq, 8ull * cap));
  return KREPP_OK;
}

static int alloc_records(krepp_batch* b, uint32_t cap)
{
  free_records(b);
  const size_t stride = b->p.hdist_th + 1;
  b->rec_cap = cap;
  CU(cudaMalloc(&b->d_rec_read, 4ull * cap)); CU(cudaMalloc(&b->d_rec_slot, 4ull * cap)); CU(cudaMalloc(&b->d_rec_hist, 4ull * cap * stride));
  CU(cudaMalloc(&b->d_rec_flags, 4ull * cap)); CU(cudaMalloc(&b->d_rec_match, 4ull * cap)); CU(cudaMalloc(&b->d_rec_hdmin, 4ull * cap));
  CU(cudaMalloc(&b->d_rec_work, 4ull * cap)); CU(cudaMalloc(&b->d_rec_alias, 4ull * cap));
  CU(cudaMalloc(&b->d_rec_d, 8ull * cap)); CU(cudaMalloc(&b->d_rec_v, 8ull * cap)); CU(cudaMalloc(&b->d_rec_chisq, 8ull * cap));
  CU(cudaMalloc(&b->d_out_rec, sizeof(krepp_record_t) * (size_t)cap));
  if (b->out_rows & KREPP_OUT_BRIEF) CU(cudaMalloc(&b->d_out_brief, sizeof(krepp_brief_t) * (size_t)cap));
  if (b->out_rows & KREPP_OUT_DIST) CU(cudaMalloc(&b->d_dist_rows, (size_t)b->dist_row_bytes * cap));
  if (b->d_voff) { CU(cudaMalloc(&b->v_rec_read, 4ull * cap)); CU(cudaMalloc(&b->v_rec_slot, 4ull * cap)); CU(cudaMalloc(&b->v_rec_hist, 4ull * cap * stride)); } // segments' rows
  // the page-locked host copies are sized when a wait first needs them (host_rows): page-locking is slow and most callers want one form only
  return KREPP_OK;
}

// Buffers of the segmented path, allocated by the first batch that holds a long read.
static int alloc_segments(krepp_batch* b)
{
  if (b->d_voff) return KREPP_OK;
  const HostIndex& h = b->ix->host;
  const size_t nv = b->cap_vreads, stride = b->p.hdist_th + 1;
  CU(cudaMallocHost(&b->h_voff, 16ull * nv)); CU(cudaMallocHost(&b->h_vbegin, 4ull * (b->max_reads + 1ull)));
  CU(cudaMalloc(&b->d_voff, 16ull * nv)); CU(cudaMalloc(&b->d_vbegin, 4ull * (b->max_reads + 1ull)));
  CU(cudaMalloc(&b->v_onmers, 4ull * nv)); CU(cudaMalloc(&b->v_wn, 8ull * nv)); CU(cudaMalloc(&b->v_hdfilt, 8ull * nv));
  CU(cudaMalloc(&b->v_rec_begin, 4ull * nv)); CU(cudaMalloc(&b->v_rec_count, 4ull * nv));
  CU(cudaMalloc(&b->v_rec_read, 4ull * b->rec_cap)); CU(cudaMalloc(&b->v_rec_slot, 4ull * b->rec_cap)); CU(cudaMalloc(&b->v_rec_hist, 4ull * b->rec_cap * stride));
  CU(cudaMalloc(&b->d_vcounters, 32)); CU(cudaMalloc(&b->d_seg_claim, 4));
  // one dense table of 2 * nleaves histograms per warp of segment_combine_kernel, at most 256 MB of them
  const size_t per_warp = 4ull * 2ull * std::max<uint32_t>(h.tree.nleaves, 1) * stride;
  b->seg_ctas = (int)std::max<size_t>(1, std::min<size_t>((size_t)b->ix->sms * 4, (256ull << 20) / (per_warp * 8)));
  CU(cudaMalloc(&b->d_seg_scratch, per_warp * 8 * (size_t)b->seg_ctas));
  CU(cudaMemsetAsync(b->d_seg_scratch, 0, per_warp * 8 * (size_t)b->seg_ctas, b->stream));
  return KREPP_OK;
}

// Cuts the batch's long reads into segments (host offsets, already relative to the first base).  Sets seg_nv.
static int plan_segments(krepp_batch* b, const uint64_t* off, uint32_t n_reads)
{
  b->seg_nv = 0;
  const uint64_t S = b->seg_windows, k = b->ix->host.k;
  if (!S || b->ix->host.nshards > 1 || b->d_tap) return KREPP_OK; // (the parity tap names lookups by read and position: whole reads only)
  bool any = false;
  for (uint32_t r = 0; r < n_reads && !any; ++r) any = off[r + 1] - off[r] >= 2 * S + k;
  if (!any) return KREPP_OK;
  if (int rc = alloc_segments(b)) return rc;
  uint64_t nv = 0;
  for (uint32_t r = 0; r < n_reads; ++r) {
    b->h_vbegin[r] = (uint32_t)nv;
    const uint64_t len = off[r + 1] - off[r], W = len >= k ? len - k + 1 : 0, nseg = W > 2 * S ? (W + S - 1) / S : 1;
    if (nv + nseg > b->cap_vreads) return fail(KREPP_ERR_CAPACITY, "batch has more read segments than the slot was sized for");
    for (uint64_t j = 0; j < nseg; ++j) { // windows [j * S, (j + 1) * S) of the read: S + k - 1 bases
      b->h_voff[2 * nv] = off[r] + j * S;
      b->h_voff[2 * nv + 1] = j + 1 < nseg ? off[r] + (j + 1) * S + k - 1 : off[r + 1];
      ++nv;
    }
  }
  b->h_vbegin[n_reads] = (uint32_t)nv;
  b->seg_nv = (uint32_t)nv;
  CU(cudaMemcpyAsync(b->d_voff, b->h_voff, 16ull * nv, cudaMemcpyHostToDevice, b->stream));
  CU(cudaMemcpyAsync(b->d_vbegin, b->h_vbegin, 4ull * (n_reads + 1ull), cudaMemcpyHostToDevice, b->stream));
  return KREPP_OK;
}

static int alloc_tuples(krepp_batch* b, uint64_t cap)
{
  if (cap > 0xFFFFFFF0ull) return fail(KREPP_ERR_CAPACITY, "batch produces too many lookups; submit fewer reads per batch");
  if (b->so.tuples) cudaFree(b->so.tuples);
  b->so.tuples = nullptr; b->so.cap_lookups = (uint32_t)cap;
  CU(cudaMalloc(&b->so.tuples, 16ull * cap));
  // coarse bins of the two-level lookup sort (sorted.cu lookup_partition_kernel): at most 512 bins of a power-of-two number of
  // rows, each with room for a quarter more than an even share of the lookups.  Opt-in (KREPP_LOOKUP=binned): measured on B200
  // (r07e/f, config 3) it only ties the two-pass sort -- 28.0 + 32.3 ms against 14.0 + 47.1 ms per 10 M reads: the partition's
  // append rounds double the lookup kernel, and the bin sort's 16-byte scatter into 3.9 MB windows (148 of them open at a time,
  // far more than L2 holds) reaches DRAM as partly written sectors (8.3 GB moved for 6 GB, 2.5 TB/s).
  if (b->so.binned) cudaFree(b->so.binned);
  b->so.binned = nullptr; b->so.nbins = 0;
  const HostIndex& h = b->ix->host;
  const char* env = getenv("KREPP_LOOKUP");
  uint32_t shift = 0;
  while ((((uint64_t)h.nrows - 1) >> shift) + 1 > 512) ++shift;
  if (env && !strcmp(env, "binned") && !b->bins_off && h.nrows && (1u << shift) <= 8192u) {
    const uint32_t nbins = (uint32_t)((((uint64_t)h.nrows - 1) >> shift) + 1);
    const uint64_t per = ((cap + cap / 4) / nbins + 64 + 3) / 4 * 4;
    if (per < (1ull << 31)) {
      if (!b->so.bin_cursor) CU(cudaMalloc(&b->so.bin_cursor, 4ull * 512));
      CU(cudaMalloc(&b->so.binned, 16ull * per * nbins));
      b->so.nbins = nbins; b->so.bin_cap = (uint32_t)per; b->so.bin_shift = shift;
    }
  }
  return KREPP_OK;
}

static int alloc_hits(krepp_batch* b, uint64_t cap)
{
  if (cap > 0xFFFFFFF0ull) return fail(KREPP_ERR_CAPACITY, "batch produces too many hit entries; submit fewer reads per batch");
  if (b->so.hits_tmp) cudaFree(b->so.hits_tmp);
  if (b->so.hits) cudaFree(b->so.hits);
  b->so.hits_tmp = b->so.hits = nullptr; b->so.cap_hits = (uint32_t)cap;
  CU(cudaMalloc(&b->so.hits_tmp, 16ull * cap)); CU(cudaMalloc(&b->so.hits, 16ull * cap));
  return KREPP_OK;
}

// Sort scratch of the resolve kernel for reads whose leaf hits exceed shared memory: one region of `cap` 64-bit keys (a power of
// two) per warp of its grid.  The default grid gets 8,192 keys per warp; a read that needs more makes the host grow the
// regions to its demand, and once they are large the grid is cut down so that the scratch stays within ~2 GB (such reads --
// contigs, or short reads on an index of many near-identical genomes -- are then resolved by fewer warps at a time).
static int alloc_keys(krepp_batch* b, uint64_t need)
{
  uint64_t cap = 8192;
  while (cap < need) cap <<= 1;
  if (cap > (1ull << 28)) return fail(KREPP_ERR_CAPACITY, "a read has %llu leaf hits, more than the sort scratch of the bucket-sorted chain can hold", (unsigned long long)need);
  const uint64_t per_cta = 8ull * cap * (uint64_t)sorted_resolve_warps_per_cta(), budget = 2ull << 30;
  const uint64_t dflt = (uint64_t)sorted_resolve_warps(b->ix->sms) / (uint64_t)sorted_resolve_warps_per_cta();
  const uint64_t ctas = std::max<uint64_t>(1, std::min<uint64_t>(dflt, budget / per_cta));
  uint64_t* fresh = nullptr;
  CU(cudaMalloc(&fresh, per_cta * ctas));
  if (b->so.keys_g) cudaFree(b->so.keys_g);
  b->so.keys_g = fresh; b->so.cap_keys_g = (uint32_t)cap; b->so.res_ctas = (uint32_t)ctas;
  return KREPP_OK;
}

static int alloc_sorted(krepp_batch* b)
{
  const HostIndex& h = b->ix->host;
  SortArgs& so = b->so;
  so.nrows = h.nrows;
  CU(cudaMalloc(&so.row_count, 4ull * h.nrows)); CU(cudaMalloc(&so.row_begin, 4ull * (h.nrows + 1))); CU(cudaMalloc(&so.row_cursor, 4ull * h.nrows));
  CU(cudaMalloc(&so.hit_count, 4ull * b->cap_vreads)); CU(cudaMalloc(&so.hit_begin, 4ull * (b->cap_vreads + 1ull))); CU(cudaMalloc(&so.hit_cursor, 4ull * b->cap_vreads));
  const uint64_t nmax = std::max<uint64_t>(h.nrows, b->cap_vreads);
  CU(cudaMalloc(&so.partials, 4ull * (nmax / 4096 + 2)));
  CU(cudaMalloc(&so.sc, 32));
  CU(cudaMallocHost(&b->h_sc, 4ull * (16 + 2 * (KREPP_MAX_SHARDS + 1))));
  if (const char* env = getenv("KREPP_SORT_WIDE")) so.extra_rank_bits = (uint32_t)std::min(24, std::max(0, atoi(env)));
  if (int rc = alloc_keys(b, 8192)) return rc;
  if (b->max_bases >= (1ull << 31)) return fail(KREPP_ERR_CAPACITY, "the bucket-sorted chain counts a batch's lookups in 32 bits: at most 2^31 - 1 bases per batch");
  // every base starts at most one window = two lookups, of which (r+1)/m are eligible on average; grown on demand
  uint32_t present = 0;
  for (uint32_t res = 0; res < h.m; ++res) present += h.res_numer[res] != 0;
  const uint64_t want = (uint64_t)((double)b->max_bases * 2.0 * present / (double)h.m * 1.1) + 4096;
  if (int rc = alloc_tuples(b, std::min<uint64_t>(want, 0xFFFFFFF0ull))) return rc;
  if (int rc = alloc_hits(b, std::min<uint64_t>(std::max<uint64_t>(96ull * b->max_reads, 65536), 0xFFFFFFF0ull))) return rc;
  return KREPP_OK;
}

// Per-warp scratch of the fused match kernel (match.cu): Hamming-histogram accumulators of 2 * nleaves * (th + 1) words per
// resident warp and friends -- gigabytes on an index of many thousand references.  A slot that runs the bucket-sorted chain
// never touches it, and a bucket-range shard cannot run the fused kernel at all, so it is allocated by the first batch that
// takes the fused path (a small-bucket index, KREPP_PIPELINE=fused, or the fallback for a read outside the chain's limits).
static int fused_scratch(krepp_batch* b)
{
  if (b->d_acc) return KREPP_OK;
  const HostIndex& h = b->ix->host;
  const size_t warps = (size_t)b->ix->resident_warps, nslots = 2ull * h.tree.nleaves, stride = b->p.hdist_th + 1;
  const size_t nbm = (nslots + 31) / 32;
  b->stack_cap = 32 * (h.max_expand_depth + 2) + 64;
  CU(cudaMalloc(&b->d_acc, 4 * warps * nslots * stride)); CU(cudaMemsetAsync(b->d_acc, 0, 4 * warps * nslots * stride, b->stream));
  CU(cudaMalloc(&b->d_bitmap, 4 * warps * nbm)); CU(cudaMemsetAsync(b->d_bitmap, 0, 4 * warps * nbm, b->stream));
  CU(cudaMalloc(&b->d_marker, 4 * warps * h.tree.nleaves)); CU(cudaMemsetAsync(b->d_marker, 0xFF, 4 * warps * h.tree.nleaves, b->stream));
  CU(cudaMalloc(&b->d_stack, 4 * warps * b->stack_cap));
  CU(cudaMalloc(&b->d_tagctr, 4 * warps)); CU(cudaMemsetAsync(b->d_tagctr, 0xFF, 4 * warps, b->stream));
  return KREPP_OK;
}

int krepp_batch_create(krepp_index_t* ix, const krepp_params_t* p, uint32_t max_reads, uint64_t max_bases, krepp_batch_t** out)
{
  if (!ix || !p || !out || !max_reads) return fail(KREPP_ERR_ARG, "krepp_batch_create: bad argument");
  *out = nullptr;
  if (p->hdist_th > (uint32_t)kMaxTh) return fail(KREPP_ERR_ARG, "--hdist-th %u exceeds %d", p->hdist_th, kMaxTh);
  if (p->place && p->hdist_th < p->tau) return fail(KREPP_ERR_ARG, "The threshold tau must be less than HD threshold --hdist-th!");
  if (p->place && !ix->host.wbackbone) return fail(KREPP_ERR_ARG, "Given index lacks a tree and no backbone tree is provided..."); // ref src/krepp.cpp:61-63
  if (ix->host.is_geometry) return fail(KREPP_ERR_ARG, "this handle carries an LSH geometry only (krepp_geometry_open): there is nothing to query");
  if (ix->device == KREPP_DEVICE_NONE) return fail(KREPP_ERR_CUDA, "this index handle was opened without a device (KREPP_DEVICE_NONE); queries need a GPU");
  if (cudaSetDevice(ix->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice(%d) failed", ix->device);
  auto* b = new krepp_batch;
  *out = b; // destroyed by the caller on failure
  b->ix = ix; b->device = ix->device; b->p = *p; b->max_reads = max_reads; b->max_bases = max_bases;
  const HostIndex& h = ix->host;
  b->dist_row_bytes = h.tree.nleaves <= 65536u ? 4u : 8u;
  { // long reads are cut into segments of this many k-mer windows (KREPP_SEGMENT_WINDOWS; 0 = never)
    const char* env = getenv("KREPP_SEGMENT_WINDOWS");
    b->seg_windows = env ? (uint32_t)std::max(0, atoi(env)) : 1024u;
    if (b->seg_windows && b->seg_windows < 32) b->seg_windows = 32;
    const uint64_t nv = (uint64_t)max_reads + (b->seg_windows ? max_bases / b->seg_windows : 0) + 1;
    if (nv > (1ull << 30)) b->seg_windows = 0;
    b->cap_vreads = b->seg_windows ? (uint32_t)nv : max_reads;
  }
  llh_tables(b->tab, h.k, h.h, p->hdist_th); // HDistHistLLH tables (ref src/hdhistllh.hpp:51-69), exact integer arithmetic then converted
  CU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&b->ev0)); CU(cudaEventCreate(&b->ev1)); CU(cudaEventCreate(&b->evm0)); CU(cudaEventCreate(&b->evm1));
  for (auto& ev : b->clk.ev) CU(cudaEventCreate(&ev));
  CU(cudaMallocHost(&b->h_bases, max_bases + 64)); CU(cudaMallocHost(&b->h_offsets, 8ull * (max_reads + 1)));
  CU(cudaMalloc(&b->d_bases, max_bases + 64)); CU(cudaMalloc(&b->d_offsets, 8ull * (max_reads + 1)));
  CU(cudaMalloc(&b->d_onmers, 4ull * max_reads)); CU(cudaMalloc(&b->d_wn, 8ull * max_reads)); CU(cudaMalloc(&b->d_hdfilt, 8ull * max_reads));
  CU(cudaMalloc(&b->d_rec_begin, 4ull * max_reads)); CU(cudaMalloc(&b->d_rec_count, 4ull * max_reads)); CU(cudaMalloc(&b->d_closest, 4ull * max_reads)); CU(cudaMalloc(&b->d_nsel, 4ull * max_reads));
  CU(cudaMalloc(&b->d_counters, 32)); CU(cudaMalloc(&b->d_stats, 64)); // stats: [0..3] live, [4..7] snapshot taken before a mode B finish
  CU(cudaMallocHost(&b->h_counters, 32)); CU(cudaMallocHost(&b->h_stats, 32));
  CU(cudaMalloc(&b->d_out_read, sizeof(krepp_read_summary_t) * (size_t)max_reads));
  CU(cudaMallocHost(&b->h_read, sizeof(krepp_read_summary_t) * (size_t)max_reads));
  // (the fused kernel's per-warp scratch is allocated by the first batch that runs it: fused_scratch)
  CU(cudaMalloc(&b->d_dist_cnt, 4ull * max_reads)); CU(cudaMalloc(&b->d_dist_begin, 4ull * (max_reads + 1ull))); CU(cudaMalloc(&b->d_dist_out_begin, 4ull * (max_reads + 1ull)));
  CU(cudaMalloc(&b->d_dist_partials, 4ull * (max_reads / 4096 + 2)));
  CU(cudaMallocHost(&b->h_dist_begin, 4ull * (max_reads + 1ull)));
  const uint64_t want = std::max<uint64_t>(4ull * max_reads, 4096);
  if (int rc = alloc_records(b, (uint32_t)std::min<uint64_t>(want, 0x7FFFFFFFull))) return rc;
  { // solve memo (see SolveArgs): a table of 8 slots per read, between 2^16 and 2^24 slots; KREPP_MEMO=0 turns it off
    b->memo_bits = std::min<uint32_t>(12, (64 - 29) / (p->hdist_th + 1));
    const char* env = getenv("KREPP_MEMO");
    if (b->memo_bits >= 4 && !(env && !strcmp(env, "0"))) {
      uint64_t slots = 1ull << 16;
      while (slots < 8ull * max_reads && slots < (1ull << 24)) slots <<= 1;
      b->memo_mask = (uint32_t)(slots - 1);
      CU(cudaMalloc(&b->d_memo_key, 8 * slots)); CU(cudaMalloc(&b->d_memo_owner, 4 * slots));
    }
  }
  b->sorted = ix->sorted_default;
  if (const char* env = getenv("KREPP_PIPELINE")) {
    if (!strcmp(env, "sorted")) { if (!ix->sorted_ok) return fail(KREPP_ERR_UNSUPPORTED, "KREPP_PIPELINE=sorted: the flattened colour lists of this index are too large"); b->sorted = true; }
    else if (!strcmp(env, "fused")) b->sorted = false;
  }
  if (h.nshards > 1) { // the fused kernel walks the whole table: a shard can only serve the bucket-sorted chain
    if (!ix->sorted_ok) return fail(KREPP_ERR_UNSUPPORTED, "sharded indexes need the flattened colour lists, which are too large for this index");
    if (max_reads > (1u << 30)) return fail(KREPP_ERR_CAPACITY, "at most 2^30 reads per batch");
    b->sorted = true;
  }
  if (b->sorted) { if (int rc = alloc_sorted(b)) return rc; }
  if (p->place) {
    const size_t nn = h.tree.nnodes, nbm_nodes = (nn + 32) / 32;
    b->place_warps = (uint32_t)ix->sms * 8u * (uint32_t)kPlaceWarpsPerCta; // 32 resident warps per SM (64 registers; 40 at 48 registers spill and lose): the collect kernel is latency-bound
    const size_t pw = b->place_warps;
    CU(cudaMalloc(&b->d_place_begin, 4ull * max_reads)); CU(cudaMalloc(&b->d_place_count, 4ull * max_reads));
    CU(cudaMalloc(&b->d_node_bitmap, 4 * pw * nbm_nodes)); CU(cudaMemset(b->d_node_bitmap, 0, 4 * pw * nbm_nodes));
    CU(cudaMalloc(&b->d_node_list, 4 * pw * nn)); CU(cudaMalloc(&b->d_node_order, 4 * pw * nn));
    b->chain_cap = 4096; // doubles per warp: sum over a read's selected references of their depth; reads beyond it take the slow walk
    CU(cudaMalloc(&b->d_sel, 4 * pw * 3 * (size_t)std::max<uint32_t>(h.tree.nleaves, 1))); CU(cudaMalloc(&b->d_chain, 8 * pw * (size_t)b->chain_cap));
    CU(cudaMalloc(&b->d_pn_begin, 4ull * max_reads)); CU(cudaMalloc(&b->d_pn_count, 4ull * max_reads));
    if (int rc = alloc_place_nodes(b, std::max<uint64_t>(32ull * max_reads, 4096))) return rc; // grown to the demand when a batch needs more
    if (int rc = alloc_placements(b, (uint32_t)std::min<uint64_t>(std::max<uint64_t>(8ull * max_reads, 4096), 0x7FFFFFFFull))) return rc;
  }
  return KREPP_OK;
}

void krepp_batch_destroy(krepp_batch_t* b)
{
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  free_records(b);
  for (void* p : {(void*)b->d_dist_cnt, (void*)b->d_dist_begin, (void*)b->d_dist_out_begin, (void*)b->d_dist_partials}) if (p) cudaFree(p);
  if (b->h_dist_begin) cudaFreeHost(b->h_dist_begin);
  if (b->d_seek) cudaFree(b->d_seek);
  if (b->h_seek) cudaFreeHost(b->h_seek);
  for (void* p : {(void*)b->d_voff, (void*)b->d_vbegin, (void*)b->v_onmers, (void*)b->v_wn, (void*)b->v_hdfilt, (void*)b->v_rec_begin, (void*)b->v_rec_count, (void*)b->d_vcounters,
                  (void*)b->d_seg_scratch, (void*)b->d_seg_claim})
    if (p) cudaFree(p);
  if (b->h_voff) cudaFreeHost(b->h_voff);
  if (b->h_vbegin) cudaFreeHost(b->h_vbegin);
  for (void* p : {(void*)b->d_bases, (void*)b->d_offsets, (void*)b->d_onmers, (void*)b->d_wn, (void*)b->d_hdfilt, (void*)b->d_rec_begin,
                  (void*)b->d_rec_count, (void*)b->d_closest, (void*)b->d_nsel, (void*)b->d_memo_key, (void*)b->d_memo_owner, (void*)b->d_counters, (void*)b->d_stats, (void*)b->d_acc, (void*)b->d_bitmap,
                  (void*)b->d_marker, (void*)b->d_stack, (void*)b->d_tagctr, (void*)b->d_out_read, (void*)b->d_tap, (void*)b->d_tap_count, (void*)b->d_place_begin,
                  (void*)b->d_place_count, (void*)b->d_node_bitmap, (void*)b->d_node_list, (void*)b->d_node_order, (void*)b->d_sel, (void*)b->d_chain, (void*)b->d_pn_begin, (void*)b->d_pn_count, (void*)b->d_pn_read,
                  (void*)b->d_pn_se, (void*)b->d_pn_flags, (void*)b->d_pn_work, (void*)b->d_pn_mc, (void*)b->d_pn_uc, (void*)b->d_pn_rho, (void*)b->d_pn_d,
                  (void*)b->d_pn_v, (void*)b->d_pn_chisq, (void*)b->d_place})
    if (p) cudaFree(p);
  for (void* p : {(void*)b->so.binned, (void*)b->so.bin_cursor, (void*)b->so.row_count, (void*)b->so.row_begin, (void*)b->so.row_cursor, (void*)b->so.tuples, (void*)b->so.hits_tmp, (void*)b->so.hits,
                  (void*)b->so.hit_count, (void*)b->so.hit_begin, (void*)b->so.hit_cursor, (void*)b->so.partials, (void*)b->so.sc, (void*)b->so.keys_g})
    if (p) cudaFree(p);
  if (b->h_sc) cudaFreeHost(b->h_sc);
  for (void* p : {(void*)b->h_bases, (void*)b->h_offsets, (void*)b->h_read, (void*)b->h_counters, (void*)b->h_stats, (void*)b->h_place})
    if (p) cudaFreeHost(p);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->evm0) cudaEventDestroy(b->evm0);
  if (b->evm1) cudaEventDestroy(b->evm1);
  for (auto ev : b->clk.ev) if (ev) cudaEventDestroy(ev);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

static MatchArgs match_args(krepp_batch* b)
{
  MatchArgs m{};
  m.bases = b->in_bases; m.offsets = b->in_offsets; m.n_bases = b->n_bases; m.n_reads = b->n_reads; m.th = b->p.hdist_th; m.keep_all = b->keep_all ? 1u : 0u;
  m.onmers = b->d_onmers; m.wn = b->d_wn; m.hdfilt = b->d_hdfilt; m.rec_begin = b->d_rec_begin; m.rec_count = b->d_rec_count;
  m.rec_read = b->d_rec_read; m.rec_slot = b->d_rec_slot; m.rec_hist = b->d_rec_hist; m.rec_cap = b->rec_cap; m.counters = b->d_counters;
  m.acc = b->d_acc; m.bitmap = b->d_bitmap; m.marker = b->d_marker; m.stack = b->d_stack; m.stack_cap = b->stack_cap; m.tagctr = b->d_tagctr; m.stats = b->d_stats;
  m.tap = b->d_tap; m.tap_count = b->d_tap_count; m.tap_cap = b->tap_cap;
  if (b->seg_nv) { // the match step sees the segments as its reads and writes their rows to the staging arrays (segment_combine_kernel follows)
    m.offsets = b->d_voff; m.off_pairs = 1; m.n_reads = b->seg_nv; m.keep_all = 1u; // (the gate needs the whole read: segment_combine_kernel applies it)
    m.onmers = b->v_onmers; m.wn = b->v_wn; m.hdfilt = b->v_hdfilt; m.rec_begin = b->v_rec_begin; m.rec_count = b->v_rec_count;
    m.rec_read = b->v_rec_read; m.rec_slot = b->v_rec_slot; m.rec_hist = b->v_rec_hist; m.counters = b->d_vcounters;
  }
  return m;
}

// Enqueues all kernels of one batch on the slot's stream (inputs already on the device).  In mode B's finish phase
// (krepp_shard_finish) the match step starts from the hit entries the shard owners returned.
static int enqueue(krepp_batch* b)
{
  krepp_index* ix = b->ix;
  const HostIndex& h = ix->host;
  cudaStream_t s = b->stream;
  CU(cudaMemsetAsync(b->d_counters, 0, 32, s));
  if (b->seg_nv) { CU(cudaMemsetAsync(b->d_vcounters, 0, 32, s)); CU(cudaMemsetAsync(b->d_seg_claim, 0, 4, s)); }
  if (b->shard_hits) CU(cudaMemcpyAsync(b->d_stats, b->d_stats + 4, 32, cudaMemcpyDeviceToDevice, s)); // the lookup / join phases' counts
  else CU(cudaMemsetAsync(b->d_stats, 0, 32, s));
  if (b->d_tap_count && !b->shard_hits) CU(cudaMemsetAsync(b->d_tap_count, 0, 8, s));
  const bool fused = !b->shard_hits && !(b->sorted && !b->fused_once);
  if (fused) { if (int rc = fused_scratch(b)) return rc; }
  MatchArgs m = match_args(b);
  CU(cudaEventRecord(b->evm0, s));
  if (!b->shard_hits) { b->clk.n = 0; b->clk.tick("start", s); }
  uint32_t match_launches = 1;
  if (b->shard_hits) {
    SortArgs so = b->so;
    so.hits_tmp = const_cast<uint4*>(static_cast<const uint4*>(b->shard_hits));
    b->h_sc[9] = (uint32_t)b->shard_n_hits;
    CU(cudaMemcpyAsync(so.sc, b->h_sc + 9, 4, cudaMemcpyHostToDevice, s));
    CU(launch_shard_finish(ix->dev, m, so, ix->sms, s, &b->clk));
    CU(cudaMemcpyAsync(b->h_sc + 5, so.sc + 5, 4, cudaMemcpyDeviceToHost, s));
    match_launches = 6;
  } else if (b->sorted && !b->fused_once) {
    CU(launch_match_sorted(ix->dev, m, b->so, ix->sms, b->d_tap != nullptr, s, &match_launches, &b->clk));
    CU(cudaMemcpyAsync(b->h_sc, b->so.sc, 32, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(b->h_sc + 8, b->so.row_begin + b->so.nrows, 4, cudaMemcpyDeviceToHost, s));
  } else {
    CU(launch_match(ix->dev, m, ix->resident_warps, ix->staged, b->d_tap != nullptr, s));
    b->clk.tick("match_kernel", s);
    (void)fused;
  }
  if (b->seg_nv) {
    SegArgs g{};
    g.n_reads = b->n_reads; g.vbegin = b->d_vbegin;
    g.v_onmers = b->v_onmers; g.v_wn = b->v_wn; g.v_hdfilt = b->v_hdfilt; g.v_rec_begin = b->v_rec_begin; g.v_rec_count = b->v_rec_count;
    g.v_rec_slot = b->v_rec_slot; g.v_rec_hist = b->v_rec_hist; g.v_counters = b->d_vcounters;
    g.onmers = b->d_onmers; g.wn = b->d_wn; g.hdfilt = b->d_hdfilt; g.rec_begin = b->d_rec_begin; g.rec_count = b->d_rec_count;
    g.rec_read = b->d_rec_read; g.rec_slot = b->d_rec_slot; g.rec_hist = b->d_rec_hist; g.counters = b->d_counters;
    g.rec_cap = b->rec_cap; g.th = b->p.hdist_th; g.keep_all = b->keep_all ? 1u : 0u; g.nleaves = h.tree.nleaves;
    g.leaf_rank = ix->dev.leaf_rank; g.leaf_se = ix->dev.leaf_se; g.scratch = b->d_seg_scratch; g.claim = b->d_seg_claim;
    CU(launch_segment_combine(g, b->seg_ctas, s));
    b->clk.tick("segment_combine_kernel", s);
    ++match_launches;
  }
  CU(cudaEventRecord(b->evm1, s));
  SolveArgs sa{};
  sa.n_reads = b->n_reads; sa.th = b->p.hdist_th; sa.k = h.k; sa.h = h.h; sa.n_records = b->rec_cap; sa.counters = b->d_counters; sa.work = b->d_rec_work;
  sa.onmers = b->d_onmers; sa.hdfilt = b->d_hdfilt; sa.rec_begin = b->d_rec_begin; sa.rec_count = b->d_rec_count;
  sa.rec_read = b->d_rec_read; sa.rec_slot = b->d_rec_slot; sa.rec_hist = b->d_rec_hist; sa.rho = ix->dev.rho;
  sa.rec_d = b->d_rec_d; sa.rec_v = b->d_rec_v; sa.rec_chisq = b->d_rec_chisq; sa.rec_flags = b->d_rec_flags; sa.rec_match = b->d_rec_match;
  sa.rec_hdmin = b->d_rec_hdmin; sa.closest = b->d_closest; sa.nsel = b->d_nsel;
  sa.memo_key = b->d_memo_key; sa.memo_owner = b->d_memo_owner; sa.memo_mask = b->memo_mask; sa.memo_bits = b->memo_bits; sa.rec_alias = b->d_rec_alias;
  sa.want_chisq = (!b->p.no_filter || b->p.summarize || b->p.place) ? 1 : 0;
  CU(launch_solve(sa, b->tab, ix->sms, s, &b->clk));
  b->launches = 3 + match_launches + (sa.want_chisq ? 1 : 0) + (sa.memo_mask ? 1 : 0);
  if (b->p.place) {
    PlaceArgs pa{};
    pa.s = sa; pa.offsets = b->in_offsets; pa.tau = b->p.tau; pa.no_filter = b->p.no_filter; pa.chisq_value = b->p.chisq;
    pa.parent = ix->dev.parent; pa.nchildren = ix->dev.nchildren; pa.eff = ix->dev.eff_nchildren; pa.subtree = ix->dev.subtree; pa.depth = ix->dev.depth; pa.logw = getenv("KREPP_PLACE_ORDERED") ? nullptr : ix->dev.logw; pa.blen = ix->dev.blen; pa.leaf_rank = ix->dev.leaf_rank;
    pa.nnodes = h.tree.nnodes; pa.nleaves = h.tree.nleaves; pa.sel = b->d_sel; pa.chain = b->d_chain; pa.chain_cap = b->chain_cap;
    pa.node_bitmap = b->d_node_bitmap; pa.node_list = b->d_node_list; pa.node_order = b->d_node_order;
    pa.node_cap = b->node_cap; pa.pn_read = b->d_pn_read; pa.pn_se = b->d_pn_se; pa.pn_flags = b->d_pn_flags; pa.pn_work = b->d_pn_work;
    pa.pn_mc = b->d_pn_mc; pa.pn_uc = b->d_pn_uc; pa.pn_rho = b->d_pn_rho; pa.pn_d = b->d_pn_d; pa.pn_v = b->d_pn_v; pa.pn_chisq = b->d_pn_chisq;
    pa.pn_begin = b->d_pn_begin; pa.pn_count = b->d_pn_count;
    pa.placements = b->d_place; pa.place_cap = b->place_cap; pa.counters = b->d_counters;
    pa.place_begin = b->d_place_begin; pa.place_count = b->d_place_count;
    CU(launch_place(pa, b->tab, (int)(b->place_warps / kPlaceWarpsPerCta), ix->sms, s, &b->clk));
    b->launches += 4;
  }
  // Output rows in the forms the caller asked for (krepp_batch_set_output): the public AoS structs, and / or the printed rows of
  // `krepp dist` chosen, ordered and rounded here so that 4 bytes per read and 4-8 per printed row leave the device.
  const bool want_full = (b->out_rows & KREPP_OUT_RECORDS) != 0, want_brief = (b->out_rows & KREPP_OUT_BRIEF) != 0;
  const bool want_sum = (b->out_rows & KREPP_OUT_SUMMARIES) != 0, want_dist = (b->out_rows & KREPP_OUT_DIST) != 0;
  if (want_full || want_brief || want_sum) {
    finalize_kernel<<<ix->sms * 4, 128, 0, s>>>(sa, want_full ? b->d_out_rec : nullptr, want_sum ? b->d_out_read : nullptr, b->d_wn, b->d_place_begin, b->d_place_count,
                                                want_brief ? b->d_out_brief : nullptr, b->p.chisq);
    CU(cudaGetLastError());
    ++b->launches;
  }
  if (want_dist) {
    DistOut d{};
    d.n_reads = b->n_reads; d.rec_begin = b->d_rec_begin; d.rec_count = b->d_rec_count; d.rec_slot = b->d_rec_slot; d.rec_flags = b->d_rec_flags;
    d.rec_d = b->d_rec_d; d.rec_chisq = b->d_rec_chisq; d.closest = b->d_closest; d.leaf_rank = ix->dev.leaf_rank; d.counters = b->d_counters;
    d.summarize = b->p.summarize; d.multi = b->p.multi; d.no_filter = b->p.no_filter; d.has_max = std::isnan(b->p.dist_max) ? 0 : 1;
    d.dist_max = b->p.dist_max; d.chisq_value = b->p.chisq;
    d.cnt = b->d_dist_cnt; d.begin = b->d_dist_begin; d.out_begin = b->d_dist_out_begin; d.rows = b->d_dist_rows; d.row_bytes = b->dist_row_bytes;
    if (b->n_reads) {
      dist_count_kernel<<<ix->sms * 8, 128, 0, s>>>(d);
      CU(cudaGetLastError());
      CU(exclusive_scan(b->d_dist_cnt, b->n_reads, b->d_dist_partials, b->d_dist_begin, nullptr, s));
      dist_emit_kernel<<<ix->sms * 8, 128, 0, s>>>(d);
      CU(cudaGetLastError());
      b->launches += 5;
    }
  }
  const bool want_seek = (b->out_rows & KREPP_OUT_SEEK) != 0;
  if (want_seek && b->n_reads) {
    CU(launch_seek(sa, b->tab, b->d_seek, ix->sms, s));
    ++b->launches;
  }
  b->clk.tick("finalize_kernel / dist rows", s);
  CU(cudaEventRecord(b->ev1, s)); // kernels only: [ev0, ev1] excludes the host<->device copies on both sides
  CU(cudaMemcpyAsync(b->h_counters, b->d_counters, 32, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(b->h_stats, b->d_stats, 32, cudaMemcpyDeviceToHost, s));
  b->summaries_copied = want_sum;
  if (want_sum) CU(cudaMemcpyAsync(b->h_read, b->d_out_read, sizeof(krepp_read_summary_t) * (size_t)b->n_reads, cudaMemcpyDeviceToHost, s));
  if (want_dist && b->n_reads) CU(cudaMemcpyAsync(b->h_dist_begin, b->d_dist_out_begin, 4ull * (b->n_reads + 1ull), cudaMemcpyDeviceToHost, s));
  if (want_seek && b->n_reads) CU(cudaMemcpyAsync(b->h_seek, b->d_seek, 8ull * b->n_reads, cudaMemcpyDeviceToHost, s));
  return KREPP_OK;
}

int krepp_batch_host_buffers(krepp_batch_t* b, char** bases, uint64_t** offsets)
{
  if (!b) return fail(KREPP_ERR_ARG, "null batch");
  if (bases) *bases = b->h_bases;
  if (offsets) *offsets = b->h_offsets;
  return KREPP_OK;
}

int krepp_batch_submit(krepp_batch_t* b, const char* bases, const uint64_t* offsets, uint32_t n_reads)
{
  if (!b || !bases || !offsets) return fail(KREPP_ERR_ARG, "krepp_batch_submit: null argument");
  if (n_reads > b->max_reads) return fail(KREPP_ERR_CAPACITY, "batch of %u reads exceeds the slot capacity of %u", n_reads, b->max_reads);
  const uint64_t nb = n_reads ? offsets[n_reads] - offsets[0] : 0;
  if (nb > b->max_bases) return fail(KREPP_ERR_CAPACITY, "batch of %llu bases exceeds the slot capacity of %llu", (unsigned long long)nb, (unsigned long long)b->max_bases);
  if (b->ix->host.nshards > 1) return fail(KREPP_ERR_UNSUPPORTED, "this index handle holds one bucket-range shard: use krepp_shard_lookup / _join / _finish");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream)); // the previous batch of this slot must be finished before its buffers are reused
  // Page-locked caller memory (cudaMallocHost / cudaHostRegister) is copied to the device straight from where it lies;
  // pageable memory goes through the slot's own pinned staging buffer first.
  const char* src = b->h_bases;
  if (bases != b->h_bases) {
    cudaPointerAttributes at{};
    if (nb && cudaPointerGetAttributes(&at, bases + offsets[0]) == cudaSuccess && at.type == cudaMemoryTypeHost) src = bases + offsets[0];
    else { cudaGetLastError(); std::memcpy(b->h_bases, bases + offsets[0], nb); }
  }
  if (offsets != b->h_offsets || offsets[0] != 0) {
    const uint64_t o0 = offsets[0];
    for (uint32_t i = 0; i <= n_reads; ++i) b->h_offsets[i] = offsets[i] - o0;
  }
  b->n_reads = n_reads; b->n_bases = nb; b->device_input = false; b->fused_once = false; b->shard_hits = nullptr;
  b->in_bases = b->d_bases; b->in_offsets = b->d_offsets;
  if (int rc = plan_segments(b, b->h_offsets, n_reads)) return rc;
  CU(cudaMemcpyAsync(b->d_bases, src, nb, cudaMemcpyHostToDevice, b->stream));
  CU(cudaMemcpyAsync(b->d_offsets, b->h_offsets, 8ull * (n_reads + 1), cudaMemcpyHostToDevice, b->stream));
  CU(cudaEventRecord(b->ev0, b->stream));
  if (int rc = enqueue(b)) return rc;
  b->submitted = true; b->pending = true;
  return KREPP_OK;
}

int krepp_batch_submit_device(krepp_batch_t* b, const char* d_bases, const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases)
{
  if (!b || !d_bases || !d_offsets) return fail(KREPP_ERR_ARG, "krepp_batch_submit_device: null argument");
  if (n_reads > b->max_reads) return fail(KREPP_ERR_CAPACITY, "batch of %u reads exceeds the slot capacity of %u", n_reads, b->max_reads);
  if (reinterpret_cast<uintptr_t>(d_bases) & 15) return fail(KREPP_ERR_ARG, "device bases pointer must be 16-byte aligned");
  if (b->ix->host.nshards > 1) return fail(KREPP_ERR_UNSUPPORTED, "this index handle holds one bucket-range shard: use krepp_shard_lookup / _join / _finish");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  b->n_reads = n_reads; b->n_bases = n_bases; b->device_input = true; b->fused_once = false; b->shard_hits = nullptr;
  b->in_bases = d_bases; b->in_offsets = d_offsets;
  b->seg_nv = 0; // the offsets live on the device: such batches are not cut (their callers cut long sequences themselves)
  CU(cudaEventRecord(b->ev0, b->stream));
  if (int rc = enqueue(b)) return rc;
  b->submitted = true; b->pending = true;
  return KREPP_OK;
}

static int wait_impl(krepp_batch_t* b, krepp_results_t* out, uint32_t rows);
int krepp_batch_wait(krepp_batch_t* b, krepp_results_t* out) { return wait_impl(b, out, b ? b->out_rows : 0u); }
int krepp_batch_wait_device(krepp_batch_t* b, krepp_results_t* out) { return wait_impl(b, out, 0u); }

// page-locked host arrays for the row forms this wait copies back (freed with the device arrays when those grow)
static int host_rows(krepp_batch* b, uint32_t rows)
{
  const size_t cap = b->rec_cap, stride = b->p.hdist_th + 1;
  if ((rows & KREPP_OUT_RECORDS) && !b->h_rec) CU(cudaMallocHost(&b->h_rec, sizeof(krepp_record_t) * cap));
  if ((rows & KREPP_OUT_HIST) && !b->h_hist) CU(cudaMallocHost(&b->h_hist, 4ull * cap * stride));
  if ((rows & KREPP_OUT_BRIEF) && !b->h_brief) CU(cudaMallocHost(&b->h_brief, sizeof(krepp_brief_t) * cap));
  if ((rows & KREPP_OUT_DIST) && !b->h_dist_rows) CU(cudaMallocHost(&b->h_dist_rows, (size_t)b->dist_row_bytes * cap));
  if ((rows & KREPP_OUT_PLACEMENTS) && b->p.place && !b->h_place) CU(cudaMallocHost(&b->h_place, sizeof(krepp_placement_t) * (size_t)b->place_cap));
  return KREPP_OK;
}

int krepp_batch_set_output(krepp_batch_t* b, uint32_t rows)
{
  if (!b || (rows & ~(uint32_t)(KREPP_OUT_ALL | KREPP_OUT_BRIEF | KREPP_OUT_DIST | KREPP_OUT_SEEK))) return fail(KREPP_ERR_ARG, "krepp_batch_set_output: bad argument");
  if ((rows & KREPP_OUT_SEEK) && !b->ix->host.is_sketch) return fail(KREPP_ERR_ARG, "krepp_batch_set_output: KREPP_OUT_SEEK needs a sketch handle (krepp_sketch_open)");
  if (b->pending) return fail(KREPP_ERR_ARG, "krepp_batch_set_output: a batch is pending on this slot (the rows are assembled by the submit); call it before krepp_batch_submit or after krepp_batch_wait");
  if ((rows & KREPP_OUT_DIST) && b->p.place) return fail(KREPP_ERR_ARG, "krepp_batch_set_output: KREPP_OUT_DIST rows are those of `dist`; this slot runs `place`");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  b->out_rows = rows;
  if ((rows & KREPP_OUT_BRIEF) && !b->d_out_brief) CU(cudaMalloc(&b->d_out_brief, sizeof(krepp_brief_t) * (size_t)b->rec_cap));
  if ((rows & KREPP_OUT_DIST) && !b->d_dist_rows) CU(cudaMalloc(&b->d_dist_rows, (size_t)b->dist_row_bytes * b->rec_cap));
  if ((rows & KREPP_OUT_SEEK) && !b->d_seek) { CU(cudaMalloc(&b->d_seek, 8ull * (b->max_reads + 1ull))); CU(cudaMallocHost(&b->h_seek, 8ull * (b->max_reads + 1ull))); }
  return KREPP_OK;
}

int krepp_batch_reserve(krepp_batch_t* b, uint64_t n_records, uint64_t n_hits, uint64_t n_nodes, uint64_t n_placements)
{
  if (!b) return fail(KREPP_ERR_ARG, "krepp_batch_reserve: null batch");
  if (b->pending) return fail(KREPP_ERR_ARG, "krepp_batch_reserve: a batch is pending on this slot");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  if (n_records > 0x7FFFFFFFull || n_placements > 0x7FFFFFFFull) return fail(KREPP_ERR_CAPACITY, "krepp_batch_reserve: at most 2^31 - 1 rows");
  if (n_records > b->rec_cap) { if (int rc = alloc_records(b, (uint32_t)n_records)) return rc; }
  if (b->sorted && n_hits > b->so.cap_hits) { if (int rc = alloc_hits(b, n_hits)) return rc; }
  if (b->p.place && n_nodes > b->node_cap) { if (int rc = alloc_place_nodes(b, n_nodes)) return rc; }
  if (b->p.place && n_placements > b->place_cap) { if (int rc = alloc_placements(b, (uint32_t)n_placements)) return rc; }
  return host_rows(b, b->out_rows);
}

static int wait_impl(krepp_batch_t* b, krepp_results_t* out, uint32_t rows)
{
  if (!b || !out) return fail(KREPP_ERR_ARG, "krepp_batch_wait: null argument");
  if (!b->submitted) return fail(KREPP_ERR_ARG, "krepp_batch_wait: nothing was submitted on this slot");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  for (int attempt = 0;; ++attempt) {
    CU(cudaStreamSynchronize(b->stream));
    if (b->h_counters[2] & kErrStackOverflow) return fail(KREPP_ERR_CAPACITY, "colour expansion stack overflow on the device");
    if (b->h_counters[2] & kErrShardData) return fail(KREPP_ERR_ARG, "krepp_shard_finish: a hit entry names a read outside the batch");
    if (b->shard_hits && (b->h_counters[2] & kErrSortFallback))
      return fail(KREPP_ERR_CAPACITY, "a read has 2^26 or more lookups, which the bucket-sorted chain cannot index, and a sharded index has no fused kernel to fall back to");
    if (b->h_counters[2] & kErrHitWrap) return fail(KREPP_ERR_CAPACITY, "batch produces 2^32 or more hit entries; submit fewer reads per batch");
    if (!(b->h_counters[2] & (kErrRecOverflow | kErrPlaceOverflow | kErrNodeOverflow | kErrLookupOverflow | kErrHitOverflow | kErrSortFallback | kErrKeysOverflow | kErrBinOverflow))) break;
    // a result buffer was too small: grow it to what the kernels asked for and run the batch again
    if (attempt >= 8) return fail(KREPP_ERR_CAPACITY, "result buffer overflow persists");
    if (b->h_counters[2] & kErrRecOverflow) {
      const uint64_t want = std::max<uint64_t>((uint64_t)b->h_counters[0] + b->h_counters[0] / 4, (uint64_t)b->rec_cap + 4096); // exact demand + 25 %
      if (want > 0x7FFFFFFFull) return fail(KREPP_ERR_CAPACITY, "batch produces too many records; submit fewer reads per batch");
      if (int rc = alloc_records(b, (uint32_t)want)) return rc;
    }
    if (b->h_counters[2] & kErrBinOverflow) { b->bins_off = true; b->so.nbins = 0; } // lookups piled on few rows: the two-pass sort sizes every row exactly
    if (b->h_counters[2] & kErrLookupOverflow) { // the batch's lookup list: exact demand + 10 %
      const uint64_t need = b->h_sc[8];
      if (int rc = alloc_tuples(b, need + need / 10 + 4096)) return rc;
    }
    if (b->h_counters[2] & kErrHitOverflow) {
      const uint64_t need = b->h_sc[0];
      if (int rc = alloc_hits(b, need + need / 4 + 4096)) return rc;
    }
    if (b->h_counters[2] & kErrKeysOverflow) { // a read with more leaf hits than a warp's sort scratch: grow it to the demand (h_sc[5])
      const int rc = alloc_keys(b, b->h_sc[5]);
      if (rc && !b->shard_hits) { b->fused_once = true; cudaGetLastError(); } // beyond any scratch: the fused kernel handles reads of any size (no such fallback on a shard)
      else if (rc) return rc;
    }
    if (b->h_counters[2] & kErrSortFallback) b->fused_once = true; // a read with 2^26 or more lookups: this batch goes through the fused kernel
    if (b->h_counters[2] & kErrNodeOverflow) { // counters[5] = tree nodes the batch touches
      const uint64_t need = b->h_counters[5];
      if (int rc = alloc_place_nodes(b, need + need / 8 + 4096)) return rc;
    }
    if (b->h_counters[2] & kErrPlaceOverflow) {
      const uint64_t want = std::max<uint64_t>((uint64_t)b->h_counters[3] + b->h_counters[3] / 4, (uint64_t)b->place_cap + 4096);
      if (want > 0x7FFFFFFFull) return fail(KREPP_ERR_CAPACITY, "batch produces too many placements; submit fewer reads per batch");
      if (int rc = alloc_placements(b, (uint32_t)want)) return rc;
    }
    if (int rc = enqueue(b)) return rc;
  }
  const uint32_t nrec = b->h_counters[0];
  const size_t stride = b->p.hdist_th + 1;
  if (int rc = host_rows(b, rows)) return rc;
  if (nrec && (rows & KREPP_OUT_RECORDS)) CU(cudaMemcpyAsync(b->h_rec, b->d_out_rec, sizeof(krepp_record_t) * (size_t)nrec, cudaMemcpyDeviceToHost, b->stream));
  if (nrec && (rows & KREPP_OUT_BRIEF)) CU(cudaMemcpyAsync(b->h_brief, b->d_out_brief, sizeof(krepp_brief_t) * (size_t)nrec, cudaMemcpyDeviceToHost, b->stream));
  if (nrec && (rows & KREPP_OUT_HIST)) CU(cudaMemcpyAsync(b->h_hist, b->d_rec_hist, 4ull * nrec * stride, cudaMemcpyDeviceToHost, b->stream));
  const uint32_t nplace = b->p.place ? b->h_counters[3] : 0;
  if (nplace && (rows & KREPP_OUT_PLACEMENTS)) CU(cudaMemcpyAsync(b->h_place, b->d_place, sizeof(krepp_placement_t) * (size_t)nplace, cudaMemcpyDeviceToHost, b->stream));
  const bool have_dist = (rows & KREPP_OUT_DIST) && (b->out_rows & KREPP_OUT_DIST);
  const uint64_t ndist = have_dist && b->n_reads ? KREPP_DIST_BEGIN(b->h_dist_begin[b->n_reads]) : 0;
  if (have_dist && !b->n_reads) b->h_dist_begin[0] = 0;
  if (ndist) CU(cudaMemcpyAsync(b->h_dist_rows, b->d_dist_rows, (size_t)b->dist_row_bytes * ndist, cudaMemcpyDeviceToHost, b->stream));
  const bool want_sum = (rows & KREPP_OUT_SUMMARIES) || rows == 0; // krepp_batch_wait_device always hands out the summaries
  if (want_sum && !b->summaries_copied) {
    if (!(b->out_rows & KREPP_OUT_SUMMARIES)) { // they were not assembled by the submit: do it now
      SolveArgs sa{};
      sa.n_reads = b->n_reads; sa.n_records = 0; sa.counters = b->d_counters; sa.onmers = b->d_onmers; sa.hdfilt = b->d_hdfilt;
      sa.rec_begin = b->d_rec_begin; sa.rec_count = b->d_rec_count; sa.closest = b->d_closest; sa.nsel = b->d_nsel;
      finalize_kernel<<<b->ix->sms * 4, 128, 0, b->stream>>>(sa, nullptr, b->d_out_read, b->d_wn, b->d_place_begin, b->d_place_count, nullptr, b->p.chisq);
      CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(b->h_read, b->d_out_read, sizeof(krepp_read_summary_t) * (size_t)b->n_reads, cudaMemcpyDeviceToHost, b->stream));
    b->summaries_copied = true;
  }
  CU(cudaStreamSynchronize(b->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  out->n_reads = b->n_reads; out->hist_stride = (uint32_t)stride; out->n_records = nrec; out->n_placements = nplace;
  out->reads = want_sum ? b->h_read : nullptr;
  out->records = (rows & KREPP_OUT_RECORDS) ? b->h_rec : nullptr; out->hist = (rows & KREPP_OUT_HIST) ? b->h_hist : nullptr;
  out->placements = nplace && (rows & KREPP_OUT_PLACEMENTS) ? b->h_place : nullptr;
  out->brief = (rows & KREPP_OUT_BRIEF) ? b->h_brief : nullptr;
  out->seek_dist = ((rows & KREPP_OUT_SEEK) && (b->out_rows & KREPP_OUT_SEEK)) ? b->h_seek : nullptr;
  out->dist_begin = have_dist ? b->h_dist_begin : nullptr; out->dist_rows = have_dist ? b->h_dist_rows : nullptr;
  out->n_dist_rows = ndist; out->dist_row_bytes = b->dist_row_bytes;
  float mms = 0;
  CU(cudaEventElapsedTime(&mms, b->evm0, b->evm1));
  out->gpu_ms = ms; out->match_ms = mms; out->gpu_launches = b->launches;
  b->pending = false;
  return KREPP_OK;
}

// ------------------------------------------------------------------------------------------------ mode B (bucket-range shards)

int krepp_shard_lookup(krepp_batch_t* b, const char* d_bases, const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases,
                       void* d_tuples, uint64_t cap_tuples, uint32_t* d_row_begin, uint64_t* send_offsets)
{
  if (!b || !d_bases || !d_offsets || !d_tuples || !d_row_begin || !send_offsets) return fail(KREPP_ERR_ARG, "krepp_shard_lookup: null argument");
  if (!b->sorted) return fail(KREPP_ERR_UNSUPPORTED, "krepp_shard_lookup: this slot does not run the bucket-sorted chain");
  if (n_reads > b->max_reads) return fail(KREPP_ERR_CAPACITY, "batch of %u reads exceeds the slot capacity of %u", n_reads, b->max_reads);
  if (reinterpret_cast<uintptr_t>(d_bases) & 15) return fail(KREPP_ERR_ARG, "device bases pointer must be 16-byte aligned");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  const HostIndex& h = b->ix->host;
  cudaStream_t s = b->stream;
  CU(cudaStreamSynchronize(s));
  b->n_reads = n_reads; b->n_bases = n_bases; b->device_input = true; b->fused_once = false; b->shard_hits = nullptr; b->submitted = false; b->seg_nv = 0;
  b->in_bases = d_bases; b->in_offsets = d_offsets;
  CU(cudaMemsetAsync(b->d_counters, 0, 32, s));
  CU(cudaMemsetAsync(b->d_stats, 0, 32, s));
  if (b->d_tap_count) CU(cudaMemsetAsync(b->d_tap_count, 0, 8, s));
  MatchArgs m = match_args(b);
  SortArgs so = b->so;
  so.tuples = static_cast<uint4*>(d_tuples); so.cap_lookups = (uint32_t)std::min<uint64_t>(cap_tuples, 0xFFFFFFF0ull); so.row_begin = d_row_begin;
  CU(cudaEventRecord(b->ev0, s));
  b->clk.n = 0; b->clk.tick("start", s);
  uint32_t* hb = b->h_sc + 16;
  for (;;) {
    CU(launch_shard_lookup(b->ix->dev, m, so, b->ix->sms, b->d_tap != nullptr, s, &b->clk));
    for (uint32_t g = 0; g <= h.nshards; ++g) CU(cudaMemcpyAsync(hb + g, d_row_begin + h.row_splits[g], 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(b->h_counters, b->d_counters, 32, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (!(b->h_counters[2] & kErrBinOverflow)) break;
    b->bins_off = true; b->so.nbins = 0; so.nbins = 0; // lookups piled on few rows: again with the two-pass sort
    CU(cudaMemsetAsync(b->d_counters, 0, 32, s));
    CU(cudaMemsetAsync(b->d_stats, 0, 32, s));
    if (b->d_tap_count) CU(cudaMemsetAsync(b->d_tap_count, 0, 8, s));
    b->clk.n = 0; b->clk.tick("start", s);
  }
  for (uint32_t g = 0; g <= h.nshards; ++g) send_offsets[g] = hb[g];
  if (b->h_counters[2] & kErrSortFallback) return fail(KREPP_ERR_CAPACITY, "a read has more lookups than the bucket-sorted chain holds");
  if (b->h_counters[2] & kErrLookupOverflow) return fail(KREPP_ERR_CAPACITY, "the batch has %llu lookups but the tuple buffer holds %llu", (unsigned long long)hb[h.nshards], (unsigned long long)cap_tuples);
  return KREPP_OK;
}

int krepp_shard_join(krepp_batch_t* b, uint32_t n_sources, const void* const* d_tuples, const uint32_t* const* d_row_begin,
                     void* d_hits, uint64_t cap_hits, uint64_t* hit_offsets)
{
  if (!b || !d_hits || !hit_offsets || (n_sources && (!d_tuples || !d_row_begin))) return fail(KREPP_ERR_ARG, "krepp_shard_join: null argument");
  if (!b->sorted) return fail(KREPP_ERR_UNSUPPORTED, "krepp_shard_join: this slot does not run the bucket-sorted chain");
  if (n_sources > KREPP_MAX_SHARDS) return fail(KREPP_ERR_ARG, "krepp_shard_join: at most %d sources", KREPP_MAX_SHARDS);
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t s = b->stream;
  CU(cudaStreamSynchronize(s));
  CU(cudaMemsetAsync(b->d_counters + 2, 0, 4, s));
  CU(cudaMemsetAsync(b->so.sc, 0, 4, s));
  uint32_t* hb = b->h_sc + 16 + KREPP_MAX_SHARDS + 1;
  hb[0] = 0;
  b->clk.tick("(exchange of lookups)", s);
  for (uint32_t src = 0; src < n_sources; ++src) {
    SortArgs so = b->so;
    so.nrows = b->ix->dev.nrows_local;
    so.tuples = const_cast<uint4*>(static_cast<const uint4*>(d_tuples[src])); so.cap_lookups = 0xFFFFFFFFu;
    so.row_begin = const_cast<uint32_t*>(d_row_begin[src]);
    so.hits_tmp = static_cast<uint4*>(d_hits); so.cap_hits = (uint32_t)std::min<uint64_t>(cap_hits, 0xFFFFFFF0ull);
    CU(launch_shard_join(b->ix->dev, so, b->p.hdist_th, b->d_counters, b->d_stats, b->ix->sms, s));
    CU(cudaMemcpyAsync(hb + src + 1, b->so.sc, 4, cudaMemcpyDeviceToHost, s));
  }
  b->clk.tick("join_kernel (as shard owner)", s);
  CU(cudaMemcpyAsync(b->h_counters, b->d_counters, 32, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  for (uint32_t src = 0; src <= n_sources; ++src) hit_offsets[src] = hb[src];
  if (b->h_counters[2] & kErrHitOverflow)
    return fail(KREPP_ERR_CAPACITY, "the joins produce %llu hit entries but the buffer holds %llu", (unsigned long long)hb[n_sources], (unsigned long long)cap_hits);
  return KREPP_OK;
}

int krepp_shard_finish(krepp_batch_t* b, const void* d_hits, uint64_t n_hits)
{
  if (!b || (!d_hits && n_hits)) return fail(KREPP_ERR_ARG, "krepp_shard_finish: null argument");
  if (!b->sorted || !b->in_bases) return fail(KREPP_ERR_ARG, "krepp_shard_finish: krepp_shard_lookup has not run on this slot");
  if (n_hits > 0xFFFFFFF0ull) return fail(KREPP_ERR_CAPACITY, "batch produces too many hit entries; submit fewer reads per batch");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  if (n_hits > b->so.cap_hits) { if (int rc = alloc_hits(b, n_hits + n_hits / 8 + 4096)) return rc; }
  static const uint4 none = {0, 0, 0, 0};
  b->shard_hits = n_hits ? d_hits : &none; b->shard_n_hits = n_hits; // (with no hits the pointer is never dereferenced)
  CU(cudaMemcpyAsync(b->d_stats + 4, b->d_stats, 32, cudaMemcpyDeviceToDevice, b->stream));
  b->clk.tick("(exchange of hit entries)", b->stream);
  if (int rc = enqueue(b)) return rc;
  b->submitted = true; b->pending = true;
  return KREPP_OK;
}

int krepp_device_alloc(int device, uint64_t bytes, void** out)
{
  if (!out) return fail(KREPP_ERR_ARG, "krepp_device_alloc: null argument");
  *out = nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  CU(cudaMalloc(out, bytes ? bytes : 16));
  return KREPP_OK;
}

void krepp_device_free(int device, void* p)
{
  if (!p) return;
  cudaSetDevice(device);
  cudaFree(p);
}

int krepp_device_copy(int dst_device, void* dst, int src_device, const void* src, uint64_t bytes)
{
  if (!bytes) return KREPP_OK;
  if (!dst || !src) return fail(KREPP_ERR_ARG, "krepp_device_copy: null argument");
  if (dst_device == KREPP_DEVICE_NONE && src_device == KREPP_DEVICE_NONE) { std::memcpy(dst, src, bytes); return KREPP_OK; }
  if (cudaSetDevice(dst_device == KREPP_DEVICE_NONE ? src_device : dst_device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  if (dst_device == KREPP_DEVICE_NONE) CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  else if (src_device == KREPP_DEVICE_NONE) CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  else {
    // device-to-device copies return before they have run (and the legacy stream they run on does not order the slots' non-blocking
    // streams), so wait for them: the caller launches kernels on other streams that read `dst` next
    if (src_device == dst_device) CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
    else CU(cudaMemcpyPeer(dst, dst_device, src, src_device, bytes)); // NVLink when peer access is possible, staged through the host otherwise
    CU(cudaStreamSynchronize(cudaStreamLegacy));
  }
  return KREPP_OK;
}

int krepp_batch_enable_tap(krepp_batch_t* b, int stage, uint64_t capacity_items)
{
  if (!b || (stage != 1 && stage != 2) || (stage == 1 && !capacity_items)) return fail(KREPP_ERR_ARG, "krepp_batch_enable_tap: bad argument");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  if (stage == 2) { b->keep_all = capacity_items != 0; return KREPP_OK; }
  if (b->d_tap) { cudaFree(b->d_tap); b->d_tap = nullptr; }
  if (!b->d_tap_count) CU(cudaMalloc(&b->d_tap_count, 8));
  CU(cudaMalloc(&b->d_tap, sizeof(uint4) * capacity_items));
  b->tap_cap = capacity_items;
  return KREPP_OK;
}

int krepp_batch_read_tap(krepp_batch_t* b, int stage, uint32_t* out, uint64_t cap_items, uint64_t* n)
{
  if (!b || stage != 1 || !n || !b->d_tap) return fail(KREPP_ERR_ARG, "krepp_batch_read_tap: bad argument or tap not enabled");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  unsigned long long cnt = 0;
  CU(cudaMemcpy(&cnt, b->d_tap_count, 8, cudaMemcpyDeviceToHost));
  *n = cnt;
  const uint64_t take = std::min<uint64_t>(std::min<uint64_t>(cnt, b->tap_cap), cap_items);
  if (out && take) CU(cudaMemcpy(out, b->d_tap, sizeof(uint4) * take, cudaMemcpyDeviceToHost));
  return KREPP_OK;
}

int krepp_batch_stage_times(krepp_batch_t* b, uint32_t cap, float* ms, const char** names, uint32_t* n)
{
  if (!b || !n) return fail(KREPP_ERR_ARG, "krepp_batch_stage_times: null argument");
  if (cudaSetDevice(b->device) != cudaSuccess) return fail(KREPP_ERR_CUDA, "cudaSetDevice failed");
  CU(cudaStreamSynchronize(b->stream));
  const uint32_t have = b->clk.n > 1 ? (uint32_t)b->clk.n - 1 : 0;
  *n = have;
  for (uint32_t i = 0; i < have && i < cap; ++i) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, b->clk.ev[i], b->clk.ev[i + 1]));
    if (ms) ms[i] = t;
    if (names) names[i] = b->clk.name[i + 1];
  }
  return KREPP_OK;
}

int krepp_batch_algorithmic_bytes(krepp_batch_t* b, uint64_t* bytes, uint64_t* lookups, uint64_t* entries)
{
  if (!b) return fail(KREPP_ERR_ARG, "null batch");
  if (bytes) *bytes = b->h_stats[0];
  if (lookups) *lookups = b->h_stats[1];
  if (entries) *entries = b->h_stats[2];
  return KREPP_OK;
}

} // extern "C"

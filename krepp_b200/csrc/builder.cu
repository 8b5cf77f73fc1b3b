// builder.cu -- device stage of `krepp index` and the krepp_builder_* entry points.
//
// The reference unions leaf tables pairwise up the guide tree, row by row, with a hash-map lookup per shared k-mer
// (IndexMultiple::build_for_subtree ref src/krepp.cpp:248-309, DynHT::union_table / union_row ref src/table.cpp:191-234).  On
// the GPU the union of ALL leaf tables is one sort:
//
//   1. every genome's leaf table (sorted unique row << 32 | encoding keys, minimizer.cu) is already in HBM; the tables are laid
//      side by side in leaf order with the leaf's rank as the value of every key            (copies + fill_leaf_kernel)
//   2. one stable radix sort by key (cub::DeviceRadixSort::SortPairs, a library primitive for a plain sort): a run of equal keys
//      is one k-mer of the library and the run's values are the references that hold it, ascending
//   3. run heads and run numbers (head_kernel + cub::DeviceScan), run starts (run_start_kernel)
//   4. one thread per run sums a 64-bit mix of its leaves (set_hash_kernel) -- the runs with equal sums are candidates for
//      "the same reference set"; the runs are sorted by that sum, the first of every group is its representative, and every
//      other run is COMPARED with the representative leaf by leaf (set_assign_kernel): a hash that merges two different sets is
//      reported, never used.  (The reference names sets by such sums and resolves clashes with a nonce, ref src/record.cpp:82-113.)
//   5. the representatives' leaf lists are gathered (set_gather_kernel): the distinct reference sets of the library, typically
//      fifty times fewer than k-mers.
//
// What leaves the device: per distinct k-mer its key and its set id, and the sets.  The host stage (library_writer.cpp) turns
// sets into colours along the tree and writes the files.  HBM: 24 bytes per (k-mer, reference) pair while sorting.
#include "../../include/krepp_b200.h"

#include "builder.hpp"

#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>

using namespace krepp;

namespace {

__global__ void fill_leaf_kernel(uint32_t* vals, uint64_t n, uint32_t leaf)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) vals[i] = leaf;
}

__global__ void head_kernel(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t* __restrict__ head)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// run_no[i] = inclusive sum of the heads: entry i belongs to run run_no[i] - 1
__global__ void run_start_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ run_no, uint64_t n, uint32_t* __restrict__ run_start, uint32_t n_runs)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (head[i]) run_start[run_no[i] - 1] = (uint32_t)i;
    if (i == 0) run_start[n_runs] = (uint32_t)n;
  }
}

__device__ __forceinline__ unsigned long long leaf_mix(uint32_t leaf, unsigned long long salt)
{ // splitmix64 finaliser
  unsigned long long x = ((unsigned long long)leaf + 1) * 0x9E3779B97F4A7C15ull + salt;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
  return x;
}

// thread per run: the run's key and the sum of its leaves' mixes
__global__ void set_hash_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ run_start, uint32_t n_runs,
                                unsigned long long salt, unsigned long long* __restrict__ run_key, unsigned long long* __restrict__ run_hash, uint32_t* __restrict__ order)
{
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_runs; r += gridDim.x * blockDim.x) {
    const uint32_t a = run_start[r], b = run_start[r + 1];
    unsigned long long sum = 0;
    for (uint32_t i = a; i < b; ++i) sum += leaf_mix(vals[i], salt);
    run_key[r] = keys[a]; run_hash[r] = sum; order[r] = r;
  }
}

__global__ void group_head_kernel(const unsigned long long* __restrict__ hash_sorted, uint32_t n_runs, uint32_t* __restrict__ ghead)
{
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += gridDim.x * blockDim.x) ghead[j] = (j == 0 || hash_sorted[j] != hash_sorted[j - 1]) ? 1u : 0u;
}

__global__ void set_rep_kernel(const uint32_t* __restrict__ ghead, const uint32_t* __restrict__ group_no, const uint32_t* __restrict__ order_sorted,
                               const uint32_t* __restrict__ run_start, uint32_t n_runs, uint32_t* __restrict__ rep, unsigned long long* __restrict__ set_len)
{
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += gridDim.x * blockDim.x)
    if (ghead[j]) { const uint32_t r = order_sorted[j]; rep[group_no[j] - 1] = r; set_len[group_no[j] - 1] = run_start[r + 1] - run_start[r]; }
}

// thread per run (in hash order): its set is its group's; the run must equal the group's representative leaf for leaf
__global__ void set_assign_kernel(const uint32_t* __restrict__ group_no, const uint32_t* __restrict__ order_sorted, const uint32_t* __restrict__ rep,
                                  const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ vals, uint32_t n_runs, uint32_t* __restrict__ set_of,
                                  uint32_t* __restrict__ clash)
{
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += gridDim.x * blockDim.x) {
    const uint32_t r = order_sorted[j], s = group_no[j] - 1, q = rep[s];
    set_of[r] = s;
    if (q == r) continue;
    const uint32_t a = run_start[r], len = run_start[r + 1] - a, c = run_start[q];
    bool same = len == run_start[q + 1] - c;
    for (uint32_t i = 0; same && i < len; ++i) same = vals[a + i] == vals[c + i];
    if (!same) atomicAdd(clash, 1u);
  }
}

__global__ void set_gather_kernel(const uint32_t* __restrict__ rep, const unsigned long long* __restrict__ set_begin, const uint32_t* __restrict__ run_start,
                                  const uint32_t* __restrict__ vals, uint32_t n_sets, uint32_t* __restrict__ set_leaves)
{ // warp per set
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t s = warp; s < n_sets; s += nwarps) {
    const uint32_t a = run_start[rep[s]];
    const unsigned long long o = set_begin[s], len = set_begin[s + 1] - o;
    for (unsigned long long i = lane; i < len; i += 32) set_leaves[o + i] = vals[a + i];
  }
}

#define B_CU(expr)                                                                                                       \
  do {                                                                                                                   \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess) { rc = set_error(KREPP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); goto done; } \
  } while (0)

void free_tables(krepp_builder* b)
{
  if (b->geom->device == KREPP_DEVICE_NONE) return;
  cudaSetDevice(b->geom->device);
  for (auto& t : b->tables) if (t.keys) cudaFree(t.keys);
  b->tables.clear();
  b->scratch.release();
}

} // namespace

extern "C" int krepp_builder_create(const krepp_index_t* geom, const char* nwk_text, const char* const* names, uint32_t n_names, krepp_builder_t** out)
{
  if (!geom || !out || (n_names && !names)) return set_error(KREPP_ERR_ARG, "krepp_builder_create: null argument");
  *out = nullptr;
  if (!geom->host.is_geometry) return set_error(KREPP_ERR_ARG, "krepp_builder_create: the handle must carry an LSH geometry (krepp_geometry_open)");
  auto* b = new krepp_builder;
  b->geom = geom;
  for (uint32_t i = 0; i < n_names; ++i) b->names.emplace_back(names[i] ? names[i] : "");
  b->with_tree = nwk_text != nullptr;
  if (b->with_tree) { b->nwk_text = nwk_text; if (!b->nwk_text.empty() && b->nwk_text.back() == '\n') b->nwk_text.pop_back(); } // Tree::split_nwk drops one, ref src/phytree.cpp:91-93
  else if (b->names.empty()) { delete b; return set_error(KREPP_ERR_ARG, "krepp_builder_create: neither a guide tree nor reference names"); }
  const std::string err = b->tree.parse(b->with_tree ? b->nwk_text : generated_newick(b->names));
  if (!err.empty()) { delete b; return set_error(KREPP_ERR_IO, "%s", err.c_str()); }
  for (uint32_t rank = 0; rank < b->tree.nleaves; ++rank) {
    const std::string& nm = b->tree.name[b->tree.leaf_se[rank]];
    if (!b->leaf_by_name.emplace(nm, rank).second) {
      const int rc = set_error(KREPP_ERR_UNSUPPORTED, "the guide tree has two leaves named %s", nm.c_str()); // (nm lives in *b: the message first)
      delete b;
      return rc;
    }
  }
  b->leaf_rho.assign(b->tree.nleaves, 0.0);
  b->leaf_added.assign(b->tree.nleaves, 0);
  *out = b;
  return KREPP_OK;
}

extern "C" void krepp_builder_destroy(krepp_builder_t* b)
{
  if (!b) return;
  free_tables(b);
  delete b;
}

extern "C" uint32_t krepp_builder_leaf_rank(const krepp_builder_t* b, const char* name)
{
  if (!b || !name) return 0xFFFFFFFFu;
  const auto it = b->leaf_by_name.find(name);
  return it == b->leaf_by_name.end() ? 0xFFFFFFFFu : it->second;
}

extern "C" uint32_t krepp_builder_nleaves(const krepp_builder_t* b) { return b ? b->tree.nleaves : 0; }

extern "C" int krepp_builder_has_leaf(const krepp_builder_t* b, const char* name)
{
  if (!b || !name) return -1;
  return b->leaf_by_name.count(name) ? 1 : 0;
}

extern "C" int krepp_builder_add_genome(krepp_builder_t* b, const char* name, const char* bases, const uint64_t* offsets, uint32_t n_seqs, uint64_t* n_keys, double* rho)
{
  if (!b || !name || !bases || !offsets) return set_error(KREPP_ERR_ARG, "krepp_builder_add_genome: null argument");
  const auto it = b->leaf_by_name.find(name);
  if (it == b->leaf_by_name.end()) return set_error(KREPP_ERR_ARG, "krepp_builder_add_genome: %s is not a leaf of the build tree", name);
  if (b->leaf_added[it->second]) return set_error(KREPP_ERR_ARG, "krepp_builder_add_genome: %s was added before", name);
  krepp_builder::DevTable t;
  t.leaf = it->second;
  double est[2] = {0, 0};
  if (int rc = extract_to_device(b->geom, bases, offsets, n_seqs, &t.keys, &t.n, est, &b->scratch)) return rc;
  b->leaf_added[t.leaf] = 1;
  b->leaf_rho[t.leaf] = est[1] / est[0]; // RSeq::compute_rho ref src/rqseq.hpp:79 (0/0 = NaN for a genome without a single window, as there)
  b->tables.push_back(t);
  b->have_union = false;
  if (n_keys) *n_keys = t.n;
  if (rho) *rho = b->leaf_rho[t.leaf];
  return KREPP_OK;
}

extern "C" int krepp_builder_union(krepp_builder_t* b, uint64_t* n_kmers, uint64_t* n_sets_out)
{
  if (!b) return set_error(KREPP_ERR_ARG, "krepp_builder_union: null argument");
  if (b->geom->device == KREPP_DEVICE_NONE) return set_error(KREPP_ERR_CUDA, "the index-side kernels need a handle opened on a GPU (there is no CPU fallback)");
  if (cudaSetDevice(b->geom->device) != cudaSuccess) return set_error(KREPP_ERR_CUDA, "cudaSetDevice(%d) failed", b->geom->device);
  uint64_t N = 0;
  for (const auto& t : b->tables) N += t.n;
  if (!N) return set_error(KREPP_ERR_ARG, "No k-mers to index!"); // ref src/krepp.cpp:183
  if (N > 0x7FFFFFFFull) return set_error(KREPP_ERR_CAPACITY, "%llu (k-mer, reference) pairs: one build holds fewer than 2^31; split the library by LSH residue (-m / -r)", (unsigned long long)N);
  std::sort(b->tables.begin(), b->tables.end(), [](const krepp_builder::DevTable& x, const krepp_builder::DevTable& y) { return x.leaf < y.leaf; });
  const HostIndex& h = b->geom->host;
  int rc = KREPP_OK;
  const int grid = std::max(1, b->geom->sms) * 8, block = 256;
  unsigned long long *d_k0 = nullptr, *d_k1 = nullptr, *d_run_key = nullptr, *d_hash = nullptr, *d_hash_s = nullptr, *d_set_len = nullptr, *d_set_begin = nullptr;
  uint32_t *d_v0 = nullptr, *d_v1 = nullptr, *d_head = nullptr, *d_run_no = nullptr, *d_run_start = nullptr, *d_order = nullptr, *d_order_s = nullptr, *d_ghead = nullptr,
           *d_group_no = nullptr, *d_rep = nullptr, *d_set_of = nullptr, *d_clash = nullptr, *d_set_leaves = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0, need = 0;
  uint32_t n_runs = 0, n_sets = 0, clash = 0;
  unsigned long long total_leaves = 0, salt = 0;
  if (const char* env = getenv("KREPP_COLOUR_SALT")) salt = strtoull(env, nullptr, 0);
  uint32_t row_bits = 1;
  while ((1ull << row_bits) < h.nrows) ++row_bits;
  {
    // 1. tables side by side in leaf order
    B_CU(cudaMalloc(&d_k0, 8 * N)); B_CU(cudaMalloc(&d_v0, 4 * N)); B_CU(cudaMalloc(&d_k1, 8 * N)); B_CU(cudaMalloc(&d_v1, 4 * N));
    uint64_t at = 0;
    for (const auto& t : b->tables) {
      if (!t.n) continue;
      B_CU(cudaMemcpyAsync(d_k0 + at, t.keys, 8 * t.n, cudaMemcpyDeviceToDevice, 0));
      fill_leaf_kernel<<<grid, block>>>(d_v0 + at, t.n, t.leaf);
      at += t.n;
    }
    B_CU(cudaGetLastError());
    // 2. the union
    B_CU(cub::DeviceRadixSort::SortPairs(nullptr, need, d_k0, d_k1, d_v0, d_v1, (int)N, 0, 32 + (int)row_bits));
    tmp_bytes = need;
    B_CU(cub::DeviceScan::InclusiveSum(nullptr, need, d_v0, d_v0, (int)N));
    tmp_bytes = std::max(tmp_bytes, need);
    B_CU(cub::DeviceScan::InclusiveSum(nullptr, need, d_k0, d_k0, (int)N));
    tmp_bytes = std::max(tmp_bytes, need);
    B_CU(cudaMalloc(&d_tmp, tmp_bytes));
    need = tmp_bytes;
    B_CU(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_k0, d_k1, d_v0, d_v1, (int)N, 0, 32 + (int)row_bits));
    B_CU(cudaDeviceSynchronize());
    free_tables(b);                       // the leaf tables have been consumed
    cudaFree(d_k0); d_k0 = nullptr;
    // 3. runs (d_v0 is reused as the head flags)
    d_head = d_v0; d_v0 = nullptr;
    B_CU(cudaMalloc(&d_run_no, 4 * N));
    head_kernel<<<grid, block>>>(d_k1, N, d_head);
    B_CU(cudaGetLastError());
    need = tmp_bytes;
    B_CU(cub::DeviceScan::InclusiveSum(d_tmp, need, d_head, d_run_no, (int)N));
    B_CU(cudaMemcpy(&n_runs, d_run_no + (N - 1), 4, cudaMemcpyDeviceToHost));
    B_CU(cudaMalloc(&d_run_start, 4ull * (n_runs + 1)));
    run_start_kernel<<<grid, block>>>(d_head, d_run_no, N, d_run_start, n_runs);
    B_CU(cudaGetLastError());
    B_CU(cudaDeviceSynchronize());
    cudaFree(d_head); d_head = nullptr; cudaFree(d_run_no); d_run_no = nullptr;
    // 4. candidate sets by hash, representatives, verification
    B_CU(cudaMalloc(&d_run_key, 8ull * n_runs)); B_CU(cudaMalloc(&d_hash, 8ull * n_runs)); B_CU(cudaMalloc(&d_hash_s, 8ull * n_runs));
    B_CU(cudaMalloc(&d_order, 4ull * n_runs)); B_CU(cudaMalloc(&d_order_s, 4ull * n_runs)); B_CU(cudaMalloc(&d_ghead, 4ull * n_runs)); B_CU(cudaMalloc(&d_group_no, 4ull * n_runs));
    B_CU(cudaMalloc(&d_set_of, 4ull * n_runs)); B_CU(cudaMalloc(&d_clash, 4)); B_CU(cudaMemset(d_clash, 0, 4));
    set_hash_kernel<<<grid, block>>>(d_k1, d_v1, d_run_start, n_runs, salt, d_run_key, d_hash, d_order);
    B_CU(cudaGetLastError());
    need = tmp_bytes; // (n_runs <= N and the pairs are the same width: the first query covers it)
    B_CU(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_hash, d_hash_s, d_order, d_order_s, (int)n_runs));
    group_head_kernel<<<grid, block>>>(d_hash_s, n_runs, d_ghead);
    B_CU(cudaGetLastError());
    need = tmp_bytes;
    B_CU(cub::DeviceScan::InclusiveSum(d_tmp, need, d_ghead, d_group_no, (int)n_runs));
    B_CU(cudaMemcpy(&n_sets, d_group_no + (n_runs - 1), 4, cudaMemcpyDeviceToHost));
    B_CU(cudaMalloc(&d_rep, 4ull * n_sets)); B_CU(cudaMalloc(&d_set_len, 8ull * (n_sets + 1))); B_CU(cudaMalloc(&d_set_begin, 8ull * (n_sets + 1)));
    B_CU(cudaMemset(d_set_len, 0, 8ull * (n_sets + 1)));
    set_rep_kernel<<<grid, block>>>(d_ghead, d_group_no, d_order_s, d_run_start, n_runs, d_rep, d_set_len);
    B_CU(cudaGetLastError());
    set_assign_kernel<<<grid, block>>>(d_group_no, d_order_s, d_rep, d_run_start, d_v1, n_runs, d_set_of, d_clash);
    B_CU(cudaGetLastError());
    B_CU(cudaMemcpy(&clash, d_clash, 4, cudaMemcpyDeviceToHost));
    if (clash) { rc = set_error(KREPP_ERR_UNSUPPORTED, "%u k-mers have a reference set whose 64-bit sum equals that of a different set; build again with another KREPP_COLOUR_SALT", clash); goto done; }
    // 5. the distinct sets
    need = tmp_bytes;
    B_CU(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_set_len, d_set_begin, (int)(n_sets + 1)));
    B_CU(cudaMemcpy(&total_leaves, d_set_begin + n_sets, 8, cudaMemcpyDeviceToHost));
    B_CU(cudaMalloc(&d_set_leaves, 4ull * std::max<unsigned long long>(total_leaves, 1)));
    set_gather_kernel<<<grid, block>>>(d_rep, d_set_begin, d_run_start, d_v1, n_sets, d_set_leaves);
    B_CU(cudaGetLastError());
    b->keys.resize(n_runs); b->set_of.resize(n_runs); b->set_begin.resize((size_t)n_sets + 1); b->set_leaves.resize(total_leaves);
    B_CU(cudaMemcpy(b->keys.data(), d_run_key, 8ull * n_runs, cudaMemcpyDeviceToHost));
    B_CU(cudaMemcpy(b->set_of.data(), d_set_of, 4ull * n_runs, cudaMemcpyDeviceToHost));
    B_CU(cudaMemcpy(b->set_begin.data(), d_set_begin, 8ull * (n_sets + 1), cudaMemcpyDeviceToHost));
    if (total_leaves) B_CU(cudaMemcpy(b->set_leaves.data(), d_set_leaves, 4ull * total_leaves, cudaMemcpyDeviceToHost));
    b->have_union = true;
    if (n_kmers) *n_kmers = n_runs;
    if (n_sets_out) *n_sets_out = n_sets;
  }
done:
  for (void* p : {(void*)d_k0, (void*)d_k1, (void*)d_v0, (void*)d_v1, (void*)d_head, (void*)d_run_no, (void*)d_run_start, (void*)d_run_key, (void*)d_hash, (void*)d_hash_s,
                  (void*)d_order, (void*)d_order_s, (void*)d_ghead, (void*)d_group_no, (void*)d_rep, (void*)d_set_len, (void*)d_set_begin, (void*)d_set_of, (void*)d_clash,
                  (void*)d_set_leaves, d_tmp})
    if (p) cudaFree(p);
  return rc;
}

extern "C" int krepp_builder_set_union(krepp_builder_t* b, uint64_t n_kmers, const uint64_t* keys, const uint32_t* set_of, uint64_t n_sets, const uint64_t* set_begin,
                                       const uint32_t* set_leaves, const double* leaf_rho)
{
  if (!b || (n_kmers && (!keys || !set_of)) || !set_begin || (n_sets && set_begin[n_sets] && !set_leaves)) return set_error(KREPP_ERR_ARG, "krepp_builder_set_union: null argument");
  b->keys.assign(keys, keys + n_kmers);
  b->set_of.assign(set_of, set_of + n_kmers);
  b->set_begin.assign(set_begin, set_begin + n_sets + 1);
  b->set_leaves.assign(set_leaves, set_leaves + set_begin[n_sets]);
  if (leaf_rho) b->leaf_rho.assign(leaf_rho, leaf_rho + b->tree.nleaves);
  b->have_union = true;
  return KREPP_OK;
}

extern "C" int krepp_builder_write(krepp_builder_t* b, const char* index_dir, uint32_t seed, uint64_t* n_kmers, uint32_t* n_subsets)
{
  if (!b || !index_dir) return set_error(KREPP_ERR_ARG, "krepp_builder_write: null argument");
  const std::string err = write_library(*b, index_dir, seed, n_kmers, n_subsets);
  if (!err.empty()) return set_error(err.rfind("Failed to", 0) == 0 ? KREPP_ERR_IO : KREPP_ERR_ARG, "%s", err.c_str());
  return KREPP_OK;
}

// Host-side image of an on-disk krepp index (the files written by `krepp index`), flattened for upload to HBM.
//
// Format (little-endian, packed, one suffix "-m{m}r{r}-{frac|no_frac}" per partial index):
//   metadata-*  u8 k, u8 w, u8 h, u32 m, u32 r, u8 frac, u32 nrows, u8 ppos[h] (descending), u8 npos[k-h] (ascending)
//               (ref src/krepp.cpp:18-29, src/index.cpp:58-72)
//   cmer-*      u64 nkmers, nkmers x {u32 enc, u32 se}                      (ref src/table.cpp:65-70)
//   inc-*       u32 nrows, nrows x u64 cumulative bucket end               (ref src/table.cpp:71-74)
//   crecord-*   u32 nnodes, u32 nsubsets, nsubsets x {u32,u32}, nnodes x f64 rho   (ref src/record.cpp:203-211)
//   tree-*      Newick text                                                 (ref src/phytree.cpp:394-404)
#pragma once
#include <cstdint>
#include <new>
#include <string>
#include <vector>

namespace krepp {

// Allocator whose resize() leaves trivially constructible elements uninitialised: the k-mer table is gigabytes that are
// overwritten from the file right away, and zero-filling them first costs as much as reading them.
template <class T>
struct NoInitAlloc {
  using value_type = T;
  NoInitAlloc() = default;
  template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(::operator new(n * sizeof(T))); }
  void deallocate(T* p, size_t) { ::operator delete(p); }
  template <class U, class... A> void construct(U* p, A&&... a)
  {
    if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U;
    else ::new (static_cast<void*>(p)) U(static_cast<A&&>(a)...);
  }
  template <class U> bool operator==(const NoInitAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const NoInitAlloc<U>&) const { return false; }
};

struct BitRun { uint8_t src, width, dst; }; // ((x >> src) & ((1<<width)-1)) << dst

// a node while a tree is being built (creation order; se is handed out when the node closes, post-order)
struct TreeNodeTmp { uint32_t se = 0, parent_tmp = 0xffffffffu, nch = 0, card = 0; bool leaf = true; double blen = 0; std::string name; std::vector<uint32_t> kids; };

struct HostTree {
  uint32_t nnodes = 0, root = 0, nleaves = 0;
  std::vector<uint32_t> parent, nchildren, card, first_child, next_sibling; // [nnodes+1], by se
  std::vector<uint32_t> eff_nchildren;   // children with an indexed reference below (ref Tree::compute_eff_nchildren src/phytree.cpp:450-473); = nchildren
                                         // unless a query tree replaced the index's own (place -t)
  void compute_logw();                   // from eff_nchildren
  std::vector<uint32_t> depth;           // ancestors of se (0 for the root)
  std::vector<uint32_t> logw;            // sum of log2(nchildren) over the proper ancestors of se when every one of them has a power-of-two
                                         // number of children (a leaf's weight at an ancestor g is then exactly 2^-(logw[leaf] - logw[g])), else 0xffffffff
  std::vector<uint32_t> subtree;         // nodes in the subtree rooted at se; post-order => it spans se in (se-subtree, se]
  std::vector<uint8_t> is_leaf;
  std::vector<double> blen;              // NaN when absent
  std::vector<std::string> name;         // "" when unlabeled
  std::vector<std::string> shown;        // node_name(se, false) of every node, built once (the writers print one per output row)
  size_t max_shown = 0;                  // longest of them
  std::vector<uint32_t> leaf_rank;       // se -> 0-based rank among leaves by ascending se (0xffffffff for non-leaves)
  std::vector<uint32_t> leaf_se;         // rank -> se
  // Parses Newick text with the reference's conventions (ref src/phytree.cpp:84-215): post-order, 1-based se.
  // Returns an empty string on success, else the error message.
  std::string parse(const std::string& newick);
  // Builds the tree of a Greengenes/GTDB style lineage file, `place -l` (ref src/phytree.cpp:320-370).
  std::string parse_lineages(const std::string& text);
  void adopt(const std::vector<TreeNodeTmp>& tmp, uint32_t root_tmp, uint32_t next_se);
  std::string node_name(uint32_t se, bool return_na) const; // ref src/phytree.hpp:133-144
  std::string jplace_newick() const;                         // ref src/phytree.cpp:47-64
};

// The h hash positions (descending) and the k - h kept positions (ascending) `krepp index` / `krepp sketch` draw for a new
// library (ref LSHF::get_random_positions src/lshf.cpp:125-147); seeded = --seed was given.
void lsh_positions(uint32_t k, uint32_t h, bool seeded, uint32_t seed, std::vector<uint8_t>& ppos, std::vector<uint8_t>& npos);

// The balanced tree the reference generates over a list of reference ids when a library has no guide tree (Node::generate_tree
// ref src/phytree.cpp:217-253: halves, the SECOND half first, every branch length 1), as Newick for HostTree::parse.
std::string generated_newick(const std::vector<std::string>& names);

struct HostIndex {
  uint32_t k = 0, w = 0, h = 0, m = 0, r = 0, frac = 0, nrows = 0;
  uint64_t nkmers = 0;
  std::vector<uint8_t> ppos, npos;
  uint64_t mask_hash_bp = 0, mask_drop_lr = 0, mask_drop_bp = 0;
  std::vector<BitRun> hash_runs, drop_runs;   // pext plans over the 2-bit (bp) k-mer word
  std::vector<int32_t> res_numer;             // [m]: 0 = residue absent, else numerator (ref src/index.cpp:144-157)
  std::vector<uint32_t> res_base;             // [m]: first row of the partial library that holds the residue (0 with one partial)
  std::vector<uint64_t, NoInitAlloc<uint64_t>> cmer; // nkmers x (enc | se<<32)
  std::vector<uint64_t> inc;                  // nrows
  std::vector<uint32_t> inc32;                // nrows, present when nkmers < 2^32 (what the device scans with)
  uint32_t cr_nnodes = 0, nsubsets = 0;
  std::vector<uint64_t> pse;                  // nsubsets x (first | second<<32)
  std::vector<double> rho;                    // cr_nnodes, already scaled by make_rho_partial (ref src/index.cpp:188-201)
  std::vector<uint8_t> kind;                  // [nsubsets]: 0 drop (null node), 1 leaf, 2 expand through pse
  std::vector<uint32_t> col_rank;             // [cr_nnodes]: leaf rank (in `tree`) of the reference whose colour id this is, 0xffffffff otherwise
  uint32_t max_expand_depth = 0;              // deepest colour DAG expansion (bounds the device stack)
  uint32_t max_colour_leaves = 0;
  // Flattened colours for the bucket-sorted pipeline: colour id se expands (the walk of ref src/query.cpp:369-387, null
  // nodes dropped, a leaf reached twice kept once) to the leaf ranks cleaf[cbeg[se] .. cbeg[se+1]), ascending.  Empty
  // when the lists would exceed kMaxFlatLeaves entries; the fused kernel, which walks the DAG on the device, is used then.
  std::vector<uint32_t> cbeg, cleaf;
  static constexpr uint64_t kMaxFlatLeaves = 1ull << 30;
  double mean_bucket = 0, size_biased_bucket = 0;
  HostTree tree;
  bool is_geometry = false;                   // LSH geometry only (krepp_geometry_open): what the index-side kernels need, no table and no tree
  bool is_sketch = false;                     // loaded from the sketch file of one genome (`krepp sketch`): one reference, one leaf; queried by `krepp seek`
  bool wbackbone = true;                      // false: no tree-* file, the tree was generated from reflist-* (dist only; ref src/krepp.cpp:59-63)
  // Bucket-range shard held by this image (SURVEY.md 8e mode B): rows [row0, row1) of the table, entries [ent0, ent0 +
  // cmer.size()) of cmer-*; `inc32` then holds row1 - row0 ends relative to ent0.  One shard = the whole table.
  uint32_t shard = 0, nshards = 1, row0 = 0, row1 = 0;
  uint64_t ent0 = 0;
  std::vector<uint32_t> row_splits;           // [nshards + 1] first row of every shard; equal cmer bytes per shard
  std::string set_masks();                    // masks and pext plans from k, h, ppos, npos
  // A handle that carries only the LSH geometry of an index or sketch still to be built (ref BaseLSH src/krepp.hpp:28-98).
  std::string set_geometry(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, bool frac, const std::vector<uint8_t>& ppos, const std::vector<uint8_t>& npos);
  // Returns "" on success, else an error message (the reference's wording where it has one).
  // with_table = false: everything but the k-mer table itself (cmer stays empty) -- enough to plan a sharding (plan_shards).
  // qtree_path (place -t, ref src/krepp.cpp:48-64 ensure_backbone, src/phytree.cpp:421-448 map_to_qtree): a Newick file whose tree
  // replaces the index's own for everything after the colour expansion -- references are matched by leaf name, references the
  // query tree does not have are dropped, and every node number (records, placements, jplace tree) is the query tree's.
  std::string load(const std::string& dir, uint32_t shard = 0, uint32_t nshards = 1, bool with_table = true, const std::string& qtree_path = "", bool lineages = false);
  // Device bytes of the parts every shard replicates (colour record, flattened colour lists, tree, hash tables), and of the
  // largest shard's slice of the table when it is split into n bucket-range shards (the split rule of load()).
  uint64_t replicated_device_bytes() const;
  uint64_t shard_table_device_bytes(uint32_t n) const;
};

} // namespace krepp

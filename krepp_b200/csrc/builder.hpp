// The library builder behind krepp_builder_* (include/krepp_b200.h): `krepp index` = IndexMultiple::build_index / save_index
// (ref src/krepp.cpp:164-309).  builder.cu holds the device stage (leaf tables in HBM, the union as one sort, the distinct
// reference sets); library_writer.cpp holds the host stage (sets -> colour DAG along the guide tree, the files).
#pragma once
#include "handles.hpp"

#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace krepp {
// Device buffers of the minimizer pass kept between the genomes of one build (a genome's pass is a few milliseconds; allocating
// and freeing its ten buffers every time costs as much again).
struct MzScratch {
  static constexpr int kSlots = 10;
  void* p[kSlots] = {};
  size_t cap[kSlots] = {};
  void release();
};
} // namespace krepp

struct krepp_builder {
  const krepp_index* geom = nullptr;
  krepp::MzScratch scratch;
  krepp::HostTree tree;              // the build tree: the guide tree, or the balanced tree generated over the names
  std::string nwk_text;              // guide tree text as given (written to tree-* verbatim); unused without one
  bool with_tree = false;
  std::vector<std::string> names;    // input_map.tsv order (reflist-*)
  std::unordered_map<std::string, uint32_t> leaf_by_name; // leaf name -> leaf rank
  std::vector<double> leaf_rho;      // by leaf rank; 0 where no genome was added (the reference's sh_to_rho default)
  std::vector<uint8_t> leaf_added;
  struct DevTable { unsigned long long* keys = nullptr; uint64_t n = 0; uint32_t leaf = 0; };
  std::vector<DevTable> tables;      // leaf tables resident in HBM (device of `geom`)
  // the union: one entry per distinct k-mer, ascending by key (row << 32 | encoding)
  bool have_union = false;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> set_of;      // per k-mer: which distinct reference set
  std::vector<uint64_t> set_begin;   // [n_sets + 1]
  std::vector<uint32_t> set_leaves;  // leaf ranks, ascending within a set
};

namespace krepp {

// minimizer.cu: one genome's leaf table left on the device (exact-size allocation owned by the caller) and the two HyperLogLog
// estimates whose ratio is rho.
int extract_to_device(const krepp_index* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, unsigned long long** d_keys, uint64_t* n_keys,
                      double est[2], MzScratch* scratch);

// The colour record of a library: ids 1..tree.nnodes are the tree's nodes, the ids above them the reference sets that are not
// whole subtrees; pse[id] = the two ids a colour splits into (first | second << 32), (0, self) for a leaf.
struct ColourRecord {
  std::vector<uint64_t> pse;          // [nsubsets], [0] = (0, 0)
  std::vector<uint32_t> set_colour;   // per distinct reference set: its colour id
};
// Decomposes every set along the tree (see library_writer.cpp).  Sets are CSR over leaf ranks, ascending within a set.
std::string colour_sets(const HostTree& tree, uint64_t n_sets, const uint64_t* set_begin, const uint32_t* set_leaves, ColourRecord* out);

// Host stage of the build: colours, table, the seven files.  Returns "" or the error message.
std::string write_library(const krepp_builder& b, const std::string& index_dir, uint32_t seed, uint64_t* n_kmers, uint32_t* n_subsets);

} // namespace krepp

// oracle/shim/gpubuilder.hpp -- TEST INFRASTRUCTURE: the index-side binding of INTEGRATION.md section 6, compiled for real.
//
// `krepp index` of the reference with IndexMultiple::build_index / save_index (src/krepp.cpp:164-246) replaced by the library
// builder of include/krepp_b200.h.  Everything around them runs as it is in the reference: CLI11 parsing and validation, the LSH
// position draw (BaseLSH::set_lshf -- its positions are handed to the builder), read_input_file, obtain_build_tree, and the
// genomes are read by the reference's own RSeq / kseq reader (src/rqseq.cpp:12-49, src/rqseq.hpp:66-86).  oracle/shim/patch_krepp.py
// redirects the two calls in main (src/krepp.cpp:729,732); tests/test_gpu_shim.py compares the library this binary writes with
// the stock reference binary's.
#pragma once
#include "krepp_b200.h"
#include "krepp.hpp"

struct GpuLibraryBuild {
  krepp_index_t* geom = nullptr;
  krepp_builder_t* b = nullptr;
};

inline GpuLibraryBuild& gpu_library_build()
{
  static GpuLibraryBuild g;
  return g;
}

// in place of IndexMultiple::build_index (src/krepp.cpp:164-185): leaf tables + rho per genome and the union, on the GPU
inline void gpu_build_index(IndexMultiple& im)
{
  GpuLibraryBuild& g = gpu_library_build();
  const vec<uint8_t> ppos = im.lshf->get_ppos();
  if (krepp_geometry_open_positions(im.k, im.w, im.h, im.m, im.r, im.frac ? 1 : 0, ppos.data(), 0, &g.geom)) error_exit(krepp_last_error());
  std::vector<const char*> names;
  for (auto& n : im.names_v) names.push_back(n.c_str());
  const bool with_tree = !im.nwk_path.empty(); // without one the builder generates the tree of Tree::generate_tree itself
  if (krepp_builder_create(g.geom, with_tree ? im.tree->nwk_str.c_str() : nullptr, names.data(), (uint32_t)names.size(), &g.b)) error_exit(krepp_last_error());
  for (auto& [name, path] : im.name_to_path) {
    if (krepp_builder_has_leaf(g.b, name.c_str()) != 1) continue; // build_for_subtree only visits the tree's leaves (src/krepp.cpp:248-252)
    RSeq rs(path, im.lshf, im.w, im.r, im.frac, im.sdust_t, im.sdust_w);
    std::string bases;
    std::vector<uint64_t> offsets{0};
    while (rs.read_next_seq()) { // DynHT::fill_table's loop (src/table.cpp:250-255); sequences shorter than w are skipped by the builder too
      rs.set_curr_seq();
      bases.append(rs.seq, rs.len);
      offsets.push_back(bases.size());
    }
    if (bases.empty()) bases.push_back('N');
    if (krepp_builder_add_genome(g.b, name.c_str(), bases.data(), offsets.data(), (uint32_t)offsets.size() - 1, nullptr, nullptr)) error_exit(krepp_last_error());
  }
  uint64_t nk = 0, nsets = 0;
  if (krepp_builder_union(g.b, &nk, &nsets)) error_exit(krepp_last_error());
}

// in place of IndexMultiple::save_index (src/krepp.cpp:206-246)
inline void gpu_save_index(IndexMultiple& im)
{
  GpuLibraryBuild& g = gpu_library_build();
  uint64_t nk = 0;
  uint32_t nsub = 0;
  if (krepp_builder_write(g.b, im.index_dir.c_str(), seed, &nk, &nsub)) error_exit(krepp_last_error());
  if (im.nwk_path.empty()) std::cerr << "Skipped saving a backbone for the index!" << std::endl;
  krepp_builder_destroy(g.b);
  krepp_index_close(g.geom);
}

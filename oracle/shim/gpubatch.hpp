// oracle/shim/gpubatch.hpp -- TEST INFRASTRUCTURE: the binding INTEGRATION.md section 2 describes, compiled for real.
//
// A stand-in for the reference's IBatch (src/query.hpp:46-97) with the same constructor arguments and the same entry points
// (estimate_distances, place_sequences, get_summary), implemented over the C ABI of include/krepp_b200.h.  oracle/Makefile
// (`make shim`) builds oracle/_ref/krepp_gpu from the reference's own sources with the two
// `std::make_shared<IBatch>(...)` calls of src/krepp.cpp (:367, :460) redirected here by oracle/shim/patch_krepp.py -- nothing
// else of the reference changes: its CLI parsing, QSeq reader, OpenMP tasks, output framing and jplace writer all run as they
// are.  tests/test_gpu_shim.py diffs that binary's output against the stock reference binary.
#pragma once
#include "krepp_b200.h"
#include "query.hpp"

class GpuIndex
{ // one per process, beside the reference's Index (which still supplies the tree for names and the summary map)
public:
  explicit GpuIndex(const std::string& dir, int device = 0)
  {
    if (krepp_index_open(dir.c_str(), device, &ix)) error_exit(krepp_last_error()); // src/common.cpp:20-24 convention
  }
  ~GpuIndex() { krepp_index_close(ix); }
  krepp_index_t* ix = nullptr;
};

inline GpuIndex& gpu_index(const std::string& dir)
{
  static GpuIndex gi(dir);
  return gi;
}

class GpuBatch
{
public:
  GpuBatch(GpuIndex& gi, index_sptr_t index, qseq_sptr_t qs, uint32_t hdist_th, double chisq_value, double dist_max, uint32_t tau, bool no_filter,
           bool multi, bool summarize, bool place)
    : gi(gi)
    , index(index)
    , summarize(summarize)
  {
    krepp_params_default(&p, place);
    p.hdist_th = hdist_th; p.chisq = chisq_value; p.dist_max = dist_max; p.tau = tau;
    p.no_filter = no_filter; p.multi = multi; p.summarize = summarize;
    std::swap(qs->seq_batch, seq_batch); // IBatch steals the batch the same way (src/query.cpp:32-33)
    std::swap(qs->identifer_batch, identifer_batch);
    uint64_t nb = 0;
    offsets.push_back(0);
    for (auto& s : seq_batch) { bases += s; nb += s.size(); offsets.push_back(nb); }
    for (auto& n : identifer_batch) { name_off.push_back(names.size()); names += n; names.push_back('\0'); }
    if (name_off.empty()) name_off.push_back(0);
    if (krepp_batch_create(gi.ix, &p, seq_batch.size() ? seq_batch.size() : 1, nb ? nb : 1, &b)) error_exit(krepp_last_error());
  }
  ~GpuBatch() { krepp_batch_destroy(b); }
  void estimate_distances(strstream& out) { run(out, false, false); }             // src/query.cpp:141-156
  void place_sequences(strstream& out, bool tabular) { run(out, true, tabular); } // src/query.cpp:198-216
  const parallel_flat_phmap<node_sptr_t, double>& get_summary() { return node_to_wcount; } // src/query.hpp:64

private:
  void run(strstream& out, bool place, bool tabular)
  {
    krepp_results_t res;
    if (krepp_batch_submit(b, bases.data(), offsets.data(), seq_batch.size()) || krepp_batch_wait(b, &res)) error_exit(krepp_last_error());
    krepp_index_info_t info;
    krepp_index_info(gi.ix, &info);
    std::vector<double> w(summarize ? info.nnodes + 1 : 0, 0.0);
    int prev = 0;
    std::string text(1 << 20, '\0');
    for (;;) {
      std::fill(w.begin(), w.end(), 0.0);
      prev = 0;
      const size_t n = place ? krepp_format_place(gi.ix, &p, &res, names.data(), name_off.data(), tabular, &prev, summarize ? w.data() : nullptr, text.data(), text.size())
                             : krepp_format_dist(gi.ix, &p, &res, names.data(), name_off.data(), summarize ? w.data() : nullptr, text.data(), text.size());
      if (n <= text.size()) { out.write(text.data(), n); break; }
      text.resize(n);
    }
    if (summarize) // the per-node weights under the reference's own node objects (src/query.cpp:160-171,232,298,323)
      for (uint32_t se = 1; se <= info.nnodes; ++se)
        if (w[se] != 0.0) node_to_wcount[index->get_tree()->get_node(se)] += w[se];
  }
  GpuIndex& gi;
  index_sptr_t index;
  bool summarize;
  krepp_params_t p;
  krepp_batch_t* b = nullptr;
  vec<std::string> seq_batch, identifer_batch;
  std::string bases, names;
  std::vector<uint64_t> offsets, name_off;
  parallel_flat_phmap<node_sptr_t, double> node_to_wcount = {};
};

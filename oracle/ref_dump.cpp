// oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Stage-level dump harness around the UNMODIFIED reference (compiled from /root/reference by oracle/Makefile).
// This file is our own driver: it includes the reference's headers with private/protected opened up and calls
// the reference's own objects (Index, QSeq, IBatch, IMers, Minfo, LSHF) so that every number printed below is
// produced by reference code.  The only logic restated here is *visiting order*: the reference iterates
// pointer-keyed hash maps (src/query.hpp:43,95), which makes `closest` ties and `place` output nondeterministic
// (SURVEY section 0, fact 6).  The "D"/"P" lines therefore re-walk src/query.cpp:96-139 and :218-333 in a fixed
// order (forward-strand leaves by ascending se, then reverse-strand leaves by ascending se; ancestors by ascending
// se), calling the reference's Minfo::optimize_likelihood / likelihood_ratio / add for all arithmetic.  The "H"
// lines come from the reference's own summarize_matches (hash order) and are used to check untied reads.
//
// Output (text, one record per line, doubles as %.17g):
//   I k h m nrows nnodes mask_hash_bp mask_drop_lr
//   R idx name len onmers wn_or wn_rc filt_or filt_rc        filt_* = raw per-strand min Hamming distance (0xffffffff: none)
//   L idx strand pos rix enc32                                (--lookups) every eligible lookup of the read
//   M idx strand leaf_se match hdist_min rho hist[0..th]      per-(strand, leaf) histogram before summarize
//   H idx leaf_se d v                                         reference summarize_matches (hash order), node_to_minfo
//   C idx closest_se                                          reference closest (hash-order ties)
//   D idx leaf_se strand d v chisq is_closest                 deterministic-order summarize
//   P idx se edge d v chisq lwr pendant distal                (--place) deterministic-order placement candidates
#include <bits/stdc++.h>
#define private public
#define protected public
#include "query.hpp"
#undef private
#undef protected
#include <boost/math/tools/minima.hpp>

static std::vector<std::string> find_suffixes(const std::string& dir)
{
  std::set<std::string> s;
  for (const auto& e : std::filesystem::directory_iterator(dir)) {
    std::string fn = e.path().filename();
    if (fn.rfind("metadata-", 0) == 0 && e.path().extension().empty()) s.insert(fn.substr(8));
  }
  return std::vector<std::string>(s.begin(), s.end());
}

struct DetEntry {
  node_sptr_t nd;
  minfo_sptr_t mi;
  int strand;
};

static std::vector<std::pair<node_sptr_t, minfo_sptr_t>> by_se(parallel_flat_phmap<node_sptr_t, minfo_sptr_t>& m)
{
  std::vector<std::pair<node_sptr_t, minfo_sptr_t>> v(m.begin(), m.end());
  std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first->get_se() < b.first->get_se(); });
  return v;
}

int main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: ref_dump INDEX_DIR QUERY [--hdist-th N] [--tau N] [--chisq X] [--lookups] [--place] [--no-filter]\n");
    return 2;
  }
  std::string dir = argv[1], query = argv[2];
  uint32_t hdist_th = 4, tau = 2;
  double chisq_value = 2.706;
  bool lookups = false, place = false, place_no_filter = false;
  for (int a = 3; a < argc; ++a) {
    std::string s = argv[a];
    if (s == "--hdist-th") hdist_th = atoi(argv[++a]);
    else if (s == "--tau") tau = atoi(argv[++a]);
    else if (s == "--chisq") chisq_value = atof(argv[++a]);
    else if (s == "--lookups") lookups = true;
    else if (s == "--place") place = true;
    else if (s == "--no-filter") place_no_filter = true;
  }
  auto index = std::make_shared<Index>(dir);
  for (auto& sfx : find_suffixes(dir)) {
    if (std::filesystem::exists(std::filesystem::path(dir) / ("tree" + sfx))) index->load_partial_tree(sfx);
    else index->generate_partial_tree(sfx);
    index->load_partial_index(sfx);
  }
  index->make_rho_partial();
  auto lshf = index->get_lshf();
  auto tree = index->get_tree();
  printf("I %u %u %u %u %u %016lx %016lx\n", (unsigned)lshf->k, (unsigned)lshf->h, lshf->m, index->nrows, tree->get_nnodes(),
         lshf->mask_hash_bp, lshf->mask_drop_lr);
  const double nan = std::numeric_limits<double>::quiet_NaN();
  auto qs = std::make_shared<QSeq>(query);
  uint64_t base = 0;
  while (qs->read_next_batch() || !qs->is_batch_finished()) {
    // dist defaults: no_filter=true, multi=true (src/krepp.cpp:632-654)
    IBatch ib(index, qs, hdist_th, chisq_value, nan, tau, true, true, false);
    uint32_t k = ib.k;
    for (uint64_t b = 0; b < ib.batch_size; ++b) {
      const std::string& s = ib.seq_batch[b];
      uint64_t idx = base + b, len = s.size();
      auto orr = std::make_shared<IMers>(index, len, hdist_th), rc = std::make_shared<IMers>(index, len, hdist_th);
      ib.search_mers(s.data(), len, orr, rc);
      printf("R %lu %s %lu %u %u %u %u %u\n", idx, ib.identifer_batch[b].c_str(), len, ib.onmers, ib.wnmers_or, ib.wnmers_rc,
             orr->hdist_filt, rc->hdist_filt);
      if (lookups) {
        // same primitives, same order as src/query.cpp:49-93, only printing
        uint32_t i, l;
        uint64_t bp = 0, lr = 0, rcbp;
        for (i = l = 0; i < len;) {
          if (seq_nt4_table[(unsigned char)s[i]] >= 4) { l = 0, i++; continue; }
          l++, i++;
          if (l < k) continue;
          if (l == k) compute_encoding(s.data() + i - k, s.data() + i, lr, bp);
          else update_encoding(s.data() + i - 1, lr, bp);
          bp &= ib.mask_bp; lr &= ib.mask_lr;
          rcbp = revcomp_bp64(bp, k);
          uint32_t rix = lshf->compute_hash(bp);
          if (index->check_partial(rix)) printf("L %lu 0 %u %u %u\n", idx, (uint32_t)(i - k), rix, lshf->drop_ppos_lr(lr));
          rix = lshf->compute_hash(rcbp);
          if (index->check_partial(rix)) printf("L %lu 1 %u %u %u\n", idx, (uint32_t)(len - i), rix, lshf->drop_ppos_lr(conv_bp64_lr64(rcbp)));
        }
      }
      for (int st = 0; st < 2; ++st) {
        for (auto& [nd, mi] : by_se(st ? rc->leaf_to_minfo : orr->leaf_to_minfo)) {
          printf("M %lu %d %u %.17g %u %.17g", idx, st, nd->get_se(), mi->match_count, mi->hdist_min, mi->rho);
          for (double x : mi->hdisthist_v) printf(" %.17g", x);
          printf("\n");
        }
      }
      // ---- deterministic-order restatement of src/query.cpp:96-139 on a second, untouched set of IMers
      auto dor = std::make_shared<IMers>(index, len, hdist_th), drc = std::make_shared<IMers>(index, len, hdist_th);
      ib.search_mers(s.data(), len, dor, drc);
      uint32_t onmers = ib.onmers;
      // ---- reference summarize (hash order)
      ib.summarize_matches(orr, rc);
      for (auto& [nd, mi] : by_se(ib.node_to_minfo)) printf("H %lu %u %.17g %.17g\n", idx, nd->get_se(), mi->d_llh, mi->v_llh);
      printf("C %lu %u\n", idx, ib.nd_closest == tree->get_root() ? 0u : ib.nd_closest->get_se());

      uint32_t f_or = 2 * dor->hdist_filt + 1, f_rc = 2 * drc->hdist_filt + 1;
      std::map<uint32_t, DetEntry> sel; // keyed by leaf se
      node_sptr_t nd_cl = nullptr;
      minfo_sptr_t mi_cl = std::make_shared<Minfo>(hdist_th);
      for (auto& [nd, mi] : by_se(dor->leaf_to_minfo)) {
        mi->mismatch_count = onmers - mi->match_count;
        if (mi->hdist_min > f_or) continue;
        mi->optimize_likelihood(ib.llhfunc);
        if (mi->d_llh <= mi_cl->d_llh) { nd_cl = nd; mi_cl = mi; }
        sel[nd->get_se()] = {nd, mi, 0};
      }
      for (auto& [nd, mi] : by_se(drc->leaf_to_minfo)) {
        mi->mismatch_count = onmers - mi->match_count;
        if (mi->hdist_min > f_rc) continue;
        mi->optimize_likelihood(ib.llhfunc);
        if (mi->d_llh <= mi_cl->d_llh) { nd_cl = nd; mi_cl = mi; }
        sel[nd->get_se()] = {nd, mi, 1};
        if (dor->leaf_to_minfo.contains(nd)) {
          minfo_sptr_t mo = dor->leaf_to_minfo[nd];
          if ((mi->d_llh > mo->d_llh) || ((mi->d_llh == mo->d_llh) && (mi->match_count < mo->match_count))) sel[nd->get_se()] = {nd, mo, 0};
        }
      }
      if (nd_cl) {
        int st = (drc->leaf_to_minfo.contains(nd_cl) && drc->leaf_to_minfo[nd_cl] == mi_cl) ? 1 : 0;
        sel[nd_cl->get_se()] = {nd_cl, mi_cl, st};
      }
      for (auto& [se, e] : sel) {
        double chi = mi_cl->likelihood_ratio(e.mi->d_llh, ib.llhfunc);
        printf("D %lu %u %d %.17g %.17g %.17g %d\n", idx, se, e.strand, e.mi->d_llh, e.mi->v_llh, chi, e.nd == nd_cl ? 1 : 0);
      }
      if (!place) continue;
      // ---- deterministic-order restatement of src/query.cpp:218-333 (multi=true, jplace fields)
      bool no_filter = place_no_filter;
      if (sel.empty() || !(no_filter || (mi_cl->get_leq_tau(tau) > 1.0))) continue;
      mi_cl->chisq = 0;
      auto emit = [&](node_sptr_t nd, minfo_sptr_t mi) {
        printf("P %lu %u %u %.17g %.17g %.17g %.17g %.17g %.17g\n", idx, nd->get_se(), nd->get_en(), mi->d_llh, mi->v_llh, mi->chisq,
               mi->lwr, mi->jukes_cantor_dist() - nd->get_midpoint_pendant(), nd->get_midpoint_pendant());
      };
      if (sel.size() == 1) { emit(nd_cl, mi_cl); continue; }
      std::map<uint32_t, std::pair<node_sptr_t, minfo_sptr_t>> pp; // ascending se
      for (auto& [se, e] : sel) {
        pp[se] = {e.nd, e.mi};
        double denom = 1.0;
        node_sptr_t par = e.nd;
        while ((par = par->get_parent())) {
          if (par->check_taxon() && e.nd->check_taxon()) denom = 1.0;
          else denom /= par->get_eff_nchildren();
          if (!pp.count(par->get_se())) pp[par->get_se()] = {par, std::make_shared<Minfo>(hdist_th)};
          pp[par->get_se()].second->add(e.mi, denom);
        }
      }
      std::vector<std::pair<node_sptr_t, minfo_sptr_t>> cand;
      for (auto& [se, pr] : pp) {
        auto& [nd, mi] = pr;
        if (nd->get_nchildren() != nd->get_eff_nchildren() || nd->get_nchildren() == 1) continue;
        if (no_filter || (mi->get_leq_tau(tau) > 1.0)) {
          if (!nd->check_leaf()) mi->optimize_likelihood(ib.llhfunc);
          mi->chisq = mi_cl->likelihood_ratio(mi->d_llh, ib.llhfunc);
          if ((mi->chisq < chisq_value) && nd->get_parent()) cand.push_back(pr);
        }
      }
      double total = 0;
      for (auto& [nd, mi] : cand) { mi->lwr = exp(-mi->chisq / 2); total = total + mi->lwr; }
      for (auto& [nd, mi] : cand) { mi->lwr = mi->lwr / total; emit(nd, mi); }
    }
    base += ib.batch_size;
  }
  return 0;
}

// library_writer.cpp -- host stage of `krepp index`: from the union of the leaf tables (one entry
// per distinct k-mer with the set of references that hold it, builder.cu) to the files of IndexMultiple::save_index
// (ref src/krepp.cpp:206-246).
//
// Colours.  The reference unions leaf tables up the guide tree (IndexMultiple::build_for_subtree ref src/krepp.cpp:248-309,
// DynHT::union_row ref src/table.cpp:214-234) and names the set of a k-mer met in two children by the SUM of the children's
// 64-bit set hashes (Record::add_subset ref src/record.cpp:82-113); a tree node's hash is the sum over its children
// (Node::add_children ref src/phytree.hpp:107-116), so a k-mer held by every reference below a node collapses to the node
// itself, and CRecord (ref src/record.cpp:156-175) stores every colour as the pair of colours it was summed from.  The result
// is a decomposition of each reference set along the tree.  Here the same decomposition is computed directly, per DISTINCT
// set and exactly (pairs are interned by value, no hash sums and so no nonce rule):
//     colour(S) at the lowest tree node g above S:  g itself when S is every leaf below g; else the colours of S restricted to
//     each child of g that S touches, folded from the right into pairs  (p1, (p2, (... , pn))).
// A node with more than two children gets the same right fold over its children as its own pair, so the tails are shared.
// (The reference gives such a node the pair (first child, sum of the others), ref src/record.cpp:172-174, which is the null id
// unless some k-mer's set is exactly the other children -- SURVEY.md 7.7, the multifurcation quirk -- and a k-mer held by every
// child gets a separate colour there, the node itself here.  The k-mers expand alike: tests/test_index_build_cpu.py checks all
// 6.9 M of the reference's own toy library, whose guide tree has multifurcations.)
// Ids: 0 = null, 1..nnodes = the tree's nodes in post-order (Record::make_compact ref src/record.cpp:132-154 numbers them
// first too), then the interned pairs in order of creation.  The reference's numbering above the nodes follows a hash map's
// iteration order and differs from run to run; nothing reads more into an id than its expansion.
#include "builder.hpp"

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <sys/stat.h>
#include <thread>
#include <algorithm>

namespace krepp {

namespace {

struct PairTable { // (first, second) -> colour id, open addressing; the ids index ColourRecord::pse.  Key and id share a 16-byte
                   // slot: a probe is one cache miss, and the interning of 2.5 M sets is little else than its misses
  struct Slot { uint64_t key; uint32_t val, pad; };
  std::vector<Slot> slot;
  uint64_t mask = 0, used = 0;
  explicit PairTable(uint64_t expect)
  {
    uint64_t cap = 1024;
    while (cap < 2 * expect) cap <<= 1;
    slot.assign(cap, Slot{0, 0, 0}); mask = cap - 1;
  }
  static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
  void grow()
  {
    std::vector<Slot> s2(slot.size() * 2, Slot{0, 0, 0});
    const uint64_t m2 = s2.size() - 1;
    for (const Slot& e : slot) {
      if (!e.val) continue;
      uint64_t at = mix(e.key) & m2;
      while (s2[at].val) at = (at + 1) & m2;
      s2[at] = e;
    }
    slot.swap(s2); mask = m2;
  }
  // the id of the pair, `fresh` when it is new (the caller then appends it under that id)
  uint32_t intern(uint64_t pair, uint32_t fresh, bool* is_new)
  {
    if (2 * (used + 1) > slot.size()) grow();
    uint64_t at = mix(pair) & mask;
    while (slot[at].val) {
      if (slot[at].key == pair) { *is_new = false; return slot[at].val; }
      at = (at + 1) & mask;
    }
    slot[at] = Slot{pair, fresh, 0}; ++used; *is_new = true;
    return fresh;
  }
};

} // namespace

std::string colour_sets(const HostTree& t, uint64_t n_sets, const uint64_t* set_begin, const uint32_t* set_leaves, ColourRecord* out)
{
  const uint32_t N = t.nnodes;
  if (!N) return "the build tree is empty";
  std::vector<uint32_t> upto(N + 1, 0); // leaves numbered <= se: the leaves below g are the ranks [upto[g - subtree[g]], upto[g])
  for (uint32_t se = 1; se <= N; ++se) upto[se] = upto[se - 1] + (t.is_leaf[se] ? 1u : 0u);
  std::vector<uint64_t>& pse = out->pse;
  pse.assign(N + 1, 0);
  PairTable pairs(n_sets + N);
  auto pair_of = [&](uint32_t a, uint32_t b) -> uint32_t {
    bool is_new = false;
    const uint64_t p = (uint64_t)a | (uint64_t)b << 32;
    if (pse.size() >= 0xFFFFFFFFull) return 0;
    const uint32_t id = pairs.intern(p, (uint32_t)pse.size(), &is_new);
    if (is_new) pse.push_back(p);
    return id;
  };
  std::vector<uint32_t> kids;
  for (uint32_t g = 1; g <= N; ++g) { // the tree's own colours: leaf = (0, self) (ref src/record.cpp:16,172-174), node = (first child, the others)
    if (t.is_leaf[g]) { pse[g] = (uint64_t)g << 32; continue; }
    kids.clear();
    for (uint32_t c = t.first_child[g]; c; c = t.next_sibling[c]) kids.push_back(c);
    if (kids.size() < 2) return "the build tree has a node with one child (ref src/phytree.cpp:165-167)";
    uint32_t acc = kids.back();
    for (size_t i = kids.size() - 1; i-- > 1;) acc = pair_of(kids[i], acc);
    pse[g] = (uint64_t)kids[0] | (uint64_t)acc << 32;
  }
  out->set_colour.assign(n_sets, 0);
  struct Frame { uint32_t lo, hi, g, at, child; size_t base; };
  std::vector<Frame> stack;
  std::vector<uint32_t> vals;
  auto leaves_below = [&](uint32_t g) { return upto[g] - upto[g - t.subtree[g]]; };
  for (uint64_t s = 0; s < n_sets; ++s) {
    const uint32_t* lv = set_leaves + set_begin[s];
    const uint64_t n64 = set_begin[s + 1] - set_begin[s];
    if (!n64 || n64 > t.nleaves) return "a reference set is empty or larger than the tree";
    const uint32_t n = (uint32_t)n64;
    for (uint32_t i = 0; i < n; ++i) if (lv[i] >= t.nleaves || (i && lv[i] <= lv[i - 1])) return "a reference set is not an ascending list of leaf ranks";
    stack.clear(); vals.clear();
    // enter(x, lo, hi): the segment lies below x.  Walks down while one child holds all of it; a node (or leaf) whose every leaf
    // is in the segment is its own colour; otherwise a frame opens at the lowest node above the segment.  Children are numbered
    // in ascending order (post-order), so the child that holds a leaf is the first one numbered at or past it.
    auto enter = [&](uint32_t x, uint32_t lo, uint32_t hi) {
      const uint32_t first = t.leaf_se[lv[lo]], last = t.leaf_se[lv[hi - 1]];
      for (;;) {
        if (hi - lo == leaves_below(x)) { vals.push_back(x); return; }
        uint32_t d = t.first_child[x];
        while (first > d) d = t.next_sibling[d];
        if (last <= d) { x = d; continue; }
        stack.push_back(Frame{lo, hi, x, lo, d, vals.size()});
        return;
      }
    };
    enter(t.root, 0, n);
    while (!stack.empty()) {
      Frame& f = stack.back();
      if (f.at < f.hi) { // the next child of f.g that the segment touches
        uint32_t d = f.child;
        while (t.leaf_se[lv[f.at]] > d) d = t.next_sibling[d];
        const uint32_t end_rank = upto[d];
        uint32_t j = f.at + 1;
        while (j < f.hi && lv[j] < end_rank) ++j;
        const uint32_t i = f.at;
        f.at = j; f.child = d; // (f may dangle after enter)
        enter(d, i, j);
      } else { // all parts are in vals[base..): right fold
        const size_t base = f.base;
        uint32_t acc = vals.back();
        for (size_t i = vals.size() - 1; i-- > base;) { acc = pair_of(vals[i], acc); if (!acc) return "The current se_t size is too small to fit all subsets observed!"; }
        vals.resize(base);
        vals.push_back(acc);
        stack.pop_back();
      }
    }
    out->set_colour[s] = vals[0];
  }
  return "";
}

namespace {

bool write_all(const std::string& path, const std::vector<std::pair<const void*, size_t>>& parts)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = true;
  for (const auto& p : parts) ok = ok && (!p.second || fwrite(p.first, 1, p.second, f) == p.second);
  return (fclose(f) == 0) && ok;
}

std::string pos_list(const std::vector<uint8_t>& v)
{ // vec_to_str, ref src/common.hpp:258-268
  std::string s = "[";
  for (size_t i = 0; i < v.size(); ++i) { if (i) s += ", "; s += std::to_string((int)v[i]); }
  return s + "]";
}

} // namespace

std::string write_library(const krepp_builder& b, const std::string& dir, uint32_t seed, uint64_t* n_kmers, uint32_t* n_subsets)
{
  const HostIndex& h = b.geom->host;
  if (!b.have_union) return "krepp_builder_write: no union yet (krepp_builder_union or krepp_builder_set_union first)";
  const uint64_t n = b.keys.size();
  if (!n) return "No k-mers to index!"; // ref src/krepp.cpp:183
  ColourRecord cr;
  const uint64_t n_sets = b.set_begin.empty() ? 0 : b.set_begin.size() - 1;
  const bool debug = getenv("KREPP_BUILD_DEBUG") && atoi(getenv("KREPP_BUILD_DEBUG"));
  auto clk = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    const auto now = std::chrono::steady_clock::now();
    if (debug) fprintf(stderr, "[library writer] %s %.3f s\n", what, std::chrono::duration<double>(now - clk).count());
    clk = now;
  };
  { std::string err = colour_sets(b.tree, n_sets, b.set_begin.data(), b.set_leaves.data(), &cr); if (!err.empty()) return err; }
  lap("colour record");
  // FlatHT (ref src/table.cpp:43-63): entries row by row, ascending encoding within a row, and the cumulative row ends
  std::vector<uint64_t, NoInitAlloc<uint64_t>> cmer(n);
  std::vector<uint64_t> inc(h.nrows, 0);
  { // in parallel over ranges of k-mers: entry i is the last of its row when the next key has another row, and is then the
    // cumulative end (i + 1) of its row and of every empty row up to the next key's
    const uint32_t nt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<unsigned>(std::thread::hardware_concurrency(), 32u), n / (1u << 20)));
    std::vector<int> bad(nt, 0);
    auto work = [&](uint32_t t) {
      const uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t key = b.keys[i], row = key >> 32;
        if (row >= h.nrows) { bad[t] = 1; return; }
        if (i && key <= b.keys[i - 1]) { bad[t] = 2; return; }
        if (b.set_of[i] >= n_sets) { bad[t] = 3; return; }
        cmer[i] = (key & 0xFFFFFFFFull) | (uint64_t)cr.set_colour[b.set_of[i]] << 32;
        const uint64_t next_row = i + 1 < n ? std::min<uint64_t>(b.keys[i + 1] >> 32, h.nrows) : h.nrows;
        for (uint64_t q = row; q < next_row; ++q) inc[q] = i + 1;
      }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (int e : bad) {
      if (e == 1) return "a k-mer fell outside the table";
      if (e == 2) return "the union is not ascending by (row, encoding)";
      if (e == 3) return "a k-mer names a reference set that does not exist";
    }
  }
  lap("table");
  // CRecord (ref src/record.cpp:156-175,213-219): nnodes = tree nodes + 1, rho by node number (leaves only)
  const uint32_t nnodes = b.tree.nnodes + 1, nsubsets = (uint32_t)cr.pse.size();
  std::vector<double> rho(nnodes, 0.0);
  for (uint32_t rank = 0; rank < b.tree.nleaves; ++rank) rho[b.tree.leaf_se[rank]] = b.leaf_rho[rank];

  if (mkdir(dir.c_str(), 0777) != 0 && errno != EEXIST) return "Failed to create the index directory " + dir;
  const std::string sfx = "-m" + std::to_string(h.m) + "r" + std::to_string(h.r) + (h.frac ? "-frac" : "-no_frac"); // ref src/krepp.cpp:586-589
  if (!write_all(dir + "/cmer" + sfx, {{&n, 8}, {cmer.data(), 8 * n}})) return "Failed to write the k-mer array of the index!";
  if (!write_all(dir + "/inc" + sfx, {{&h.nrows, 4}, {inc.data(), 8ull * h.nrows}})) return "Failed to write the offset array of the index!";
  if (!write_all(dir + "/crecord" + sfx, {{&nnodes, 4}, {&nsubsets, 4}, {cr.pse.data(), 8ull * nsubsets}, {rho.data(), 8ull * nnodes}}))
    return "Failed to write the color array of the index!";
  { std::string rl; for (const auto& nm : b.names) { rl += nm; rl += '\n'; }
    if (!write_all(dir + "/reflist" + sfx, {{rl.data(), rl.size()}})) return "Failed to write the reference list of the index!"; }
  if (b.with_tree) { if (!write_all(dir + "/tree" + sfx, {{b.nwk_text.data(), b.nwk_text.size()}})) return "Failed to write the backbone tree of the index!"; }
  else remove((dir + "/tree" + sfx).c_str()); // a tree file left by an earlier build of this suffix would be loaded as the backbone
  { // save_configuration, ref src/krepp.cpp:18-29
    const uint8_t k8 = (uint8_t)h.k, w8 = (uint8_t)h.w, h8 = (uint8_t)h.h, frac8 = h.frac ? 1 : 0;
    if (!write_all(dir + "/metadata" + sfx, {{&k8, 1}, {&w8, 1}, {&h8, 1}, {&h.m, 4}, {&h.r, 4}, {&frac8, 1}, {&h.nrows, 4}, {h.ppos.data(), h.ppos.size()}, {h.npos.data(), h.npos.size()}}))
      return "Failed to write the metadata of the index!";
  }
  { // save_info, ref src/krepp.cpp:187-204
    char date[64];
    std::time_t now = std::time(nullptr);
    std::strftime(date, sizeof date, "%Y-%m-%d %H:%M:%S", std::localtime(&now));
    std::string info = "krepp version: v0.8.3+b200\ndate: " + std::string(date) + "\nseed: " + std::to_string(seed) + "\nk: " + std::to_string(h.k) + "\nw: " + std::to_string(h.w) +
                       "\nh: " + std::to_string(h.h) + "\nm: " + std::to_string(h.m) + "\nfrac: " + (h.frac ? "true" : "false") + "\nppos_v: " + pos_list(h.ppos) +
                       "\nnpos_v: " + pos_list(h.npos) + "\nnrows: " + std::to_string(h.nrows) + "\ntotal_num_kmers: " + std::to_string(n) + "\nsdust-t: 0\nsdust-w: 0\n";
    if (!write_all(dir + "/metadata" + sfx + ".txt", {{info.data(), info.size()}})) return "Failed to write the text metadata of the index!";
  }
  lap("files");
  if (n_kmers) *n_kmers = n;
  if (n_subsets) *n_subsets = nsubsets;
  return "";
}

} // namespace krepp

// Host loader for the on-disk krepp index; see index_image.hpp for the format and reference citations.
#include "index_image.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <thread>
#include <unistd.h>
#include <random>
#include <unordered_map>
#include <fstream>
#include <limits>
#include <sstream>

namespace krepp {

// ------------------------------------------------------------------------------------------------ Newick

namespace {

// Splits Newick text into label/length strings and the four structural characters.  The observable conventions of
// the reference tokenizer (ref src/phytree.cpp:84-148) are kept: one trailing newline is ignored, the text must end in
// ';', quotes (' or ") protect structural characters and a doubled quote is a literal ', [...] comments are only
// recognised (and dropped) inside quotes, and a label string -- possibly empty -- is emitted in front of every ')',
// ':' and ',' unless the preceding character is '('.
struct NewickScanner {
  std::vector<std::string> items;
  std::string error;

  bool run(std::string text)
  {
    if (text.empty()) { error = "Given Newick tree seems to be empty?!?."; return false; }
    if (text.back() == '\n') text.pop_back();
    if (text.empty() || text.back() != ';') { error = "Given Newick tree ends with a character other than ';'."; return false; }
    std::string pending;
    bool in_quote = false, prev_was_quote = false, in_comment = false;
    const size_t n = text.size();
    for (size_t i = 0; i < n; ++i) {
      const char c = text[i];
      if (in_comment) { if (c == ']') in_comment = false; continue; }
      const bool is_quote = (c == '\'' || c == '"');
      if (is_quote && prev_was_quote) { in_quote = false; pending += '\''; continue; }
      prev_was_quote = is_quote;
      if (is_quote) { in_quote = !in_quote; continue; }
      if (in_quote) {
        if (c == '[') in_comment = true; else pending += c;
        continue;
      }
      switch (c) {
        case '(': items.emplace_back(1, c); break;
        case ')': case ':': case ',':
          if (i == 0 || text[i - 1] != '(') { items.push_back(pending); pending.clear(); }
          items.emplace_back(1, c);
          break;
        case '[': case ']':
          error = "Given Newick tree contains an unquoted label or length with '[' or ']'."; return false;
        case ';':
          if (i + 1 == n) { i = n; break; }
          error = (i + 1 < n && text[i + 1] == '\n') ? "Given Newick file may contain multiple trees, encountered unexpected ';'."
                                                     : "Given Newick tree contains an unquoted label or length with ';'.";
          return false;
        default:
          if ((c == ' ' || c == '\n') && !pending.empty()) {
            error = "Given Newick tree contains an unquoted label or length with ' ' or newline."; return false;
          }
          pending += c;
      }
    }
    if (!pending.empty()) items.push_back(pending);
    return true;
  }
};

struct TreeBuilder {
  const std::vector<std::string>& it;
  size_t at = 0;
  HostTree& t;
  std::string error;
  // temporary per-node storage in creation order; se is handed out when a node closes (post-order)
  using Tmp = TreeNodeTmp;
  std::vector<Tmp> tmp;
  uint32_t next_se = 0;

  TreeBuilder(const std::vector<std::string>& items, HostTree& tree) : it(items), t(tree) {}
  bool is(const char* s) const { return at < it.size() && it[at] == s; }

  void label_and_length(Tmp& nd)
  { // optional label then optional ":length" (ref src/phytree.cpp:175-187,192-203)
    nd.name.clear();
    nd.blen = std::numeric_limits<double>::quiet_NaN();
    if (at >= it.size()) return; // unlabeled root: nothing left to read
    if (!is(",")) {
      if (!is(":")) { nd.name = it[at]; ++at; }
      if (is(":")) { nd.blen = (at + 1 < it.size()) ? std::atof(it[at + 1].c_str()) : 0.0; at += 2; }
    }
  }

  uint32_t subtree()
  {
    const uint32_t id = (uint32_t)tmp.size();
    tmp.emplace_back();
    if (!error.empty() || at >= it.size()) return id;
    if (is("(")) {
      do {
        ++at;
        const uint32_t ch = subtree();
        tmp[ch].parent_tmp = id;
        tmp[id].kids.push_back(ch);
        tmp[id].card += tmp[ch].card;
        tmp[id].leaf = false;
      } while (is(","));
      if (tmp[id].kids.size() == 1) { error = "A node has a single child in the backbone tree! Please suppress unifurcations."; return id; }
      tmp[id].se = ++next_se;
      if (is(")")) { ++at; if (is(")")) return id; } // unlabeled, length-less last child keeps blen = 0 (ref :171-174)
      label_and_length(tmp[id]);
    } else {
      label_and_length(tmp[id]);
      tmp[id].leaf = true;
      tmp[id].card = 1;
      tmp[id].se = ++next_se;
    }
    return id;
  }
};

std::string fixed5(double v)
{
  char b[64];
  snprintf(b, sizeof b, "%.5f", v);
  return b;
}

} // namespace

std::string HostTree::parse(const std::string& newick)
{
  NewickScanner sc;
  if (!sc.run(newick)) return sc.error;
  TreeBuilder tb(sc.items, *this);
  const uint32_t root_tmp = tb.subtree();
  if (!tb.error.empty()) return tb.error;
  adopt(tb.tmp, root_tmp, tb.next_se);
  return "";
}

void HostTree::adopt(const std::vector<TreeNodeTmp>& tmp, uint32_t root_tmp, uint32_t next_se)
{
  nnodes = next_se;
  const size_t N = nnodes + 1;
  parent.assign(N, 0); nchildren.assign(N, 0); card.assign(N, 0); first_child.assign(N, 0); next_sibling.assign(N, 0);
  is_leaf.assign(N, 0); blen.assign(N, std::numeric_limits<double>::quiet_NaN()); name.assign(N, "");
  leaf_rank.assign(N, 0xffffffffu); leaf_se.clear();
  for (const auto& nd : tmp) {
    if (!nd.se) continue;
    parent[nd.se] = nd.parent_tmp == 0xffffffffu ? 0 : tmp[nd.parent_tmp].se;
    nchildren[nd.se] = (uint32_t)nd.kids.size();
    card[nd.se] = nd.card; is_leaf[nd.se] = nd.leaf; blen[nd.se] = nd.blen; name[nd.se] = nd.name;
    for (size_t i = 0; i < nd.kids.size(); ++i) {
      const uint32_t c = tmp[nd.kids[i]].se;
      if (i == 0) first_child[nd.se] = c;
      if (i + 1 < nd.kids.size()) next_sibling[c] = tmp[nd.kids[i + 1]].se;
    }
  }
  root = tmp[root_tmp].se;
  for (uint32_t se = 1; se <= nnodes; ++se)
    if (is_leaf[se]) { leaf_rank[se] = (uint32_t)leaf_se.size(); leaf_se.push_back(se); }
  nleaves = (uint32_t)leaf_se.size();
  subtree.assign(N, 1); subtree[0] = 0;
  for (uint32_t se = 1; se <= nnodes; ++se) if (parent[se]) subtree[parent[se]] += subtree[se]; // children precede parents
  shown.assign(N, "");
  max_shown = 0;
  for (uint32_t se = 1; se <= nnodes; ++se) { shown[se] = name[se].empty() ? std::to_string(se - 1) : name[se]; max_shown = std::max(max_shown, shown[se].size()); }
  depth.assign(N, 0);
  for (uint32_t se = nnodes; se >= 1; --se) if (parent[se]) depth[se] = depth[parent[se]] + 1;  // parents have the larger se
  eff_nchildren = nchildren;
  compute_logw();
}

std::string HostTree::parse_lineages(const std::string& text)
{
  // ref src/phytree.cpp:320-370 Tree::parse_lineages.  One reference per line, "NAME<tab>d__A; p__B; ...[<tab>ignored]": every
  // "; " becomes ";", a taxon loses each "<any character>__" it holds (the rank prefix) and is skipped when nothing remains,
  // taxa are keyed by what remains alone (the same word at two ranks is one node, its parent the first one seen), the
  // reference hangs below the last taxon of its line.  No branch lengths (NaN), unifurcations stay; numbering is post-order
  // from a node named "root".  Nodes without a parent hang below the root in the order they first appear: the reference
  // walks a hash map there, so with more than one top-level taxon its numbering is whatever that map's order is.
  using Tmp = TreeNodeTmp;
  std::vector<Tmp> tmp(1);
  const double nan = std::numeric_limits<double>::quiet_NaN();
  tmp[0].name = "root"; tmp[0].leaf = false; tmp[0].blen = nan;
  std::unordered_map<std::string, uint32_t> by_name;
  // Node::add_children (ref src/phytree.hpp:104-113) adds the child's cardinality AS IT IS when the child is hung: a taxon is hung
  // when it is created, before anything is below it, so a node's card ends up as the number of references directly below it
  // (plus, for the root, what its late-hung children had gathered) -- `--no-multi` ranks candidates by it (src/query.cpp:311-330)
  auto hang = [&](uint32_t child, uint32_t par) { tmp[child].parent_tmp = par; tmp[par].kids.push_back(child); tmp[par].leaf = false; tmp[par].card += tmp[child].card; };
  auto node = [&](const std::string& nm, bool leaf, uint32_t par) {
    const uint32_t id = (uint32_t)tmp.size();
    tmp.emplace_back();
    tmp[id].name = nm; tmp[id].leaf = leaf; tmp[id].blen = nan; tmp[id].card = leaf ? 1 : 0;
    if (par != 0xffffffffu) hang(id, par);
    by_name.emplace(nm, id);
    return id;
  };
  for (size_t at = 0, n = text.size(); at < n;) {
    size_t eol = text.find('\n', at);
    if (eol == std::string::npos) eol = n;
    std::string line;
    line.reserve(eol - at);
    for (size_t i = at; i < eol; ++i) { line += text[i]; if (text[i] == ';' && i + 1 < eol && text[i + 1] == ' ') ++i; }
    at = eol + 1;
    const size_t tab = line.find('\t');
    if (tab == std::string::npos || tab + 1 >= line.size()) return "Failed to reference to lineage mapping!";
    const std::string ref_name = line.substr(0, tab);
    const size_t tab2 = line.find('\t', tab + 1);
    const std::string lineage = line.substr(tab + 1, tab2 == std::string::npos ? std::string::npos : tab2 - tab - 1);
    uint32_t par = 0xffffffffu;
    for (size_t p = 0; p < lineage.size();) {
      size_t e = lineage.find(';', p);
      if (e == std::string::npos) e = lineage.size();
      std::string taxon;
      for (size_t i = p; i < e;) {
        if (i + 2 < e && lineage[i + 1] == '_' && lineage[i + 2] == '_' && lineage[i] != '\r') i += 3; else taxon += lineage[i++];
      }
      p = e + 1;
      if (taxon.empty()) continue;
      auto it = by_name.find(taxon);
      par = it == by_name.end() ? node(taxon, false, par) : it->second;
    }
    if (by_name.count(ref_name)) return "The same reference appears more than once in the lineage file.";
    node(ref_name, true, par);
  }
  for (uint32_t id = 1; id < tmp.size(); ++id) if (tmp[id].parent_tmp == 0xffffffffu) hang(id, 0);
  if (tmp.size() == 1) tmp[0].leaf = true;
  // post-order numbering, children in the order they were hung
  uint32_t next_se = 0;
  std::vector<std::pair<uint32_t, uint32_t>> st{{0u, 0u}};
  while (!st.empty()) {
    const uint32_t id = st.back().first, k = st.back().second;
    if (k < tmp[id].kids.size()) { ++st.back().second; st.push_back({tmp[id].kids[k], 0u}); continue; }
    tmp[id].se = ++next_se;
    st.pop_back();
  }
  adopt(tmp, 0, next_se);
  return "";
}

void HostTree::compute_logw()
{
  logw.assign(nnodes + 1, 0);
  for (uint32_t se = nnodes; se >= 1; --se) { // parents have the larger se
    const uint32_t p = parent[se];
    if (!p) continue;
    const uint32_t nc = eff_nchildren[p];
    logw[se] = (logw[p] == 0xFFFFFFFFu || nc == 0 || (nc & (nc - 1))) ? 0xFFFFFFFFu : logw[p] + (uint32_t)__builtin_ctz(nc);
  }
}

std::string HostTree::node_name(uint32_t se, bool return_na) const
{
  if (se == 0 || se > nnodes) return "";
  if (!name[se].empty()) return name[se];
  return return_na ? std::string("NA") : std::to_string(se - 1);
}

std::string HostTree::jplace_newick() const
{
  // iterative emission of "(" children ")" name[:blen]{se-1}, ";" after the root (ref src/phytree.cpp:47-64)
  std::string out;
  struct Frame { uint32_t se; uint32_t next; bool opened; };
  std::vector<Frame> st;
  st.push_back({root, 0, false});
  while (!st.empty()) {
    Frame& f = st.back();
    if (!is_leaf[f.se] && !f.opened) { out += '('; f.opened = true; f.next = first_child[f.se]; }
    if (!is_leaf[f.se] && f.next) {
      const uint32_t c = f.next;
      if (c != first_child[f.se]) out += ',';
      f.next = next_sibling[c];
      st.push_back({c, 0, false});
      continue;
    }
    if (!is_leaf[f.se]) out += ')';
    out += name[f.se];
    if (!std::isnan(blen[f.se])) { out += ':'; out += fixed5(blen[f.se]); }
    out += '{'; out += std::to_string(f.se - 1); out += '}';
    if (f.se == root) out += ';';
    st.pop_back();
  }
  return out;
}

// ------------------------------------------------------------------------------------------------ index files

namespace {

bool slurp(const std::string& path, std::string& out)
{
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

std::vector<BitRun> runs_of(uint64_t mask)
{ // decomposes a pext mask into contiguous runs: pext(x, mask) == OR over runs of ((x >> src) & ones(width)) << dst
  std::vector<BitRun> v;
  uint8_t dst = 0;
  for (int b = 0; b < 64;) {
    if (!((mask >> b) & 1)) { ++b; continue; }
    int e = b;
    while (e < 64 && ((mask >> e) & 1)) ++e;
    v.push_back({(uint8_t)b, (uint8_t)(e - b), dst});
    dst = (uint8_t)(dst + (e - b));
    b = e;
  }
  return v;
}

} // namespace

std::string generated_newick(const std::vector<std::string>& names)
{
  std::string out;
  struct Gen {
    const std::vector<std::string>& nm; std::string& out;
    void run(size_t lo, size_t hi)
    { // a range of one name is a leaf; a longer one gets the SECOND half as its first child; every branch length is 1
      if (hi - lo == 1) { out += '\''; for (char c : nm[lo]) { if (c == '\'') out += '\''; out += c; } out += '\''; }
      else { const size_t half = lo + (hi - lo) / 2; out += '('; run(half, hi); out += ','; run(lo, half); out += ')'; }
      out += ":1";
    }
  } gen{names, out};
  if (!names.empty()) gen.run(0, names.size());
  out += ';';
  return out;
}

static std::vector<uint32_t> split_rows(const std::vector<uint64_t>& inc, uint64_t nkmers, uint32_t nrows, uint32_t nshards)
{ // shard g starts at the first row whose bucket ends beyond g/nshards of the entries: contiguous row ranges of (nearly) equal
  // cmer bytes, the same on every rank because they depend on inc-* alone
  std::vector<uint32_t> row_splits(nshards + 1, 0);
  for (uint32_t g = 1; g < nshards; ++g) {
    const uint64_t target = (uint64_t)((unsigned __int128)nkmers * g / nshards);
    const uint32_t at = (uint32_t)(std::upper_bound(inc.begin(), inc.end(), target) - inc.begin());
    row_splits[g] = std::max(row_splits[g - 1], std::min(at, nrows));
  }
  row_splits[nshards] = nrows;
  return row_splits;
}

uint64_t HostIndex::replicated_device_bytes() const
{ // what krepp_index_open_shard uploads besides cmer / inc32 (api.cu)
  const uint64_t nn = (uint64_t)tree.nnodes + 1;
  return 8ull * pse.size() + kind.size() + 8ull * kind.size() + 8ull * rho.size() + 4ull * nn * 7 + 8ull * nn + 4ull * tree.leaf_se.size() + 8ull * 256 * 16 +
         4ull * cbeg.size() + 4ull * (cleaf.size() + 1);
}

uint64_t HostIndex::shard_table_device_bytes(uint32_t n) const
{
  const std::vector<uint32_t> sp = split_rows(inc, nkmers, nrows, n);
  uint64_t worst = 0;
  for (uint32_t g = 0; g < n; ++g) {
    const uint64_t e0 = sp[g] ? inc[sp[g] - 1] : 0, e1 = sp[g + 1] ? inc[sp[g + 1] - 1] : 0;
    worst = std::max<uint64_t>(worst, 8ull * (e1 - e0 + 4) + 4ull * ((uint64_t)sp[g + 1] - sp[g] + 1));
  }
  return worst;
}

namespace {

// One partial library of the directory: the files of one suffix "-m{m}r{r}-{frac|no_frac}" (ref src/krepp.cpp:72-91).
struct Partial {
  std::string sfx;
  uint32_t k = 0, w = 0, h = 0, m = 0, r = 0, frac = 0, nrows = 0;
  std::vector<uint8_t> ppos, npos;
  uint64_t nkmers = 0;
  std::vector<uint64_t> inc, pse;
  std::vector<double> rho;
  uint32_t cr_nnodes = 0, nsubsets = 0;
  bool wbackbone = true;
  std::string newick;
  uint32_t row_base = 0, cshift = 0; // where its rows start in the merged table; what its colour ids above the tree nodes are shifted by
  uint64_t ent_base = 0;             // where its entries start in the merged table
  bool sketch = false;               // the one "partial" of a sketch file: 4-byte entries at byte 8 of `sfx` (the file's path)
};

// The sketch of one genome (`krepp sketch`, ref src/sketch.cpp:3-23 Sketch::load_full_sketch; table part SFlatHT::load
// src/table.cpp:23-33) as a partial library with ONE reference: u64 nkmers, nkmers x u32 enc, u32 nrows, nrows x u64 inc, then the
// metadata record of an index (k, w, h, m, r, frac, nrows, ppos, npos) and the genome's rho.  The colour of every entry is 1,
// the only leaf of a one-node tree, so the same match / resolve / solve chain produces the two per-strand histograms of
// SBatch::search_mers (ref src/seek.cpp:55-120) as its records.
std::string read_sketch(const std::string& path, Partial& q)
{
  q.sfx = path; q.sketch = true;
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return "Failed to open " + path;
  f.read(reinterpret_cast<char*>(&q.nkmers), 8);
  if (!f.good() || q.nkmers > (1ull << 40)) return "Failed to read the sketch file!";
  f.seekg((std::streamoff)(8 + 4 * q.nkmers));
  uint32_t nr = 0;
  f.read(reinterpret_cast<char*>(&nr), 4);
  if (!f.good()) return "Failed to read the sketch file!";
  q.inc.resize(nr);
  f.read(reinterpret_cast<char*>(q.inc.data()), (std::streamsize)(8ull * nr));
  unsigned char md[16];
  f.read(reinterpret_cast<char*>(md), 16);
  if (!f.good()) return "Failed to read the sketch file!";
  q.k = md[0]; q.w = md[1]; q.h = md[2];
  std::memcpy(&q.m, md + 3, 4); std::memcpy(&q.r, md + 7, 4); q.frac = md[11]; std::memcpy(&q.nrows, md + 12, 4);
  if (q.k == 0 || q.k > 32 || q.h == 0 || q.h >= q.k || q.k - q.h > 16 || q.m == 0) return "Failed to read the sketch file!";
  q.nrows = nr;
  q.ppos.resize(q.h); q.npos.resize(q.k - q.h);
  f.read(reinterpret_cast<char*>(q.ppos.data()), q.h);
  f.read(reinterpret_cast<char*>(q.npos.data()), q.k - q.h);
  double rho = 0;
  f.read(reinterpret_cast<char*>(&rho), 8);
  if (!f.good()) return "Failed to read the sketch file!";
  uint64_t prev = 0;
  for (uint64_t v : q.inc) { if (v < prev || v > q.nkmers) return "Failed to read the sketch file!"; prev = v; }
  std::string stem = path.substr(path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
  q.newick = "'";
  for (char c : stem) { if (c == '\'') q.newick += '\''; q.newick += c; }
  q.newick += "';";
  q.wbackbone = false;
  q.cr_nnodes = 2; q.nsubsets = 2;
  q.pse = {0ull, 1ull << 32}; // a leaf's record is (0, itself) (ref src/record.cpp: leaves)
  q.rho = {0.0, rho};
  return "";
}

std::string read_partial(const std::string& dir, Partial& q)
{
  std::string buf;
  const std::string& sfx = q.sfx;
  if (!slurp(dir + "/metadata" + sfx, buf)) return "Failed to open " + dir + "/metadata" + sfx;
  if (buf.size() < 16) return "Failed to read the metadata of a partial skecth!";
  const unsigned char* md = reinterpret_cast<const unsigned char*>(buf.data());
  q.k = md[0]; q.w = md[1]; q.h = md[2];
  std::memcpy(&q.m, md + 3, 4); std::memcpy(&q.r, md + 7, 4); q.frac = md[11]; std::memcpy(&q.nrows, md + 12, 4);
  if (q.k == 0 || q.k > 32 || q.h == 0 || q.h >= q.k || q.k - q.h > 16 || q.m == 0 || buf.size() < 16 + (size_t)q.k) return "Failed to read the metadata of a partial skecth!";
  q.ppos.assign(md + 16, md + 16 + q.h);
  q.npos.assign(md + 16 + q.h, md + 16 + q.k);
  // ---- tree: the backbone tree file, else the balanced tree the reference generates over reflist-* (ref src/index.cpp:3-27,
  //      Node::generate_tree src/phytree.cpp:217-253), written as Newick so that the same parser numbers it
  q.wbackbone = slurp(dir + "/tree" + sfx, q.newick);
  if (!q.wbackbone) {
    std::string rl;
    if (!slurp(dir + "/reflist" + sfx, rl)) return "Unable to open reference list file for an index without a tree.";
    std::vector<std::string> names;
    for (size_t at = 0; at < rl.size();) { // std::getline: one name per line, a last line without newline counts
      size_t e = rl.find('\n', at);
      if (e == std::string::npos) e = rl.size();
      names.push_back(rl.substr(at, e - at));
      at = e + 1;
    }
    if (names.empty()) return "Unable to open reference list file for an index without a tree.";
    q.newick = generated_newick(names);
  }
  { // ---- inc, and the size of cmer
    std::ifstream f(dir + "/inc" + sfx, std::ios::binary);
    if (!f.is_open()) return "Failed to open " + dir + "/inc" + sfx;
    uint32_t nr = 0;
    f.read(reinterpret_cast<char*>(&nr), 4);
    if (!f.good()) return "Failed to read the offset array of a partial index!";
    q.nrows = nr;
    q.inc.resize(q.nrows);
    f.read(reinterpret_cast<char*>(q.inc.data()), (std::streamsize)((uint64_t)q.nrows * 8));
    if (!f.good() && q.nrows) return "Failed to read the offset array of a partial index!";
  }
  {
    std::ifstream f(dir + "/cmer" + sfx, std::ios::binary);
    if (!f.is_open()) return "Failed to open " + dir + "/cmer" + sfx;
    f.read(reinterpret_cast<char*>(&q.nkmers), 8);
    if (!f.good()) return "Failed to read the k-mer vector of a partial index!";
    uint64_t prev = 0;
    for (uint64_t v : q.inc) { if (v < prev || v > q.nkmers) return "Failed to read the offset array of a partial index!"; prev = v; }
  }
  // ---- crecord
  if (!slurp(dir + "/crecord" + sfx, buf)) return "Failed to open " + dir + "/crecord" + sfx;
  if (buf.size() < 8) return "Failed to read the color array of a partial index!";
  std::memcpy(&q.cr_nnodes, buf.data(), 4); std::memcpy(&q.nsubsets, buf.data() + 4, 4);
  if (buf.size() < 8 + 8ull * q.nsubsets + 8ull * q.cr_nnodes) return "Failed to read the color array of a partial index!";
  q.pse.resize(q.nsubsets);
  std::memcpy(q.pse.data(), buf.data() + 8, 8ull * q.nsubsets);
  q.rho.resize(q.cr_nnodes);
  std::memcpy(q.rho.data(), buf.data() + 8 + 8ull * q.nsubsets, 8ull * q.cr_nnodes);
  return "";
}

} // namespace

std::string HostIndex::set_masks()
{ // ref src/lshf.cpp:39-52
  mask_hash_bp = mask_drop_lr = mask_drop_bp = 0;
  for (uint8_t p : npos) { if (p >= k) return "Failed to read the metadata of a partial skecth!"; mask_drop_lr += 0x0000000100000001ull << p; mask_drop_bp += 3ull << (2 * p); }
  for (uint32_t i = 0; i < 16 - (k - h); ++i) mask_drop_lr += 1ull << (i + k);
  for (uint8_t p : ppos) { if (p >= k) return "Failed to read the metadata of a partial skecth!"; mask_hash_bp += 3ull << (2 * p); }
  if ((mask_hash_bp & mask_drop_bp) || __builtin_popcountll(mask_hash_bp | mask_drop_bp) != 2 * (int)k) return "Failed to read the metadata of a partial skecth!";
  hash_runs = runs_of(mask_hash_bp);
  drop_runs = runs_of(mask_drop_bp);
  return "";
}

void lsh_positions(uint32_t k, uint32_t h, bool seeded, uint32_t seed, std::vector<uint8_t>& ppos, std::vector<uint8_t>& npos)
{
  // ref src/lshf.cpp:125-147 LSHF::get_random_positions over the global std::mt19937 `gen` (src/common.cpp:7: default-constructed,
  // re-seeded only by --seed, src/krepp.cpp:688-692).  std::uniform_int_distribution<uint8_t>(0, k - 1) over a 32-bit engine is,
  // in libstdc++ (bits/uniform_int_dist.h, _S_nd), Lemire's multiply-and-reject: the high word of draw * k, redrawn while the low
  // word falls below 2^32 mod k.
  std::mt19937 gen;
  if (seeded) gen.seed(seed);
  auto draw = [&]() {
    uint64_t product = (uint64_t)(uint32_t)gen() * k;
    uint32_t low = (uint32_t)product;
    if (low < k) { const uint32_t threshold = (0u - k) % k; while (low < threshold) { product = (uint64_t)(uint32_t)gen() * k; low = (uint32_t)product; } }
    return (uint8_t)(product >> 32);
  };
  ppos.clear(); npos.clear();
  while (ppos.size() < h) { const uint8_t n = draw(); if (!std::count(ppos.begin(), ppos.end(), n)) ppos.push_back(n); }
  std::sort(ppos.begin(), ppos.end());
  for (uint32_t i = 0, at = 0; i < k; ++i) { if (at < h && i == ppos[at]) ++at; else npos.push_back((uint8_t)i); }
  std::sort(ppos.begin(), ppos.end(), std::greater<uint8_t>());
}

std::string HostIndex::set_geometry(uint32_t k_, uint32_t w_, uint32_t h_, uint32_t m_, uint32_t r_, bool frac_, const std::vector<uint8_t>& ppos_, const std::vector<uint8_t>& npos_)
{
  // validate_configuration (ref src/krepp.hpp:59-85), its messages
  if (w_ < k_) return "The minimum minimizer window size (-w) is k (-k).";
  if (h_ < 3) return "The minimum number of LSH positions (-h) is 3.";
  if (h_ > 15) return "The maximum number of LSH positions (-h) is 15.";
  if (k_ > 31) return "The maximum allowed k-mer length (-k) is 31.";
  if (k_ < 19) return "The minimum allowed k-mer length (-k) is 19.";
  if (k_ - h_ > 16) return "For compact k-mer encodings, h must be >= k-16.";
  if (!m_ || ppos_.size() != h_ || npos_.size() != k_ - h_) return "Invalid configuration!";
  k = k_; w = w_; h = h_; m = m_; r = r_; frac = frac_ ? 1 : 0; ppos = ppos_; npos = npos_;
  if (std::string err = set_masks(); !err.empty()) return err;
  { // BaseLSH::set_nrows (ref src/krepp.cpp:5-16)
    const uint32_t hash_size = 1u << (2 * h), full_residue = hash_size % m;
    if (frac) { nrows = (hash_size / m) * (r + 1); nrows = full_residue > r ? nrows + (r + 1) : nrows + full_residue; }
    else { nrows = hash_size / m; nrows = full_residue > r ? nrows + 1 : nrows; }
  }
  res_numer.assign(m, 0); res_base.assign(m, 0);
  for (uint32_t res = 0; res < m; ++res) if (frac ? res <= r : res == r) res_numer[res] = frac ? (int32_t)(r + 1) : 1;
  is_geometry = true;
  nkmers = 0; cr_nnodes = 1; nsubsets = 1;
  row0 = 0; row1 = 0; shard = 0; nshards = 1;
  return "";
}

std::string HostIndex::load(const std::string& dir, uint32_t shard_id, uint32_t shard_count, bool with_table, const std::string& qtree_path, bool lineages)
{
  if (!shard_count || shard_id >= shard_count) return "Bad shard arguments for the index!";
  shard = shard_id; nshards = shard_count;
  struct stat dst;
  is_sketch = stat(dir.c_str(), &dst) == 0 && S_ISREG(dst.st_mode);
  std::vector<Partial> parts;
  if (is_sketch) {
    parts.resize(1);
    if (std::string err = read_sketch(dir, parts[0]); !err.empty()) return err;
  } else {
    // group files by suffix the way TargetIndex::load_index does (ref src/krepp.cpp:72-91)
    std::vector<std::string> suffixes;
    DIR* d = opendir(dir.c_str());
    if (!d) return "Failed to open " + dir;
    while (dirent* e = readdir(d)) {
      std::string fn = e->d_name;
      if (fn.rfind("metadata-", 0) == 0 && fn.find('.') == std::string::npos) suffixes.push_back(fn.substr(8));
    }
    closedir(d);
    if (suffixes.empty()) return "There is no partial index in " + dir;
    std::sort(suffixes.begin(), suffixes.end());

    // ---- every partial library of the directory (ref src/krepp.cpp:92-106).  Several of them -- separate `krepp index` runs over
    //      disjoint hash residues into one directory -- become ONE image: their tables are laid one after the other (a residue's
    //      rows start at its partial's row base), colour ids above the tree nodes are shifted per partial so that they stay
    //      distinct, and the colour records are appended in the same order.  What the reference keeps per partial and the image
    //      cannot: a rho array of its own -- partials built from the same genomes carry the same whole-genome estimates, and a
    //      directory whose partials disagree is refused rather than approximated.
    parts.resize(suffixes.size());
    for (size_t i = 0; i < parts.size(); ++i) {
      parts[i].sfx = suffixes[i];
      if (std::string err = read_partial(dir, parts[i]); !err.empty()) return err;
    }
  }
  const Partial& p0 = parts[0];
  k = p0.k; w = p0.w; h = p0.h; m = p0.m; r = p0.r; frac = p0.frac;
  ppos = p0.ppos; npos = p0.npos;
  wbackbone = p0.wbackbone;
  if (std::string err = tree.parse(p0.newick); !err.empty()) return err;
  for (size_t i = 1; i < parts.size(); ++i) {
    const Partial& q = parts[i];
    if (q.k != k || q.h != h || q.m != m || q.ppos != ppos || q.npos != npos) return "Partial indexes are incompatible, not built using the same LSH function!"; // ref src/lshf.cpp:159-180
    HostTree t2;
    if (std::string err = t2.parse(q.newick); !err.empty()) return err;
    if (t2.nnodes != tree.nnodes || t2.name != tree.name || q.wbackbone != wbackbone) return "Partial indexes are incompatible, not built on the same backbone tree!"; // ref src/phytree.cpp:10-36
  }
  if (std::string err = set_masks(); !err.empty()) return err;
  // residues -> numerator and row base (ref src/index.cpp:144-157 r_to_flatht / r_to_numerator, :160-168 bucket_indices)
  res_numer.assign(m, 0);
  res_base.assign(m, 0);
  {
    uint64_t rows = 0, ents = 0;
    uint32_t extra = 0;
    const uint64_t hash_size = 1ull << (2 * h);
    for (Partial& q : parts) {
      if (q.cr_nnodes != tree.nnodes + 1 || q.nsubsets < q.cr_nnodes) return "The colour record does not match the backbone tree of the index!";
      q.row_base = (uint32_t)rows; q.ent_base = ents; q.cshift = extra;
      for (uint32_t res = 0; res < m; ++res) {
        const bool mine = q.frac ? res <= q.r : res == q.r;
        if (!mine) continue;
        if (res_numer[res]) return "Partial indexes of " + dir + " overlap: residue " + std::to_string(res) + " is held by two of them";
        res_numer[res] = q.frac ? (int32_t)(q.r + 1) : 1;
        res_base[res] = q.row_base;
        // every rix the hash can produce must address a row of the partial (ref src/krepp.cpp:5-16 set_nrows)
        if (res < hash_size) {
          const uint64_t max_rix = (hash_size - 1 - res) / m * m + res;
          const uint64_t off = res_numer[res] > 1 ? (max_rix / m) * res_numer[res] + res : max_rix / m;
          if (off + 1 > q.nrows) return "Failed to read the offset array of a partial index!";
        }
      }
      rows += q.nrows; ents += q.nkmers; extra += q.nsubsets - q.cr_nnodes;
      if (rows > 0xFFFFFFFFull) return "The partial indexes of " + dir + " have more than 2^32 rows together";
    }
    nrows = (uint32_t)rows; nkmers = ents;
    cr_nnodes = tree.nnodes + 1;
    if ((uint64_t)cr_nnodes + extra > 0xFFFFFFFFull) return "The colour records of " + dir + " do not fit 32-bit ids";
    nsubsets = cr_nnodes + extra;
  }
  // merged offsets, colour records and rho
  inc.clear(); inc.reserve(nrows);
  for (const Partial& q : parts) for (uint64_t v : q.inc) inc.push_back(v + q.ent_base);
  pse.assign(nsubsets, 0);
  rho = p0.rho;
  for (const Partial& q : parts) {
    auto remap = [&](uint32_t c) { return c >= cr_nnodes ? c + q.cshift : c; };
    for (uint32_t se = 0; se < q.nsubsets; ++se) {
      const uint64_t v = (uint64_t)remap((uint32_t)q.pse[se]) | (uint64_t)remap((uint32_t)(q.pse[se] >> 32)) << 32;
      if (se < cr_nnodes) { if (&q != &p0 && v != pse[se]) return "Partial indexes are incompatible, their colour records disagree on the backbone tree!"; pse[se] = v; }
      else pse[se + q.cshift] = v;
    }
    if (q.rho != rho)
      return "The partial libraries of " + dir + " carry different rho estimates; the GPU path keeps one rho per reference and does not load such a directory";
  }
  {
    double s1 = 0, s2 = 0;
    uint64_t prev = 0;
    for (uint64_t v : inc) { const double len = (double)(v - prev); s1 += len; s2 += len * len; prev = v; }
    mean_bucket = nrows ? s1 / nrows : 0;
    size_biased_bucket = s1 > 0 ? s2 / s1 : 0;
  }
  // ---- the shard's slice of the (merged) table
  {
    row_splits = split_rows(inc, nkmers, nrows, nshards);
    row0 = row_splits[shard]; row1 = row_splits[shard + 1];
    ent0 = row0 ? inc[row0 - 1] : 0;
    const uint64_t ent1 = row1 ? inc[row1 - 1] : 0;
    cmer.resize(with_table ? ent1 - ent0 : 0);
    for (const Partial& q : parts) { // every partial's entries that fall into [ent0, ent1), read by several threads, each pread()ing its own range
      if (!with_table) break;
      const uint64_t lo_e = std::max<uint64_t>(ent0, q.ent_base), hi_e = std::min<uint64_t>(ent1, q.ent_base + q.nkmers);
      if (lo_e >= hi_e) continue;
      if (q.sketch) { // 4-byte entries widened with the colour of the sketch's one reference
        const int fd = open(q.sfx.c_str(), O_RDONLY);
        if (fd < 0) return "Failed to open " + q.sfx;
        std::vector<uint32_t> enc(hi_e - lo_e);
        size_t done = 0;
        const size_t bytes = enc.size() * 4;
        while (done < bytes) {
          const ssize_t got = pread(fd, reinterpret_cast<char*>(enc.data()) + done, bytes - done, (off_t)(8 + 4 * (lo_e - q.ent_base) + done));
          if (got <= 0) { close(fd); return "Failed to read the sketch file!"; }
          done += (size_t)got;
        }
        close(fd);
        for (size_t i = 0; i < enc.size(); ++i) cmer[lo_e - ent0 + i] = (uint64_t)enc[i] | 1ull << 32;
        continue;
      }
      const int fd = open((dir + "/cmer" + q.sfx).c_str(), O_RDONLY);
      if (fd < 0) return "Failed to open " + dir + "/cmer" + q.sfx;
      const size_t bytes = (size_t)(hi_e - lo_e) * 8, nth = std::max<size_t>(1, std::min<size_t>(8, bytes >> 24));
      std::vector<char> ok(nth, 1);
      std::vector<std::thread> th;
      auto work = [&](size_t t) {
        size_t lo = bytes * t / nth / 8 * 8, hi = bytes * (t + 1) / nth / 8 * 8;
        if (t + 1 == nth) hi = bytes;
        char* dst = reinterpret_cast<char*>(cmer.data() + (lo_e - ent0));
        const size_t first = lo;
        while (lo < hi) {
          const ssize_t got = pread(fd, dst + lo, hi - lo, (off_t)(8 + 8 * (lo_e - q.ent_base) + lo));
          if (got <= 0) { ok[t] = 0; return; }
          lo += (size_t)got;
        }
        if (q.cshift) // this partial's colour ids above the tree nodes move up
          for (uint64_t* e = reinterpret_cast<uint64_t*>(dst + first), *end = reinterpret_cast<uint64_t*>(dst + hi); e < end; ++e)
            if ((uint32_t)(*e >> 32) >= cr_nnodes) *e += (uint64_t)q.cshift << 32;
      };
      for (size_t t = 1; t < nth; ++t) th.emplace_back(work, t);
      work(0);
      for (auto& x : th) x.join();
      close(fd);
      for (char o : ok) if (!o) return "Failed to read the k-mer vector of a partial index!";
    }
    if (ent1 - ent0 < (1ull << 32)) { inc32.resize(row1 - row0); for (uint32_t i = row0; i < row1; ++i) inc32[i - row0] = (uint32_t)(inc[i] - ent0); }
  }
  // ---- which reference (leaf of `tree`) a colour id below cr_nnodes stands for; with a query tree (place -t) the index's own
  //      tree is only the numbering of the colour ids, and everything downstream lives on the query tree
  //      (ref src/phytree.cpp:421-448 map_to_qtree: leaves matched by name, unmatched index leaves become null nodes)
  std::vector<uint8_t> idx_is_leaf(tree.is_leaf.begin(), tree.is_leaf.end());
  col_rank.assign(cr_nnodes, 0xFFFFFFFFu);
  if (qtree_path.empty()) {
    for (uint32_t se = 1; se < cr_nnodes; ++se) if (tree.is_leaf[se]) col_rank[se] = tree.leaf_rank[se];
  } else {
    std::string text;
    if (!slurp(qtree_path, text)) return "Error opening " + qtree_path;
    HostTree qt;
    if (std::string err = lineages ? qt.parse_lineages(text) : qt.parse(text); !err.empty()) return err;
    std::unordered_map<std::string, uint32_t> name_to_se;
    for (uint32_t se = 1; se < cr_nnodes; ++se) if (tree.is_leaf[se]) name_to_se[tree.name[se]] = se;
    std::vector<double> qrho(qt.nnodes + 1, 0.0);
    std::vector<uint8_t> covered(qt.nnodes + 1, 0);
    for (uint32_t q = 1; q <= qt.nnodes; ++q) {
      if (!qt.is_leaf[q] || qt.name[q].empty()) continue;
      auto it = name_to_se.find(qt.name[q]);
      if (it == name_to_se.end()) continue;
      col_rank[it->second] = qt.leaf_rank[q];
      qrho[q] = rho[it->second];
      for (uint32_t a = q; a && !covered[a]; a = qt.parent[a]) covered[a] = 1; // ref src/phytree.cpp:450-473 compute_eff_nchildren
    }
    qt.eff_nchildren.assign(qt.nnodes + 1, 0);
    for (uint32_t q = 1; q <= qt.nnodes; ++q) if (covered[q] && qt.parent[q]) ++qt.eff_nchildren[qt.parent[q]];
    qt.compute_logw();
    tree = std::move(qt);
    rho.swap(qrho);
    wbackbone = true;
  }
  // make_rho_partial: rho *= (#residues present)/m (ref src/index.cpp:188-201, src/record.cpp:304-309)
  {
    uint32_t present = 0;
    for (uint32_t res = 0; res < m; ++res) present += res_numer[res] != 0;
    const double ratio_m = (double)present / (double)m;
    for (double& x : rho) x *= ratio_m;
  }
  // classify colour ids the way add_matching_mer walks them (ref src/query.cpp:369-387)
  kind.assign(nsubsets, 2);
  kind[0] = 0;
  for (uint32_t se = 1; se < cr_nnodes; ++se) kind[se] = idx_is_leaf[se] ? (col_rank[se] != 0xFFFFFFFFu ? 1 : 0) : 2;
  { // every entry's colour id must exist (several threads: the table is most of the index)
    const size_t n = cmer.size(), nth = n < (1u << 22) ? 1 : 8;
    std::vector<char> bad(nth, 0);
    std::vector<std::thread> th;
    auto scan = [&](size_t t) { for (size_t i = n * t / nth, e = n * (t + 1) / nth; i < e; ++i) if ((uint32_t)(cmer[i] >> 32) >= nsubsets) { bad[t] = 1; return; } };
    for (size_t t = 1; t < nth; ++t) th.emplace_back(scan, t);
    scan(0);
    for (auto& x : th) x.join();
    for (char b : bad) if (b) return "Failed to read the k-mer vector of a partial index!";
  }
  // expansion depth / leaf count per colour (iterative post-order over the DAG; also rejects cycles)
  {
    std::vector<uint32_t> depth(nsubsets, 0), leaves(nsubsets, 0);
    std::vector<uint8_t> state(nsubsets, 0); // 0 new, 1 open, 2 done
    std::vector<uint32_t> st, order; // order: colour ids, children before parents
    order.reserve(nsubsets);
    for (uint32_t s0 = 0; s0 < nsubsets; ++s0) {
      if (state[s0]) continue;
      st.push_back(s0);
      while (!st.empty()) {
        const uint32_t se = st.back();
        if (state[se] == 2) { st.pop_back(); continue; } // pushed twice before it was finished (both children of one node)
        if (kind[se] != 2) { state[se] = 2; depth[se] = 0; leaves[se] = kind[se] == 1; order.push_back(se); st.pop_back(); continue; }
        const uint32_t a = (uint32_t)pse[se], b = (uint32_t)(pse[se] >> 32);
        if (a >= nsubsets || b >= nsubsets) return "The colour record of the index is corrupt (child id out of range)!";
        if (state[se] == 0) {
          state[se] = 1;
          bool pushed = false;
          for (uint32_t c : {a, b}) {
            if (state[c] == 1) return "The colour record of the index is corrupt (cycle)!";
            if (state[c] == 0) { st.push_back(c); pushed = true; }
          }
          if (pushed) continue;
        }
        state[se] = 2;
        depth[se] = 1 + std::max(depth[a], depth[b]);
        leaves[se] = leaves[a] + leaves[b];
        order.push_back(se);
        st.pop_back();
      }
    }
    max_expand_depth = *std::max_element(depth.begin(), depth.end());
    max_colour_leaves = *std::max_element(leaves.begin(), leaves.end());
    // flattened leaf lists, built children first; duplicates (a leaf reachable along two paths) are merged away
    uint64_t total = 0;
    for (uint32_t c : leaves) total += c;
    if (total <= kMaxFlatLeaves && tree.nleaves) {
      // Every colour gets room for its list WITH duplicates (leaves[], an upper bound), so the lists of one DAG level -- whose
      // children all sit on lower levels and are finished -- can be built independently, by several threads; the lists are then
      // packed in colour-id order.
      std::vector<uint64_t> start(nsubsets + 1, 0);
      for (uint32_t se = 0; se < nsubsets; ++se) start[se + 1] = start[se] + leaves[se];
      std::vector<uint32_t, NoInitAlloc<uint32_t>> flat;
      flat.resize(total);
      std::vector<uint32_t> len(nsubsets, 0);
      std::vector<uint32_t> level_begin(max_expand_depth + 2, 0), bylevel(nsubsets);
      for (uint32_t se = 0; se < nsubsets; ++se) ++level_begin[depth[se] + 1];
      for (uint32_t d = 0; d <= max_expand_depth; ++d) level_begin[d + 1] += level_begin[d];
      { std::vector<uint32_t> cur(level_begin.begin(), level_begin.end() - 1); for (uint32_t se = 0; se < nsubsets; ++se) bylevel[cur[depth[se]]++] = se; }
      auto build = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
          const uint32_t se = bylevel[i];
          uint32_t* out = flat.data() + start[se];
          if (kind[se] == 1) { out[0] = col_rank[se]; len[se] = 1; }
          else if (kind[se] == 2) { // sorted union of the two children's lists
            const uint32_t ca = (uint32_t)pse[se], cb = (uint32_t)(pse[se] >> 32);
            const uint32_t *pa = flat.data() + start[ca], *ea = pa + len[ca], *pb = flat.data() + start[cb], *eb = pb + len[cb];
            uint32_t n = 0;
            while (pa < ea && pb < eb) {
              const uint32_t x = *pa, y = *pb;
              out[n++] = x < y ? x : y;
              pa += x <= y; pb += y <= x;
            }
            while (pa < ea) out[n++] = *pa++;
            while (pb < eb) out[n++] = *pb++;
            len[se] = n;
          }
        }
      };
      auto in_parallel = [&](uint32_t lo, uint32_t hi, auto&& fn) {
        const uint32_t n = hi - lo, nth = n < (1u << 14) ? 1u : 8u;
        std::vector<std::thread> th;
        for (uint32_t t = 1; t < nth; ++t) th.emplace_back(fn, lo + (uint64_t)n * t / nth, lo + (uint64_t)n * (t + 1) / nth);
        fn(lo, lo + (uint64_t)n / nth);
        for (auto& x : th) x.join();
      };
      for (uint32_t d = 0; d <= max_expand_depth; ++d) in_parallel(level_begin[d], level_begin[d + 1], build);
      cbeg.assign(nsubsets + 1, 0);
      for (uint32_t se = 0; se < nsubsets; ++se) cbeg[se + 1] = cbeg[se] + len[se];
      cleaf.resize(cbeg[nsubsets]);
      in_parallel(0, nsubsets, [&](uint32_t lo, uint32_t hi) {
        for (uint32_t se = lo; se < hi; ++se) std::copy(flat.data() + start[se], flat.data() + start[se] + len[se], cleaf.begin() + cbeg[se]);
      });
    }
  }
  return "";
}

} // namespace krepp

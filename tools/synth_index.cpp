// tools/synth_index.cpp -- MEASUREMENT INFRASTRUCTURE (not part of the product, not part of the oracle).
//
// Generates the synthetic workloads of BASELINE.json configs 3-5 directly on the machine that runs the benchmark
// (nothing persists on a GPU box between calls and a 2.4 GB index cannot be shipped): N genomes evolved down a random
// rooted binary tree with JC69 substitutions, a krepp index over them in the on-disk format of
// `krepp index` (section 8 row a16: metadata-/cmer-/inc-/crecord-/tree-/reflist-), and 150 bp reads sampled from the
// genomes with per-read substitution rate U(0, max_sub) and random strand.
//
// The index is built the way the reference builds it -- window-w minimizers of every genome by the murmur-fmix64 order,
// LSH residue filter, per-bucket sort/unique by the 32-bit residual encoding (ref src/rqseq.cpp:51-144,
// src/table.cpp:110-166), colours as additive 64-bit leaf-hash sums decomposed along the tree (ref
// src/record.cpp:82-107,132-176) -- but as one flat parallel pass instead of the reference's recursive table unions,
// which need tens of GB and many minutes at this size.  tests/test_synth_index.py checks that the reference's own
// `krepp index` on the same genomes and tree gives the identical inc-* file and enc column, and that the reference's
// `krepp dist` prints the same distances on either index.
//
// usage: synth_index --out DIR [--genomes 1000] [--length 3000000] [--seed 7] [--blen 0.02] [--threads T]
//                    [--reads 10000000] [--read-len 150] [--max-sub 0.15] [--fastq-reads 0] [--fasta]
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

namespace {

struct Rng { // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n); }
};

inline uint64_t fmix64(uint64_t h)
{ // ref src/common.hpp:147-155 xur64_hash
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}

struct Node { int left = -1, right = -1, parent = -1; double blen = 0; uint32_t se = 0, lo = 0, hi = 0; int leaf = -1; uint64_t sh = 0; };

struct Cfg {
  std::string out;
  uint32_t genomes = 1000, length = 3000000, k = 27, w = 35, h = 11, m = 4, r = 1, read_len = 150;
  uint64_t seed = 7, reads = 10000000, fastq_reads = 0;
  double blen = 0.02, max_sub = 0.15;
  int threads = 0;
  bool fasta = false;
};

std::vector<Node> nodes;
std::vector<int> leaf_node; // rank -> node

int grow(Rng& rng, uint32_t n, int parent, double blen_mean)
{
  const int id = (int)nodes.size();
  nodes.emplace_back();
  nodes[id].parent = parent;
  nodes[id].blen = parent < 0 ? 0.0 : -std::log(1.0 - rng.uniform()) * blen_mean + 1e-4;
  if (n == 1) { nodes[id].leaf = (int)leaf_node.size(); nodes[id].lo = (uint32_t)leaf_node.size(); leaf_node.push_back(id); nodes[id].hi = nodes[id].lo + 1; return id; }
  const uint32_t nl = 1 + (uint32_t)rng.below(n - 1);
  const uint32_t lo = (uint32_t)leaf_node.size();
  const int l = grow(rng, nl, id, blen_mean);
  const int r = grow(rng, n - nl, id, blen_mean);
  nodes[id].left = l; nodes[id].right = r; nodes[id].lo = lo; nodes[id].hi = (uint32_t)leaf_node.size();
  return id;
}

void number_postorder(int id, uint32_t& se)
{ // iterative post-order (children before parents, left subtree first) == the order Node::parse hands out se
  std::vector<std::pair<int, int>> st{{id, 0}};
  while (!st.empty()) {
    auto& [n, stage] = st.back();
    if (nodes[n].left < 0) { nodes[n].se = ++se; st.pop_back(); continue; }
    if (stage == 0) { stage = 1; st.push_back({nodes[n].left, 0}); }
    else if (stage == 1) { stage = 2; st.push_back({nodes[n].right, 0}); }
    else { nodes[n].se = ++se; st.pop_back(); }
  }
}

std::string leaf_name(int rank) { char b[32]; snprintf(b, sizeof b, "G%06d", rank); return b; }

void newick(int root, std::string& out)
{
  struct F { int n; int stage; };
  std::vector<F> st{{root, 0}};
  char b[64];
  while (!st.empty()) {
    F& f = st.back();
    const Node& nd = nodes[f.n];
    if (nd.left < 0) {
      out += leaf_name(nd.leaf);
      snprintf(b, sizeof b, ":%.6f", nd.blen); out += b;
      st.pop_back();
      continue;
    }
    if (f.stage == 0) { out += '('; f.stage = 1; st.push_back({nd.left, 0}); }
    else if (f.stage == 1) { out += ','; f.stage = 2; st.push_back({nd.right, 0}); }
    else {
      out += ')';
      // the reference's Newick reader needs a label and a length on the root (see tools/synth.py random_genomes)
      if (nd.parent < 0) out += "root:0.0;"; else { snprintf(b, sizeof b, ":%.6f", nd.blen); out += b; }
      st.pop_back();
    }
  }
}

// JC69: a site differs after time t with probability 0.75 (1 - exp(-4t/3)); geometric skipping between hits
void mutate(std::vector<uint8_t>& s, double t, Rng& rng)
{
  const double p = 0.75 * (1.0 - std::exp(-4.0 * t / 3.0));
  if (p <= 0) return;
  const double lq = std::log(1.0 - p);
  double pos = std::floor(std::log(1.0 - rng.uniform()) / lq);
  while (pos < (double)s.size()) {
    const size_t i = (size_t)pos;
    s[i] = (uint8_t)((s[i] + 1 + rng.below(3)) & 3);
    pos += 1.0 + std::floor(std::log(1.0 - rng.uniform()) / lq);
  }
}

std::vector<uint8_t> pack2(const std::vector<uint8_t>& s)
{
  std::vector<uint8_t> p((s.size() + 3) / 4, 0);
  for (size_t i = 0; i < s.size(); ++i) p[i >> 2] |= (uint8_t)(s[i] << (2 * (i & 3)));
  return p;
}

void write_file(const std::string& path, const void* data, size_t n)
{
  FILE* f = fopen(path.c_str(), "wb");
  if (!f || (n && fwrite(data, 1, n, f) != n)) { fprintf(stderr, "synth_index: cannot write %s\n", path.c_str()); exit(1); }
  fclose(f);
}

} // namespace

int main(int argc, char** argv)
{
  Cfg c;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--out") c.out = val();
    else if (a == "--genomes") c.genomes = (uint32_t)atol(val());
    else if (a == "--length") c.length = (uint32_t)atol(val());
    else if (a == "--seed") c.seed = strtoull(val(), nullptr, 10);
    else if (a == "--blen") c.blen = atof(val());
    else if (a == "--threads") c.threads = atoi(val());
    else if (a == "--reads") c.reads = strtoull(val(), nullptr, 10);
    else if (a == "--read-len") c.read_len = (uint32_t)atol(val());
    else if (a == "--max-sub") c.max_sub = atof(val());
    else if (a == "--fastq-reads") c.fastq_reads = strtoull(val(), nullptr, 10);
    else if (a == "--fasta") c.fasta = true;
    else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (c.out.empty() || c.genomes < 2 || c.length < c.w) { fprintf(stderr, "usage: synth_index --out DIR [--genomes N>=2] [--length L] ...\n"); return 2; }
  if (c.threads > 0) omp_set_num_threads(c.threads);
  mkdir(c.out.c_str(), 0755);
  const std::string idx = c.out + "/index";
  mkdir(idx.c_str(), 0755);
  const double t0 = omp_get_wtime();

  // ---- tree
  Rng trng(c.seed);
  const int root = grow(trng, c.genomes, -1, c.blen);
  uint32_t nnodes = 0;
  number_postorder(root, nnodes);
  for (size_t r = 0; r < leaf_node.size(); ++r) nodes[leaf_node[r]].sh = fmix64(0x51ED270B1ull + r * 0x9E3779B97F4A7C15ull) | 1ull;
  { // node colour hash = sum of the leaves below (ref src/record.cpp:63-66 sum_children_sh)
    std::vector<int> order(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) order[nodes[i].se - 1] = (int)i;
    for (int n : order) if (nodes[n].left >= 0) nodes[n].sh = nodes[nodes[n].left].sh + nodes[nodes[n].right].sh;
  }
  std::string nwk;
  newick(root, nwk);
  nwk += "\n";

  // ---- genomes: evolve down the tree, keep the leaves 2-bit packed
  std::vector<std::vector<uint8_t>> packed(c.genomes);
  {
    Rng grng(c.seed * 0x2545F4914F6CDD1Dull + 11);
    std::vector<uint8_t> rootseq(c.length);
    for (auto& b : rootseq) b = (uint8_t)(grng.next() >> 62);
    struct F { int n; std::vector<uint8_t> seq; int stage; };
    std::vector<F> st;
    st.push_back({root, std::move(rootseq), 0});
    while (!st.empty()) {
      F& f = st.back();
      const Node& nd = nodes[f.n];
      if (nd.left < 0) { packed[nd.leaf] = pack2(f.seq); st.pop_back(); continue; }
      if (f.stage < 2) {
        const int ch = f.stage == 0 ? nd.left : nd.right;
        ++f.stage;
        std::vector<uint8_t> s = f.seq;
        mutate(s, nodes[ch].blen, grng);
        st.push_back({ch, std::move(s), 0}); // invalidates f
      } else st.pop_back();
    }
  }
  fprintf(stderr, "[synth_index] %u genomes x %u bp evolved on a random binary tree (%u nodes): %.1f s\n", c.genomes, c.length, nnodes, omp_get_wtime() - t0);

  // ---- LSH geometry (positions as the reference's default-seeded mt19937 draws them for k=27, h=11)
  std::vector<uint8_t> ppos;
  if (c.k == 27 && c.h == 11) ppos = {26, 24, 22, 21, 17, 14, 8, 7, 5, 3, 2};
  else { Rng prng(c.seed + 99); while (ppos.size() < c.h) { const uint8_t p = (uint8_t)prng.below(c.k); if (!std::count(ppos.begin(), ppos.end(), p)) ppos.push_back(p); } std::sort(ppos.rbegin(), ppos.rend()); }
  std::vector<uint8_t> npos;
  for (uint8_t p = 0; p < c.k; ++p) if (!std::count(ppos.begin(), ppos.end(), p)) npos.push_back(p);
  std::vector<int> hrank(c.k, -1), nrank(c.k, -1);
  { std::vector<uint8_t> asc = ppos; std::sort(asc.begin(), asc.end()); for (size_t j = 0; j < asc.size(); ++j) hrank[asc[j]] = (int)j; for (size_t j = 0; j < npos.size(); ++j) nrank[npos[j]] = (int)j; }
  const uint64_t hash_size = 1ull << (2 * c.h);
  const uint32_t res = (uint32_t)(hash_size % c.m);
  const uint32_t nrows = (uint32_t)((hash_size / c.m) * (c.r + 1) + (res > c.r ? c.r + 1 : res)); // frac (ref src/krepp.cpp:5-16)
  const uint64_t mask_bp = c.k == 32 ? ~0ull : ((1ull << (2 * c.k)) - 1);
  // byte tables: k-mer word byte -> contribution to rix and enc32
  const uint32_t nbytes = (2 * c.k + 7) / 8;
  std::vector<uint32_t> lut_rix(nbytes * 256, 0), lut_enc(nbytes * 256, 0);
  for (uint32_t b = 0; b < nbytes; ++b)
    for (uint32_t v = 0; v < 256; ++v)
      for (uint32_t s = 0; s < 4; ++s) {
        const uint32_t p = 4 * b + s, code = (v >> (2 * s)) & 3;
        if (p >= c.k) continue;
        if (hrank[p] >= 0) lut_rix[b * 256 + v] |= code << (2 * hrank[p]);
        if (nrank[p] >= 0) lut_enc[b * 256 + v] |= (code & 1) << nrank[p] | (code >> 1) << (16 + nrank[p]);
      }

  // ---- per genome: minimizers -> sorted unique (row << 32 | enc); reads sampled on the way
  const uint32_t ldiff = c.w - c.k + 1;
  std::vector<std::vector<uint64_t>> keys(c.genomes);
  std::vector<double> rho(c.genomes, 0);
  std::vector<uint8_t> reads((size_t)c.reads * c.read_len);
  uint64_t perm_a = (uint64_t)((double)c.reads * 0.6180339887) | 1;
  auto gcd = [](uint64_t a, uint64_t b) { while (b) { const uint64_t t = a % b; a = b; b = t; } return a; };
  while (c.reads && gcd(perm_a, c.reads) != 1) perm_a += 2;
#pragma omp parallel for schedule(dynamic, 1)
  for (uint32_t g = 0; g < c.genomes; ++g) {
    std::vector<uint8_t> s(c.length);
    for (uint32_t i = 0; i < c.length; ++i) s[i] = (packed[g][i >> 2] >> (2 * (i & 3))) & 3;
    std::vector<uint64_t>& out = keys[g];
    out.reserve(c.length / 8);
    std::vector<uint64_t> ring_bp(ldiff, 0), ring_z(ldiff, fmix64(0));
    uint64_t bp = 0, kix = 0, nk = 0, nchange = 0, prev = ~0ull;
    for (uint32_t i = 0; i < c.length; ++i) { // no N in synthetic genomes: the valid run is the whole sequence
      bp = ((bp << 2) | s[i]) & mask_bp;
      if (i + 1 < c.k) continue;
      ++nk;
      const uint32_t slot = (uint32_t)(kix % ldiff);
      ring_bp[slot] = bp; ring_z[slot] = fmix64(bp);
      ++kix;
      if (i + 1 < c.w) continue;
      uint32_t best = 0;
      for (uint32_t j = 1; j < ldiff; ++j) if (ring_z[j] < ring_z[best]) best = j;
      const uint64_t mz = ring_bp[best];
      if (mz != prev) { ++nchange; prev = mz; } else continue; // the same minimizer again adds nothing after unique
      uint32_t rix = 0, enc = 0;
      for (uint32_t b = 0; b < nbytes; ++b) { const uint32_t v = (uint32_t)(mz >> (8 * b)) & 0xFF; rix |= lut_rix[b * 256 + v]; enc |= lut_enc[b * 256 + v]; }
      const uint32_t rr = rix % c.m;
      if (rr <= c.r) out.push_back((uint64_t)((rix / c.m) * (c.r + 1) + rr) << 32 | enc);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    rho[g] = nk ? (double)nchange / (double)nk : 0; // stands in for the reference's HLL ratio (ref src/rqseq.hpp:79)
    // reads of this genome (equal-length genomes => uniform genome choice)
    Rng rr(c.seed * 1315423911ull + g * 2654435761ull + 5);
    const uint64_t base = c.reads / c.genomes, extra = c.reads % c.genomes;
    const uint64_t cnt = base + (g < extra ? 1 : 0), first = g * base + std::min<uint64_t>(g, extra);
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t j = 0; j < cnt; ++j) {
      const uint64_t slot = c.reads ? ((first + j) * perm_a + 12345) % c.reads : 0;
      uint8_t* dst = reads.data() + slot * c.read_len;
      const uint32_t start = (uint32_t)rr.below(c.length - c.read_len + 1);
      const double d = rr.uniform() * c.max_sub;
      const bool flip = rr.next() & 1;
      for (uint32_t x = 0; x < c.read_len; ++x) {
        uint8_t code = s[start + x];
        if (rr.uniform() < d) code = (uint8_t)(rr.next() >> 62);
        if (flip) dst[c.read_len - 1 - x] = (uint8_t)ACGT[3 - code]; else dst[x] = (uint8_t)ACGT[code];
      }
    }
    if (c.fasta) {
      std::string fa = ">" + leaf_name((int)g) + "\n";
      for (uint32_t i = 0; i < c.length; i += 80) { for (uint32_t x = i; x < std::min(c.length, i + 80); ++x) fa += ACGT[s[x]]; fa += '\n'; }
      mkdir((c.out + "/genomes").c_str(), 0755);
      write_file(c.out + "/genomes/" + leaf_name((int)g) + ".fna", fa.data(), fa.size());
    }
    std::vector<uint8_t>().swap(packed[g]);
  }
  fprintf(stderr, "[synth_index] minimizers + %llu reads: %.1f s\n", (unsigned long long)c.reads, omp_get_wtime() - t0);

  // ---- gather by row: counting sort of (enc << 32 | leaf rank)
  std::vector<uint64_t> row_start((size_t)nrows + 1, 0);
  for (uint32_t g = 0; g < c.genomes; ++g) for (uint64_t key : keys[g]) ++row_start[(key >> 32) + 1];
  for (uint32_t r = 0; r < nrows; ++r) row_start[r + 1] += row_start[r];
  const uint64_t total = row_start[nrows];
  std::vector<uint64_t> items(total);
  {
    std::vector<std::atomic<uint32_t>> fill(nrows);
    for (auto& f : fill) f.store(0, std::memory_order_relaxed);
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t g = 0; g < c.genomes; ++g) {
      for (uint64_t key : keys[g]) {
        const uint32_t row = (uint32_t)(key >> 32);
        items[row_start[row] + fill[row].fetch_add(1, std::memory_order_relaxed)] = (key << 32) | g;
      }
      std::vector<uint64_t>().swap(keys[g]);
    }
  }
  // ---- per row: sort, count distinct encodings
  std::vector<uint64_t> inc(nrows, 0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (uint32_t r = 0; r < nrows; ++r) {
    uint64_t* b = items.data() + row_start[r];
    uint64_t* e = items.data() + row_start[r + 1];
    std::sort(b, e);
    uint64_t n = 0;
    for (uint64_t* p = b; p < e; ++p) n += (p == b) || ((p[0] >> 32) != (p[-1] >> 32));
    inc[r] = n;
  }
  for (uint32_t r = 1; r < nrows; ++r) inc[r] += inc[r - 1];
  const uint64_t nkmers = nrows ? inc[nrows - 1] : 0;
  fprintf(stderr, "[synth_index] %llu (k-mer, genome) pairs -> %llu index entries, mean bucket %.1f: %.1f s\n", (unsigned long long)total,
          (unsigned long long)nkmers, (double)nkmers / nrows, omp_get_wtime() - t0);

  // ---- colours: leaf sets -> ids through additive hashes, decomposed along the tree (clades are rank intervals)
  constexpr int kShards = 1024;
  struct Shard { std::mutex mu; std::unordered_map<uint64_t, uint32_t> map; };
  std::vector<Shard> shards(kShards);
  std::vector<std::pair<uint32_t, uint32_t>> pse(nnodes + 1, {0, 0});
  std::mutex pse_mu;
  for (const Node& nd : nodes) {
    pse[nd.se] = nd.left < 0 ? std::make_pair(0u, nd.se) : std::make_pair(nodes[nd.left].se, nodes[nd.right].se);
    shards[fmix64(nd.sh) % kShards].map[nd.sh] = nd.se;
  }
  auto combine = [&](uint32_t a, uint32_t b, uint64_t sh) -> uint32_t {
    Shard& s = shards[fmix64(sh) % kShards];
    std::lock_guard<std::mutex> l(s.mu);
    auto it = s.map.find(sh);
    if (it != s.map.end()) return it->second;
    uint32_t id;
    { std::lock_guard<std::mutex> l2(pse_mu); id = (uint32_t)pse.size(); pse.emplace_back(a, b); }
    s.map.emplace(sh, id);
    return id;
  };
  auto lookup = [&](uint64_t sh, uint32_t& id) -> bool {
    Shard& s = shards[fmix64(sh) % kShards];
    std::lock_guard<std::mutex> l(s.mu);
    auto it = s.map.find(sh);
    if (it == s.map.end()) return false;
    id = it->second;
    return true;
  };
  // returns (colour id, hash) of the sorted rank set S[0..n) which lies inside node `nd`
  struct Dec {
    decltype(combine)& comb;
    std::pair<uint32_t, uint64_t> operator()(const uint32_t* S, uint32_t n, int nd) const
    {
      for (;;) {
        const Node& N = nodes[nd];
        if (n == N.hi - N.lo) return {N.se, N.sh};
        const uint32_t mid = nodes[N.right].lo;
        if (S[n - 1] < mid) { nd = N.left; continue; }
        if (S[0] >= mid) { nd = N.right; continue; }
        const uint32_t cut = (uint32_t)(std::lower_bound(S, S + n, mid) - S);
        const auto a = (*this)(S, cut, N.left), b = (*this)(S + cut, n - cut, N.right);
        return {comb(a.first, b.first, a.second + b.second), a.second + b.second};
      }
    }
  } decompose{combine};

  std::vector<uint64_t> cmer(nkmers);
#pragma omp parallel
  {
    std::vector<uint32_t> S;
    constexpr uint32_t kCache = 1u << 14;
    std::vector<uint64_t> ck(kCache, 0);
    std::vector<uint32_t> cv(kCache, 0);
#pragma omp for schedule(dynamic, 1024)
    for (uint32_t r = 0; r < nrows; ++r) {
      const uint64_t* b = items.data() + row_start[r];
      const uint64_t* e = items.data() + row_start[r + 1];
      uint64_t at = r ? inc[r - 1] : 0;
      while (b < e) {
        const uint32_t enc = (uint32_t)(b[0] >> 32);
        const uint64_t* q = b;
        while (q < e && (uint32_t)(q[0] >> 32) == enc) ++q;
        uint32_t id;
        if (q - b == 1) id = nodes[leaf_node[(uint32_t)b[0]]].se;
        else {
          uint64_t sh = 0;
          for (const uint64_t* p = b; p < q; ++p) sh += nodes[leaf_node[(uint32_t)p[0]]].sh;
          const uint32_t slot = (uint32_t)(fmix64(sh) & (kCache - 1));
          if (ck[slot] == sh) id = cv[slot];
          else {
            if (!lookup(sh, id)) {
              S.clear();
              for (const uint64_t* p = b; p < q; ++p) S.push_back((uint32_t)p[0]);
              id = decompose(S.data(), (uint32_t)S.size(), root).first;
            }
            ck[slot] = sh; cv[slot] = id;
          }
        }
        cmer[at++] = (uint64_t)id << 32 | enc;
        b = q;
      }
    }
  }
  const uint32_t nsubsets = (uint32_t)pse.size();
  fprintf(stderr, "[synth_index] %u colours (%u tree nodes): %.1f s\n", nsubsets, nnodes, omp_get_wtime() - t0);

  // ---- files
  const std::string sfx = "-m" + std::to_string(c.m) + "r" + std::to_string(c.r) + "-frac";
  {
    std::vector<uint8_t> md;
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; ++i) md.push_back((uint8_t)(v >> (8 * i))); };
    md.push_back((uint8_t)c.k); md.push_back((uint8_t)c.w); md.push_back((uint8_t)c.h); u32(c.m); u32(c.r); md.push_back(1); u32(nrows);
    md.insert(md.end(), ppos.begin(), ppos.end()); md.insert(md.end(), npos.begin(), npos.end());
    write_file(idx + "/metadata" + sfx, md.data(), md.size());
  }
  {
    FILE* f = fopen((idx + "/cmer" + sfx).c_str(), "wb");
    fwrite(&nkmers, 8, 1, f);
    // on disk: {u32 enc, u32 se}; cmer[] holds se << 32 | enc which is exactly that pair in little-endian
    if (fwrite(cmer.data(), 8, nkmers, f) != nkmers) { fprintf(stderr, "synth_index: short write\n"); return 1; }
    fclose(f);
    f = fopen((idx + "/inc" + sfx).c_str(), "wb");
    fwrite(&nrows, 4, 1, f);
    fwrite(inc.data(), 8, nrows, f);
    fclose(f);
    f = fopen((idx + "/crecord" + sfx).c_str(), "wb");
    const uint32_t crn = nnodes + 1;
    fwrite(&crn, 4, 1, f); fwrite(&nsubsets, 4, 1, f);
    fwrite(pse.data(), 8, nsubsets, f);
    std::vector<double> rho_se(crn, 0.0);
    for (uint32_t g = 0; g < c.genomes; ++g) rho_se[nodes[leaf_node[g]].se] = rho[g];
    fwrite(rho_se.data(), 8, crn, f);
    fclose(f);
  }
  write_file(idx + "/tree" + sfx, nwk.data(), nwk.size());
  write_file(c.out + "/tree.nwk", nwk.data(), nwk.size());
  {
    std::string names, map;
    for (uint32_t g = 0; g < c.genomes; ++g) { names += leaf_name((int)g) + "\n"; map += leaf_name((int)g) + "\tgenomes/" + leaf_name((int)g) + ".fna\n"; }
    write_file(idx + "/reflist" + sfx, names.data(), names.size());
    if (c.fasta) write_file(c.out + "/input_map.tsv", map.data(), map.size());
  }
  // ---- reads: raw matrix for the benchmark, FASTQ head for the CLIs
  write_file(c.out + "/reads.u8", reads.data(), reads.size());
  if (c.fastq_reads) {
    const uint64_t n = std::min<uint64_t>(c.fastq_reads, c.reads);
    std::string fq;
    fq.reserve(n * (2 * c.read_len + 24));
    const std::string qual(c.read_len, 'I');
    for (uint64_t i = 0; i < n; ++i) {
      fq += "@r" + std::to_string(i) + "\n";
      fq.append(reinterpret_cast<const char*>(reads.data() + i * c.read_len), c.read_len);
      fq += "\n+\n" + qual + "\n";
    }
    write_file(c.out + "/reads.fq", fq.data(), fq.size());
  }
  {
    char b[512];
    snprintf(b, sizeof b, "{\"genomes\": %u, \"length\": %u, \"seed\": %llu, \"blen\": %g, \"reads\": %llu, \"read_len\": %u, \"max_sub\": %g, "
             "\"nkmers\": %llu, \"nrows\": %u, \"nsubsets\": %u, \"nnodes\": %u, \"mean_bucket\": %.3f, \"build_s\": %.1f}\n",
             c.genomes, c.length, (unsigned long long)c.seed, c.blen, (unsigned long long)c.reads, c.read_len, c.max_sub,
             (unsigned long long)nkmers, nrows, nsubsets, nnodes, (double)nkmers / nrows, omp_get_wtime() - t0);
    write_file(c.out + "/workload.json", b, strlen(b));
    fputs(b, stderr);
  }
  return 0;
}

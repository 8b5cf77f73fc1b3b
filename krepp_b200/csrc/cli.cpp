// cli.cpp -- `krepp_b200 dist|place`: the reference's command line for the query path (ref src/krepp.cpp:593-716,738-763)
// over the C ABI of include/krepp_b200.h.  Same flags, defaults and validation as `krepp dist` / `krepp place`, same
// output framing; reads are reported in input order (the reference's order is unspecified).
//
// Pipeline: one producer thread parses FASTA/FASTQ straight into the pinned buffers of the next free batch slot and
// submits it (H2D + kernels + D2H are asynchronous on the slot's stream); one consumer thread waits for slots in
// submission order, formats them with --num-threads workers and writes the text.  Slots are spread round-robin over the
// GPUs given by --num-gpus / --devices (index replicated per GPU, no data-path collective: SURVEY.md section 8e mode A).
#include "../../include/krepp_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <limits>
#include <mutex>
#include <string>
#include <atomic>
#include <fcntl.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <unordered_map>
#include <vector>

namespace {

[[noreturn]] void error_exit(const std::string& msg)
{ // ref src/common.cpp:20-24
  fprintf(stderr, "[ERROR] %s\n", msg.c_str());
  exit(EXIT_FAILURE);
}

struct Options {
  std::string sub, query, index_dir, output_path, nwk_path, lineage_path;
  uint32_t hdist_th = 4, tau = 2, num_threads = 1, seed = 0;
  double chisq = 2.706, dist_max = std::numeric_limits<double>::quiet_NaN();
  bool multi = true, filter = false, summarize = false, tabular = false, verbose = false, seed_set = false;
  // sketch (ref src/krepp.hpp:35-46 set_sketch_defaults, src/krepp.cpp:516-540)
  uint32_t sk_k = 26, sk_w = 32, sk_h = 10, sk_m = 4, sk_r = 1, sdust_t = 0, sdust_w = 0;
  bool sk_frac = true, sk_w_set = false, sk_k_set = false, sk_h_set = false;
  // additions of this implementation
  std::vector<int> devices;
  bool shard_index = false; // --shard-index: every device holds one bucket-range shard of the table (SURVEY.md 8e mode B)
  uint32_t batch_reads = 1u << 18, slots_per_gpu = 3;
  uint64_t batch_bases = 64ull << 20;
};

const char* kUsage =
  "krepp_b200: B200-native query path of krepp (k-mer-based distance estimation & phylogenetic placement).\n"
  "Usage: krepp_b200 [--num-threads N] [--seed S] [--verbose] {dist|place} -i INDEX_DIR -q QUERY [options]\n"
  "       krepp_b200 [--num-threads N] seek -i,--sketch-path SKETCH_FILE -q QUERY [-o PATH] [--hdist-th N]\n"
  "       krepp_b200 [--seed S] sketch -i,--input-file FASTA -o,--output-path SKETCH_FILE [-k 26] [-w k+6] [-h k-16] [-m 4] [-r 1] [--frac/--no-frac]\n"
  "       krepp_b200 [--num-threads N] [--seed S] index -i,--input-file MAP.tsv -o,--index-dir DIR [-t,--nwk-file NWK] [-k 29] [-w k+6] [-h k-16] [-m 4] [-r 1] [--frac/--no-frac]\n"
  "  common:  -q,--query PATH   -i,--index-dir DIR   -o,--output-path PATH   --hdist-th N [4]   --chisq X [2.706]\n"
  "           --summarize/--no-summarize [false]\n"
  "  dist:    --dist-max X   --multi/--no-multi [true]   --filter/--no-filter [false]\n"
  "  place:   --tau N [2]   --multi/--no-multi [true]   --filter/--no-filter [true]   --tabular/--no-tabular [false]\n"
  "           -t,--nwk-file PATH  place on this tree instead of the index's backbone\n"
  "           -l,--lineage-file PATH  place on the taxonomy of a Greengenes/GTDB style lineage file (not together with -t)\n"
  "  GPU:     --num-gpus N [1] | --devices 0,1,..   --batch-reads N [262144]   --batch-bases N [67108864]   --slots N [3]\n"
  "           --shard-index   split the index by LSH bucket range over the devices instead of replicating it (for an index\n"
  "                           larger than one GPU's memory; lookups and hits travel between the GPUs by peer copies)\n";

bool exists(const std::string& p, bool dir)
{
  struct stat st;
  if (stat(p.c_str(), &st) != 0) return false;
  return dir ? S_ISDIR(st.st_mode) : S_ISREG(st.st_mode);
}

Options parse(int argc, char** argv)
{
  Options o;
  bool filter_set = false;
  int num_gpus = 0;
  std::vector<std::string> a(argv + 1, argv + argc);
  auto need = [&](size_t& i, const std::string& name) -> std::string {
    const size_t eq = a[i].find('=');
    if (a[i].rfind("--", 0) == 0 && eq != std::string::npos) return a[i].substr(eq + 1);
    if (i + 1 >= a.size()) error_exit("Option " + name + " requires an argument.");
    return a[++i];
  };
  for (size_t i = 0; i < a.size(); ++i) {
    std::string k = a[i];
    if (k.rfind("--", 0) == 0 && k.find('=') != std::string::npos) k = k.substr(0, k.find('='));
    const bool build_sub = o.sub == "sketch" || o.sub == "index"; // the subcommands that set up a new LSH geometry
    if (k == "dist" || k == "place" || k == "seek" || k == "sketch" || k == "index") { if (!o.sub.empty()) error_exit("Only one subcommand may be given."); o.sub = k; }
    else if (k == "inspect") error_exit("Subcommand '" + k + "' is not part of the GPU query path; use the reference krepp binary for it.");
    else if (k == "--help") { fputs(kUsage, stdout); exit(0); }
    else if (k == "--verbose") o.verbose = true;
    else if (k == "--no-verbose") o.verbose = false;
    else if (k == "--seed") { o.seed = (uint32_t)strtoul(need(i, k).c_str(), nullptr, 10); o.seed_set = true; }
    else if (build_sub && (k == "-i" || k == "--input-file")) o.index_dir = need(i, k);
    else if (o.sub == "index" && (k == "-o" || k == "--index-dir")) o.output_path = need(i, k);
    else if (build_sub && (k == "-k" || k == "--kmer-len")) { o.sk_k = (uint32_t)atoi(need(i, k).c_str()); o.sk_k_set = true; if (o.sk_k < 19 || o.sk_k > 31) error_exit("--kmer-len: Value " + std::to_string(o.sk_k) + " not in range [19 - 31]"); }
    else if (build_sub && (k == "-w" || k == "--win-len")) { o.sk_w = (uint32_t)atoi(need(i, k).c_str()); o.sk_w_set = true; }
    else if (build_sub && (k == "-h" || k == "--num-positions")) { o.sk_h = (uint32_t)atoi(need(i, k).c_str()); o.sk_h_set = true; }
    else if (build_sub && (k == "-m" || k == "--modulo-lsh")) { o.sk_m = (uint32_t)atoi(need(i, k).c_str()); if (!o.sk_m) error_exit("--modulo-lsh: Number less or equal to 0"); }
    else if (build_sub && (k == "-r" || k == "--residue-lsh")) o.sk_r = (uint32_t)atoi(need(i, k).c_str());
    else if (build_sub && k == "--frac") o.sk_frac = true;
    else if (build_sub && k == "--no-frac") o.sk_frac = false;
    else if (build_sub && k == "--sdust-t") o.sdust_t = (uint32_t)atoi(need(i, k).c_str());
    else if (build_sub && k == "--sdust-w") o.sdust_w = (uint32_t)atoi(need(i, k).c_str());
    else if (k == "--num-threads") o.num_threads = (uint32_t)strtoul(need(i, k).c_str(), nullptr, 10);
    else if (k == "-q" || k == "--query") o.query = need(i, k);
    else if (k == "-i" || k == "--index-dir" || k == "--sketch-path") o.index_dir = need(i, k);
    else if (k == "-o" || k == "--output-path") o.output_path = need(i, k);
    else if (k == "--hdist-th") { const std::string v = need(i, k); if (v.empty() || v[0] == '-') error_exit("--hdist-th: Number less than 0"); o.hdist_th = (uint32_t)strtoul(v.c_str(), nullptr, 10); }
    else if (k == "--chisq") { o.chisq = atof(need(i, k).c_str()); if (!(o.chisq > 0)) error_exit("--chisq: Number less or equal to 0"); }
    else if (k == "--summarize") o.summarize = true;
    else if (k == "--no-summarize") o.summarize = false;
    else if (k == "--dist-max") { o.dist_max = atof(need(i, k).c_str()); if (!(o.dist_max >= 1e-8 && o.dist_max <= 0.33)) error_exit("--dist-max: Value not in range [1e-08 - 0.33]"); }
    else if (k == "--multi") o.multi = true;
    else if (k == "--no-multi") o.multi = false;
    else if (k == "--filter") { o.filter = true; filter_set = true; }
    else if (k == "--no-filter") { o.filter = false; filter_set = true; }
    else if (k == "--tau") { const std::string v = need(i, k); if (v.empty() || v[0] == '-') error_exit("--tau: Number less than 0"); o.tau = (uint32_t)strtoul(v.c_str(), nullptr, 10); }
    else if (k == "--tabular") o.tabular = true;
    else if (k == "--no-tabular") o.tabular = false;
    else if (k == "-t" || k == "--nwk-file") o.nwk_path = need(i, k);
    else if (k == "-l" || k == "--lineage-file") o.lineage_path = need(i, k);
    else if (k == "--num-gpus") num_gpus = atoi(need(i, k).c_str());
    else if (k == "--devices") { const std::string v = need(i, k); size_t p = 0; while (p < v.size()) { o.devices.push_back(atoi(v.c_str() + p)); p = v.find(',', p); if (p == std::string::npos) break; ++p; } }
    else if (k == "--batch-reads") o.batch_reads = (uint32_t)strtoul(need(i, k).c_str(), nullptr, 10);
    else if (k == "--batch-bases") o.batch_bases = strtoull(need(i, k).c_str(), nullptr, 10);
    else if (k == "--shard-index") o.shard_index = true;
    else if (k == "--slots") o.slots_per_gpu = (uint32_t)strtoul(need(i, k).c_str(), nullptr, 10);
    else error_exit("The following argument was not expected: " + a[i]);
  }
  if (o.sub.empty()) { fputs(kUsage, stderr); error_exit("A subcommand is required"); }
  if (o.sub == "index") { // ref src/krepp.cpp:560-591 (defaults: set_index_defaults src/krepp.hpp:47-58)
    if (o.index_dir.empty()) error_exit("--input-file is required");
    if (o.output_path.empty()) error_exit("--index-dir is required");
    if (!exists(o.index_dir, false)) error_exit("--input-file: File does not exist: " + o.index_dir);
    if (!o.nwk_path.empty() && !exists(o.nwk_path, false)) error_exit("--nwk-file: File does not exist: " + o.nwk_path);
    if (!o.sk_k_set) o.sk_k = 29;
    if (!o.sk_h_set) o.sk_h = 13;
    if (!o.sk_w_set) { o.sk_w = o.sk_k + 6; o.sk_h = o.sk_k - 16; }
    if (o.sdust_t && o.sdust_w) error_exit("--sdust-t / --sdust-w (dustmasker) are not part of the GPU path; build such a library with the reference binary");
    if (!o.lineage_path.empty() || !o.query.empty()) error_exit("The following argument was not expected for index");
    if (o.devices.empty()) o.devices.push_back(0);
    return o;
  }
  if (o.sub == "sketch") { // ref src/krepp.cpp:516-540
    if (o.index_dir.empty()) error_exit("--input-file is required");
    if (o.output_path.empty()) error_exit("--output-path is required");
    if (!exists(o.index_dir, false)) error_exit("--input-file: File does not exist: " + o.index_dir);
    if (!o.sk_w_set) { o.sk_w = o.sk_k + 6; o.sk_h = o.sk_k - 16; }
    if (o.sdust_t && o.sdust_w) error_exit("--sdust-t / --sdust-w (dustmasker) are not part of the GPU path; build such a sketch with the reference binary");
    if (o.devices.empty()) o.devices.push_back(0);
    return o;
  }
  if (o.query.empty()) error_exit("--query is required");
  if (o.index_dir.empty()) error_exit(o.sub == "seek" ? "--sketch-path is required" : "--index-dir is required");
  if (!exists(o.query, false)) error_exit("--query: File does not exist: " + o.query);
  if (o.sub == "seek") { // ref src/krepp.cpp:543-559: -q, -i, -o, --hdist-th and nothing else
    if (!exists(o.index_dir, false)) error_exit("--sketch-path: File does not exist: " + o.index_dir);
    if (o.tabular || o.summarize || filter_set || o.shard_index || !o.nwk_path.empty() || !o.lineage_path.empty()) error_exit("The following argument was not expected for seek");
  } else if (!exists(o.index_dir, true)) error_exit("--index-dir: Directory does not exist: " + o.index_dir);
  if (o.sub == "place" && !filter_set) o.filter = true; // ref src/krepp.cpp:614
  if (o.sub == "dist" && (o.tabular || !o.nwk_path.empty() || !o.lineage_path.empty())) error_exit("The following argument was not expected for dist");
  if (!o.lineage_path.empty() && !o.nwk_path.empty()) error_exit("--nwk-file excludes --lineage-file"); // ref src/krepp.cpp:598-603 (CLI11 excludes)
  if (!o.lineage_path.empty() && !exists(o.lineage_path, false)) error_exit("--lineage-file: File does not exist: " + o.lineage_path);
  if (!o.nwk_path.empty() && !exists(o.nwk_path, false)) error_exit("--nwk-file: File does not exist: " + o.nwk_path);
  if (o.devices.empty()) for (int d = 0; d < (num_gpus > 0 ? num_gpus : 1); ++d) o.devices.push_back(d);
  if (!o.num_threads) o.num_threads = 1;
  if (!o.batch_reads || !o.batch_bases || !o.slots_per_gpu) error_exit("--batch-reads, --batch-bases and --slots must be positive");
  return o;
}

struct Slot {
  krepp_batch_t* batch = nullptr;
  char* bases = nullptr;
  uint64_t* offsets = nullptr;
  std::vector<char> names;
  std::vector<uint64_t> name_off;
  uint32_t n = 0;
  int gpu = 0; // index into devices
};

template <class T>
class Channel {
public:
  void push(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
  bool pop(T& v)
  {
    std::unique_lock<std::mutex> l(m_);
    cv_.wait(l, [&] { return !q_.empty() || closed_; });
    if (q_.empty()) return false;
    v = q_.front();
    q_.pop_front();
    return true;
  }
  void close() { { std::lock_guard<std::mutex> l(m_); closed_ = true; } cv_.notify_all(); }

private:
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<T> q_;
  bool closed_ = false;
};

void check(int rc) { if (rc != KREPP_OK) error_exit(krepp_last_error()); }

} // namespace

// ------------------------------------------------------------------------------------------------ --shard-index (mode B)
// All shards live in this process, one per entry of --devices (an entry may repeat: several shards on one GPU).  A round takes
// one sub-batch of reads per device and walks the three phases of include/krepp_b200.h "bucket-range shards", one thread per
// device inside a phase and a join between phases; runs of lookup tuples and of hit entries change device by
// krepp_device_copy (peer copies).  Output order is input order.
struct ShardDev {
  int dev = 0;
  krepp_index_t* ix = nullptr;
  krepp_batch_t* slot = nullptr;
  char* h_bases = nullptr; uint64_t* h_offsets = nullptr;
  std::vector<char> names; std::vector<uint64_t> name_off;
  uint32_t n = 0; uint64_t nb = 0;
  void *d_bases = nullptr, *d_offsets = nullptr, *d_tuples = nullptr, *d_rowbegin = nullptr, *recv_tuples = nullptr, *recv_rb = nullptr, *d_hits = nullptr,
       *recv_hits = nullptr;
  uint64_t cap_tuples = 0, cap_recv_tuples = 0, cap_hits = 0, cap_recv_hits = 0;
  std::vector<uint64_t> so, ho;
  uint32_t row0 = 0, row1 = 0;
  std::vector<char> text; size_t text_len = 0;
  std::vector<double> w;
};

template <class F>
static void each_device(size_t n, F&& f)
{
  std::vector<std::thread> th;
  for (size_t g = 1; g < n; ++g) th.emplace_back(f, g);
  f(0);
  for (auto& t : th) t.join();
}

static void grow(int dev, void*& p, uint64_t& cap, uint64_t want_items)
{ // 16-byte items; contents are not kept
  if (want_items <= cap) return;
  krepp_device_free(dev, p);
  cap = want_items + want_items / 8 + 4096;
  check(krepp_device_alloc(dev, 16 * cap, &p));
}

// the index with the tree the command line asks for: -l (lineages) or -t (ref src/krepp.cpp:742-748)
static int open_for(const Options& o, int dev, uint32_t shard, uint32_t nshards, krepp_index_t** out)
{
  if (!o.lineage_path.empty()) return krepp_index_open_lineages(o.index_dir.c_str(), dev, shard, nshards, o.lineage_path.c_str(), out);
  return krepp_index_open_tree(o.index_dir.c_str(), dev, shard, nshards, o.nwk_path.empty() ? nullptr : o.nwk_path.c_str(), out);
}

static int run_sharded(const Options& o, const krepp_params_t& p, bool place, const std::string& invocation, FILE* out)
{
  const size_t N = o.devices.size();
  std::vector<ShardDev> D(N);
  each_device(N, [&](size_t g) { D[g].dev = o.devices[g]; check(open_for(o, D[g].dev, (uint32_t)g, (uint32_t)N, &D[g].ix)); });
  krepp_index_info_t info;
  check(krepp_index_info(D[0].ix, &info));
  std::vector<uint32_t> splits(N + 1);
  { krepp_shard_info_t si; check(krepp_index_shard_info(D[0].ix, &si, splits.data(), (uint32_t)N + 1)); }
  for (size_t g = 0; g < N; ++g) {
    ShardDev& d = D[g];
    d.row0 = splits[g]; d.row1 = splits[g + 1];
    check(krepp_batch_create(d.ix, &p, o.batch_reads, o.batch_bases, &d.slot));
    check(krepp_batch_set_output(d.slot, place ? (KREPP_OUT_PLACEMENTS | KREPP_OUT_SUMMARIES) : KREPP_OUT_DIST));
    check(krepp_batch_host_buffers(d.slot, &d.h_bases, &d.h_offsets));
    d.names.resize(64ull * o.batch_reads); d.name_off.resize(o.batch_reads);
    check(krepp_device_alloc(d.dev, o.batch_bases + 64, &d.d_bases));
    check(krepp_device_alloc(d.dev, 8ull * (o.batch_reads + 1), &d.d_offsets));
    check(krepp_device_alloc(d.dev, 4ull * (info.nrows + 1), &d.d_rowbegin));
    check(krepp_device_alloc(d.dev, 4ull * N * (d.row1 - d.row0 + 1), &d.recv_rb));
    d.so.assign(N + 1, 0); d.ho.assign(N + 1, 0);
    d.text.resize(1 << 20);
  }
  std::vector<char> text(1 << 20);
  { size_t n = krepp_format_header(D[0].ix, &p, o.tabular, invocation.c_str(), text.data(), text.size());
    if (n > text.size()) { text.resize(n); n = krepp_format_header(D[0].ix, &p, o.tabular, invocation.c_str(), text.data(), text.size()); }
    if (n && fwrite(text.data(), 1, n, out) != n) error_exit("Failed to write the output"); }
  krepp_reader_t* reader = nullptr;
  check(krepp_reader_open(o.query.c_str(), &reader));
  uint64_t total_queries = 0;
  std::vector<double> wcount(info.nnodes + 1, 0.0);
  int has_previous = 0, eof = 0;
  const bool jplace = place && !o.tabular && !p.summarize;
  while (!eof) {
    uint64_t round_reads = 0;
    for (size_t g = 0; g < N; ++g) { // one sub-batch per device, in input order
      ShardDev& d = D[g];
      d.n = 0; d.nb = 0;
      if (!eof) check(krepp_reader_next(reader, d.h_bases, o.batch_bases, d.h_offsets, o.batch_reads, d.names.data(), d.names.size(), d.name_off.data(), &d.n, &eof));
      if (!d.n) d.h_offsets[0] = 0;
      d.nb = d.h_offsets[d.n];
      round_reads += d.n;
    }
    if (!round_reads) break;
    total_queries += round_reads;
    // phase 1, home: reads -> tuples grouped by row
    each_device(N, [&](size_t g) {
      ShardDev& d = D[g];
      check(krepp_device_copy(d.dev, d.d_bases, KREPP_DEVICE_NONE, d.h_bases, d.nb));
      check(krepp_device_copy(d.dev, d.d_offsets, KREPP_DEVICE_NONE, d.h_offsets, 8ull * (d.n + 1)));
      for (;;) {
        if (!d.cap_tuples) grow(d.dev, d.d_tuples, d.cap_tuples, std::max<uint64_t>((uint64_t)o.batch_bases / 2, 4096));
        const int rc = krepp_shard_lookup(d.slot, (const char*)d.d_bases, (const uint64_t*)d.d_offsets, d.n, d.nb + 64, d.d_tuples, d.cap_tuples, (uint32_t*)d.d_rowbegin, d.so.data());
        if (rc == KREPP_ERR_CAPACITY && d.so[N] > d.cap_tuples) { grow(d.dev, d.d_tuples, d.cap_tuples, d.so[N]); continue; }
        check(rc);
        break;
      }
    });
    // phase 2, owner: every sender's run against this shard
    each_device(N, [&](size_t g) {
      ShardDev& d = D[g];
      const uint64_t nloc = (uint64_t)d.row1 - d.row0 + 1;
      uint64_t total = 0;
      for (size_t s = 0; s < N; ++s) total += D[s].so[g + 1] - D[s].so[g];
      grow(d.dev, d.recv_tuples, d.cap_recv_tuples, total);
      std::vector<const void*> tp(N), rp(N);
      uint64_t at = 0;
      for (size_t s = 0; s < N; ++s) {
        const uint64_t cnt = D[s].so[g + 1] - D[s].so[g];
        check(krepp_device_copy(d.dev, (char*)d.recv_tuples + 16 * at, D[s].dev, (const char*)D[s].d_tuples + 16 * D[s].so[g], 16 * cnt));
        check(krepp_device_copy(d.dev, (char*)d.recv_rb + 4 * nloc * s, D[s].dev, (const char*)D[s].d_rowbegin + 4ull * d.row0, 4 * nloc));
        tp[s] = (const char*)d.recv_tuples + 16 * at;
        rp[s] = (const char*)d.recv_rb + 4 * nloc * s;
        at += cnt;
      }
      for (;;) {
        if (!d.cap_hits) grow(d.dev, d.d_hits, d.cap_hits, std::max<uint64_t>(96ull * o.batch_reads, 65536));
        const int rc = krepp_shard_join(d.slot, (uint32_t)N, tp.data(), reinterpret_cast<const uint32_t* const*>(rp.data()), d.d_hits, d.cap_hits, d.ho.data());
        if (rc == KREPP_ERR_CAPACITY && d.ho[N] > d.cap_hits) { grow(d.dev, d.d_hits, d.cap_hits, d.ho[N]); continue; }
        check(rc);
        break;
      }
    });
    // phase 3, home again: the batch's hit entries from all owners -> results -> text
    each_device(N, [&](size_t g) {
      ShardDev& d = D[g];
      uint64_t total = 0;
      for (size_t ow = 0; ow < N; ++ow) total += D[ow].ho[g + 1] - D[ow].ho[g];
      grow(d.dev, d.recv_hits, d.cap_recv_hits, total);
      uint64_t at = 0;
      for (size_t ow = 0; ow < N; ++ow) {
        const uint64_t cnt = D[ow].ho[g + 1] - D[ow].ho[g];
        check(krepp_device_copy(d.dev, (char*)d.recv_hits + 16 * at, D[ow].dev, (const char*)D[ow].d_hits + 16 * D[ow].ho[g], 16 * cnt));
        at += cnt;
      }
      check(krepp_shard_finish(d.slot, d.recv_hits, total));
      krepp_results_t res;
      check(krepp_batch_wait(d.slot, &res));
      double* w = nullptr;
      if (p.summarize) { d.w.assign(info.nnodes + 1, 0.0); w = d.w.data(); }
      for (;;) {
        int prev = 0;
        const size_t n = place ? krepp_format_place(d.ix, &p, &res, d.names.data(), d.name_off.data(), o.tabular, &prev, w, d.text.data(), d.text.size())
                               : krepp_format_dist(d.ix, &p, &res, d.names.data(), d.name_off.data(), w, d.text.data(), d.text.size());
        if (n <= d.text.size()) { d.text_len = n; break; }
        d.text.resize(n + n / 4);
        if (w) d.w.assign(info.nnodes + 1, 0.0);
      }
    });
    for (size_t g = 0; g < N; ++g) {
      ShardDev& d = D[g];
      if (p.summarize) { for (uint32_t se = 0; se <= info.nnodes; ++se) wcount[se] += d.w[se]; continue; }
      if (!d.text_len) continue;
      if (jplace && has_previous) fwrite(",\n", 1, 2, out); // ref src/krepp.cpp:476-481
      if (fwrite(d.text.data(), 1, d.text_len, out) != d.text_len) error_exit("Failed to write the output");
      has_previous = 1;
    }
  }
  krepp_reader_close(reader);
  { size_t n = krepp_format_footer(D[0].ix, &p, o.tabular, wcount.data(), total_queries, invocation.c_str(), text.data(), text.size());
    if (n > text.size()) { text.resize(n); n = krepp_format_footer(D[0].ix, &p, o.tabular, wcount.data(), total_queries, invocation.c_str(), text.data(), text.size()); }
    if (n && fwrite(text.data(), 1, n, out) != n) error_exit("Failed to write the output"); }
  fflush(out);
  fprintf(stderr, "Total number of sequences queried: %llu\n", (unsigned long long)total_queries);
  for (ShardDev& d : D) {
    krepp_batch_destroy(d.slot);
    for (void* q : {d.d_bases, d.d_offsets, d.d_tuples, d.d_rowbegin, d.recv_tuples, d.recv_rb, d.d_hits, d.recv_hits}) krepp_device_free(d.dev, q);
    krepp_index_close(d.ix);
  }
  return 0;
}

// Every sequence of a FASTA/FASTQ file (plain or gzip) in one batch: what RSeq hands extract_mers sequence by sequence
// (ref src/rqseq.cpp:39-49).  A batch that ends before the input does is read again with more room.
static uint32_t read_whole_file(const std::string& path, std::vector<char>& bases, std::vector<uint64_t>& offsets)
{
  struct stat st;
  if (stat(path.c_str(), &st) != 0) error_exit("Failed to open the file at " + path); // ref src/rqseq.cpp:36-38
  const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
  uint64_t cap_bases = (uint64_t)st.st_size * (gz ? 8 : 1) + (1u << 20);
  uint32_t cap_seqs = 1u << 16;
  std::vector<char> names;
  std::vector<uint64_t> name_off;
  uint32_t n = 0;
  for (;;) {
    bases.resize(cap_bases + 64); offsets.resize((size_t)cap_seqs + 1); names.resize(64ull * cap_seqs); name_off.resize(cap_seqs);
    krepp_reader_t* rd = nullptr;
    if (krepp_reader_open(path.c_str(), &rd) != KREPP_OK) error_exit("Failed to open the file at " + path);
    int eof = 0;
    const int rc = krepp_reader_next(rd, bases.data(), cap_bases, offsets.data(), cap_seqs, names.data(), names.size(), name_off.data(), &n, &eof);
    krepp_reader_close(rd);
    if (rc == KREPP_OK && eof) break;
    if (rc != KREPP_OK && rc != KREPP_ERR_CAPACITY) error_exit(krepp_last_error());
    if (cap_bases > (1ull << 40)) error_exit("the input does not fit in memory");
    cap_bases *= 2; cap_seqs *= 2;
  }
  return n;
}

static krepp_index_t* open_geometry(const Options& o)
{
  krepp_index_t* geom = nullptr;
  if (krepp_geometry_open(o.sk_k, o.sk_w, o.sk_h, o.sk_m, o.sk_r, o.sk_frac ? 1 : 0, o.seed_set ? (int64_t)o.seed : -1, o.devices[0], &geom) != KREPP_OK) {
    const std::string msg = krepp_last_error();
    if (msg.find("(-") != std::string::npos || msg.find("h must be") != std::string::npos) { fprintf(stderr, "%s\n", msg.c_str()); error_exit("Invalid configuration!"); } // ref src/krepp.hpp:59-85
    error_exit(msg);
  }
  return geom;
}

// `krepp index` (ref src/krepp.cpp:131-309,723-736): reference genomes + guide tree -> a library directory.  Host threads read
// (and inflate) the genomes; each genome's leaf table and rho are computed on the GPU and stay there; the union over the tree is
// one sort on the GPU; the colour record and the files are written by the host (krepp_builder_*, include/krepp_b200.h).
static int run_index(const Options& o)
{
  fprintf(stderr, "Reading the tree and initializing the index...\n");
  const auto t0 = std::chrono::system_clock::now();
  krepp_index_t* geom = open_geometry(o);
  std::vector<std::string> names;
  std::vector<std::pair<std::string, std::string>> todo; // (name, path), one per distinct name: a later line overrides (ref src/krepp.cpp:157-158)
  std::unordered_map<std::string, size_t> todo_at;
  { // IndexMultiple::read_input_file ref src/krepp.cpp:147-162
    FILE* f = fopen(o.index_dir.c_str(), "r");
    if (!f) error_exit("Error opening " + o.index_dir);
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    fclose(f);
    for (size_t at = 0; at < text.size();) {
      size_t e = text.find('\n', at);
      if (e == std::string::npos) e = text.size();
      const std::string line = text.substr(at, e - at);
      at = e + 1;
      const size_t tab = line.find('\t');
      if (tab == std::string::npos) error_exit("Failed to read the reference name to path/URL mapping!");
      const size_t tab2 = line.find('\t', tab + 1);
      const std::string name = line.substr(0, tab), path = line.substr(tab + 1, tab2 == std::string::npos ? std::string::npos : tab2 - tab - 1);
      if (path.empty()) error_exit("Failed to read the reference name to path/URL mapping!");
      names.push_back(name);
      const auto known = todo_at.find(name);
      if (known != todo_at.end()) todo[known->second].second = path;
      else { todo_at.emplace(name, todo.size()); todo.emplace_back(name, path); }
    }
  }
  std::string nwk;
  if (o.nwk_path.empty()) fprintf(stderr, "No tree has given as a guide, the color index could be suboptimal.\n");
  else {
    FILE* f = fopen(o.nwk_path.c_str(), "r");
    if (!f) error_exit("Error opening " + o.nwk_path);
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) nwk.append(buf, got);
    fclose(f);
  }
  std::vector<const char*> name_ptrs;
  for (const auto& nm : names) name_ptrs.push_back(nm.c_str());
  krepp_builder_t* b = nullptr;
  check(krepp_builder_create(geom, o.nwk_path.empty() ? nullptr : nwk.c_str(), name_ptrs.data(), (uint32_t)name_ptrs.size(), &b));
  const auto t1 = std::chrono::system_clock::now();
  fprintf(stderr, "Building the index...\n");
  { // readers in parallel, one genome on the GPU at a time
    std::atomic<size_t> next{0};
    std::mutex gpu;
    uint32_t done = 0;
    auto work = [&] {
      std::vector<char> bases;
      std::vector<uint64_t> offsets;
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= todo.size()) return;
        if (krepp_builder_has_leaf(b, todo[i].first.c_str()) != 1) continue; // not on the tree: never visited (ref src/krepp.cpp:248-252)
        const uint32_t n = read_whole_file(todo[i].second, bases, offsets);
        std::lock_guard<std::mutex> lock(gpu);
        uint64_t nk = 0;
        double rho = 0;
        check(krepp_builder_add_genome(b, todo[i].first.c_str(), bases.data(), offsets.data(), n, &nk, &rho));
        ++done;
        if (o.verbose) fprintf(stderr, "Leaf node: %s\tsize: %llu\trho: %g\tprogress: %u/%u\n", todo[i].first.c_str(), (unsigned long long)nk, rho, done, krepp_builder_nleaves(b));
      }
    };
    std::vector<std::thread> th;
    const uint32_t nt = (uint32_t)std::max<size_t>(1, std::min<size_t>(o.num_threads, todo.size()));
    for (uint32_t t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  uint64_t nk = 0, nsets = 0;
  const auto t2 = std::chrono::system_clock::now();
  check(krepp_builder_union(b, &nk, &nsets));
  const auto t3 = std::chrono::system_clock::now();
  if (o.verbose) {
    const std::chrono::duration<float> a = t1 - t0, g = t2 - t1, u = t3 - t2;
    fprintf(stderr, "\n[stages] set-up (device, tree) %.3f s, genomes (read + leaf tables on the GPU) %.3f s, union on the GPU %.3f s", a.count(), g.count(), u.count());
  }
  const std::chrono::duration<float> es_b = std::chrono::system_clock::now() - t0;
  fprintf(stderr, "\nFinished indexing, elapsed: %g sec\n", es_b.count());
  uint32_t nsub = 0;
  check(krepp_builder_write(b, o.output_path.c_str(), o.seed_set ? o.seed : 0, &nk, &nsub));
  if (o.nwk_path.empty()) fprintf(stderr, "Skipped saving a backbone for the index!\n");
  if (o.verbose) fprintf(stderr, "k-mers: %llu, distinct reference sets: %llu, colours: %u\n", (unsigned long long)nk, (unsigned long long)nsets, nsub);
  const std::chrono::duration<float> es_s = std::chrono::system_clock::now() - t0 - es_b;
  fprintf(stderr, "Done converting & saving, elapsed: %g sec\n", es_s.count());
  const auto t4 = std::chrono::system_clock::now();
  krepp_builder_destroy(b);
  krepp_index_close(geom);
  if (o.verbose) { const std::chrono::duration<float> d = std::chrono::system_clock::now() - t4; fprintf(stderr, "[stages] releasing the builder %.3f s\n", d.count()); }
  return 0;
}

// `krepp sketch` (ref src/krepp.cpp:110-128,773-782): one FASTA/FASTQ file -> the sketch file `seek` reads.
static int run_sketch(const Options& o)
{
  fprintf(stderr, "Initializing the sketch...\n");
  const auto t0 = std::chrono::system_clock::now();
  krepp_index_t* geom = open_geometry(o);
  std::vector<char> bases;
  std::vector<uint64_t> offsets;
  const uint32_t n = read_whole_file(o.index_dir, bases, offsets);
  uint64_t nk = 0;
  double rho = 0;
  check(krepp_sketch_write(geom, bases.data(), offsets.data(), n, o.output_path.c_str(), &nk, &rho));
  fprintf(stderr, "Total number of k-mers included in the sketch: %llu\n", (unsigned long long)nk);
  fprintf(stderr, "Subsampling rate (rho) is: %g\n", rho);
  const std::chrono::duration<float> es = std::chrono::system_clock::now() - t0;
  fprintf(stderr, "Done sketching & saving, elapsed: %g sec\n", es.count());
  krepp_index_close(geom);
  return 0;
}

int main(int argc, char** argv)
{
  fprintf(stderr, "krepp_b200 version: v0.8.3+b200\n");
  const Options o = parse(argc, argv);
  std::string invocation;
  for (int i = 0; i < argc; ++i) invocation += std::string(argv[i]) + (i + 1 < argc ? " " : "");
  const auto tstart = std::chrono::system_clock::now();
  { std::time_t t = std::chrono::system_clock::to_time_t(tstart); fprintf(stderr, "Invocation: %s\n%s", invocation.c_str(), std::ctime(&t)); }

  if (o.sub == "sketch") return run_sketch(o);
  if (o.sub == "index") {
    const int rc = run_index(o);
    const auto tend = std::chrono::system_clock::now();
    if (o.verbose) { const std::chrono::duration<float> d = tend - tstart; fprintf(stderr, "[stages] main() %.3f s\n", d.count()); }
    std::time_t t = std::chrono::system_clock::to_time_t(tend);
    fprintf(stderr, "%s", std::ctime(&t)); // ref src/krepp.cpp:795-797
    return rc;
  }
  const bool place = o.sub == "place", seek = o.sub == "seek";
  krepp_params_t p;
  krepp_params_default(&p, place ? 1 : 0);
  p.hdist_th = o.hdist_th; p.chisq = o.chisq; p.dist_max = o.dist_max; p.tau = o.tau;
  p.no_filter = o.filter ? 0 : 1; p.multi = o.multi ? 1 : 0; p.summarize = o.summarize ? 1 : 0;
  if (place && o.hdist_th < o.tau) { // ref src/krepp.hpp:192-199
    fprintf(stderr, "The threshold tau must be less than HD threshold --hdist-th!\n");
    error_exit("Invalid configuration!");
  }

  fprintf(stderr, seek ? "Loading the sketch...\n" : place ? "Loading the index and the backbone tree...\n" : "Loading the index and initializing...\n");
  if (o.shard_index) {
    FILE* sout = stdout;
    if (!o.output_path.empty()) { sout = fopen(o.output_path.c_str(), "wb"); if (!sout) error_exit("Failed to open the output file at " + o.output_path); }
    const auto tq = std::chrono::system_clock::now();
    run_sharded(o, p, place, invocation, sout);
    if (sout != stdout) fclose(sout);
    const std::chrono::duration<float> es = std::chrono::system_clock::now() - tq;
    fprintf(stderr, place ? "Done placing queries, elapsed: %g sec\n" : "Done estimating distances, elapsed: %g sec\n", es.count());
    return 0;
  }
  std::vector<krepp_index_t*> index(o.devices.size(), nullptr);
  { // one replica of the index image per GPU, uploaded concurrently
    std::vector<std::thread> th;
    std::vector<std::string> err(o.devices.size());
    for (size_t g = 0; g < o.devices.size(); ++g)
      th.emplace_back([&, g] { if ((seek ? krepp_sketch_open(o.index_dir.c_str(), o.devices[g], &index[g]) : open_for(o, o.devices[g], 0, 1, &index[g])) != KREPP_OK) err[g] = krepp_last_error(); });
    for (auto& t : th) t.join();
    for (auto& e : err) if (!e.empty()) error_exit(e);
  }
  krepp_index_info_t info;
  check(krepp_index_info(index[0], &info));

  FILE* out = stdout;
  if (!o.output_path.empty()) { out = fopen(o.output_path.c_str(), "wb"); if (!out) error_exit("Failed to open the output file at " + o.output_path); }
  std::vector<char> obuf(8 << 20);
  setvbuf(out, obuf.data(), _IOFBF, obuf.size());

  std::vector<Slot> slots(o.devices.size() * o.slots_per_gpu);
  Channel<Slot*> free_q, busy_q;
  for (size_t i = 0; i < slots.size(); ++i) {
    Slot& s = slots[i];
    s.gpu = (int)(i % o.devices.size());
    check(krepp_batch_create(index[s.gpu], &p, o.batch_reads, o.batch_bases, &s.batch));
    // the writers never read the histograms, and `dist` needs only the rows it prints, which the device selects and rounds
    check(krepp_batch_set_output(s.batch, seek ? KREPP_OUT_SEEK : place ? (KREPP_OUT_PLACEMENTS | KREPP_OUT_SUMMARIES) : KREPP_OUT_DIST));
    check(krepp_batch_host_buffers(s.batch, &s.bases, &s.offsets));
    { // result buffers sized once, before the clock starts: a large-bucket index (many genomes) gives ~19 records, ~56 hit entries
      // and, when placing, ~100 tree nodes per 150 bp read; the buffers still grow if a batch needs more
      const uint64_t n = o.batch_reads, big = info.size_biased_bucket > 24.0 ? 1 : 0;
      check(krepp_batch_reserve(s.batch, (big ? 24 : 6) * n, (big ? 96 : 24) * n, place ? (big ? 128 : 48) * n : 0, place ? (big ? 12 : 6) * n : 0));
    }
    s.names.resize(64ull * o.batch_reads);
    s.name_off.resize(o.batch_reads);
    free_q.push(&s);
  }
  for (size_t g = 0; g < o.devices.size(); ++g) { // one tiny batch per GPU loads the kernels (the CUDA runtime loads a kernel at its first launch)
    Slot& s = slots[g];
    static const char prime[] = "ACGTTGCAAGCTTAGGCATCGATCGGATTACAGGCTTAACGTAGCTAGGCTAACGGTATCGATCGTAGCTAGCTAGGATCCGATTACGATCGGCTAGCTAGGCTAACGTACGATCGTAGCTAGCTAACGGATCGATCGTAGCTAGCATCGATCG";
    const uint64_t offs[2] = {0, sizeof prime - 1};
    krepp_results_t res;
    check(krepp_batch_submit(s.batch, prime, offs, 1));
    check(krepp_batch_wait(s.batch, &res));
  }

  fprintf(stderr, seek ? "Seeking query sequences in the sktech...\n" : place ? (o.lineage_path.empty() ? "Placing given sequences on the backbone tree...\n" : "Placing given sequences on the taxonomic lineage...\n") : "Estimating distances between given sequences and references...\n");
  const auto tquery = std::chrono::system_clock::now();
  std::vector<char> text(1 << 20);
  auto emit = [&](size_t n) { if (n && fwrite(text.data(), 1, n, out) != n) error_exit("Failed to write the output"); };
  if (!seek) { // header / begin_jplace (`seek` prints rows only: the reference never writes the header it builds, src/krepp.cpp:321-324)
    size_t n = krepp_format_header(index[0], &p, o.tabular, invocation.c_str(), text.data(), text.size());
    if (n > text.size()) { text.resize(n); n = krepp_format_header(index[0], &p, o.tabular, invocation.c_str(), text.data(), text.size()); }
    emit(n);
  }

  uint64_t total_queries = 0;
  std::vector<double> wcount(info.nnodes + 1, 0.0);

  // The text of a batch is formatted by --num-threads workers, each into its own buffer.  When the output is a regular file the
  // workers also write: once all of them know their lengths, every worker pwrite()s its part at its own offset, so neither the
  // formatting nor the copy into the page cache is serial.  Otherwise (a pipe, a terminal) the parts go to a writer thread of
  // their own through one of two buffer sets, so that the next batch is formatted while this one is written.
  struct TextSet { std::vector<std::vector<char>> part; std::vector<size_t> len; };
  const uint32_t T = o.num_threads;
  std::vector<TextSet> sets(2);
  Channel<TextSet*> sets_free, sets_full;
  for (TextSet& ts : sets) { ts.part.assign(T, std::vector<char>(1 << 20)); ts.len.assign(T, 0); sets_free.push(&ts); }
  const bool jplace = place && !o.tabular && !p.summarize;
  bool direct = false;
  off_t file_off = 0;
  { struct stat st; fflush(out); direct = !(getenv("KREPP_OUT_DIRECT") && !atoi(getenv("KREPP_OUT_DIRECT"))) && fstat(fileno(out), &st) == 0 && S_ISREG(st.st_mode) && !(fcntl(fileno(out), F_GETFL) & O_APPEND); if (direct) file_off = lseek(fileno(out), 0, SEEK_CUR); if (file_off < 0) direct = false; }
  double t_read = 0, t_wait = 0, t_format = 0, t_submit = 0; // --verbose: seconds the reader / consumer spent in each stage
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  int wrote_any = 0; // (direct mode) a placement was already written: the next one is preceded by ",\n" (ref src/krepp.cpp:476-481)

  std::thread writer([&] {
    int has_previous = 0;
    TextSet* ts = nullptr;
    while (sets_full.pop(ts)) {
      for (uint32_t t = 0; t < T; ++t) {
        if (!ts->len[t]) continue;
        if (jplace && has_previous) fwrite(",\n", 1, 2, out); // ref src/krepp.cpp:476-481
        if (fwrite(ts->part[t].data(), 1, ts->len[t], out) != ts->len[t]) error_exit("Failed to write the output");
        has_previous = 1;
      }
      sets_free.push(ts);
    }
  });

  std::thread consumer([&] {
    std::vector<std::vector<double>> part_w(T);
    Slot* s = nullptr;
    while (busy_q.pop(s)) {
      krepp_results_t res;
      const auto tw0 = now();
      check(krepp_batch_wait(s->batch, &res));
      const auto tw1 = now();
      t_wait += secs(tw0, tw1);
      TextSet* ts = nullptr;
      sets_free.pop(ts);
      std::atomic<uint32_t> formatted{0};
      std::atomic<int> failed{0};
      // split the batch's reads over T workers; every worker formats its range into its own buffer
      auto work = [&](uint32_t t) {
        const uint32_t lo = (uint32_t)((uint64_t)res.n_reads * t / T), hi = (uint32_t)((uint64_t)res.n_reads * (t + 1) / T);
        krepp_results_t sub = res;
        if (res.reads) sub.reads = res.reads + lo;
        if (res.dist_begin) sub.dist_begin = res.dist_begin + lo;
        if (res.seek_dist) sub.seek_dist = res.seek_dist + lo;
        sub.n_reads = hi - lo;
        double* w = nullptr;
        if (p.summarize) { part_w[t].assign(info.nnodes + 1, 0.0); w = part_w[t].data(); }
        std::vector<char>& buf = ts->part[t];
        for (;;) {
          int prev = 0;
          const size_t n = seek ? krepp_format_seek(&sub, s->names.data(), s->name_off.data() + lo, buf.data(), buf.size())
                           : place ? krepp_format_place(index[0], &p, &sub, s->names.data(), s->name_off.data() + lo, o.tabular, &prev, w, buf.data(), buf.size())
                                 : krepp_format_dist(index[0], &p, &sub, s->names.data(), s->name_off.data() + lo, w, buf.data(), buf.size());
          if (n <= buf.size()) { ts->len[t] = n; break; }
          buf.resize(n + n / 4);
          if (w) part_w[t].assign(info.nnodes + 1, 0.0);
        }
        if (!direct || p.summarize) return;
        // direct mode: wait until every part's length is known, then write this one at its offset
        formatted.fetch_add(1, std::memory_order_release);
        while (formatted.load(std::memory_order_acquire) < T) std::this_thread::yield();
        off_t at = file_off;
        int before = wrote_any;
        for (uint32_t u = 0; u < t; ++u) if (ts->len[u]) { at += (off_t)ts->len[u] + ((jplace && before) ? 2 : 0); before = 1; }
        if (!ts->len[t]) return;
        if (jplace && before) { if (pwrite(fileno(out), ",\n", 2, at) != 2) failed = 1; at += 2; }
        size_t done = 0;
        while (done < ts->len[t]) {
          const ssize_t k = pwrite(fileno(out), buf.data() + done, ts->len[t] - done, at + (off_t)done);
          if (k <= 0) { failed = 1; break; }
          done += (size_t)k;
        }
      };
      if (T == 1) work(0);
      else { std::vector<std::thread> th; for (uint32_t t = 1; t < T; ++t) th.emplace_back(work, t); work(0); for (auto& x : th) x.join(); }
      if (failed) error_exit("Failed to write the output");
      t_format += secs(tw1, now());
      free_q.push(s); // results and names are no longer needed: the reader may fill the slot again
      if (p.summarize) {
        for (uint32_t t = 0; t < T; ++t) { for (uint32_t se = 0; se <= info.nnodes; ++se) wcount[se] += part_w[t][se]; ts->len[t] = 0; }
        sets_free.push(ts);
      } else if (direct) {
        for (uint32_t t = 0; t < T; ++t) if (ts->len[t]) { file_off += (off_t)ts->len[t] + ((jplace && wrote_any) ? 2 : 0); wrote_any = 1; ts->len[t] = 0; }
        sets_free.push(ts);
      } else sets_full.push(ts);
    }
    sets_full.close();
  });

  krepp_reader_t* reader = nullptr;
  check(krepp_reader_open(o.query.c_str(), &reader));
  check(krepp_reader_set_threads(reader, o.num_threads)); // plain four-line FASTQ is framed chunk-parallel
  for (;;) {
    Slot* s = nullptr;
    free_q.pop(s);
    int eof = 0;
    const auto tr0 = now();
    check(krepp_reader_next(reader, s->bases, o.batch_bases, s->offsets, o.batch_reads, s->names.data(), s->names.size(), s->name_off.data(), &s->n, &eof));
    const auto tr1 = now();
    t_read += secs(tr0, tr1);
    if (s->n) {
      total_queries += s->n;
      check(krepp_batch_submit(s->batch, s->bases, s->offsets, s->n));
      t_submit += secs(tr1, now());
      busy_q.push(s);
    } else free_q.push(s);
    if (eof) break;
  }
  krepp_reader_close(reader);
  busy_q.close();
  consumer.join();
  writer.join();

  if (direct) fseeko(out, file_off, SEEK_SET); // the batches were written behind stdio's back
  if (!seek) { // --summarize table / end_jplace
    size_t n = krepp_format_footer(index[0], &p, o.tabular, wcount.data(), total_queries, invocation.c_str(), text.data(), text.size());
    if (n > text.size()) { text.resize(n); n = krepp_format_footer(index[0], &p, o.tabular, wcount.data(), total_queries, invocation.c_str(), text.data(), text.size()); }
    emit(n);
  }
  fflush(out);
  if (out != stdout) fclose(out);
  const std::chrono::duration<float> es = std::chrono::system_clock::now() - tquery;
  fprintf(stderr, seek ? "Done seeking sequences, elapsed: %g sec\n" : place ? "Done placing queries, elapsed: %g sec\n" : "Done estimating distances, elapsed: %g sec\n", es.count());
  fprintf(stderr, "Total number of sequences queried: %llu\n", (unsigned long long)total_queries);
  if (o.verbose) fprintf(stderr, "[stages] reader %.3f s, submit %.3f s (producer thread); waiting for the GPU %.3f s, formatting + writing %.3f s (consumer thread)%s\n",
                         t_read, t_submit, t_wait, t_format, direct ? "; output written by the formatter threads (pwrite)" : "");
  for (Slot& s : slots) krepp_batch_destroy(s.batch);
  for (krepp_index_t* ix : index) krepp_index_close(ix);
  { std::time_t t = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now()); fprintf(stderr, "%s", std::ctime(&t)); }
  return 0;
}

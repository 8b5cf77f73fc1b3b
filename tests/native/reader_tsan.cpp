// tests/native/reader_tsan.cpp -- TEST INFRASTRUCTURE: the host reader's chunk-parallel FASTQ framing under ThreadSanitizer.
// Built by tests/test_reader_tsan_cpu.py from krepp_b200/csrc/host_io.cpp + index_image.cpp with -fsanitize=thread; reads the
// file given on the command line batch by batch with 1 and with N threads and checks that both give the same records.
// usage: reader_tsan FILE THREADS BATCH_READS
//        reader_tsan --load INDEX_DIR [NSHARDS]   the index loader (threads read the table and flatten the colour lists), whole and per shard
//        reader_tsan --write OUT_DIR N            the library writer's table stage (threads over ranges of k-mers) on a synthetic union of N k-mers
#include "../../include/krepp_b200.h"
#include "../../krepp_b200/csrc/builder.hpp"
#include "../../krepp_b200/csrc/index_image.hpp"
#include <cstring>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace krepp {
int set_error(int code, const char* fmt, ...)
{ // (the library's lives in api.cu, which needs the CUDA runtime)
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
  return code;
}
} // namespace krepp

static bool read_all(const char* path, uint32_t threads, uint32_t batch_reads, std::string& seqs, std::string& names, uint64_t& n_total)
{
  krepp_reader_t* r = nullptr;
  if (krepp_reader_open(path, &r) != KREPP_OK) return false;
  krepp_reader_set_threads(r, threads);
  const uint64_t max_bases = 400ull * batch_reads;
  std::vector<char> bases(max_bases + 64), nm(64ull * batch_reads);
  std::vector<uint64_t> off(batch_reads + 1), noff(batch_reads);
  n_total = 0;
  for (;;) {
    uint32_t n = 0;
    int eof = 0;
    if (krepp_reader_next(r, bases.data(), max_bases, off.data(), batch_reads, nm.data(), nm.size(), noff.data(), &n, &eof) != KREPP_OK) { krepp_reader_close(r); return false; }
    for (uint32_t i = 0; i < n; ++i) {
      seqs.append(bases.data() + off[i], off[i + 1] - off[i]); seqs.push_back('\n');
      names.append(nm.data() + noff[i]); names.push_back('\n');
    }
    n_total += n;
    if (eof) break;
  }
  krepp_reader_close(r);
  return true;
}

static int load_index(const char* dir, uint32_t nshards)
{
  uint64_t total = 0;
  for (uint32_t s = 0; s < nshards; ++s) {
    krepp::HostIndex h;
    const std::string err = h.load(dir, s, nshards);
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 5; }
    total += h.cmer.size();
    if (s == 0) printf("k-mers %llu, colour ids %u, flattened colour lists %zu entries\n", (unsigned long long)h.nkmers, h.nsubsets, h.cleaf.size());
  }
  printf("%u shard(s) hold %llu entries\n", nshards, (unsigned long long)total);
  return 0;
}

static int write_library_check(const char* dir, uint64_t n)
{
  krepp_index geom;
  std::vector<uint8_t> ppos, npos;
  krepp::lsh_positions(21, 7, false, 0, ppos, npos);
  std::string err = geom.host.set_geometry(21, 25, 7, 4, 1, true, ppos, npos);
  if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 6; }
  krepp_builder b;
  b.geom = &geom; b.with_tree = true; b.nwk_text = "((A:1,B:1):1,(C:1,D:1):1);"; b.names = {"A", "B", "C", "D"};
  if (!(err = b.tree.parse(b.nwk_text)).empty()) { fprintf(stderr, "%s\n", err.c_str()); return 6; }
  b.leaf_rho.assign(4, 0.2);
  b.set_begin.push_back(0);
  for (uint32_t mask = 1; mask < 16; ++mask) { for (uint32_t l = 0; l < 4; ++l) if (mask >> l & 1) b.set_leaves.push_back(l); b.set_begin.push_back(b.set_leaves.size()); }
  const uint64_t nrows = geom.host.nrows;
  std::vector<uint64_t> want(nrows, 0);
  b.keys.resize(n); b.set_of.resize(n);
  for (uint64_t i = 0; i < n; ++i) { // rows 0, 3, 6, ... get runs of entries; the rows between stay empty
    const uint64_t row = (i * nrows / n) / 3 * 3;
    b.keys[i] = row << 32 | (uint32_t)i; b.set_of[i] = (uint32_t)(i % 15);
    ++want[row];
  }
  for (uint64_t r = 1; r < nrows; ++r) want[r] += want[r - 1];
  b.have_union = true;
  uint64_t nk = 0;
  uint32_t nsub = 0;
  if (!(err = krepp::write_library(b, dir, 0, &nk, &nsub)).empty()) { fprintf(stderr, "%s\n", err.c_str()); return 6; }
  std::vector<uint64_t> inc(nrows);
  FILE* f = fopen((std::string(dir) + "/inc-m4r1-frac").c_str(), "rb");
  uint32_t nr = 0;
  const bool ok = f && fread(&nr, 4, 1, f) == 1 && nr == nrows && fread(inc.data(), 8, nrows, f) == nrows;
  if (f) fclose(f);
  if (!ok || inc != want) { fprintf(stderr, "offsets differ from the sequential count\n"); return 7; }
  printf("%llu k-mers written, %u colour ids, offsets as counted sequentially\n", (unsigned long long)nk, nsub);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc >= 4 && !strcmp(argv[1], "--write")) return write_library_check(argv[2], strtoull(argv[3], nullptr, 10));
  if (argc >= 3 && !strcmp(argv[1], "--load")) return load_index(argv[2], argc > 3 ? (uint32_t)atoi(argv[3]) : 1);
  if (argc < 4) return 2;
  const uint32_t threads = (uint32_t)atoi(argv[2]), batch = (uint32_t)atoi(argv[3]);
  std::string s1, n1, sN, nN;
  uint64_t c1 = 0, cN = 0;
  if (!read_all(argv[1], 1, batch, s1, n1, c1) || !read_all(argv[1], threads, batch, sN, nN, cN)) { fprintf(stderr, "read failed\n"); return 3; }
  if (c1 != cN || s1 != sN || n1 != nN) { fprintf(stderr, "records differ: %llu vs %llu\n", (unsigned long long)c1, (unsigned long long)cN); return 4; }
  printf("%llu records, %u threads: same as one thread\n", (unsigned long long)c1, threads);
  return 0;
}

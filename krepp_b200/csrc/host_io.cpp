// host_io.cpp -- the host steps either side of the GPU path, behind the same C ABI (include/krepp_b200.h):
//   * FASTA/FASTQ batch reader with the record framing of kseq_read (ref src/kseq.h:177-216) as driven by
//     QSeq::read_next_batch (ref src/rqseq.cpp:180-197), parsing straight into a batch slot's pinned buffers;
//   * text formatting of the result structs: report_distances (ref src/query.cpp:158-196), report_placement
//     (ref src/query.cpp:218-333, text and --no-multi / --summarize selection only; all arithmetic is done on the GPU),
//     headers and jplace framing (ref src/krepp.cpp:311-319,396-432).
// Pure host code; nothing here touches the device.
#include "../../include/krepp_b200.h"

#include "handles.hpp"

#include <zlib.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <thread>
#include <vector>

using namespace krepp;

#define KREPP_VERSION_STRING "v0.8.3+b200"

// ------------------------------------------------------------------------------------------------ reader

namespace {

// byte classes inside a sequence block (ref src/kseq.h:193-201): 0 dropped (not isgraph), 1 sequence character,
// 2 '>' or '@' (next header), 3 '+' (quality block follows)
struct ByteClass {
  uint8_t seq[256];
  uint8_t space[256];
  ByteClass()
  {
    for (int c = 0; c < 256; ++c) {
      seq[c] = (c > 32 && c < 127) ? 1 : 0;
      space[c] = (c == ' ' || (c >= 9 && c <= 13)) ? 1 : 0;
    }
    seq[(int)'>'] = seq[(int)'@'] = 2;
    seq[(int)'+'] = 3;
  }
};
const ByteClass kClass;

enum class St { Seek, Name, Comment, Seq, Plus, Qual, QualTail, Done };

} // namespace

struct krepp_reader {
  gzFile f = nullptr;   // gzip input (zlib inflates; its "transparent" mode for plain files is a slow copy, hence:)
  int fd = -1;          // plain input: read(2) straight into the buffer
  std::vector<unsigned char> buf;
  size_t at = 0, end = 0;
  bool eof = false;
  St st = St::Seek;
  std::string name, seq;   // the record being parsed
  uint64_t qual_left = 0;
  bool fresh = false;        // name/seq still hold the record returned last; cleared when parsing resumes
  bool have_pending = false; // a complete record that did not fit the previous batch
  bool fast = true;          // four-line FASTQ fast path (KREPP_READER_FAST=0 leaves every record to the state machine)
  std::string pend_name, pend_seq;
  // plain regular files are read with pread at explicit offsets, and with several threads (krepp_reader_set_threads) whole
  // batches of four-line FASTQ are framed chunk-parallel from a window the threads pread together (parallel_fastq below)
  bool regular = false;
  size_t file_len = 0, file_at = 0; // file_at: file offset of the first byte not yet handed to buf
  unsigned char* win = nullptr;     // the parallel path's window of the file (malloc'd: no zero fill, first touched by the pread threads)
  size_t win_cap = 0;
  uint32_t threads = 1;
};

namespace {

// up to `want` bytes of input at `dst`; fewer only at the end of the input (or on a read error)
size_t read_input(krepp_reader* r, unsigned char* dst, size_t want)
{
  size_t got = 0;
  while (got < want) {
    long n;
    if (r->regular) { n = (long)pread(r->fd, dst + got, want - got, (off_t)r->file_at); if (n > 0) r->file_at += (size_t)n; }
    else if (r->fd >= 0) n = (long)read(r->fd, dst + got, want - got);
    else n = gzread(r->f, dst + got, (unsigned)(want - got));
    if (n <= 0) break;
    got += (size_t)n;
  }
  return got;
}

bool refill(krepp_reader* r)
{
  if (r->eof) return false;
  const long n = (long)read_input(r, r->buf.data(), r->buf.size());
  r->at = 0;
  r->end = n > 0 ? (size_t)n : 0;
  if (n < (long)r->buf.size()) r->eof = true;
  return r->end > 0;
}

// Advances the state machine until one complete record sits in r->name / r->seq (returns true) or the input ends.
bool next_record(krepp_reader* r)
{
  if (r->fresh) { r->name.clear(); r->seq.clear(); r->fresh = false; }
  for (;;) {
    if (r->at >= r->end && !refill(r)) {
      // end of input: a FASTA record in progress is complete, a FASTQ record is complete only when all of its quality
      // characters were seen (ref src/kseq.h:203-213: -2 "truncated quality string" drops the record)
      const St s = r->st;
      r->st = St::Done;
      // a header cut short by EOF is still a record with an empty sequence, unless nothing followed '>' / '@'
      // (ref src/kseq.h:189-191: ks_getuntil returns -1 only when the stream is already exhausted)
      const bool rec = s == St::Seq || s == St::QualTail || (s == St::Qual && r->qual_left == 0) || s == St::Comment ||
                       (s == St::Name && !r->name.empty());
      r->fresh = rec;
      return rec;
    }
    const unsigned char* p = r->buf.data() + r->at;
    const unsigned char* e = r->buf.data() + r->end;
    switch (r->st) {
      case St::Seek: {
        while (p < e && *p != '>' && *p != '@') ++p;
        if (p < e) { ++p; r->st = St::Name; r->name.clear(); r->seq.clear(); }
        break;
      }
      case St::Name: {
        const unsigned char* b = p;
        while (p < e && !kClass.space[*p]) ++p;
        r->name.append(reinterpret_cast<const char*>(b), p - b);
        if (p < e) { r->st = (*p == '\n') ? St::Seq : St::Comment; ++p; }
        break;
      }
      case St::Comment: {
        const void* nl = memchr(p, '\n', e - p);
        if (nl) { p = static_cast<const unsigned char*>(nl) + 1; r->st = St::Seq; } else p = e;
        break;
      }
      case St::Seq: {
        // run of sequence characters: copy whole spans between bytes of another class
        while (p < e) {
          const unsigned char* b = p;
          while (p < e && kClass.seq[*p] == 1) ++p;
          if (p > b) r->seq.append(reinterpret_cast<const char*>(b), p - b);
          if (p == e) break;
          const uint8_t c = kClass.seq[*p];
          ++p;
          if (c == 2) { r->at = p - r->buf.data(); r->st = St::Name; r->fresh = true; return true; } // header char already consumed
          if (c == 3) { r->st = St::Plus; break; }
        }
        break;
      }
      case St::Plus: {
        const void* nl = memchr(p, '\n', e - p);
        if (nl) { p = static_cast<const unsigned char*>(nl) + 1; r->st = St::Qual; r->qual_left = r->seq.size(); } else p = e;
        break;
      }
      case St::Qual: {
        while (p < e && r->qual_left) { if (*p >= 33 && *p <= 127) --r->qual_left; ++p; }
        if (!r->qual_left) r->st = St::QualTail;
        break;
      }
      case St::QualTail: { // kseq consumes one more character after the last quality character (ref src/kseq.h:208)
        ++p;
        r->at = p - r->buf.data();
        r->st = St::Seek;
        r->fresh = true;
        return true;
      }
      case St::Done: return false;
    }
    r->at = p - r->buf.data();
  }
}

} // namespace

namespace {

inline uint64_t load8(const unsigned char* p) { uint64_t x; memcpy(&x, p, 8); return x; }
// true when every byte of [p, p + n) is a sequence character: 33..126 and none of '>', '@', '+' (ByteClass::seq == 1)
inline bool all_seq_chars(const unsigned char* p, size_t n)
{
  const uint64_t k01 = 0x0101010101010101ull, k80 = 0x8080808080808080ull;
  auto has = [&](uint64_t x, unsigned char c) { const uint64_t y = x ^ (k01 * c); return ((y - k01) & ~y & k80) != 0; }; // some byte == c
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const uint64_t x = load8(p + i);
    if (x & k80) return false;                                   // a byte >= 128
    if (((x - k01 * 33) & ~x & k80) != 0) return false;          // a byte < 33 (no byte has its top bit set here)
    if (has(x, 127) || has(x, '>') || has(x, '@') || has(x, '+')) return false;
  }
  for (; i < n; ++i) if (kClass.seq[p[i]] != 1) return false;
  return true;
}
// true when every byte of [p, p + n) is a quality character as the Qual state counts them: 33..127
inline bool all_qual_chars(const unsigned char* p, size_t n)
{
  const uint64_t k01 = 0x0101010101010101ull, k80 = 0x8080808080808080ull;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const uint64_t x = load8(p + i);
    if ((x & k80) || ((x - k01 * 33) & ~x & k80)) return false;
  }
  for (; i < n; ++i) if (p[i] < 33 || p[i] > 127) return false;
  return true;
}

// Fast path of krepp_reader_next for the overwhelmingly common record: a four-line FASTQ record lying whole in the buffer.
// With the parser in state Seek at an '@', it is recognised by three memchr()s and two range checks and copied straight into
// the batch arrays; the conditions are exactly those under which the state machine of next_record() would walk
// Name[/Comment] -> Seq (one unbroken run of sequence characters, then the newline, then '+') -> Plus -> Qual (as many
// quality characters as bases, back to back) -> QualTail (one more byte) -> Seek.  Anything else -- FASTA, wrapped lines,
// stray characters, a record cut by the end of the buffer -- returns false and is left to the state machine.
bool fast_fastq(krepp_reader* r, char* bases, uint64_t max_bases, uint64_t& nb, char* names, uint64_t max_name_bytes, uint64_t& nn,
                uint64_t* offsets, uint64_t* name_offsets, uint32_t& n)
{
  const unsigned char* const b0 = r->buf.data();
  const unsigned char* p = b0 + r->at;
  const unsigned char* const e = b0 + r->end;
  if (p >= e || *p != '@') return false;
  const unsigned char* nl1 = static_cast<const unsigned char*>(memchr(p + 1, '\n', e - (p + 1)));
  if (!nl1) return false;
  const unsigned char* ne = p + 1;
  while (!kClass.space[*ne]) ++ne;                                  // stops at nl1 at the latest
  const size_t name_len = ne - (p + 1);
  if (name_len == 0) return false;
  const unsigned char* sq = nl1 + 1;
  const unsigned char* nl2 = static_cast<const unsigned char*>(memchr(sq, '\n', e - sq));
  if (!nl2 || nl2 + 1 >= e || nl2[1] != '+') return false;
  const size_t len = nl2 - sq;
  if (!all_seq_chars(sq, len)) return false;
  const unsigned char* nl3 = static_cast<const unsigned char*>(memchr(nl2 + 1, '\n', e - (nl2 + 1)));
  if (!nl3) return false;
  const unsigned char* q = nl3 + 1;
  if ((size_t)(e - q) < len + 1 || !all_qual_chars(q, len)) return false; // the byte after the last quality character is consumed too
  if (nb + len > max_bases || nn + name_len + 1 > max_name_bytes) return false; // does not fit: the general path keeps it for the next batch
  memcpy(bases + nb, sq, len);
  nb += len;
  name_offsets[n] = nn;
  memcpy(names + nn, p + 1, name_len);
  names[nn + name_len] = 0;
  nn += name_len + 1;
  offsets[++n] = nb;
  r->at = (q + len + 1) - b0;
  return true;
}

} // namespace

namespace {

struct FqRec { uint32_t name_len, len; uint64_t name_off, seq_off; };

// The conditions of fast_fastq on raw pointers: a four-line FASTQ record starting at p ('@'), lying whole in [p, e) together with
// the byte that follows its last quality character.  Returns the position after that byte, or nullptr.
inline const unsigned char* probe_fastq(const unsigned char* base, const unsigned char* p, const unsigned char* e, FqRec* rec)
{
  if (p >= e || *p != '@') return nullptr;
  const unsigned char* nl1 = static_cast<const unsigned char*>(memchr(p + 1, '\n', e - (p + 1)));
  if (!nl1) return nullptr;
  const unsigned char* ne = p + 1;
  while (!kClass.space[*ne]) ++ne;
  const size_t name_len = ne - (p + 1);
  if (name_len == 0) return nullptr;
  const unsigned char* sq = nl1 + 1;
  const unsigned char* nl2 = static_cast<const unsigned char*>(memchr(sq, '\n', e - sq));
  if (!nl2 || nl2 + 1 >= e || nl2[1] != '+') return nullptr;
  const size_t len = nl2 - sq;
  if (len > 0xFFFFFFFFull || name_len > 0xFFFFFFFFull || !all_seq_chars(sq, len)) return nullptr;
  const unsigned char* nl3 = static_cast<const unsigned char*>(memchr(nl2 + 1, '\n', e - (nl2 + 1)));
  if (!nl3) return nullptr;
  const unsigned char* q = nl3 + 1;
  if ((size_t)(e - q) < len + 1 || !all_qual_chars(q, len)) return nullptr;
  rec->name_len = (uint32_t)name_len; rec->len = (uint32_t)len; rec->name_off = (uint64_t)(p + 1 - base); rec->seq_off = (uint64_t)(sq - base);
  return q + len + 1;
}

// Chunk-parallel framing of one batch from the mapped file.  The parser stands in state Seek at file offset `pos`.  The window
// that should hold the batch is cut into one chunk per thread; every thread but the first looks for a record start in its chunk
// (a line starting with '@' from which probe_fastq succeeds: in four-line FASTQ a quality line that starts with '@' is followed
// by a header line, which is no sequence line, so it cannot be mistaken) and frames records up to the next thread's start.  The
// result is accepted as far as the chunks stitch: thread t must end exactly where thread t + 1 began, which by induction from
// the true boundary `pos` makes every start a true boundary -- the records are then the ones the sequential state machine walks,
// because probe_fastq's conditions are those under which it walks a record as Name -> Seq -> Plus -> Qual -> QualTail -> Seek.
// Returns the number of records taken (0: nothing usable here, leave the batch to the sequential path).
uint32_t parallel_fastq(krepp_reader* r, size_t pos, char* bases, uint64_t max_bases, uint64_t* offsets, uint32_t max_reads, char* names,
                        uint64_t max_name_bytes, uint64_t* name_offsets, size_t* new_pos)
{
  // the first record (from a small read) sizes the window; then the threads pread the window together.  Offsets below are
  // relative to the window: `base` is file offset `pos`.
  if (pos >= r->file_len) return 0;
  unsigned char head[4096];
  const size_t hn = (size_t)std::max<ssize_t>(0, pread(r->fd, head, std::min<size_t>(sizeof head, r->file_len - pos), (off_t)pos));
  FqRec first;
  const unsigned char* after = probe_fastq(head, head, head + hn, &first);
  if (!after) return 0;
  const size_t rec_bytes = (size_t)(after - head);
  uint64_t want = max_reads;
  if (first.len) want = std::min<uint64_t>(want, max_bases / first.len);
  want = std::min<uint64_t>(want, max_name_bytes / (first.name_len + 1ull));
  if (want < 1024) return 0; // small batches: not worth the threads
  // a little more than the batch should need, so that its last record (and the byte after it) lies inside the window
  const size_t window = (size_t)std::min<uint64_t>((uint64_t)r->file_len - pos, want * rec_bytes + (1u << 16));
  const uint32_t T = (uint32_t)std::max<size_t>(1, std::min<size_t>(r->threads, window >> 18));
  static const bool dbg = getenv("KREPP_READER_DEBUG") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  if (r->win_cap < window) { // sized with some room so that later windows (slightly different record sizes) fit without a new allocation
    free(r->win);
    r->win_cap = window + window / 8 + (1u << 20);
    r->win = static_cast<unsigned char*>(malloc(r->win_cap));
    if (!r->win) { r->win_cap = 0; return 0; }
  }
  {
    auto fill = [&](uint32_t t) {
      size_t lo = window / T * t, hi = t + 1 == T ? window : window / T * (t + 1);
      while (lo < hi) { const ssize_t got = pread(r->fd, r->win + lo, hi - lo, (off_t)(pos + lo)); if (got <= 0) break; lo += (size_t)got; }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < T; ++t) th.emplace_back(fill, t);
    fill(0);
    for (auto& x : th) x.join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  const unsigned char* base = r->win;
  const unsigned char* const fe = base + window;
  pos = 0; // window-relative from here on
  std::vector<size_t> start(T, 0), stop(T, 0);
  std::vector<std::vector<FqRec>> recs(T);
  std::vector<uint64_t> tb(T, 0), tn(T, 0);
  // every thread but the first finds the first record start after its chunk's nominal boundary, then frames records while they
  // start at or before the next chunk's nominal boundary -- so it ends on the start the next thread finds, if both are right
  auto frame = [&](uint32_t t) {
    size_t at = window / T * t;
    const size_t lim = t + 1 == T ? window : window / T * (t + 1);
    FqRec rec;
    if (t) {
      for (;;) {
        const unsigned char* nl = static_cast<const unsigned char*>(memchr(base + at, '\n', (size_t)(fe - (base + at))));
        if (!nl || nl + 1 >= fe) { at = window; break; }
        at = (size_t)(nl + 1 - base);
        if (base[at] == '@' && probe_fastq(base, base + at, fe, &rec)) break;
      }
    }
    start[t] = at;
    std::vector<FqRec>& v = recs[t];
    v.reserve((lim > at ? lim - at : 0) / std::max<size_t>(rec_bytes, 1) + 16);
    uint64_t b = 0, nn = 0;
    while (at <= lim && at < window) {
      const unsigned char* nx = probe_fastq(base, base + at, fe, &rec);
      if (!nx) break;
      v.push_back(rec);
      b += rec.len; nn += rec.name_len + 1ull;
      at = (size_t)(nx - base);
    }
    stop[t] = at; tb[t] = b; tn[t] = nn;
  };
  {
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < T; ++t) th.emplace_back(frame, t);
    frame(0);
    for (auto& x : th) x.join();
  }
  const auto t2 = std::chrono::steady_clock::now();
  // 3. how far the chunks stitch, and how many records the batch limits admit
  std::vector<uint64_t> keep(T, 0), rbase(T, 0), bbase(T, 0), nbase(T, 0);
  uint64_t nr = 0, nb = 0, nn = 0;
  size_t endpos = pos;
  for (uint32_t t = 0; t < T; ++t) {
    if (t && stop[t - 1] != start[t]) break; // the previous chunk did not end on this chunk's start: what follows is not proven
    rbase[t] = nr; bbase[t] = nb; nbase[t] = nn;
    uint64_t k = recs[t].size();
    if (nr + k > max_reads || nb + tb[t] > max_bases || nn + tn[t] > max_name_bytes) { // the cut falls inside this chunk
      k = 0;
      uint64_t b = nb, m = nn;
      while (k < recs[t].size() && nr + k < max_reads && b + recs[t][k].len <= max_bases && m + recs[t][k].name_len + 1ull <= max_name_bytes) {
        b += recs[t][k].len; m += recs[t][k].name_len + 1ull; ++k;
      }
      keep[t] = k; nr += k; nb = b; nn = m;
      endpos = k < recs[t].size() ? (size_t)recs[t][k].name_off - 1 : stop[t];
      break;
    }
    keep[t] = k; nr += k; nb += tb[t]; nn += tn[t];
    endpos = stop[t];
  }
  if (!nr) return 0;
  auto copy = [&](uint32_t t) { // 4. into the batch arrays
    uint64_t b = bbase[t], m = nbase[t], i = rbase[t];
    for (uint64_t k = 0; k < keep[t]; ++k, ++i) {
      const FqRec& rc = recs[t][k];
      memcpy(bases + b, base + rc.seq_off, rc.len);
      b += rc.len;
      name_offsets[i] = m;
      memcpy(names + m, base + rc.name_off, rc.name_len);
      names[m + rc.name_len] = 0;
      m += rc.name_len + 1ull;
      offsets[i + 1] = b;
    }
  };
  {
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < T; ++t) if (keep[t]) th.emplace_back(copy, t);
    copy(0);
    for (auto& x : th) x.join();
  }
  if (dbg) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[reader] %u threads, window %zu B: pread %.2f ms, frame %.2f ms, stitch+copy %.2f ms, %llu reads\n", T, window, ms(t0, t1), ms(t1, t2), ms(t2, t3), (unsigned long long)nr);
  }
  *new_pos = endpos; // window-relative
  return (uint32_t)nr;
}

} // namespace

extern "C" int krepp_reader_set_threads(krepp_reader_t* r, uint32_t threads)
{
  if (!r) return set_error(KREPP_ERR_ARG, "krepp_reader_set_threads: null reader");
  r->threads = threads ? threads : 1;
  return KREPP_OK;
}

extern "C" int krepp_reader_open(const char* path, krepp_reader_t** out)
{
  if (!path || !out) return set_error(KREPP_ERR_ARG, "krepp_reader_open: null argument");
  *out = nullptr;
  int fd = open(path, O_RDONLY);
  if (fd < 0) return set_error(KREPP_ERR_IO, "Failed to open the file at %s", path);
  unsigned char magic[2] = {0, 0};
  const bool gz = pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
  gzFile f = nullptr;
  if (gz) {
    close(fd);
    fd = -1;
    f = gzopen(path, "rb");
    if (!f) return set_error(KREPP_ERR_IO, "Failed to open the file at %s", path);
    gzbuffer(f, 1 << 20);
  }
  auto* r = new krepp_reader;
  r->f = f; r->fd = fd;
  r->buf.resize(4 << 20);
  if (fd >= 0) { // plain input: a regular file is read with pread (a pipe keeps read(2))
    struct stat st;
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) { r->regular = true; r->file_len = (size_t)st.st_size; }
  }
  if (const char* env = getenv("KREPP_READER_FAST")) r->fast = strcmp(env, "0") != 0;
  *out = r;
  return KREPP_OK;
}

extern "C" void krepp_reader_close(krepp_reader_t* r)
{
  if (!r) return;
  if (r->f) gzclose(r->f);
  if (r->fd >= 0) close(r->fd);
  free(r->win);
  delete r;
}

extern "C" int krepp_reader_next(krepp_reader_t* r, char* bases, uint64_t max_bases, uint64_t* offsets, uint32_t max_reads,
                                 char* names, uint64_t max_name_bytes, uint64_t* name_offsets, uint32_t* n_reads, int* eof)
{
  if (!r || !bases || !offsets || !names || !name_offsets || !n_reads || !eof) return set_error(KREPP_ERR_ARG, "krepp_reader_next: null argument");
  uint32_t n = 0;
  uint64_t nb = 0, nn = 0;
  offsets[0] = 0;
  *eof = 0;
  auto fits = [&](const std::string& nm, const std::string& sq) {
    return n < max_reads && nb + sq.size() <= max_bases && nn + nm.size() + 1 <= max_name_bytes;
  };
  auto put = [&](const std::string& nm, const std::string& sq) {
    memcpy(bases + nb, sq.data(), sq.size());
    nb += sq.size();
    name_offsets[n] = nn;
    memcpy(names + nn, nm.c_str(), nm.size() + 1);
    nn += nm.size() + 1;
    offsets[++n] = nb;
  };
  if (r->have_pending) {
    if (!fits(r->pend_name, r->pend_seq))
      return set_error(KREPP_ERR_CAPACITY, "a single sequence of %zu bases (name of %zu bytes) does not fit the batch buffers", r->pend_seq.size(), r->pend_name.size());
    put(r->pend_name, r->pend_seq);
    r->have_pending = false;
  }
  if (r->regular && r->threads > 1 && r->fast && n == 0 && r->st == St::Seek) { // whole batch chunk-parallel
    const size_t pos = r->file_at - (r->end - r->at);
    size_t rel = 0;
    const uint32_t got = parallel_fastq(r, pos, bases, max_bases, offsets, max_reads, names, max_name_bytes, name_offsets, &rel);
    if (got) {
      r->file_at = pos + rel; r->at = r->end = 0; r->eof = false; // the buffer's read-ahead is dropped: the next bytes are read at file_at
      *n_reads = got;
      return KREPP_OK; // (the end of the input is reported by the next call, which finds nothing left)
    }
  }
  for (;;) {
    if (n >= max_reads) break;
    if (r->fast && r->st == St::Seek && fast_fastq(r, bases, max_bases, nb, names, max_name_bytes, nn, offsets, name_offsets, n)) continue;
    if (!next_record(r)) { *eof = 1; break; }
    if (!fits(r->name, r->seq)) {
      if (n == 0) return set_error(KREPP_ERR_CAPACITY, "a single sequence of %zu bases (name of %zu bytes) does not fit the batch buffers", r->seq.size(), r->name.size());
      r->pend_name.swap(r->name);
      r->pend_seq.swap(r->seq);
      r->have_pending = true;
      break;
    }
    put(r->name, r->seq);
  }
  if (r->st == St::Done && !r->have_pending) *eof = 1;
  *n_reads = n;
  return KREPP_OK;
}

// ------------------------------------------------------------------------------------------------ formatting

namespace {

struct Out {
  char* buf;
  size_t cap, len = 0;
  Out(char* b, size_t c) : buf(b), cap(b ? c : 0) {}
  void put(const char* s, size_t n)
  {
    if (len + n <= cap) memcpy(buf + len, s, n);
    len += n;
  }
  void put(const char* s) { put(s, strlen(s)); }
  void put(const std::string& s) { put(s.data(), s.size()); }
  void ch(char c) { if (len < cap) buf[len] = c; ++len; }
  void u32(uint32_t v)
  {
    char t[12];
    int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) ch(t[--n]);
  }
  // std::fixed << precision(5): the exact binary value rounded to 5 decimals.  Fast path: scale, split, and fall back
  // to printf whenever the scaled value sits close to a rounding boundary (or is large / non-finite).
  void fixed5(double v)
  {
    const double a = fabs(v);
    if (a < 1000.0) {
      const double s = a * 100000.0;
      const double fl = floor(s);
      const double fr = s - fl;
      if (fabs(fr - 0.5) > 1e-6) {
        uint64_t q = (uint64_t)fl + (fr > 0.5 ? 1 : 0);
        const uint32_t ip = (uint32_t)(q / 100000), fp = (uint32_t)(q % 100000);
        if (std::signbit(v)) ch('-');
        u32(ip);
        ch('.');
        char t[5];
        uint32_t x = fp;
        for (int i = 4; i >= 0; --i) { t[i] = (char)('0' + x % 10); x /= 10; }
        put(t, 5);
        return;
      }
    }
    char t[400];
    const int n = snprintf(t, sizeof t, "%.5f", v);
    put(t, (size_t)n);
  }
};

inline const char* name_of(const char* names, const uint64_t* name_offsets, uint32_t i) { return names + name_offsets[i]; }

// Selected records (node_to_minfo) of one read by ascending leaf se: records are stored as forward leaves by ascending
// se, then reverse leaves by ascending se, and at most one of the two strands of a leaf is selected.
// field access for the two row forms (krepp_record_t, and the 16-byte krepp_brief_t that `dist` front ends ask for)
inline uint32_t r_se(const krepp_record_t& r) { return r.leaf_se; }
inline uint32_t r_se(const krepp_brief_t& r) { return KREPP_BRIEF_SE(r.ref); }
inline uint32_t r_strand(const krepp_record_t& r) { return r.strand; }
inline uint32_t r_strand(const krepp_brief_t& r) { return KREPP_BRIEF_STRAND(r.ref); }
inline uint32_t r_flags(const krepp_record_t& r) { return r.flags; }
inline uint32_t r_flags(const krepp_brief_t& r) { return KREPP_BRIEF_FLAGS(r.ref); }
inline bool r_chisq_ok(const krepp_record_t& r, const krepp_params_t* p) { return r.chisq < p->chisq; }
inline bool r_chisq_ok(const krepp_brief_t& r, const krepp_params_t*) { return KREPP_BRIEF_CHISQ_OK(r.ref) != 0; }

template <class R, class F>
void for_selected(const R* rec, const krepp_read_summary_t& s, F&& f)
{
  uint32_t i = s.rec_begin, ie = s.rec_begin + s.rec_count;
  uint32_t nf = 0;
  while (i + nf < ie && r_strand(rec[i + nf]) == 0) ++nf;
  uint32_t j = i + nf;
  const uint32_t je = ie;
  ie = i + nf;
  while (i < ie || j < je) {
    const uint32_t si = i < ie ? r_se(rec[i]) : 0xFFFFFFFFu, sj = j < je ? r_se(rec[j]) : 0xFFFFFFFFu;
    uint32_t pick;
    if (si < sj) pick = i++;
    else if (sj < si) pick = j++;
    else { pick = (r_flags(rec[i]) & KREPP_REC_SELECTED) ? i : j; ++i; ++j; }
    if (r_flags(rec[pick]) & KREPP_REC_SELECTED) f(rec[pick]);
  }
}
template <class F>
void for_selected(const krepp_results_t* res, const krepp_read_summary_t& s, F&& f) { for_selected(res->records, s, f); }

// IBatch::report_distances over either row form (ref src/query.cpp:158-196)
template <class R>
size_t format_dist_rows(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res, const R* rows, const char* names,
                        const uint64_t* name_offsets, double* wcount, char* buf, size_t cap)
{
  Out o(buf, cap);
  const HostTree& t = ix->host.tree;
  const bool has_max = !std::isnan(p->dist_max);
  std::vector<uint32_t> keep;
  for (uint32_t r = 0; r < res->n_reads; ++r) {
    const krepp_read_summary_t& s = res->reads[r];
    const char* id = name_of(names, name_offsets, r);
    const size_t id_len = strlen(id);
    if (p->summarize) { // ref src/query.cpp:160-171
      if (!wcount) continue;
      keep.clear();
      for_selected(rows, s, [&](const R& rec) {
        if (r_chisq_ok(rec, p) && (!has_max || rec.d_llh < p->dist_max)) keep.push_back(r_se(rec));
      });
      for (uint32_t se : keep) wcount[se] += 1.0 / (double)keep.size();
      continue;
    }
    if (s.closest < 0 || (has_max && rows[s.closest].d_llh > p->dist_max)) { // ref :173-176
      o.put(id); o.put("\tNA\tNaN\n");
      continue;
    }
    if (p->multi) {
      for_selected(rows, s, [&](const R& rec) {
        if (!p->no_filter && !r_chisq_ok(rec, p)) return;
        if (has_max && !(rec.d_llh < p->dist_max)) return;
        o.put(id, id_len); o.ch('\t'); o.put(t.shown[r_se(rec)]); o.ch('\t'); o.fixed5(rec.d_llh); o.ch('\n');
      });
    } else {
      const R& rec = rows[s.closest];
      o.put(id, id_len); o.ch('\t'); o.put(t.shown[r_se(rec)]); o.ch('\t'); o.fixed5(rec.d_llh); o.ch('\n');
    }
  }
  return o.len;
}

// The same text from the rows the device already chose, ordered and rounded (KREPP_OUT_DIST, include/krepp_b200.h): per read one
// word of dist_begin and its rows; nothing is decided here.
template <class Row>
size_t format_dist_compact(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res, const Row* rows, const char* names,
                           const uint64_t* name_offsets, double* wcount, char* buf, size_t cap)
{
  Out o(buf, cap);
  const HostTree& t = ix->host.tree;
  for (uint32_t r = 0; r < res->n_reads; ++r) {
    const uint32_t w0 = res->dist_begin[r], b = KREPP_DIST_BEGIN(w0), e = KREPP_DIST_BEGIN(res->dist_begin[r + 1]);
    auto se_of = [&](const Row& row) -> uint32_t { if constexpr (sizeof(Row) == 4) return t.leaf_se[row >> 16]; else return (uint32_t)row; };
    if (p->summarize) { // ref src/query.cpp:160-171
      if (wcount) for (uint32_t i = b; i < e; ++i) wcount[se_of(rows[i])] += 1.0 / (double)(e - b);
      continue;
    }
    const char* id = name_of(names, name_offsets, r);
    const size_t id_len = strlen(id);
    if (KREPP_DIST_NA(w0)) { o.put(id, id_len); o.put("\tNA\tNaN\n"); continue; }
    // every row is id, tab, reference name, tab, d.ddddd, newline: when the read's rows certainly fit, they are written through a
    // bare pointer (one capacity check per read instead of one per piece)
    const size_t per_row = id_len + t.max_shown + 11;
    if (o.len + (size_t)(e - b) * per_row <= o.cap) {
      char* p = o.buf + o.len;
      for (uint32_t i = b; i < e; ++i) {
        uint32_t units;
        if constexpr (sizeof(Row) == 4) units = rows[i] & 0xFFFFu; else units = (uint32_t)(rows[i] >> 32);
        const std::string& nm = t.shown[se_of(rows[i])];
        memcpy(p, id, id_len); p += id_len;
        *p++ = '\t';
        memcpy(p, nm.data(), nm.size()); p += nm.size();
        *p++ = '\t';
        uint32_t ip = units / 100000u, x = units - ip * 100000u;
        if (ip < 10) *p++ = (char)('0' + ip);
        else { char tmp[12]; int n = 0; do { tmp[n++] = (char)('0' + ip % 10); ip /= 10; } while (ip); while (n) *p++ = tmp[--n]; }
        *p++ = '.';
        p[4] = (char)('0' + x % 10); x /= 10; p[3] = (char)('0' + x % 10); x /= 10; p[2] = (char)('0' + x % 10); x /= 10; p[1] = (char)('0' + x % 10); x /= 10; p[0] = (char)('0' + x);
        p += 5;
        *p++ = '\n';
      }
      o.len = (size_t)(p - o.buf);
      continue;
    }
    for (uint32_t i = b; i < e; ++i) {
      uint32_t units;
      if constexpr (sizeof(Row) == 4) units = rows[i] & 0xFFFFu; else units = (uint32_t)(rows[i] >> 32);
      o.put(id, id_len); o.ch('\t'); o.put(t.shown[se_of(rows[i])]); o.ch('\t');
      o.u32(units / 100000u); o.ch('.');
      char d5[5];
      uint32_t x = units % 100000u;
      for (int k = 4; k >= 0; --k) { d5[k] = (char)('0' + x % 10); x /= 10; }
      o.put(d5, 5); o.ch('\n');
    }
  }
  return o.len;
}

} // namespace

extern "C" size_t krepp_format_header(const krepp_index_t* ix, const krepp_params_t* p, int tabular, const char* invocation, char* buf, size_t cap)
{
  if (!ix || !p) return 0;
  Out o(buf, cap);
  const std::string inv = invocation ? invocation : "";
  if (!p->place) { // header_dreport (ref src/krepp.cpp:311-319)
    o.put("# software: krepp\tversion: " KREPP_VERSION_STRING "\tinvocation :" + inv);
    o.put(p->summarize ? "\nREFERENCE_NAME\tWEIGHTED_COUNT\tSEQUENCE_ABUNDANCE\n" : "\nSEQ_ID\tREFERENCE_NAME\tDIST\n");
  } else if (p->summarize || tabular) { // header_preport (ref src/krepp.cpp:396-408)
    o.put("# software: krepp\tversion: " KREPP_VERSION_STRING "\tinvocation :" + inv);
    o.put("\n# ");
    o.put(ix->host.tree.jplace_newick());
    o.put(p->summarize ? "\nDISTAL_NODE\tEDGE_NUM\tWEIGHTED_COUNT\tSEQUENCE_ABUNDANCE\n" : "\nSEQ_ID\tDISTAL_NODE\tEDGE_NUM\tLWR\tDIST\n");
  } else { // begin_jplace (ref src/krepp.cpp:426-432)
    o.put("{\n\t\"version\" : 3,\n\t\"fields\" : [\"edge_num\", \"pendant_length\", \"distal_length\", \"likelihood\", \"like_weight_ratio\", \"distance\"],\n\t\"placements\" : [\n");
  }
  return o.len;
}

extern "C" size_t krepp_format_dist(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res, const char* names,
                                    const uint64_t* name_offsets, double* wcount, char* buf, size_t cap)
{
  if (!ix || !p || !res || !names || !name_offsets) return 0;
  if (res->dist_begin) {
    if (res->dist_row_bytes == 4) return format_dist_compact(ix, p, res, static_cast<const uint32_t*>(res->dist_rows), names, name_offsets, wcount, buf, cap);
    return format_dist_compact(ix, p, res, static_cast<const uint64_t*>(res->dist_rows), names, name_offsets, wcount, buf, cap);
  }
  if (res->records || !res->brief) return format_dist_rows(ix, p, res, res->records, names, name_offsets, wcount, buf, cap);
  return format_dist_rows(ix, p, res, res->brief, names, name_offsets, wcount, buf, cap);
}

extern "C" size_t krepp_format_seek(const krepp_results_t* res, const char* names, const uint64_t* name_offsets, char* buf, size_t cap)
{ // SBatch::seek_sequences' rows (ref src/seek.cpp:41-49): "<id>\t<distance>" at std::fixed precision 5, "<id>\tNaN" when nothing matched
  if (!res || !names || !name_offsets || !res->seek_dist) return 0;
  Out o(buf, cap);
  for (uint32_t r = 0; r < res->n_reads; ++r) {
    o.put(name_of(names, name_offsets, r)); o.ch('\t');
    const double d = res->seek_dist[r];
    if (d != d) o.put("NaN"); else o.fixed5(d);
    o.ch('\n');
  }
  return o.len;
}

extern "C" size_t krepp_format_place(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res, const char* names,
                                     const uint64_t* name_offsets, int tabular, int* has_previous, double* wcount, char* buf, size_t cap)
{
  if (!ix || !p || !res || !names || !name_offsets) return 0;
  Out o(buf, cap);
  const HostTree& t = ix->host.tree;
  int prev = has_previous ? *has_previous : 0;
  auto jplace_row = [&](const krepp_placement_t& q) { // PP_JPLACE_FIELDS (ref src/query.hpp:202-204)
    o.ch('['); o.u32(q.se - 1); o.put(", "); o.fixed5(q.pendant); o.put(", "); o.fixed5(q.distal); o.put(", "); o.fixed5(q.loglik);
    o.put(", "); o.fixed5(q.lwr); o.put(", "); o.fixed5(q.d_llh); o.ch(']');
  };
  auto tab_row = [&](const char* id, const krepp_placement_t& q) { // PP_TABULAR_FIELDS (ref src/query.hpp:206)
    o.put(id); o.ch('\t'); o.put(t.node_name(q.se, true)); o.ch('\t'); o.u32(q.se - 1); o.ch('\t'); o.fixed5(q.lwr); o.ch('\t'); o.fixed5(q.d_llh); o.ch('\n');
  };
  for (uint32_t r = 0; r < res->n_reads; ++r) {
    const krepp_read_summary_t& s = res->reads[r];
    if (!s.place_count) continue; // report_placement returned false (ref src/query.cpp:220-222)
    const char* id = name_of(names, name_offsets, r);
    const krepp_placement_t* q = res->placements + s.place_begin;
    const uint32_t nsel = s.n_selected; // (the records themselves are not needed here: place front ends leave them in HBM)
    const bool text = !tabular && !p->summarize;
    if (text) {
      if (prev) o.put(",\n");
      o.put("\t\t\t{\"n\" : [\""); o.put(id); o.put("\"], \"p\" : [");
      prev = 1;
    }
    if (nsel == 1) { // single reference: the closest itself (ref src/query.cpp:231-241)
      if (p->summarize) { if (wcount) wcount[q[0].se] += 1.0; }
      else if (tabular) tab_row(id, q[0]);
      else { jplace_row(q[0]); o.put("]}"); }
      continue;
    }
    if (p->multi) { // ref src/query.cpp:293-310
      for (uint32_t i = 0; i < s.place_count; ++i) {
        if (p->summarize) { if (wcount) wcount[q[i].se] += 1.0 / (double)s.place_count; }
        else if (tabular) tab_row(id, q[i]);
        else { if (i) o.ch(','); o.put("\n\t\t\t\t"); jplace_row(q[i]); }
      }
      if (text) o.put("]\n\t\t\t}");
    } else { // largest clade, then smallest distance (ref src/query.cpp:311-330; remaining ties: highest se)
      uint32_t best = 0;
      for (uint32_t i = 1; i < s.place_count; ++i) {
        const uint32_t cb = t.card[q[best].se], ci = t.card[q[i].se];
        if (ci > cb || (ci == cb && q[i].d_llh <= q[best].d_llh)) best = i;
      }
      if (p->summarize) { if (wcount) wcount[q[best].se] += 1.0; }
      else if (tabular) tab_row(id, q[best]);
      else { jplace_row(q[best]); o.put("]}"); }
    }
  }
  if (has_previous) *has_previous = prev;
  return o.len;
}

extern "C" size_t krepp_format_footer(const krepp_index_t* ix, const krepp_params_t* p, int tabular, const double* wcount, uint64_t total_queries,
                                      const char* invocation, char* buf, size_t cap)
{
  if (!ix || !p) return 0;
  Out o(buf, cap);
  const HostTree& t = ix->host.tree;
  if (p->summarize) { // ref src/krepp.cpp:385-392,492-497 (rows by ascending se; the reference's order is unspecified)
    if (!wcount) return 0;
    double total = 0;
    for (uint32_t se = 1; se <= t.nnodes; ++se) total += wcount[se];
    for (uint32_t se = 1; se <= t.nnodes; ++se) {
      if (wcount[se] == 0) continue;
      if (p->place) { o.put(t.node_name(se, true)); o.ch('\t'); o.u32(se - 1); }
      else o.put(t.node_name(se, false));
      o.ch('\t'); o.fixed5(wcount[se]); o.ch('\t'); o.fixed5(wcount[se] / total); o.ch('\n');
    }
  } else if (p->place && !tabular) { // end_jplace (ref src/krepp.cpp:410-424)
    o.put("],\n\t\"metadata\" : {\n\t\t\"software\" : \"krepp\",\n\t\t\"version\" : \"" KREPP_VERSION_STRING "\",\n"
          "\t\t\"repository\" : \"https://github.com/bo1929/krepp\",\n\t\t\"num_queries\" : \"");
    o.put(std::to_string(total_queries));
    o.put("\",\n\t\t\"invocation\" : \"");
    o.put(invocation ? invocation : "");
    o.put("\"\n\t},\n\t\"tree\" : \"");
    o.put(t.jplace_newick());
    o.put("\"\n}");
  }
  return o.len;
}

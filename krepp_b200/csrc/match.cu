// match.cu -- K1..K3 of the query path fused into one persistent kernel (sm_100a).
//
// One warp owns one read at a time (reads are claimed from a global counter, so ragged lengths balance) and walks it
// in tiles of 128 k-mer windows:
//   A0. 128-bit loads of the tile's ASCII bases -> 2-bit codes + validity bits, packed MSB-first in shared memory
//       (ref src/common.cpp:10-18 seq_nt4_table, src/common.hpp:225-243 compute/update_encoding).
//   A1. every lane extracts 4 windows from the packed stream, forms the reverse complement (ref src/common.hpp:177-186)
//       and the LSH bucket id rix = pext(bp, mask_hash_bp) (ref src/lshf.cpp:62) for both strands (ref
//       src/query.cpp:82-91); eligible lookups (ref src/index.hpp:27) are compacted into a shared-memory list.
//   A2. the compacted list is completed with all lanes busy: residual q = pext(lr, mask_drop_lr) (ref src/lshf.cpp:64-69)
//       and the bucket range [inc[off-1], inc[off]) (ref src/index.cpp:160-168, src/table.hpp:121-136); all range loads
//       of a tile are in flight together; empty buckets are dropped from the list.
//   B.  bucket scans with XOR/OR/popc (ref src/common.hpp:175, src/query.cpp:361-368), two strategies chosen by the
//       host from the index's bucket occupancy:
//       * small buckets (toy index): every lane pulls whole lookups from the list and scans its bucket with 128-bit
//         loads;
//       * large buckets (1,000-genome index, ~100 entries = ~800 contiguous bytes per lookup): the buckets of the
//         lookup list are streamed through a per-warp shared-memory ring by 1-D bulk copies (cp.async.bulk, completion
//         on an mbarrier per ring slot), several lookups ahead of the scan, so that each warp keeps kilobytes of HBM
//         reads in flight instead of one dependent load; lanes then scan the staged entries out of shared memory.
//       Hits expand their colour through the se->(se,se) DAG to leaves on the device (ref src/query.cpp:369-387) and
//       bump the per-(strand, leaf) Hamming histogram (ref src/query.hpp:153-176).
//
// The reference's Minfo::update_match keeps, per (strand, leaf, position), the MINIMUM Hamming distance over all
// matching entries.  All entries that can match one (read, strand, position) live in one bucket, so that minimum is
// formed inside a single lookup.  A lookup with one hit entry commits directly.  Otherwise every lookup gets a tag
// (a per-warp counter that only decreases) and each leaf reached does atomicMin(marker[leaf], tag << 5 | hd): the old
// value tells whether this is the leaf's first hit in this lookup (histogram[hd] += 1) or an improvement over an
// earlier, larger distance (histogram[old] -= 1, histogram[hd] += 1).  The updates telescope to exactly one count at
// the minimum whatever order the lanes run in, with one pass and no reset of the marker array.  Histograms live in a
// per-warp accumulator and are emitted as records (read, strand<<31|leaf_se, hist[0..th]) in (strand, leaf) order when
// the read is finished.
#include "device.cuh"
#include "match_common.cuh"
#include "solve.cuh"

namespace krepp {

constexpr int kWarpsLane = 8, kWarpsStaged = 14; // warps per CTA: the staged path runs one big CTA per SM so that the LUT is staged once
__host__ __device__ constexpr int warps_per_cta(bool staged) { return staged ? kWarpsStaged : kWarpsLane; }
// staged path: a ROUND is four lookups, one per group of eight lanes; each lookup's bucket (<= kChunk entries; longer
// buckets take the whole-warp path) is bulk-copied into its own ring slot, two rounds are in flight per warp
constexpr int kChunk = 128;                    // bucket entries per ring slot (16 per lane of the group)
constexpr int kSlotEntries = kChunk + 2;       // + 16-byte alignment slack at either end of a bucket
constexpr int kGroups = 4, kRoundsInFlight = 2, kSlots = kGroups * kRoundsInFlight;
constexpr int kLongCap = 16;                   // pending lookups with more than kChunk entries
constexpr int kHitCap = 96;                    // hit entries queued per warp between two resolutions
constexpr int kTabSize = 512;                  // (lookup, leaf) -> min distance table, open addressing (about 2 leaves per hit)
constexpr uint32_t kTabEmpty = 0xFFFFFFFFu;    // lookup id 127 is never handed out
constexpr uint32_t kTabMaxLeaves = 1u << 19;   // leaf ranks must fit the table key
constexpr uint32_t kTagStart = 0x07FFFFFEu;    // marker tags count down from here; 0x07FFFFFF is the rest value's tag
constexpr uint32_t kInfoLeaf = 0x80000000u, kInfoExpand = 0x40000000u; // DevIndex::cnode[].x: leaf | rank, or expand | first child (.y = second child)

struct __align__(16) WarpStage {              // staged path only
  unsigned long long bar[kRoundsInFlight];    // one mbarrier per round in flight
  uint2 ent[kSlots][kSlotEntries];            // ring slots, 16-byte aligned
  uint2 hitq[kHitCap];                        // queued hit entries: {colour id, lookup id << 25 | strand << 24 | hd}
  uint32_t tab[kTabSize];                     // lookup id << 25 | strand << 24 | leaf rank << 5 | min hd
  uint32_t lg_a[kLongCap], lg_l[kLongCap], lg_q[kLongCap]; // lookups with more than kChunk entries (first, len | strand << 31, q)
  uint32_t ovf, pad[3];
};
static_assert(sizeof(WarpStage) % 16 == 0 && (kSlotEntries * 8) % 16 == 0, "ring slots must stay 16-byte aligned");
__host__ __device__ inline size_t stage_offset(uint32_t k) { return (lut_chunks(k) * 256 * sizeof(uint4) + kWarpsStaged * sizeof(WarpSmem) + 15) & ~(size_t)15; }
__host__ __device__ inline size_t smem_bytes(uint32_t k, bool staged) { return staged ? stage_offset(k) + kWarpsStaged * sizeof(WarpStage) : lut_chunks(k) * 256 * sizeof(uint4) + kWarpsLane * sizeof(WarpSmem); }

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// The index fields the colour expansion and rescan helpers need.  Cold paths are compiled out of line (the staged
// kernel is large enough for instruction-cache misses to show up as stall_no_inst) and take this small view by value:
// handing them the kernel's DevIndex parameter by reference would force a per-thread copy of the whole struct.
struct IxView {
  const uint2* cmer;
  const uint2* cnode;
  uint32_t local_expand;
};

struct WarpCtx {
  uint32_t* acc;      // [2*nleaves*(th+1)]
  uint32_t* bitmap;   // [ceil(2*nleaves/32)]
  uint32_t* marker;   // [nleaves]
  uint32_t* stack;
  uint32_t stack_cap;
  uint32_t stride;    // th+1
  uint32_t nleaves;
  uint32_t* err;
};

__device__ __forceinline__ void commit(const WarpCtx& w, uint32_t strand, uint32_t rank, uint32_t hd)
{
  const uint32_t slot = strand * w.nleaves + rank;
  atomicAdd(&w.acc[slot * w.stride + hd], 1u);
  atomicOr(&w.bitmap[slot >> 5], 1u << (slot & 31));
}

// One leaf reached by a hit of the lookup tagged `tagbase` (= tag << 5): keeps the minimum distance per leaf and moves
// the histogram count along with it (see the header).  Any number of lanes may run this concurrently.
__device__ __forceinline__ void leaf_hit(const WarpCtx& w, uint32_t strand, uint32_t rank, uint32_t hd, uint32_t tagbase)
{
  const uint32_t old = atomicMin(&w.marker[rank], tagbase | hd);
  if ((old ^ tagbase) >> 5) commit(w, strand, rank, hd);            // first hit of this leaf in this lookup
  else if ((old & 31u) > hd) {                                        // a smaller distance than the one counted so far
    uint32_t* h = w.acc + (size_t)(strand * w.nleaves + rank) * w.stride;
    atomicSub(&h[old & 31u], 1u);
    atomicAdd(&h[hd], 1u);
  }
}

// Lane-local colour expansion: depth-first with a small private stack.  The host only enables it (ix.local_expand)
// when the deepest colour DAG of the index fits.  tagbase == 0xffffffff: the lookup has this single hit, commit directly.
__device__ __forceinline__ void expand_local(const IxView& ix, const WarpCtx& w, uint32_t se, uint32_t strand, uint32_t hd, uint32_t tagbase)
{
  uint32_t st[kLocalStack];
  int sp = 0;
  st[sp++] = se;
  while (sp) {
    const uint32_t s = st[--sp];
    const uint2 cn = __ldg(&ix.cnode[s]);
    const uint32_t ci = cn.x;
    if (ci & kInfoLeaf) { if (tagbase == 0xFFFFFFFFu) commit(w, strand, ci & 0x3FFFFFFFu, hd); else leaf_hit(w, strand, ci & 0x3FFFFFFFu, hd, tagbase); }
    else if (ci & kInfoExpand) { st[sp++] = cn.y; st[sp++] = ci & 0x3FFFFFFFu; }
  }
}

// Warp-cooperative colour expansion over the per-warp HBM stack, for colour DAGs too deep for the private stack.
__device__ __noinline__ void expand_coop(const IxView ix, const WarpCtx w, uint32_t se, uint32_t strand, uint32_t hd, uint32_t tagbase)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t size = 1;
  if (lane == 0) w.stack[0] = se;
  __syncwarp();
  while (size) {
    const uint32_t take = min(size, 32u);
    uint32_t s = 0, ci = 0;
    uint2 cn = make_uint2(0u, 0u);
    if (lane < take) { s = w.stack[size - 1 - lane]; cn = __ldg(&ix.cnode[s]); ci = cn.x; }
    __syncwarp();
    size -= take;
    if (ci & kInfoLeaf) { if (tagbase == 0xFFFFFFFFu) commit(w, strand, ci & 0x3FFFFFFFu, hd); else leaf_hit(w, strand, ci & 0x3FFFFFFFu, hd, tagbase); }
    const bool expand = !(ci & kInfoLeaf) && (ci & kInfoExpand);
    const uint32_t ex = __ballot_sync(0xFFFFFFFFu, expand);
    const uint32_t nex = __popc(ex);
    if (size + 2 * nex > w.stack_cap) { if (lane == 0) atomicOr(w.err, kErrStackOverflow); return; }
    if (expand) {
      const uint32_t at = size + 2 * __popc(ex & ((1u << lane) - 1));
      w.stack[at] = ci & 0x3FFFFFFFu; w.stack[at + 1] = cn.y;
    }
    size += 2 * nex;
    __syncwarp();
  }
}

// Hits of one lookup held one per lane (hit lanes have hd <= th): expand all of them.
__device__ __forceinline__ void expand_hits(const IxView& ix, const WarpCtx& w, bool hit, uint32_t se, uint32_t hd, uint32_t strand, uint32_t tagbase)
{
  if (ix.local_expand) { if (hit) expand_local(ix, w, se, strand, hd, tagbase); }
  else {
    uint32_t hits = __ballot_sync(0xFFFFFFFFu, hit);
    while (hits) {
      const int src = __ffs(hits) - 1;
      hits &= hits - 1;
      expand_coop(ix, w, __shfl_sync(0xFFFFFFFFu, se, src), strand, __shfl_sync(0xFFFFFFFFu, hd, src), tagbase);
    }
  }
}

// A fresh tag for the next lookup that needs the marker array (warp-uniform).  Tags only decrease, so a new lookup's
// atomicMin always wins over whatever an older lookup left behind; when the counter runs out the markers are reset.
__device__ __forceinline__ uint32_t next_tag(const WarpCtx& w, uint32_t& tag)
{
  if (tag <= 1u) {
    for (uint32_t i = threadIdx.x & 31; i < w.nleaves; i += 32) w.marker[i] = 0xFFFFFFFFu;
    __syncwarp();
    tag = kTagStart;
  } else --tag;
  return tag << 5;
}

// Staged path: one leaf reached by a queued hit goes into the warp's shared-memory table keyed by (lookup, strand, leaf);
// the value keeps the minimum distance.  Any number of lanes may insert concurrently (shared-memory atomics only).
__device__ __forceinline__ void tab_insert(WarpStage* stg, uint32_t v)
{
  uint32_t slot = ((v >> 5) * 0x9E3779B1u) >> 23; // 9 bits = kTabSize
  for (int probes = 0; probes < 96; ++probes) {
    const uint32_t old = atomicCAS(&stg->tab[slot], kTabEmpty, v);
    if (old == kTabEmpty) return;
    if (((old ^ v) >> 5) == 0) { atomicMin(&stg->tab[slot], v); return; }
    slot = (slot + 1) & (kTabSize - 1);
  }
  stg->ovf = 1; // too crowded: the whole batch is redone through the marker path
}

__device__ __forceinline__ void expand_to_table(const IxView& ix, WarpStage* stg, uint32_t se, uint32_t meta)
{
  uint32_t st[kLocalStack];
  int sp = 0;
  st[sp++] = se;
  while (sp) {
    const uint32_t s = st[--sp];
    const uint2 cn = __ldg(&ix.cnode[s]);
    const uint32_t ci = cn.x;
    if (ci & kInfoLeaf) tab_insert(stg, (meta & 0xFF000000u) | ((ci & 0x7FFFFu) << 5) | (meta & 31u));
    else if (ci & kInfoExpand) { st[sp++] = cn.y; st[sp++] = ci & 0x3FFFFFFFu; }
  }
}

// More distinct (lookup, leaf) pairs than the table holds (rare): the queue, which is ordered by lookup id, is replayed
// lookup by lookup through the markers instead.  Returns the updated tag counter.
__device__ __noinline__ uint32_t resolve_overflow(const IxView ix, const WarpCtx w, WarpStage* stg, uint32_t n_hits, uint32_t tag)
{
  const uint32_t lane = threadIdx.x & 31;
  if (lane == 0) stg->ovf = 0;
  uint32_t pos = 0, last_id = 0xFFFFFFFFu, tagbase = 0;
  while (pos < n_hits) {
    const uint32_t id = stg->hitq[pos].y >> 25;
    if (id != last_id) { tagbase = next_tag(w, tag); last_id = id; }
    const uint32_t i = pos + lane;
    uint2 h = make_uint2(0u, 0u);
    if (i < n_hits) h = stg->hitq[i];
    const uint32_t same = __ballot_sync(0xFFFFFFFFu, i < n_hits && (h.y >> 25) == id);
    const uint32_t run = __ffs(~same) - 1; // entries of this lookup at the head of the window (>= 1)
    if (lane < run) expand_local(ix, w, h.x, (h.y >> 24) & 1u, h.y & 31u, tagbase);
    __syncwarp();
    pos += run;
  }
  return tag;
}

// Resolves the queued hits of a warp: colours are expanded with all lanes busy, leaves are deduplicated per lookup in the
// shared-memory table, and every surviving (lookup, leaf) bumps the histogram once at its minimum distance.
__device__ __forceinline__ void resolve_hits(const IxView& ix, const WarpCtx& w, WarpStage* stg, uint32_t n_hits, uint32_t n_ids, uint32_t& tag)
{
  const uint32_t lane = threadIdx.x & 31;
  __syncwarp();
  for (uint32_t i = lane; i < n_hits; i += 32) { const uint2 h = stg->hitq[i]; expand_to_table(ix, stg, h.x, h.y); }
  __syncwarp();
  const bool overflow = *reinterpret_cast<volatile uint32_t*>(&stg->ovf) != 0;
  for (uint32_t slot = lane; slot < (uint32_t)kTabSize; slot += 32) {
    const uint32_t v = stg->tab[slot];
    if (v == kTabEmpty) continue;
    stg->tab[slot] = kTabEmpty;
    if (!overflow) commit(w, (v >> 24) & 1u, (v >> 5) & 0x7FFFFu, v & 31u);
  }
  __syncwarp();
  if (overflow) tag = resolve_overflow(ix, w, stg, n_hits, tag);
}

// Small-bucket path: one lookup with several hit entries (or a colour too deep for the private stack), rescanned by the
// whole warp in a single pass.
__device__ __forceinline__ uint32_t careful_lookup(const IxView& ix, const WarpCtx& w, uint32_t begin, uint32_t len, uint32_t q, uint32_t strand, uint32_t th, uint32_t tagbase)
{
  const uint32_t lane = threadIdx.x & 31;
  uint32_t best = 0xFFFFFFFFu; // this lane's smallest distance among the hit entries
  __syncwarp(); // lanes still expanding hits of the previous lookup must not meet this lookup's (smaller) tag in the markers
  for (uint32_t base = 0; base < len; base += 32) {
    uint32_t se = 0, hd = 0xFFFFFFFFu;
    if (base + lane < len) {
      const uint2 e = __ldg(&ix.cmer[(size_t)begin + base + lane]);
      const uint32_t z = e.x ^ q;
      hd = __popc((z | (z >> 16)) & 0xFFFFu);
      se = e.y;
    }
    if (hd <= th) best = min(best, hd);
    expand_hits(ix, w, hd <= th, se, hd, strand, tagbase);
  }
  return best;
}

// Staged path, rare: lookups whose bucket does not fit the ring are scanned by the whole warp straight from HBM and
// deduplicated through the markers.  Returns {tag, filt0, filt1} updated.
__device__ __noinline__ uint3 long_lookups(const IxView ix, const WarpCtx w, const WarpStage* stg, uint32_t n_long, uint32_t th, uint32_t tag,
                                           uint32_t filt0, uint32_t filt1)
{
  __syncwarp();
  for (uint32_t j = 0; j < n_long; ++j) {
    const uint32_t l = stg->lg_l[j];
    const uint32_t best = careful_lookup(ix, w, stg->lg_a[j], l & 0x7FFFFFFFu, stg->lg_q[j], l >> 31, th, next_tag(w, tag));
    if (l >> 31) filt1 = min(filt1, best); else filt0 = min(filt0, best);
  }
  __syncwarp();
  return make_uint3(tag, filt0, filt1);
}

// Staged path, rare: a round with more hit entries than the queue holds (or an index whose colour DAG is too deep for
// the table path) is resolved lookup by lookup under marker tags.  Lane state of the round comes in by value.
__device__ __noinline__ uint3 marker_round(const IxView ix, const WarpCtx w, const uint2* slot, uint32_t hmask, uint32_t nbits, uint32_t cq, uint32_t cs,
                                           uint32_t shared_id, uint32_t tag, uint32_t filt0, uint32_t filt1)
{
  const uint32_t lane = threadIdx.x & 31, gl = lane & 7u, grp = lane >> 3;
  const bool one_lookup = __shfl_sync(0xFFFFFFFFu, shared_id, 0) != 0; // a long bucket spread over the groups
  uint32_t tagbase = 0;
  for (uint32_t g = 0; g < (uint32_t)kGroups; ++g) {
    if (!__any_sync(0xFFFFFFFFu, grp == g && hmask != 0)) continue;
    if (!one_lookup || !tagbase) tagbase = next_tag(w, tag);
    const uint32_t gcs = __shfl_sync(0xFFFFFFFFu, cs, 8 * g);
    __syncwarp();
    for (uint32_t t = 0; t < nbits; ++t) {
      const bool hit = grp == g && ((hmask >> (nbits - 1u - t)) & 1u);
      if (!__any_sync(0xFFFFFFFFu, hit)) continue;
      uint32_t se = 0, hd = 0;
      if (hit) { const uint2 e = slot[2u * gl + 16u * (t >> 1) + (t & 1u)]; const uint32_t z = e.x ^ cq; hd = __popc((z | (z >> 16)) & 0xFFFFu); se = e.y; if (cs) filt1 = min(filt1, hd); else filt0 = min(filt0, hd); }
      expand_hits(ix, w, hit, se, hd, gcs, tagbase);
    }
    __syncwarp();
  }
  return make_uint3(tag, filt0, filt1);
}

// Emits one read's records in (strand, leaf) order from the warp's accumulator and clears it (see the header); returns
// the number of records.  `list` is 512 words of the warp's shared memory.
__device__ __forceinline__ uint32_t emit_records(const DevIndex& ix, const MatchArgs& a, const WarpCtx& w, uint32_t* list, uint32_t read,
                                                 uint32_t filt0, uint32_t filt1, uint32_t& rbegin_out, bool& fits_out)
{
  const uint32_t lane = threadIdx.x & 31, nleaves = w.nleaves, stride = w.stride, nbm = (2 * nleaves + 31) >> 5;
  __syncwarp();
  uint32_t total = 0;
  if (!a.keep_all) {
    // The hdist_filt gate of summarize_matches (ref src/query.cpp:101-106,116-119) applied where the histograms are
    // born: a (strand, leaf) pair whose smallest distance exceeds 2 * hdist_filt[strand] + 1 is never solved,
    // reported or looked at again by the reference, so it is dropped here (its accumulator is cleared) instead of
    // being written out, gated, merged and copied to the host.  keep_all (parity tap 2) keeps every pair.
    const uint32_t g0 = 2u * filt0 + 1u, g1 = 2u * filt1 + 1u; // uint32 wrap kept, as in the reference
    for (uint32_t wbase = 0; wbase < nbm; wbase += 16) {
      uint32_t bits = (lane < 16 && wbase + lane < nbm) ? __ldcg(&w.bitmap[wbase + lane]) : 0u;
      const uint32_t c = __popc(bits);
      uint32_t incl = c;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
      const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, incl, 31);
      if (!cnt) continue;
      uint32_t li = incl - c;
      while (bits) { const uint32_t b = __ffs(bits) - 1; bits &= bits - 1; list[li++] = (wbase + lane) * 32 + b; }
      __syncwarp();
      for (uint32_t i = lane; i < ((cnt + 31u) & ~31u); i += 32) {
        bool pass = false;
        if (i < cnt) {
          const uint32_t slot = list[i];
          uint32_t* h = w.acc + (size_t)slot * stride;
          uint32_t hdmin = 0xFFFFFFFFu;
          for (uint32_t x = 0; x < stride; ++x) if (__ldcg(&h[x]) && hdmin == 0xFFFFFFFFu) hdmin = x;
          pass = !(hdmin > (slot >= nleaves ? g1 : g0));
          if (!pass) {
            for (uint32_t x = 0; x < stride; ++x) h[x] = 0;
            atomicAnd(&w.bitmap[slot >> 5], ~(1u << (slot & 31)));
          }
        }
        total += __popc(__ballot_sync(0xFFFFFFFFu, pass));
      }
      __syncwarp();
    }
  } else {
    for (uint32_t wbase = 0; wbase < nbm; wbase += 32) {
      const uint32_t bits = (wbase + lane < nbm) ? __ldcg(&w.bitmap[wbase + lane]) : 0u;
      uint32_t c = __popc(bits);
      for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
      total += c;
    }
  }
  uint32_t rbegin = 0;
  if (lane == 0 && total) rbegin = atomicAdd(a.counters, total);
  rbegin = __shfl_sync(0xFFFFFFFFu, rbegin, 0);
  const bool fits = (uint64_t)rbegin + total <= a.rec_cap;
  if (!fits && lane == 0) atomicOr(a.counters + 2, kErrRecOverflow);
  // records in ascending slot order; 16 bitmap words (<= 512 slots) per round: lanes 0..15 list the set bits of their
  // word in shared memory, then every lane takes whole records, so clustered leaves do not pile up on one lane
  uint32_t done = 0;
  for (uint32_t wbase = 0; wbase < nbm; wbase += 16) {
    uint32_t bits = (lane < 16 && wbase + lane < nbm) ? __ldcg(&w.bitmap[wbase + lane]) : 0u;
    const uint32_t c = __popc(bits);
    uint32_t incl = c;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
    const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (!cnt) continue;
    if (bits) w.bitmap[wbase + lane] = 0;
    uint32_t li = incl - c;
    while (bits) { const uint32_t b = __ffs(bits) - 1; bits &= bits - 1; list[li++] = (wbase + lane) * 32 + b; }
    __syncwarp();
    for (uint32_t i = lane; i < cnt; i += 32) {
      const uint32_t slot = list[i], at = rbegin + done + i;
      const uint32_t strand = slot >= nleaves, rank = slot - strand * nleaves;
      uint32_t* h = w.acc + (size_t)slot * stride;
      uint32_t hv[kMaxTh + 1];
      for (uint32_t x = 0; x < stride; ++x) hv[x] = __ldcg(&h[x]);
      for (uint32_t x = 0; x < stride; ++x) h[x] = 0;
      if (fits) {
        a.rec_read[at] = read;
        a.rec_slot[at] = strand << 31 | ix.leaf_se[rank];
        for (uint32_t x = 0; x < stride; ++x) a.rec_hist[(size_t)at * stride + x] = hv[x];
      }
    }
    done += cnt;
    __syncwarp();
  }
  rbegin_out = rbegin; fits_out = fits;
  return total;
}

template <bool STAGED, bool TAP>
__global__ void __launch_bounds__(warps_per_cta(STAGED) * 32, STAGED ? 1 : 4) match_kernel(const DevIndex ix, const MatchArgs a)
{
  constexpr int kWarpsPerCta = warps_per_cta(STAGED);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nchunks = lut_chunks(ix.k);
  uint4* lut = reinterpret_cast<uint4*>(smem_raw);
  WarpSmem* smem = reinterpret_cast<WarpSmem*>(smem_raw + nchunks * 256 * sizeof(uint4));
  for (uint32_t i = threadIdx.x; i < nchunks * 256; i += blockDim.x) lut[i] = ix.lut[i];
  const bool wide = nchunks > 7;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * kWarpsPerCta + warp;
  WarpSmem& sm = smem[warp];
  WarpStage* stg = STAGED ? reinterpret_cast<WarpStage*>(smem_raw + stage_offset(ix.k)) + warp : nullptr;
  if (STAGED) {
    for (uint32_t i = lane; i < (uint32_t)kTabSize; i += 32) stg->tab[i] = kTabEmpty;
    if (lane == 0) {
      stg->ovf = 0;
      for (int i = 0; i < kRoundsInFlight; ++i) mbar_init(smem_u32(&stg->bar[i]), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  __syncthreads();
  const uint32_t k = ix.k, th = a.th, stride = th + 1, nleaves = ix.nleaves;
  const uint32_t nslots = 2 * nleaves, nbm = (nslots + 31) >> 5;
  const uint32_t lt_mask = (1u << lane) - 1;

  WarpCtx w;
  w.acc = a.acc + (size_t)gwarp * nslots * stride;
  w.bitmap = a.bitmap + (size_t)gwarp * nbm;
  w.marker = a.marker + (size_t)gwarp * nleaves;
  w.stack = a.stack + (size_t)gwarp * a.stack_cap;
  w.stack_cap = a.stack_cap; w.stride = stride; w.nleaves = nleaves; w.err = a.counters + 2;
  IxView iv; iv.cmer = ix.cmer; iv.cnode = ix.cnode; iv.local_expand = ix.local_expand;
  uint32_t tag = min(a.tagctr[gwarp], kTagStart);   // persists across launches: markers are never cleared in between
  uint32_t p_round = 0, c_round = 0;                // rounds issued / consumed by this warp over the whole launch (mbarrier phases)
  const bool use_table = STAGED && ix.local_expand && nleaves <= kTabMaxLeaves;
  const uint32_t bar0 = STAGED ? smem_u32(&stg->bar[0]) : 0u, ent0 = STAGED ? smem_u32(&stg->ent[0][0]) : 0u;

  unsigned long long st_bytes = 0, st_lookups = 0, st_entries = 0;

  uint32_t claim = 0, claim_end = 0;
  for (;;) {
    if (claim == claim_end) { // claim the next kClaim reads for this warp
      if (lane == 0) claim = atomicAdd(a.counters + 1, kClaim);
      claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
      claim_end = min(claim + kClaim, a.n_reads);
      if (claim >= a.n_reads) break;
    }
    const uint32_t read = claim++;
    uint64_t off, len;
    read_span(a, read, off, len);
    uint32_t onmers = 0, wn0 = 0, wn1 = 0, filt0 = 0xFFFFFFFFu, filt1 = 0xFFFFFFFFu;
    st_bytes += (lane == 0) ? len : 0;

    for (uint64_t t0 = 0; t0 + k <= len; t0 += kTileWindows) {
      const uint32_t nl = tile_lookups<TAP>(ix, a, sm, lut, wide, read, off, len, t0, onmers, wn0, wn1);

      // ---- A2. bucket ranges [inc[off-1], inc[off]) with all loads of the tile in flight; empty buckets are dropped
      //          (in-place compaction: writes trail reads)
      uint32_t nout = 0, n_long = 0;
      auto flush_long = [&]() { // long buckets: scanned by the whole warp straight from HBM, deduplicated through the markers
        const uint3 r = long_lookups(iv, w, stg, n_long, th, tag, filt0, filt1);
        tag = r.x; filt0 = r.y; filt1 = r.z;
        n_long = 0;
      };
      for (uint32_t base = 0; base < nl; base += 128) {
        uint32_t qv[4], bg[4], en[4], sb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = base + 32 * u + lane;
          qv[u] = 0; bg[u] = 0; en[u] = 0; sb[u] = 0;
          if (i < nl) {
            const uint32_t ob = sm.lk_a[i], offset = ob & 0x7FFFFFFFu;
            qv[u] = sm.lk_q[i];
            bg[u] = offset ? __ldg(&ix.inc32[offset - 1]) : 0u;
            en[u] = __ldg(&ix.inc32[offset]);
            sb[u] = ob & 0x80000000u;
          }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t i = base + 32 * u + lane;
          const uint32_t blen = en[u] - bg[u];
          if (i < nl) { st_lookups += 1; st_entries += blen; }
          const bool lng = STAGED && i < nl && blen > (uint32_t)kChunk; // staged path: too long for a ring slot
          if (STAGED) {
            const uint32_t lm = __ballot_sync(0xFFFFFFFFu, lng);
            if (lm) {
              if (n_long + __popc(lm) > (uint32_t)kLongCap) { flush_long(); }
              if (lng) { const uint32_t o = n_long + __popc(lm & lt_mask); stg->lg_a[o] = bg[u]; stg->lg_l[o] = blen | sb[u]; stg->lg_q[o] = qv[u]; }
              n_long += __popc(lm);
            }
          }
          const bool keep = i < nl && blen != 0 && !lng;
          const uint32_t km = __ballot_sync(0xFFFFFFFFu, keep);
          if (keep) {
            const uint32_t o = nout + __popc(km & lt_mask);
            sm.lk_a[o] = bg[u]; sm.lk_l[o] = blen | sb[u]; sm.lk_q[o] = qv[u];
          }
          nout += __popc(km);
        }
        __syncwarp();
      }

      if (STAGED && n_long) {
        // Buckets of up to kGroups * kChunk entries become rounds of their own at the end of the list: one slot per
        // kChunk entries, all groups under the lookup's single id (bit 30 of the length word).  Anything longer, or
        // whatever does not fit the list, takes the whole-warp path.
        bool fits = ((nout + 3u) & ~3u) + (uint32_t)kGroups * n_long <= (uint32_t)kMaxLookups;
        for (uint32_t j = 0; j < n_long && fits; ++j) fits = (stg->lg_l[j] & 0x7FFFFFFFu) <= (uint32_t)(kGroups * kChunk);
        if (!fits) flush_long();
        else {
          __syncwarp();
          for (uint32_t i = nout + lane; i < ((nout + 3u) & ~3u); i += 32) { sm.lk_a[i] = 0; sm.lk_l[i] = 0; sm.lk_q[i] = 0; } // idle groups
          nout = (nout + 3u) & ~3u;
          for (uint32_t i = lane; i < (uint32_t)kGroups * n_long; i += 32) {
            const uint32_t j = i / kGroups, c = i % kGroups, l = stg->lg_l[j], blen = l & 0x7FFFFFFFu;
            const uint32_t cnt = blen > c * kChunk ? min((uint32_t)kChunk, blen - c * kChunk) : 0u;
            sm.lk_a[nout + i] = stg->lg_a[j] + c * kChunk; sm.lk_l[nout + i] = cnt | (l & 0x80000000u) | 0x40000000u; sm.lk_q[nout + i] = stg->lg_q[j];
          }
          nout += (uint32_t)kGroups * n_long;
          n_long = 0;
          __syncwarp();
        }
      }

      // ---- B. bucket scans.
      if (!STAGED) {
        // Small buckets: every lane pulls whole lookups from the list and scans its bucket two entries (one 128-bit
        // load) per step.  No warp-collective sits in this loop; lanes leave it independently.
        constexpr uint32_t kCareful = 0x40000000u;
        uint32_t idx = 0, cq = 0, cs = 0, e = 0, lo = 0, hi = 0, cnt = 0, fse = 0, fhd = 0;
        bool have = false;
        for (;;) {
          if (!have) {
            idx = atomicAdd(&sm.cursor, 1u);
            if (idx >= nout) break;
            const uint32_t l = sm.lk_l[idx];
            lo = sm.lk_a[idx]; hi = lo + (l & 0x3FFFFFFFu); cs = l >> 31; cq = sm.lk_q[idx];
            e = lo & ~1u; cnt = 0; have = true;
          }
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(ix.cmer + e));
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(ix.cmer + ((e + 2 < hi) ? e + 2 : e)));
#define KREPP_SCAN1(enc, sev, ok)                                                           \
          { const uint32_t z = (enc) ^ cq; const uint32_t hd = __popc((z | (z >> 16)) & 0xFFFFu); \
            if (hd <= th && (ok)) { if (!cnt) { fse = (sev); fhd = hd; } ++cnt;            \
              if (cs) filt1 = min(filt1, hd); else filt0 = min(filt0, hd); } }
          KREPP_SCAN1(v.x, v.y, e >= lo)
          KREPP_SCAN1(v.z, v.w, e + 1 < hi)
          KREPP_SCAN1(u.x, u.y, e + 2 < hi)
          KREPP_SCAN1(u.z, u.w, e + 3 < hi)
#undef KREPP_SCAN1
          e += 4;
          if (e >= hi) {
            have = false;
            if (cnt == 1) { // exactly one hit entry: no other entry can lower a leaf's distance
              const uint32_t ci = __ldg(&ix.cnode[fse]).x;
              if (ci & kInfoLeaf) commit(w, cs, ci & 0x3FFFFFFFu, fhd);
              else if (ci & kInfoExpand) { if (ix.local_expand) expand_local(iv, w, fse, cs, fhd, 0xFFFFFFFFu); else cnt = 2; }
            }
            if (cnt > 1) sm.lk_l[idx] |= kCareful; // several hit entries: handled by the whole warp below
          }
        }
        __syncwarp();
        for (uint32_t base = 0; base < nout; base += 32) {
          const uint32_t l = (base + lane < nout) ? sm.lk_l[base + lane] : 0u;
          uint32_t need = __ballot_sync(0xFFFFFFFFu, (l & kCareful) != 0);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint32_t ll = sm.lk_l[base + src];
            careful_lookup(iv, w, sm.lk_a[base + src], ll & 0x3FFFFFFFu, sm.lk_q[base + src], ll >> 31, th, next_tag(w, tag));
          }
        }
      } else {
        // Large buckets: rounds of four lookups, one per group of eight lanes.  The four buckets of a round are bulk-copied
        // into four ring slots under one mbarrier, two rounds ahead of the scan; every lane then compares 16 rows of
        // its group's slot.  Hit entries are queued and resolved in batches (resolve_hits) so that colour expansion and
        // deduplication run with all lanes busy.  All cursors are warp-uniform.
        constexpr uint32_t kSlotBytes = kSlotEntries * 8;
        const uint32_t gl = lane & 7u, grp = lane >> 3;
        uint32_t p_idx = 0, c_idx = 0, n_hits = 0, lk_id = 0;
        auto issue_round = [&]() {
          const uint32_t idx = p_idx + grp;
          const bool leader = gl == 0 && idx < nout;
          uint32_t a0 = 0, bytes = 0;
          if (leader) {
            const uint32_t first = sm.lk_a[idx], cnt = sm.lk_l[idx] & 0x3FFFFFFFu;
            a0 = first & ~1u;                                     // 16-byte aligned source range (cmer is padded)
            bytes = cnt ? (((first + cnt + 1) & ~1u) - a0) * 8u : 0u;
          }
          uint32_t total = bytes;
          total += __shfl_xor_sync(0xFFFFFFFFu, total, 8);
          total += __shfl_xor_sync(0xFFFFFFFFu, total, 16);
          const uint32_t bar = bar0 + 8u * (p_round & 1u);
          if (lane == 0) mbar_expect_tx(bar, total);
          if (leader && bytes) bulk_g2s(ent0 + kSlotBytes * ((p_round & 1u) * kGroups + grp), ix.cmer + a0, bytes, bar);
          p_idx = min(p_idx + (uint32_t)kGroups, nout);
          ++p_round;
        };
        while (p_idx < nout && p_round - c_round < (uint32_t)kRoundsInFlight) issue_round();
        while (c_idx < nout) {
          const uint32_t idx = c_idx + grp;
          uint32_t cnt = 0, cs = 0, cq = 0, first = 0, shared_id = 0;
          if (idx < nout) { const uint32_t l = sm.lk_l[idx]; cnt = l & 0x3FFFFFFFu; cs = l >> 31; shared_id = (l >> 30) & 1u; cq = sm.lk_q[idx]; first = sm.lk_a[idx]; }
          const uint32_t lo = first & 1u, hi = lo + cnt;           // this lookup's entries inside its (16-byte aligned) slot
          const uint32_t nrows = (__reduce_max_sync(0xFFFFFFFFu, cnt ? hi : 0u) + 15u) >> 4; // rows of 16 entries: two per lane of the group
          mbar_wait(bar0 + 8u * (c_round & 1u), (c_round >> 1) & 1u);
          const uint2* slot = &stg->ent[(c_round & 1u) * kGroups + grp][0];
          const uint4* sp = reinterpret_cast<const uint4*>(slot) + gl;
          // one bit per entry compared, shifted in from the right: the t-th entry this lane looks at (slot index
          // 2 gl + 16 (t >> 1) + (t & 1)) ends at bit nbits-1-t
          uint32_t hmask = 0;
          const uint32_t thp1 = th + 1, nbits = 2u * nrows;
#define KREPP_CMP(enc) { const uint32_t z = (enc) ^ cq; hmask = __funnelshift_l(__popc((z | (z >> 16)) & 0xFFFFu) - thp1, hmask, 1); } // hd <= th <=> sign bit of hd - (th + 1)
          uint32_t r = 0;
          for (; r + 2 <= nrows; r += 2) {
            const uint4 e0 = sp[8 * r], e1 = sp[8 * r + 8];          // past the bucket: stale bytes of the slot, masked below
            KREPP_CMP(e0.x) KREPP_CMP(e0.z) KREPP_CMP(e1.x) KREPP_CMP(e1.z)
          }
          if (r < nrows) { const uint4 e0 = sp[8 * r]; KREPP_CMP(e0.x) KREPP_CMP(e0.z) }
#undef KREPP_CMP
          { // entries this lane really owns: lo <= slot index < hi
            const uint32_t rem = hi > 2u * gl ? hi - 2u * gl : 0u;
            const uint32_t own = 2u * (rem >> 4) + min(rem & 15u, 2u);                   // t = 0 .. own-1
            uint32_t keep = own ? (0xFFFFFFFFu >> (32u - own)) << (nbits - own) : 0u;    // their (reversed) bit positions
            if (gl == 0 && lo) keep &= ~(1u << (nbits - 1u));                              // slot entry 0 belongs to the bucket before
            hmask &= cnt ? keep : 0u;
          }
          if (__any_sync(0xFFFFFFFFu, hmask != 0)) {
            const uint32_t mine = __popc(hmask), total = __reduce_add_sync(0xFFFFFFFFu, mine);
            if (!use_table || total > (uint32_t)kHitCap) {
              // marker path (out of line): the round's lookups one after the other, each under its own tag
              const uint3 r3 = marker_round(iv, w, slot, hmask, nbits, cq, cs, shared_id, tag, filt0, filt1);
              tag = r3.x; filt0 = r3.y; filt1 = r3.z;
            } else {
              if (n_hits + total > (uint32_t)kHitCap || lk_id + (uint32_t)kGroups > 126u) { resolve_hits(iv, w, stg, n_hits, lk_id, tag); n_hits = 0; lk_id = 0; }
              uint32_t incl = mine;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += t; }
              uint32_t at = n_hits + incl - mine, m = hmask;
              const uint32_t meta = (lk_id + (shared_id ? 0u : grp)) << 25 | cs << 24;
              while (m) {
                const uint32_t bit = 31u - __clz(m), t = nbits - 1u - bit; // highest bit first = ascending entry
                m ^= 1u << bit;
                const uint2 e = slot[2u * gl + 16u * (t >> 1) + (t & 1u)];
                const uint32_t z = e.x ^ cq, hd = __popc((z | (z >> 16)) & 0xFFFFu);
                if (cs) filt1 = min(filt1, hd); else filt0 = min(filt0, hd);
                stg->hitq[at++] = make_uint2(e.y, meta | hd);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.cnode + e.y)); // resolved later: have the colour node on its way
              }
              n_hits += total;
            }
          }
          lk_id += kGroups;
          __syncwarp();               // every lane is done with the round's slots: they may be refilled
          c_idx = min(c_idx + (uint32_t)kGroups, nout);
          ++c_round;
          if (p_idx < nout) issue_round();
        }
        if (n_hits) resolve_hits(iv, w, stg, n_hits, lk_id, tag);
      }
      __syncwarp();
    }

    // ---- per-read scalars
    for (int o = 16; o; o >>= 1) {
      filt0 = min(filt0, __shfl_xor_sync(0xFFFFFFFFu, filt0, o)); filt1 = min(filt1, __shfl_xor_sync(0xFFFFFFFFu, filt1, o));
    }
    // ---- emit this read's records in (strand, leaf) order and reset the accumulator
    uint32_t rbegin; bool fits;
    const uint32_t total = emit_records(ix, a, w, sm.lk_a, read, filt0, filt1, rbegin, fits); // the lookup list is dead here (512 words with lk_l)
    if (lane == 0) {
      a.onmers[read] = onmers; a.wn[2 * read] = wn0; a.wn[2 * read + 1] = wn1;
      a.hdfilt[2 * read] = filt0; a.hdfilt[2 * read + 1] = filt1;
      a.rec_begin[read] = fits ? rbegin : 0; a.rec_count[read] = fits ? total : 0;
      st_bytes += 64ull * total;
    }
    __syncwarp();
  }
  if (lane == 0) a.tagctr[gwarp] = tag;
  // ---- roofline accounting
  for (int o = 16; o; o >>= 1) {
    st_bytes += __shfl_xor_sync(0xFFFFFFFFu, st_bytes, o);
    st_lookups += __shfl_xor_sync(0xFFFFFFFFu, st_lookups, o);
    st_entries += __shfl_xor_sync(0xFFFFFFFFu, st_entries, o);
  }
  if (lane == 0) { atomicAdd(a.stats, st_bytes + 16ull * st_lookups + 8ull * st_entries); atomicAdd(a.stats + 1, st_lookups); atomicAdd(a.stats + 2, st_entries); }
}

// ------------------------------------------------------------------------------------------------ host launchers

template <bool STAGED, bool TAP>
static cudaError_t prepare(uint32_t k)
{
  return cudaFuncSetAttribute(match_kernel<STAGED, TAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(k, STAGED));
}

// Resident warps of the persistent grid; 0 when the staged variant does not fit this device's shared memory for this k
// (the caller then uses the lane-per-bucket variant).
int match_resident_warps(int device, uint32_t k, bool staged)
{
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e;
  if (staged) { e = prepare<true, false>(k); if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel<true, false>, kWarpsStaged * 32, smem_bytes(k, true)); }
  else { e = prepare<false, false>(k); if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel<false, false>, kWarpsLane * 32, smem_bytes(k, false)); }
  if (e != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
  if (per_sm < 1) { if (staged) return 0; per_sm = 1; }
  return sms * per_sm * warps_per_cta(staged);
}

template <bool STAGED>
static cudaError_t launch_s(const DevIndex& ix, const MatchArgs& a, int grid, bool tap, cudaStream_t stream)
{
  cudaError_t e = tap ? prepare<STAGED, true>(ix.k) : prepare<STAGED, false>(ix.k);
  if (e != cudaSuccess) return e;
  if (tap) match_kernel<STAGED, true><<<grid, warps_per_cta(STAGED) * 32, smem_bytes(ix.k, STAGED), stream>>>(ix, a);
  else match_kernel<STAGED, false><<<grid, warps_per_cta(STAGED) * 32, smem_bytes(ix.k, STAGED), stream>>>(ix, a);
  return cudaGetLastError();
}

cudaError_t launch_match(const DevIndex& ix, const MatchArgs& a, int resident_warps, bool staged, bool tap, cudaStream_t stream)
{
  const int grid = resident_warps / warps_per_cta(staged);
  return staged ? launch_s<true>(ix, a, grid, tap, stream) : launch_s<false>(ix, a, grid, tap, stream);
}

} // namespace krepp

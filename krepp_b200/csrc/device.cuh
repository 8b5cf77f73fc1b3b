// Device-side data model shared by the kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace krepp {

constexpr int kMaxRuns = 18;      // a pext mask over k <= 32 two-bit groups has at most 17 runs
constexpr int kMaxResidues = 64;  // m (modulo of the LSH partition, ref src/krepp.hpp:40) supported on device
constexpr int kMaxTh = 16;

constexpr int kLutChunks = 8;     // bytes of the 64-bit k-mer word

// Flat index image resident in HBM (see index_image.hpp for the on-disk format it mirrors).
struct DevIndex {
  const uint2* cmer;        // nkmers x (enc32, se)               ref FlatHT::cmer_v  src/table.hpp:143
  const uint32_t* inc32;    // nrows cumulative bucket ends, narrowed to 32 bits (nkmers < 2^32)
                            //                                     ref FlatHT::inc_v   src/table.hpp:142
  const uint2* pse;         // nsubsets x (first, second)          ref CRecord::se_to_pse src/record.hpp:103
  const uint8_t* kind;      // nsubsets: 0 drop, 1 leaf, 2 expand  ref src/query.cpp:373-386
  const uint2* cnode;       // nsubsets: colour DAG node in one 8-byte word: x = 0 drop, 0x80000000 | leaf rank, or
                            //           0x40000000 | first child with y = second child
  const double* rho;        // by se, scaled                        ref CRecord::se_to_rho src/record.hpp:104
  const uint32_t* leaf_rank;// by se
  const uint32_t* leaf_se;  // by rank
  const uint32_t* parent;   // by se (0 for root)                  ref Node::parent
  const uint32_t* nchildren;// by se
  const uint32_t* eff_nchildren; // by se: children with an indexed reference below (= nchildren unless place -t replaced the tree)
  const double* blen;       // by se
  const uint32_t* depth;    // by se: number of ancestors
  const uint32_t* logw;     // by se: log2 of the product of the ancestors' child counts when they are all powers of two (HostTree::logw)
  const uint32_t* subtree;  // by se: number of nodes in the subtree rooted there (post-order => se range (se-subtree, se])
  uint64_t nkmers;
  uint32_t nrows, nsubsets, nnodes, nleaves;
  uint32_t row0, nrows_local; // bucket-range shard (SURVEY.md 8e mode B): cmer / inc32 cover rows [row0, row0 + nrows_local) only
  uint32_t k, h, m, m_shift; // m_shift = log2(m) when m is a power of two, else 0xffffffff
  uint32_t local_expand;    // 1 when the deepest colour DAG fits the lane-private expansion stack
  const uint4* lut;         // [bytes of the k-mer word][256]: {rix fwd, q fwd, rix rc, q rc} parts (match.cu lut_pext)
  const uint32_t* cbeg;     // [nsubsets + 1] flattened colours: the leaves colour id se expands to (the walk of
  const uint32_t* cleaf;    //   ref src/query.cpp:369-387 done once at load) are cleaf[cbeg[se] .. cbeg[se+1]), as leaf ranks
  int32_t res_numer[kMaxResidues];
  uint32_t res_base[kMaxResidues];  // first row of the partial library holding the residue (several partial libraries in one directory)
};

// Per-slot buffers for one batch.
struct MatchArgs {
  const char* bases;          // concatenated ASCII reads
  const uint64_t* offsets;    // n_reads + 1; with off_pairs 2 * n_reads: read i spans [offsets[2i], offsets[2i + 1]) (segments of long reads overlap)
  uint64_t n_bases;           // bytes readable at `bases`
  uint32_t n_reads, th;
  uint32_t off_pairs;
  uint32_t keep_all;          // 1: emit every (strand, leaf) pair with a hit (parity tap 2); 0: only those passing the hdist_filt gate
  // per-read outputs
  uint32_t* onmers;           // [n]
  uint32_t* wn;               // [2n]
  uint32_t* hdfilt;           // [2n]
  uint32_t* rec_begin;        // [n]
  uint32_t* rec_count;        // [n]
  // record outputs (SoA)
  uint32_t* rec_read;         // [cap]
  uint32_t* rec_slot;         // [cap]  strand<<31 | leaf_se
  uint32_t* rec_hist;         // [cap * (th+1)]
  uint32_t rec_cap;
  uint32_t* counters;         // [0] records reserved, [1] next read to claim, [2] error flags, [3] placements, [4] solve work items
  // per-warp scratch in HBM (sized by the host from the resident warp count)
  uint32_t* acc;              // [warps][2*nleaves*(th+1)] Hamming histograms being accumulated
  uint32_t* bitmap;           // [warps][ceil(2*nleaves/32)] touched (strand, leaf) slots
  uint32_t* marker;           // [warps][nleaves] per-lookup min-hd markers: tag << 5 | hd (0xffffffff at rest)
  uint32_t* stack;            // [warps][stack_cap] colour expansion stack
  uint32_t* tagctr;           // [warps] marker tag counters (count down; persist across launches)
  uint32_t stack_cap;
  unsigned long long* stats;  // [0] algorithmic bytes, [1] lookups, [2] entries scanned
  // parity tap (stage 1)
  uint4* tap;
  unsigned long long* tap_count;
  unsigned long long tap_cap;
};

__device__ __forceinline__ void read_span(const MatchArgs& a, uint32_t read, uint64_t& off, uint64_t& len)
{
  if (a.off_pairs) { off = a.offsets[2ull * read]; len = a.offsets[2ull * read + 1] - off; }
  else { off = a.offsets[read]; len = a.offsets[read + 1] - off; }
}

// Long reads (contigs, long-read sequencing) are cut into segments of seg_windows k-mer windows, overlapping by k - 1 bases so
// that every window belongs to exactly one segment; the segments go through the match step as reads of their own and
// segment_combine_kernel (solve.cu) adds their per-(strand, reference) histograms up again -- a lookup is counted once, at its
// own smallest distance, so the sums are the histograms of the whole read -- before anything is gated or solved.
struct SegArgs {
  uint32_t n_reads;
  const uint32_t* vbegin;       // [n_reads + 1] first segment of every read
  const uint32_t *v_onmers, *v_wn, *v_hdfilt, *v_rec_begin, *v_rec_count, *v_rec_slot, *v_rec_hist, *v_counters; // what the match step wrote per segment
  uint32_t *onmers, *wn, *hdfilt, *rec_begin, *rec_count, *rec_read, *rec_slot, *rec_hist, *counters;            // the read's own
  uint32_t rec_cap, th, keep_all, nleaves;
  const uint32_t* leaf_rank;    // se -> rank among the leaves
  const uint32_t* leaf_se;      // rank -> se
  uint32_t* scratch;            // [warps][2 * nleaves * (th + 1)], zero between reads
  uint32_t* claim;              // next read to claim
};

// Bucket-sorted pipeline (sorted.cu): per-slot buffers between its kernels.
struct SortArgs {
  uint32_t nrows;
  uint32_t* row_count;    // [nrows] lookups per bucket row of this batch
  uint32_t* row_begin;    // [nrows + 1] exclusive prefix of row_count; [nrows] = lookups of the batch
  uint32_t* row_cursor;   // [nrows] fill cursors (start as row_begin)
  uint4* tuples;          // [cap_lookups] {q, read, local lookup index | strand << 31, 0}, grouped by row
  uint32_t cap_lookups;
  uint4* hits_tmp;        // [cap_hits] {read, strand << 31 | local lookup index << 5 | hd, colour id, 0} in join order
  uint4* hits;            // [cap_hits] {first leaf of the colour in cleaf, number of leaves, same meta word, 0}, grouped by read
  uint32_t cap_hits;
  uint32_t* hit_count;    // [n_reads] hit entries per read
  uint32_t* hit_begin;    // [n_reads + 1] exclusive prefix; [n_reads] = hit entries of the batch
  uint32_t* hit_cursor;   // [n_reads]
  uint32_t* partials;     // scan scratch
  uint32_t* sc;           // [0] hit entries appended, [1] next row to claim, [2] next read to claim (lookup pass 1), [3] (pass 2), [4] (resolve),
                          // [5] most leaf hits of one read that did not fit keys_g (the host grows it to that), [6] next bin to claim (bin sort)
  uint64_t* keys_g;       // [warps][cap_keys_g] per-warp sort scratch in HBM for reads whose leaf hits exceed shared memory
  uint32_t cap_keys_g;
  // two-level sort of the lookups (lookup_partition_kernel -> bin_sort_kernel): coarse bins of 2^bin_shift rows, each with room
  // for bin_cap tuples {q, read, local lookup index | strand << 31, row}; nbins = 0 switches the path off
  uint4* binned;          // [nbins * bin_cap]
  uint32_t* bin_cursor;   // [nbins] tuples appended to every bin
  uint32_t nbins, bin_cap, bin_shift;
  uint32_t res_ctas;      // CTAs of the resolve kernel (keys_g holds res_ctas * warps per CTA regions); 0 = the default grid
  uint32_t extra_rank_bits; // test knob (KREPP_SORT_WIDE): widens the leaf field of the sort keys so that the 64-bit key path runs
};

constexpr uint32_t kErrRecOverflow = 1u, kErrStackOverflow = 2u, kErrPlaceOverflow = 4u;
constexpr uint32_t kErrLookupOverflow = 8u, kErrHitOverflow = 16u, kErrSortFallback = 32u; // sorted pipeline: grow tuples / grow hits / redo the batch with the fused kernel
constexpr uint32_t kErrNodeOverflow = 128u; // placement: the batch touches more tree nodes than PlaceArgs::node_cap
constexpr uint32_t kErrKeysOverflow = 256u; // sorted pipeline: a read has more leaf hits than a warp's sort scratch (SortArgs::keys_g): grown by the host
constexpr uint32_t kErrHitWrap = 512u;      // sorted pipeline: 2^32 or more hit entries in one batch
constexpr uint32_t kErrBinOverflow = 1024u; // sorted pipeline: a coarse bin of the two-level lookup sort is full (skewed rows): the batch re-runs with the two-pass counting sort
constexpr uint32_t kErrShardData = 64u; // mode B: a hit entry names a read outside the batch (the caller mixed up its exchange buffers)
constexpr uint32_t kErrRedo = kErrRecOverflow | kErrStackOverflow | kErrLookupOverflow | kErrHitOverflow | kErrSortFallback | kErrShardData | kErrKeysOverflow | kErrHitWrap | kErrBinOverflow; // records are incomplete: later kernels skip, the host re-runs the batch

struct SolveArgs {
  uint32_t n_reads, th, k, h;
  uint32_t n_records;                // filled from counters[0] on the device when 0xffffffff
  uint32_t* counters;                // [0] records, [2] error flags, [4] length of `work`
  uint32_t* work;                    // [n_records] indices of the records that pass the hdist_filt gate (unordered)
  const uint32_t* onmers; const uint32_t* hdfilt;
  const uint32_t* rec_begin; const uint32_t* rec_count;
  const uint32_t* rec_read; const uint32_t* rec_slot; const uint32_t* rec_hist;
  const double* rho;                 // by se
  // outputs
  double* rec_d; double* rec_v; double* rec_chisq; uint32_t* rec_flags; uint32_t* rec_match; uint32_t* rec_hdmin;
  int32_t* closest;                  // [n_reads] record index or -1
  uint32_t* nsel;                    // [n_reads] selected references (entries of node_to_minfo, ref src/query.cpp:114,127-137)
  int want_chisq;
  // Identical problems are solved once per batch: the objective of a record is a function of (histogram, onmers - matches,
  // rho of the leaf) alone, and on a large index most records are weak matches that share those (about nine in ten of the
  // 1,000-genome workload's).  The gate enters each record in a hash table keyed by the exact tuple; the first to arrive
  // is solved, the others copy its result, which is bit-identical to solving them.
  unsigned long long* memo_key;      // [memo_mask + 1] packed tuple, 0 = empty (zeroed per batch)
  uint32_t* memo_owner;              // [memo_mask + 1] record that is solved for the slot
  uint32_t memo_mask;                // slots - 1 (a power of two), 0 = no table
  uint32_t memo_bits;                // bits per histogram bin in the key: (th + 1) * memo_bits + 8 (onmers) + 21 (leaf se) <= 64
  uint32_t* rec_alias;               // [n_records] slot whose owner's result the record takes, 0xffffffff = solved itself
};

// K5 (placement) arguments: everything K4 produced plus the flattened phytree.
struct PlaceArgs {
  SolveArgs s;
  const uint64_t* offsets;     // read offsets (enmers = len - k + 1, ref src/query.cpp:345-349)
  uint32_t tau; int no_filter; double chisq_value;
  // flattened tree (by se)
  const uint32_t* parent; const uint32_t* nchildren; const uint32_t* eff; const uint32_t* subtree; const uint32_t* depth; const uint32_t* logw; const double* blen; const uint32_t* leaf_rank;
  uint32_t nnodes, nleaves;
  // per-warp scratch of the collect kernel
  uint32_t* node_bitmap;       // [warps][ceil((nnodes+1)/32)]
  uint32_t* node_list;         // [warps][nnodes]
  uint32_t* node_order;        // [warps][nnodes] visiting order of the list (heaviest nodes first)
  uint32_t* sel;               // [warps][3 * nleaves] the read's selected references by ascending se: record, se, start of its chain
  double* chain;               // [warps][chain_cap] per selected leaf, level by level towards the root: prod 1/eff_nchildren
  uint32_t chain_cap;
  // tree nodes touched by the batch's reads (every selected leaf and all its ancestors), read by read, ascending se
  uint32_t node_cap;
  uint32_t* pn_read; uint32_t* pn_se; uint32_t* pn_flags;  // [node_cap]; flags: kPnSolve | kPnEligible | kPnCandidate
  double* pn_mc;               // [node_cap * (th+1)] weighted histogram of an internal node (Minfo::add, ref src/query.hpp:139-152)
  double* pn_uc; double* pn_rho; double* pn_d; double* pn_v; double* pn_chisq;  // [node_cap]
  uint32_t* pn_work;           // [node_cap] entries whose likelihood has to be maximised
  uint32_t* pn_begin; uint32_t* pn_count;  // [n_reads]
  // outputs
  void* placements;            // krepp_placement_t[place_cap]
  uint32_t place_cap;
  uint32_t* counters;          // [3] placements reserved, [2] error flags, [5] node entries reserved, [6] length of pn_work
  uint32_t* place_begin; uint32_t* place_count;  // [n_reads]
};
constexpr uint32_t kPnSolve = 1u, kPnEligible = 2u, kPnCandidate = 4u;

} // namespace krepp

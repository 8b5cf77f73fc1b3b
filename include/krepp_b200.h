/* include/krepp_b200.h -- C ABI of the B200-native krepp query path (`krepp dist` / `krepp place`).
 *
 * This is the drop-in boundary.  The reference (bo1929/krepp v0.8.3) has no plugin/FFI layer; its seam for this path
 * is the C++ class IBatch (src/query.hpp:46-97) sitting on Index (src/index.hpp:11-42).  Each entry point below names
 * the reference interface it replaces (file:line relative to the reference tree).  Plain pointers and sizes only; no
 * C++ or torch types cross this boundary; no exceptions cross it (status codes + krepp_last_error()).
 *
 * Threading: an index handle is immutable after open and may be shared; a batch handle ("slot") owns one CUDA stream
 * and its buffers and must be driven by one host thread at a time.  Several slots on one index pipeline H2D / kernels
 * / D2H.  There is NO CPU fallback: every call that needs the GPU fails with KREPP_ERR_CUDA when none is usable.
 */
#ifndef KREPP_B200_H
#define KREPP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KREPP_ABI_VERSION 2
#define KREPP_MAX_TH 16 /* k-h <= 16 positions survive in the 32-bit residual encoding (src/lshf.cpp:39-54) */

enum {
  KREPP_OK = 0,
  KREPP_ERR_ARG = 1,      /* invalid argument / configuration (src/krepp.hpp:192-204 validate_configuration_*) */
  KREPP_ERR_IO = 2,       /* index directory or file unreadable / malformed (src/index.cpp:51-158 error_exit paths) */
  KREPP_ERR_CUDA = 3,     /* no usable device, launch or copy failure */
  KREPP_ERR_CAPACITY = 4, /* a batch exceeded the slot's declared capacity */
  KREPP_ERR_UNSUPPORTED = 5
};

typedef struct krepp_index krepp_index_t;
typedef struct krepp_batch krepp_batch_t;

/* -------------------------------------------------------------------------------------------------- index */

typedef struct {
  uint32_t k, w, h, m, r, frac;    /* metadata-* (src/krepp.cpp:18-29) */
  uint32_t nrows;                  /* rows of the flat table = entries of inc-* (src/table.cpp:71-73) */
  uint64_t nkmers;                 /* entries of cmer-* (src/table.cpp:67-69) */
  uint32_t nnodes;                 /* tree nodes; se runs 1..nnodes, 0 is the null sentinel (src/phytree.hpp:53) */
  uint32_t nleaves;
  uint32_t nsubsets;               /* colour ids in crecord-* (src/record.cpp:203-211) */
  uint32_t root_se;
  uint64_t mask_hash_bp;           /* LSHF::mask_hash_bp (src/lshf.cpp:47-50) */
  uint64_t mask_drop_lr;           /* LSHF::mask_drop_lr (src/lshf.cpp:39-46) */
  uint64_t device_bytes;           /* HBM held by the index image */
  double mean_bucket, size_biased_bucket; /* occupancy statistics used to pick the scan group width */
} krepp_index_info_t;

/* Replaces TargetIndex::load_index (src/krepp.cpp:66-108) + Index::load_partial_tree/_index + make_rho_partial
 * (src/index.cpp:29-158,188-201) + FlatHT::load (src/table.cpp:65-75) + CRecord::load (src/record.cpp:203-211) +
 * Tree::load (src/phytree.cpp:394-404): reads the on-disk index written by `krepp index` unchanged and uploads the
 * flat image to `device`.  Round-1 scope: one partial suffix per directory, with a backbone tree file. */
#define KREPP_DEVICE_NONE (-1) /* parse + validate only (metadata, tree, names); batches cannot be created on it */
int krepp_index_open(const char* index_dir, int device, krepp_index_t** out);
/* The same with `place -t NWK` (TargetIndex::ensure_backbone src/krepp.cpp:48-64, Tree::map_to_qtree / compute_eff_nchildren
 * src/phytree.cpp:421-473): the Newick tree of nwk_path replaces the index's own backbone for everything after the colour
 * expansion.  References are matched to its leaves by name; references it does not have are dropped (null nodes,
 * src/query.cpp:373-376); a node weighs its children by how many of them have an indexed reference below and is a placement
 * candidate only when all of them do (src/query.cpp:250-271).  Node numbers in records, placements and the jplace tree are
 * the query tree's.  nwk_path NULL = the index's own tree; shard / nshards as for krepp_index_open_shard. */
int krepp_index_open_tree(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* nwk_path, krepp_index_t** out);
/* The LSH geometry of a library still to be built (BaseLSH set_nrows / set_lshf src/krepp.cpp:3-16, LSHF::get_random_positions
 * src/lshf.cpp:125-147, validate_configuration src/krepp.hpp:59-85): k, w, h, m, r, frac and the h hash positions the reference
 * draws from its global std::mt19937 (default-constructed; seed >= 0 is `--seed`, a negative seed means the option was not
 * given).  The handle serves the index-side calls (krepp_extract_mers, krepp_sketch_write) and krepp_index_info; it has no table
 * and cannot be queried. */
int krepp_geometry_open(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, int64_t seed, int device, krepp_index_t** out);
/* The same with the h hash positions given by the caller (any order; the reference keeps them descending and the other k - h
 * positions ascending, src/lshf.cpp:125-147): for a binding that sits beside the reference's own LSHF (LSHF::get_ppos,
 * src/lshf.hpp:17), and for adding a partial library to a directory whose positions are already fixed. */
int krepp_geometry_open_positions(uint32_t k, uint32_t w, uint32_t h, uint32_t m, uint32_t r, int frac, const uint8_t* ppos, int device, krepp_index_t** out);
/* `krepp seek -i SKETCH` (TargetSketch::load_sketch src/krepp.cpp:31-35, Sketch::load_full_sketch / make_rho_partial
 * src/sketch.cpp:3-32): the sketch file of ONE genome written by `krepp sketch` -- a table of 4-byte residual encodings without
 * colours, its LSH geometry and the genome's rho.  It is held as an index whose tree is a single leaf named after the file, so
 * krepp_index_info and the batch calls work on it unchanged; a batch on it gives at most one record per strand, and
 * KREPP_OUT_SEEK the distance SBatch::seek_sequences prints (src/seek.cpp:22-53).  `place` is refused. */
int krepp_sketch_open(const char* sketch_path, int device, krepp_index_t** out);
/* krepp_index_open_tree with `place -l FILE` in place of `-t` (TargetIndex::read_lineages src/krepp.cpp:37-46, Tree::parse_lineages src/phytree.cpp:320-370):
 * the tree is built from a Greengenes/GTDB style lineage file ("NAME<tab>d__A; p__B; ..."), the references hang below the
 * last taxon of their line, there are no branch lengths (pendant and distal lengths print as 0) and nodes with one child are
 * kept but are no placement candidates.  Works on an index without a backbone tree too (the reference skips
 * ensure_backbone with -l). */
int krepp_index_open_lineages(const char* index_dir, int device, uint32_t shard, uint32_t nshards, const char* lineage_path, krepp_index_t** out);
void krepp_index_close(krepp_index_t* ix);
int krepp_index_info(const krepp_index_t* ix, krepp_index_info_t* out);
/* Wrapping 64-bit sums over the arrays of the host image this handle holds, for checking a loader (or a shard's slice) without
 * a device: out[0] = sum of the cmer words (enc | se << 32), out[1] = sum of the bucket ends (relative to the shard's first
 * entry), out[2] = sum over colour ids c of c * (number of leaves c expands to), out[3] = sum over colour ids c and their
 * leaves l (leaf ranks, ascending) of (c + 1) * (l + 1). */
int krepp_index_host_checksums(const krepp_index_t* ix, uint64_t out[4]);
/* Node::get_name(return_na) (src/phytree.hpp:133-144): label, else to_string(se-1) or "NA".  Pointer valid until the
 * next call from the same thread. */
const char* krepp_index_node_name(const krepp_index_t* ix, uint32_t se, int return_na);
/* Flattened phytree arrays, each of length nnodes+1 and indexed by se (src/phytree.hpp:95-131,156): parent se (0 for
 * the root), number of children, leaf flag, branch length (NaN when absent).  Any output pointer may be NULL. */
int krepp_index_tree(const krepp_index_t* ix, uint32_t* parent, uint32_t* nchildren, uint8_t* is_leaf, double* blen);
/* Tree::stream_nwk_jplace (src/phytree.cpp:47-64): edge-numbered Newick, std::fixed precision 5.  Returns the number
 * of bytes needed (excluding NUL); writes at most cap bytes including the NUL. */
size_t krepp_index_jplace_tree(const krepp_index_t* ix, char* buf, size_t cap);

/* -------------------------------------------------------------------------------------------------- parameters */

/* The IBatch constructor arguments (src/query.hpp:49-57, src/query.cpp:8-38) plus the CLI's `tabular` switch. */
typedef struct {
  uint32_t hdist_th;  /* --hdist-th [4] */
  double chisq;       /* --chisq [2.706] */
  double dist_max;    /* --dist-max [NaN = unset] */
  uint32_t tau;       /* --tau [2] */
  int32_t no_filter;  /* dist default 1 (src/krepp.cpp:637-642), place default 0 (:614-617) */
  int32_t multi;      /* [1] */
  int32_t summarize;  /* [0] */
  int32_t place;      /* 0: IBatch::estimate_distances (src/query.cpp:141-156); 1: place_sequences (:198-216) */
} krepp_params_t;

void krepp_params_default(krepp_params_t* p, int place);

/* -------------------------------------------------------------------------------------------------- results */

/* One (read, strand, leaf) Hamming histogram = one Minfo of IMers::leaf_to_minfo (src/query.hpp:43,213-226). */
typedef struct {
  uint32_t read;        /* index of the read inside the batch */
  uint32_t leaf_se;     /* Node::se of the reference genome */
  uint32_t strand;      /* 0 forward, 1 reverse complement */
  uint32_t match_count; /* Minfo::match_count */
  uint32_t hdist_min;   /* Minfo::hdist_min */
  uint32_t flags;       /* KREPP_REC_* */
  double rho;           /* CRecord::se_to_rho after make_rho_partial */
  double d_llh, v_llh;  /* Minfo::optimize_likelihood (src/query.cpp:426-433); DBL_MAX / NaN when not solved */
  double chisq;         /* Minfo::likelihood_ratio vs the closest (src/query.cpp:420-424); NaN when not computed */
} krepp_record_t;
#define KREPP_REC_SOLVED 1u   /* passed the hdist_filt gate of summarize_matches (src/query.cpp:106,119) */
#define KREPP_REC_SELECTED 2u /* is the entry of IBatch::node_to_minfo for its leaf (src/query.cpp:114,127-137) */
#define KREPP_REC_CLOSEST 4u  /* is IBatch::mi_closest (src/query.cpp:110-113,123-126) */

/* One candidate placement = one PP_JPLACE_FIELDS row (src/query.hpp:202-204, src/query.cpp:284-310). */
typedef struct {
  uint32_t read, se;                               /* edge_num = se-1 (src/phytree.hpp:156) */
  double pendant, distal, loglik, lwr, d_llh, chisq;
} krepp_placement_t;

typedef struct {
  uint32_t onmers;        /* IBatch::onmers: valid k-mer windows (src/query.cpp:66) */
  uint32_t wn[2];         /* IBatch::wnmers_or / wnmers_rc: eligible lookups per strand (src/query.cpp:85,90) */
  uint32_t hdist_filt[2]; /* IMers::hdist_filt per strand before the 2x+1 of summarize_matches (0xffffffff: none) */
  uint32_t rec_begin, rec_count;     /* this read's records[]: forward leaves by ascending se, then reverse */
  uint32_t place_begin, place_count; /* this read's placements[] by ascending se */
  int32_t closest;                   /* index into records[] of mi_closest, -1 when node_to_minfo is empty */
  uint32_t n_selected;               /* entries of IBatch::node_to_minfo: references that keep a record after summarize_matches
                                        (src/query.cpp:114,127-137); report_placement's single-reference shortcut tests it (:231) */
} krepp_read_summary_t;

/* The 16-byte form of a record, for front ends that only print distances (`krepp dist`): what report_distances
 * (src/query.cpp:158-196) reads of a Minfo.  Same order and indices as records[]. */
typedef struct {
  uint32_t read;
  uint32_t ref;   /* leaf_se | strand << 27 | flags (KREPP_REC_*) << 28 | (chisq < params.chisq) << 31 */
  double d_llh;
} krepp_brief_t;
#define KREPP_BRIEF_SE(ref) ((ref) & 0x07FFFFFFu)
#define KREPP_BRIEF_STRAND(ref) (((ref) >> 27) & 1u)
#define KREPP_BRIEF_FLAGS(ref) (((ref) >> 28) & 7u)
#define KREPP_BRIEF_CHISQ_OK(ref) ((ref) >> 31)

/* The rows `krepp dist` prints, chosen, ordered and rounded on the device (KREPP_OUT_DIST): everything report_distances
 * (src/query.cpp:158-196) decides -- the NA rule, --no-multi, --filter / --chisq, --dist-max, --summarize's kept set -- is applied
 * by a kernel, and what leaves the GPU per read is one word of dist_begin plus its rows, references by ascending se, the distance
 * as the integer its five printed decimals show (round-to-nearest of the exact binary value, ties to even, as the reference's
 * std::fixed << setprecision(5)).  Read i owns rows [KREPP_DIST_BEGIN(dist_begin[i]), KREPP_DIST_BEGIN(dist_begin[i + 1]));
 * KREPP_DIST_NA(dist_begin[i]) says the read prints "NA\tNaN" instead (no match, or closest beyond --dist-max).  A row is 4 bytes
 * (leaf rank << 16 | distance units; indexes of at most 65,536 references) or 8 bytes (leaf se | (uint64) units << 32). */
#define KREPP_DIST_BEGIN(w) ((w) & 0x7FFFFFFFu)
#define KREPP_DIST_NA(w) ((w) >> 31)

typedef struct {
  uint32_t n_reads;
  uint32_t hist_stride;               /* hdist_th + 1 */
  uint64_t n_records, n_placements;
  const krepp_read_summary_t* reads;  /* [n_reads]; NULL when KREPP_OUT_SUMMARIES was not asked for */
  const krepp_record_t* records;      /* [n_records] */
  const uint32_t* hist;               /* [n_records * hist_stride]: Minfo::hdisthist_v */
  const krepp_placement_t* placements;/* [n_placements] */
  float gpu_ms;                       /* device time of all kernels of this batch, copies excluded (CUDA events on the slot's stream) */
  float match_ms;                     /* device time of the match kernel alone (same stream, CUDA events) */
  uint32_t gpu_launches;              /* kernels launched for this batch */
  const krepp_brief_t* brief;         /* [n_records] when KREPP_OUT_BRIEF was asked for, else NULL */
  const uint32_t* dist_begin;         /* [n_reads + 1] when KREPP_OUT_DIST was asked for, else NULL */
  const void* dist_rows;              /* [n_dist_rows] rows of dist_row_bytes bytes */
  uint64_t n_dist_rows;
  uint32_t dist_row_bytes;            /* 4 or 8 */
  const double* seek_dist;            /* [n_reads] when KREPP_OUT_SEEK was asked for, else NULL: SBatch::seek_sequences' distance (src/seek.cpp:22-53),
                                         the smaller of the two strands' estimates; NaN = no k-mer of the read matched the sketch */
} krepp_results_t;

/* -------------------------------------------------------------------------------------------------- batches */

/* Replaces the IBatch constructor (src/query.cpp:8-38).  max_reads / max_bases size the slot's pinned and device
 * buffers once; a larger submit returns KREPP_ERR_CAPACITY. */
int krepp_batch_create(krepp_index_t* ix, const krepp_params_t* p, uint32_t max_reads, uint64_t max_bases,
                       krepp_batch_t** out);
void krepp_batch_destroy(krepp_batch_t* b);

/* Pre-sizes the slot's result buffers (device arrays and the page-locked host arrays of the rows krepp_batch_set_output asked
 * for): room for n_records records, n_hits hit entries of the bucket-sorted chain, n_nodes tree nodes and n_placements placement
 * rows per batch (0 = leave as is).  Optional: the buffers start small and grow to the demand of the first batches, which costs
 * those batches a second pass and a few allocations; a front end that knows its index (about 19 records and 56 hit entries per
 * 150 bp read on a 1,000-genome index) reserves once, before the first submit. */
int krepp_batch_reserve(krepp_batch_t* b, uint64_t n_records, uint64_t n_hits, uint64_t n_nodes, uint64_t n_placements);

/* Replaces IBatch::estimate_distances / place_sequences up to (not including) text formatting (src/query.cpp:141-156,
 * 198-216): `bases` holds the reads' ASCII characters back to back, read i is bases[offsets[i] .. offsets[i+1]).
 * HOST buffers; the call enqueues H2D, all kernels and D2H on the slot's stream and returns without waiting.  Pageable
 * `bases` are first staged through the slot's pinned memory and may be reused as soon as the call returns; page-locked
 * `bases` (cudaMallocHost / cudaHostRegister memory) are copied to the device from where they lie and must stay
 * untouched until krepp_batch_wait returns.  `offsets` may always be reused at once. */
int krepp_batch_submit(krepp_batch_t* b, const char* bases, const uint64_t* offsets, uint32_t n_reads);

/* The slot's own pinned input buffers (max_bases + 64 bytes, max_reads + 1 offsets).  A producer that parses reads
 * straight into them and then passes the same pointers to krepp_batch_submit skips the staging copy. */
int krepp_batch_host_buffers(krepp_batch_t* b, char** bases, uint64_t** offsets);

/* Same work with inputs already resident in HBM (device pointers on the index's device); nothing is copied in.  Used
 * to measure kernel-side throughput and by callers that produce reads on the device. */
int krepp_batch_submit_device(krepp_batch_t* b, const char* d_bases, const uint64_t* d_offsets, uint32_t n_reads,
                              uint64_t n_bases);

/* Waits for the slot's stream and exposes the results (library-owned pinned host memory, valid until the next submit
 * on this slot).  Replaces reading IBatch::node_to_minfo / get_summary() (src/query.hpp:64,95). */
int krepp_batch_wait(krepp_batch_t* b, krepp_results_t* out);

/* Which row arrays krepp_batch_wait copies to the host (default: all).  The text formatters below read `records` and
 * `placements` but never `hist`; a front end that only prints distances can leave the histograms (4 * (hdist_th + 1)
 * bytes per record) in HBM and save a quarter of the device-to-host traffic.  Arrays that are not copied come back NULL. */
#define KREPP_OUT_RECORDS 1u
#define KREPP_OUT_HIST 2u
#define KREPP_OUT_PLACEMENTS 4u
#define KREPP_OUT_BRIEF 8u /* krepp_brief_t rows (16 bytes per record instead of 56 + histogram); not part of KREPP_OUT_ALL.
                              krepp_format_dist reads them when `records` is NULL. */
#define KREPP_OUT_DIST 16u /* the printed rows of `krepp dist` (see krepp_results_t): 4 bytes per read + 4 or 8 per printed row.
                              Not part of KREPP_OUT_ALL; krepp_format_dist prefers them when present. */
#define KREPP_OUT_SUMMARIES 32u /* the 44-byte krepp_read_summary_t rows; a front end that asks for KREPP_OUT_DIST alone does without */
#define KREPP_OUT_SEEK 64u /* sketch handles (krepp_sketch_open) only: one double per read, what `krepp seek` prints (seek_dist) */
#define KREPP_OUT_ALL 39u
/* Must be called while no batch is pending on the slot (before krepp_batch_submit, or after the krepp_batch_wait that follows
 * it): the kernels that assemble the rows run as part of the submit.  KREPP_ERR_ARG otherwise. */
int krepp_batch_set_output(krepp_batch_t* b, uint32_t rows);

/* The same wait (including the grow-and-rerun of a batch whose result buffers were too small) without copying the record,
 * histogram and placement rows to the host: `reads` (44 bytes per read) and the counts are valid, the three row pointers
 * are NULL and the rows stay in HBM.  For callers that only need the per-read summaries, and for timing the kernels
 * without the PCIe transfer of the rows.  (The summaries are copied here even if KREPP_OUT_SUMMARIES was not asked for.) */
int krepp_batch_wait_device(krepp_batch_t* b, krepp_results_t* out);

/* Parity taps (SURVEY.md section 8b "dump_stage").  stage 1: every eligible lookup of the last submitted batch as
 * 4 x u32 {read, strand<<31|pos, rix, enc32}, unordered (src/query.cpp:82-91 arguments of add_matching_mer).
 * Must be enabled before submit with krepp_batch_enable_tap.  Returns the number of items through *n.
 * stage 2 (a switch, capacity_items != 0 turns it on; nothing to read back): records[] also keeps the (strand, leaf)
 * pairs that fail the hdist_filt gate of summarize_matches (src/query.cpp:101-106,116-119) -- every Minfo the reference
 * holds BEFORE that gate, with KREPP_REC_SOLVED clear on the failing ones.  By default those pairs are dropped on the
 * device as soon as their read is finished, which is what the reference's node_to_minfo holds after the gate. */
int krepp_batch_enable_tap(krepp_batch_t* b, int stage, uint64_t capacity_items);
int krepp_batch_read_tap(krepp_batch_t* b, int stage, uint32_t* out, uint64_t cap_items, uint64_t* n);

/* Roofline accounting of the last waited batch: algorithmic bytes = sum over reads of
 * len + sum over eligible lookups (16 + 8*|bucket|) + 64 * records, computed on the device while matching. */
int krepp_batch_algorithmic_bytes(krepp_batch_t* b, uint64_t* bytes, uint64_t* lookups, uint64_t* entries_scanned);

/* Per-stage device times of the last waited batch (measurement only; no reference equivalent): CUDA events recorded on
 * the slot's stream after every kernel (or short kernel group) of the batch.  ms[i] / names[i] (static strings) for
 * i < min(*n, cap); *n = number of stages.  The stages of the match step (src/query.cpp:40-94,352-390) come first. */
int krepp_batch_stage_times(krepp_batch_t* b, uint32_t cap, float* ms, const char** names, uint32_t* n);

/* -------------------------------------------------------------------------------------------------- bucket-range shards
 * SURVEY.md section 8e, mode B: an index too large for one GPU is split by LSH bucket (row) range over the ranks of a
 * job.  The reference has no counterpart (it holds the whole table of src/table.hpp:121-143 in one address space); what
 * is replaced is still IBatch::search_mers + IMers::add_matching_mer (src/query.cpp:40-94,352-390), cut at the two
 * points where data changes owner.  Every entry that can match a lookup lives in the lookup's bucket, hence on exactly
 * one shard, so the per-(read, strand, leaf) histograms are complete after one round trip and identical to the
 * unsharded result.
 *
 *   home rank    krepp_shard_lookup   reads -> eligible lookups ("tuples", 16 bytes: {enc32 q, read, strand << 31 | lookup
 *                                     index, 0}) counting-sorted by row, so shard g's tuples are tuples[send_offsets[g] ..
 *                                     send_offsets[g+1]) and its row directory is row_begin[row_splits[g] .. row_splits[g+1]]
 *   (exchange)                        the caller moves both runs to rank g (an all-to-all-v; NCCL over NVLink in
 *                                     krepp_b200/dist.py) -- the library never touches another rank's memory
 *   owner rank   krepp_shard_join     per sender: its tuples against this shard's buckets -> hit entries (16 bytes: {read,
 *                                     strand << 31 | lookup index << 5 | hd, colour id, 0}), appended sender by sender
 *   (exchange)                        hit_offsets[src .. src+1) goes back to rank src
 *   home rank    krepp_shard_finish   hit entries of the batch from all owners -> colour expansion, min-hd histograms, gate,
 *                                     likelihood solve, merge: then krepp_batch_wait as for an unsharded batch
 *
 * All buffers passed here are DEVICE memory owned by the caller on the index's device.  The three calls of one batch run
 * on one slot in this order; lookup and join return when their kernels have finished (their outputs size the exchange),
 * finish only enqueues. */
#define KREPP_MAX_SHARDS 256

typedef struct {
  uint32_t shard, nshards;
  uint32_t row0, row1;              /* rows [row0, row1) of the table live on this shard */
  uint64_t first_entry, n_entries;  /* entries [first_entry, first_entry + n_entries) of cmer-* */
} krepp_shard_info_t;

/* krepp_index_open for one shard: parses every file but keeps (and uploads) only this shard's slice of cmer-* and inc-*;
 * colour record, tree and hash tables are replicated.  Shards are contiguous row ranges of (nearly) equal cmer bytes,
 * derived from inc-* alone, so every rank computes the same split.  nshards = 1 is krepp_index_open. */
int krepp_index_open_shard(const char* index_dir, int device, uint32_t shard, uint32_t nshards, krepp_index_t** out);
/* How many bucket-range shards an index needs so that every shard's device image (its slice of the table plus the replicated
 * colour record, colour lists, tree and hash tables) stays within `budget_bytes` per GPU: 1 = it can be replicated (mode A).
 * budget_bytes = 0 means "what is free on `device` now, less 25 % for the batch slots".  Reads everything but the table itself.
 * whole_bytes / shard_bytes (may be NULL): image size unsharded, and of the largest shard at *nshards.  KREPP_ERR_CAPACITY when
 * even KREPP_MAX_SHARDS shards do not fit. */
int krepp_index_plan_shards(const char* index_dir, int device, uint64_t budget_bytes, uint32_t* nshards, uint64_t* whole_bytes,
                            uint64_t* shard_bytes);
/* row_splits (may be NULL): first row of every shard, nshards + 1 values (at most cap are written). */
int krepp_index_shard_info(const krepp_index_t* ix, krepp_shard_info_t* out, uint32_t* row_splits, uint32_t cap);

/* d_tuples: room for cap_tuples tuples; d_row_begin: nrows + 1 words; send_offsets: nshards + 1 host words.  Returns
 * KREPP_ERR_CAPACITY when the batch has more lookups than cap_tuples (send_offsets[nshards] then holds the demand). */
int krepp_shard_lookup(krepp_batch_t* b, const char* d_bases, const uint64_t* d_offsets, uint32_t n_reads, uint64_t n_bases,
                       void* d_tuples, uint64_t cap_tuples, uint32_t* d_row_begin, uint64_t* send_offsets);
/* d_tuples[src] / d_row_begin[src]: sender src's tuples for this shard and its row_begin[row0 .. row1] (row1 - row0 + 1
 * words, positions still relative to the sender's list).  hit_offsets: n_sources + 1 host words.  KREPP_ERR_CAPACITY when
 * cap_hits is too small (hit_offsets[n_sources] then holds the demand; call again with a larger buffer). */
int krepp_shard_join(krepp_batch_t* b, uint32_t n_sources, const void* const* d_tuples, const uint32_t* const* d_row_begin,
                     void* d_hits, uint64_t cap_hits, uint64_t* hit_offsets);
/* d_hits must stay untouched until krepp_batch_wait has returned. */
int krepp_shard_finish(krepp_batch_t* b, const void* d_hits, uint64_t n_hits);

/* Device memory for the exchange buffers, for hosts that do not link the CUDA runtime themselves (the krepp_b200 executable:
 * all shards in one process, runs moved between GPUs by peer copies over NVLink).  krepp_device_copy is synchronous; a
 * device of KREPP_DEVICE_NONE means host memory. */
int krepp_device_alloc(int device, uint64_t bytes, void** out);
void krepp_device_free(int device, void* p);
int krepp_device_copy(int dst_device, void* dst, int src_device, const void* src, uint64_t bytes);

/* -------------------------------------------------------------------------------------------------- index side
 * What `krepp index` does to one reference genome before the colour unions: RSeq::extract_mers (src/rqseq.cpp:51-144; sdust
 * off, not canonical: every valid forward k-mer, xur64_hash minimizer of every window of w valid bases, the LSH residue filter,
 * row and 32-bit residual encoding, and the end-of-sequence emit of :112-116) followed by the per-row sort and unique of
 * DynHT::fill_table (src/table.cpp:110-117,157-166,248-260).  Geometry (k, w, h, m, r, frac, positions) is that of `ix`, which
 * must be open on a GPU.  bases / offsets: HOST memory, the genome's sequences back to back (sequences shorter than w are
 * skipped, src/rqseq.hpp:80-86).  keys: row << 32 | encoding, ascending -- one leaf table; *n_keys is always the number found;
 * KREPP_ERR_CAPACITY when cap is too small (cap = 0 just counts). */
int krepp_extract_mers(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, uint64_t* keys,
                       uint64_t cap, uint64_t* n_keys);

/* The subsampling rate of a genome as `krepp index` / `krepp sketch` estimate it (RSeq::extract_mers + compute_rho, src/rqseq.cpp:63-64,
 * 107-108,117,142-143, src/rqseq.hpp:79; hll::HyperLogLog src/hyperloglog.hpp:58-140, 12 bits): per sequence one HyperLogLog over
 * the hashes of all valid k-mers and one over the window minimizers (registers filled by the minimizer kernel), the estimates
 * summed over the sequences; rho = *n_minimizers_est / *n_kmers_est. */
int krepp_sequence_rho(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, double* n_kmers_est,
                       double* n_minimizers_est);
/* `krepp sketch` (SketchSingle::create_sketch / save_sketch src/krepp.cpp:110-128): the sequences' minimizers through
 * krepp_extract_mers, the table (SFlatHT::save src/table.cpp:35-41), the configuration (src/krepp.cpp:18-29) and rho, written
 * to out_path -- the file `krepp seek` / krepp_sketch_open read.  ix: a geometry handle (krepp_geometry_open) on a GPU. */
int krepp_sketch_write(const krepp_index_t* ix, const char* bases, const uint64_t* offsets, uint32_t n_seqs, const char* out_path,
                       uint64_t* n_kmers, double* rho);

/* `krepp index` (IndexMultiple::build_index / build_for_subtree / save_index, src/krepp.cpp:164-309): a whole library from
 * its reference genomes.  The reference builds a leaf table per genome and unions the tables up the guide tree, giving every
 * k-mer a colour -- the set of references that hold it -- as a DAG of pairs over the tree's nodes (DynHT::union_row
 * src/table.cpp:214-234, Record::add_subset src/record.cpp:82-113, CRecord src/record.cpp:156-175).  Here:
 *   krepp_builder_add_genome   GPU: the genome's leaf table (krepp_extract_mers) stays in HBM; its rho (krepp_sequence_rho).
 *   krepp_builder_union        GPU: every (k-mer, reference) pair of the library in one array, one stable radix sort by
 *                              (row, encoding); a run of equal keys is one k-mer with its references in leaf order; runs are
 *                              hashed and grouped into the DISTINCT reference sets (verified element by element, never trusted
 *                              to the hash).  Leaves the host with: per k-mer its key and the id of its set; per set its leaves.
 *   krepp_builder_write        host: each distinct set is decomposed along the guide tree (a set that is a whole subtree is the
 *                              tree node itself, as in the reference; otherwise pairs, shared between sets), colour ids are
 *                              numbered, and the seven files of src/krepp.cpp:206-246 are written in the reference's format.
 * The result is the reference's library up to the numbering of colours above the tree nodes (which the reference itself does
 * not fix from run to run): metadata-*, inc-*, the encoding column of cmer-*, reflist-*, tree-* and
 * rho are the reference's bytes, and every k-mer's colour expands to the same references.
 * geom: a geometry handle (krepp_geometry_open); on a GPU for add_genome / union, any for set_union / write.
 * nwk_text: the guide tree (-t), written to tree-* verbatim; NULL = no tree: the balanced tree the reference generates over
 * the names (Node::generate_tree src/phytree.cpp:217-253) and no tree-* file.  names: the reference ids of input_map.tsv in
 * file order (reflist-*).  A name the tree lacks is never visited and a leaf without a genome stays empty, as in the reference. */
typedef struct krepp_builder krepp_builder_t;
int krepp_builder_create(const krepp_index_t* geom, const char* nwk_text, const char* const* names, uint32_t n_names, krepp_builder_t** out);
void krepp_builder_destroy(krepp_builder_t* b);
/* 1 when `name` is a leaf of the build tree (its genome will be used), 0 when not (the reference skips it), < 0 on error */
int krepp_builder_has_leaf(const krepp_builder_t* b, const char* name);
int krepp_builder_add_genome(krepp_builder_t* b, const char* name, const char* bases, const uint64_t* offsets, uint32_t n_seqs,
                             uint64_t* n_keys, double* rho);
int krepp_builder_union(krepp_builder_t* b, uint64_t* n_kmers, uint64_t* n_sets);
/* The union handed in by the caller instead (host arrays, copied): n_kmers keys (row << 32 | encoding, ascending, distinct),
 * the set id of each, and the sets as CSR over leaf RANKS (leaves of the build tree in ascending post-order number, ascending
 * within a set); rho per leaf rank (0 where a leaf has no genome).  What krepp_builder_union leaves behind, for callers that
 * computed the union elsewhere and for the host-side tests. */
int krepp_builder_set_union(krepp_builder_t* b, uint64_t n_kmers, const uint64_t* keys, const uint32_t* set_of, uint64_t n_sets,
                            const uint64_t* set_begin, const uint32_t* set_leaves, const double* leaf_rho);
/* Leaf rank of a reference in the build tree (0xffffffff when the tree lacks it) and the number of leaves. */
uint32_t krepp_builder_leaf_rank(const krepp_builder_t* b, const char* name);
uint32_t krepp_builder_nleaves(const krepp_builder_t* b);
/* Writes the library into index_dir (created when missing) with the suffix -m{m}r{r}-{frac|no_frac} (src/krepp.cpp:586-589).
 * seed: what metadata-*.txt reports (src/krepp.cpp:192).  *n_subsets: colour ids incl. the null id (crecord's nsubsets). */
int krepp_builder_write(krepp_builder_t* b, const char* index_dir, uint32_t seed, uint64_t* n_kmers, uint32_t* n_subsets);

/* -------------------------------------------------------------------------------------------------- host I/O layer
 * The steps immediately either side of the GPU path.  Pure host code: usable
 * without a device (the index handle may have been opened with KREPP_DEVICE_NONE). */

typedef struct krepp_reader krepp_reader_t;

/* Replaces the QSeq constructor (src/rqseq.cpp:161-178): opens a FASTA/FASTQ file, plain or gzip (zlib gzopen). */
int krepp_reader_open(const char* path, krepp_reader_t** out);
void krepp_reader_close(krepp_reader_t* r);
/* Worker threads krepp_reader_next may use (default 1).  With more than one, batches of plain (not gzip) four-line FASTQ are
 * framed chunk-parallel from the mapped file; every record and the record order are those of the sequential reader (anything
 * the chunks cannot prove -- FASTA, wrapped lines, a truncated tail -- is left to it), only batch boundaries may differ. */
int krepp_reader_set_threads(krepp_reader_t* r, uint32_t threads);
/* Replaces QSeq::read_next_batch (src/rqseq.cpp:180-197) over kseq_read (src/kseq.h:177-216) with the same record
 * framing: a record starts at '>' or '@'; the name is the header up to the first whitespace; sequence characters are
 * all printable non-space bytes up to the next '>', '@' or '+'; after '+' the rest of that line is skipped and as many
 * quality characters as there were bases are consumed; a truncated quality string ends the input.  Reads are appended
 * back to back into `bases` (offsets[0] = 0 ... offsets[n]) and names, NUL-terminated, into `names`
 * (name_offsets[i] = start of name i).  The batch ends when max_reads, max_bases or max_name_bytes would be exceeded
 * (the record that did not fit opens the next batch) or at end of input (*eof = 1).  A single record larger than
 * max_bases returns KREPP_ERR_CAPACITY. */
int krepp_reader_next(krepp_reader_t* r, char* bases, uint64_t max_bases, uint64_t* offsets, uint32_t max_reads,
                      char* names, uint64_t max_name_bytes, uint64_t* name_offsets, uint32_t* n_reads, int* eof);

/* All formatters append to `buf` (capacity `cap`) and return the number of bytes the complete text needs; when that
 * exceeds cap nothing useful was written and the caller retries with a larger buffer.  Doubles are printed like the
 * reference's streams: std::fixed, precision 5 (src/query.cpp:152-153, src/krepp.cpp:351-352). */

/* header_dreport / header_preport / begin_jplace (src/krepp.cpp:311-319,396-408,426-432). */
size_t krepp_format_header(const krepp_index_t* ix, const krepp_params_t* p, int tabular, const char* invocation,
                           char* buf, size_t cap);
/* IBatch::report_distances for every read of a batch (src/query.cpp:158-196): reads in input order, references by
 * ascending se.  With p->summarize the rows are not written; the per-node weights are added to wcount[nnodes+1]
 * instead (src/query.cpp:160-171) and 0 is returned.  Reads res->records, or res->brief when records is NULL. */
size_t krepp_format_dist(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res,
                         const char* names, const uint64_t* name_offsets, double* wcount, char* buf, size_t cap);
/* IBatch::place_sequences / report_placement text for a batch (src/query.cpp:198-333): jplace "placements" entries
 * (PP_JPLACE_FIELDS, src/query.hpp:202-204) or --tabular rows (PP_TABULAR_FIELDS, :206); --no-multi picks the
 * candidate with the largest clade, then the smallest distance (src/query.cpp:311-330).  *has_previous carries the
 * "a placement was already written" state across batches (src/krepp.cpp:476-481).  With p->summarize weights go to
 * wcount (src/query.cpp:232,298,323). */
size_t krepp_format_place(const krepp_index_t* ix, const krepp_params_t* p, const krepp_results_t* res,
                          const char* names, const uint64_t* name_offsets, int tabular, int* has_previous, double* wcount,
                          char* buf, size_t cap);
/* SBatch::seek_sequences' rows for a batch on a sketch handle (src/seek.cpp:41-49): "<id>\t<distance>", "<id>\tNaN" for a read
 * without a matching k-mer; needs res->seek_dist (KREPP_OUT_SEEK).  `krepp seek` writes no header line (the reference builds one,
 * src/krepp.cpp:305-309, and never sends it to the output). */
size_t krepp_format_seek(const krepp_results_t* res, const char* names, const uint64_t* name_offsets, char* buf, size_t cap);
/* Tail of the output: the --summarize table (src/krepp.cpp:385-392,492-497) or end_jplace (src/krepp.cpp:410-424). */
size_t krepp_format_footer(const krepp_index_t* ix, const krepp_params_t* p, int tabular, const double* wcount,
                           uint64_t total_queries, const char* invocation, char* buf, size_t cap);

const char* krepp_last_error(void);
int krepp_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif

/* pgmm_b200.h -- C-ABI of libpgmm_b200.so, the B200-native replacement for pangraph's alignment hot path.
 *
 * Part 1 is the boundary the reference's FFI crate binds today: packages/minimap2-sys/minimap2.h (bindgen over
 * minimap2/minimap.h + mmpriv.h) -- the ten symbols and five struct layouts that packages/minimap2 (the safe Rust
 * wrapper) actually touches.  A build of pangraph that links this library instead of the vendored C sees the same
 * names, argument meaning, ownership rules and struct offsets (verified by tests/test_abi.py with offsetof()).
 * Part 2 adds the batched entry point a GPU needs, and the host half of the path (find_matches -> split -> filter).
 * Part 3 are stage-level entry points used by the parity tests and by bench.py.
 *
 * All entry points are plain C: pointers and sizes only.  CUDA failures abort() with a message, like the reference's
 * C, which has no error channel either.  There is no CPU fallback behind any of them.
 */
#ifndef PGMM_B200_H
#define PGMM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PGMM_API __attribute__((visibility("default")))
#else
#define PGMM_API
#endif

/* ===================== Part 1: the minimap2-sys boundary ===================== */

/* mapping flags used on pangraph's path -- reference: minimap2/minimap.h:10-47 */
#define MM_F_NO_DIAG    (0x001LL)
#define MM_F_NO_DUAL    (0x002LL)
#define MM_F_CIGAR      (0x004LL)
#define MM_F_OUT_SAM    (0x008LL)
#define MM_F_NO_QUAL    (0x010LL)
#define MM_F_OUT_CG     (0x020LL)
#define MM_F_OUT_CS     (0x040LL)
#define MM_F_SPLICE     (0x080LL)
#define MM_F_SPLICE_FOR (0x100LL)
#define MM_F_SPLICE_REV (0x200LL)
#define MM_F_NO_LJOIN   (0x400LL)
#define MM_F_OUT_CS_LONG (0x800LL)
#define MM_F_SR         (0x1000LL)
#define MM_F_FRAG_MODE  (0x2000LL)
#define MM_F_NO_PRINT_2ND (0x4000LL)
#define MM_F_2_IO_THREADS (0x8000LL)
#define MM_F_LONG_CIGAR (0x10000LL)
#define MM_F_INDEPEND_SEG (0x20000LL)
#define MM_F_SPLICE_FLANK (0x40000LL)
#define MM_F_SOFTCLIP   (0x80000LL)
#define MM_F_FOR_ONLY   (0x100000LL)
#define MM_F_REV_ONLY   (0x200000LL)
#define MM_F_HEAP_SORT  (0x400000LL)
#define MM_F_ALL_CHAINS (0x800000LL)
#define MM_F_OUT_MD     (0x1000000LL)
#define MM_F_COPY_COMMENT (0x2000000LL)
#define MM_F_EQX        (0x4000000LL)
#define MM_F_PAF_NO_HIT (0x8000000LL)
#define MM_F_NO_END_FLT (0x10000000LL)
#define MM_F_HARD_MLEVEL (0x20000000LL)
#define MM_F_SAM_HIT_ONLY (0x40000000LL)
#define MM_F_RMQ        (0x80000000LL)
#define MM_F_QSTRAND    (0x100000000LL)
#define MM_F_NO_INV     (0x200000000LL)
#define MM_F_NO_HASH_NAME (0x400000000LL)
#define MM_F_SPLICE_OLD (0x800000000LL)
#define MM_F_SECONDARY_SEQ (0x1000000000LL)

#define MM_I_HPC     0x1
#define MM_I_NO_SEQ  0x2
#define MM_I_NO_NAME 0x4

#define MM_CIGAR_MATCH 0
#define MM_CIGAR_INS   1
#define MM_CIGAR_DEL   2
#define MM_CIGAR_STR   "MIDNSHP=XB"

/* reference: minimap.h:75-80 (24 bytes) */
typedef struct {
	char *name;
	uint64_t offset;
	uint32_t len;
	uint32_t is_alt;
} mm_idx_seq_t;

/* reference: minimap.h:82-92 (80 bytes).  The caller reads b,w,k,flag,n_seq and seq[] directly
 * (packages/minimap2/src/map.rs:278-291); S/B/I/km/h are private to the implementation -- here `h` points at the
 * device-resident index object and the others stay NULL. */
typedef struct {
	int32_t b, w, k, flag;
	uint32_t n_seq;
	int32_t index;
	int32_t n_alt;
	mm_idx_seq_t *seq;
	uint32_t *S;
	struct mm_idx_bucket_s *B;
	struct mm_idx_intv_s *I;
	void *km, *h;
} mm_idx_t;

/* reference: minimap.h:95-101 (24 bytes + flexible array); cigar[i] = len<<4 | op */
typedef struct {
	uint32_t capacity;
	int32_t dp_score, dp_max, dp_max2;
	uint32_t n_ambi:30, trans_strand:2;
	uint32_t n_cigar;
	uint32_t cigar[];
} mm_extra_t;

/* reference: minimap.h:103-117 (80 bytes) */
typedef struct {
	int32_t id;
	int32_t cnt;
	int32_t rid;
	int32_t score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc;
	int32_t as;
	int32_t mlen, blen;
	int32_t n_sub;
	int32_t score0;
	uint32_t mapq:8, split:2, rev:1, inv:1, sam_pri:1, proper_frag:1, pe_thru:1, seg_split:1, seg_id:8, split_inv:1, is_alt:1, strand_retained:1, dummy:5;
	uint32_t hash;
	float div;
	mm_extra_t *p;
} mm_reg1_t;

/* reference: minimap.h:119-123 (24 bytes) */
typedef struct {
	short k, w, flag, bucket_bits;
	int64_t mini_batch_size;
	uint64_t batch_size;
} mm_idxopt_t;

/* reference: minimap.h:125-181 (248 bytes); the Rust wrapper writes fields in place
 * (packages/minimap2/src/options_args.rs:273-551) */
typedef struct {
	int64_t flag;
	int seed;
	int sdust_thres;
	int max_qlen;
	int bw, bw_long;
	int max_gap, max_gap_ref;
	int max_frag_len;
	int max_chain_skip, max_chain_iter;
	int min_cnt;
	int min_chain_score;
	float chain_gap_scale;
	float chain_skip_scale;
	int rmq_size_cap, rmq_inner_dist;
	int rmq_rescue_size;
	float rmq_rescue_ratio;
	float mask_level;
	int mask_len;
	float pri_ratio;
	int best_n;
	float alt_drop;
	int a, b, q, e, q2, e2;
	int sc_ambi;
	int noncan;
	int junc_bonus;
	int zdrop, zdrop_inv;
	int end_bonus;
	int min_dp_max;
	int min_ksw_len;
	int anchor_ext_len, anchor_ext_shift;
	float max_clip_ratio;
	int rank_min_len;
	float rank_frac;
	int pe_ori, pe_bonus;
	float mid_occ_frac;
	float q_occ_frac;
	int32_t min_mid_occ, max_mid_occ;
	int32_t mid_occ;
	int32_t max_occ, max_max_occ, occ_dist;
	int64_t mini_batch_size;
	int64_t max_sw_mat;
	int64_t cap_kalloc;
	const char *split_prefix;
} mm_mapopt_t;

/* reference: minimap.h:196-199; opaque to the caller (packages/minimap2/src/buf.rs) */
typedef struct mm_tbuf_s mm_tbuf_t;

/* replaces options.c:88-162.  preset==NULL initialises both structs to the defaults; returns 0, or -1 for an unknown
 * preset (packages/minimap2/src/options.rs:89-112 turns non-zero into an error) */
PGMM_API int mm_set_opt(const char *preset, mm_idxopt_t *io, mm_mapopt_t *mo);
/* replaces options.c:164-234; 0 when valid, the same negative codes otherwise (options.rs:123) */
PGMM_API int mm_check_opt(const mm_idxopt_t *io, const mm_mapopt_t *mo);
/* replace options.c:5-12 and :14-64 (used by the Default impls, minimap2-sys/src/lib.rs:13-31) */
PGMM_API void mm_idxopt_init(mm_idxopt_t *opt);
PGMM_API void mm_mapopt_init(mm_mapopt_t *opt);
/* replaces index.c:408-456: builds the minimizer index of n in-memory sequences ON THE GPU.  Copies names and
 * sequences (the caller frees its strings right after, packages/minimap2/src/index.rs:40-52).  NULL when n<=0. */
PGMM_API mm_idx_t *mm_idx_str(int w, int k, int is_hpc, int bucket_bits, int n, const char **seq, const char **name);
/* replaces options.c:66-80: computes mid_occ from the index (index.rs:47) */
PGMM_API void mm_mapopt_update(mm_mapopt_t *opt, const mm_idx_t *mi);
/* replaces index.c:56-79 (index.rs:71-77) */
PGMM_API void mm_idx_destroy(mm_idx_t *mi);
/* replace map.c:13-26 (buf.rs:17,36).  One per concurrent mm_map call. */
PGMM_API mm_tbuf_t *mm_tbuf_init(void);
PGMM_API void mm_tbuf_destroy(mm_tbuf_t *b);
/* replaces map.c:376-381 (packages/minimap2/src/map.rs:390).  Returns a malloc() array of *n_regs hits, each with a
 * separately malloc()ed ->p; the caller frees every ->p and then the array with free() (minimap.h:353-366,
 * map.rs:407-421).  Re-entrant for distinct tbufs. */
PGMM_API mm_reg1_t *mm_map(const mm_idx_t *mi, int l_seq, const char *seq, int *n_regs, mm_tbuf_t *b, const mm_mapopt_t *opt, const char *name);
/* replaces align.c:911-917 (map.rs:323) */
PGMM_API double mm_event_identity(const mm_reg1_t *r);

/* ===================== Part 2: batched mapping and the host half of the path ===================== */

/* All queries of one find_matches round at once (what mm_map does per query, packages/pangraph/src/align/
 * minimap2_lib/align_with_minimap2_lib.rs:64-74, done for the whole batch so the GPU sees every DP problem of the
 * round together).  n_regs[i]/regs[i] follow mm_map's ownership rules; results are independent of batch composition. */
PGMM_API void pgmm_map_batch(const mm_idx_t *mi, int n, const int *lens, const char *const *seqs, const char *const *names,
                             const mm_mapopt_t *opt, int *n_regs, mm_reg1_t **regs);

/* The same round with the inputs kept resident in HBM: pgmm_idx_upload codes the sequences and copies them to the
 * device once; pgmm_idx_build runs the sketch and index kernels on the resident bases (mm_idx_str = upload + build);
 * pgmm_map_self maps the indexed sequences against their own index -- pangraph's all-vs-all pattern -- deriving the
 * query-side buffers on the device, so no sequence crosses the bus again.  n_regs/regs have mi->n_seq entries. */
PGMM_API mm_idx_t *pgmm_idx_upload(int n, const char **seq, const char **name);
PGMM_API void pgmm_idx_build(mm_idx_t *mi, int w, int k, int bucket_bits);
PGMM_API void pgmm_map_self(const mm_idx_t *mi, const mm_mapopt_t *opt, int *n_regs, mm_reg1_t **regs);

/* One alignment record, the C image of pangraph's `Alignment` (packages/pangraph/src/align/alignment.rs:13-59).
 * Block names are the decimal BlockId values; cigar is malloc()ed (len<<4|op) and owned by the caller. */
typedef struct {
	uint64_t qry_name, ref_name;
	uint64_t qry_len, ref_len;
	uint64_t qry_start, qry_end, ref_start, ref_end;
	uint64_t matches, length, quality;
	int32_t reverse;      /* Strand::Reverse */
	int32_t has_divergence;
	double divergence;    /* PAF de */
	double align;         /* PAF AS */
	uint32_t n_cigar;
	uint32_t *cigar;
} pgmm_alignment_t;

/* AlignmentArgs of the reference (packages/pangraph/src/align/alignment_args.rs:6-37) */
typedef struct {
	uint64_t indel_len_threshold; /* -l, default 100 */
	double alpha;                 /* -a, default 100 */
	double beta;                  /* -b, default 10 */
	uint64_t sensitivity;         /* -s, 5|10|20 */
	int64_t kmer_length;          /* -K, <=0: preset default */
} pgmm_alignment_args_t;

PGMM_API void pgmm_alignment_args_default(pgmm_alignment_args_t *args);

/* align_with_minimap2_lib (align_with_minimap2_lib.rs:15-121): block ids + consensus sequences -> Alignment list in
 * query-index order.  Returns 0, or a negative code (-1 unknown sensitivity, -2 size mismatch / bad input).
 * *out is one malloc() array of *n_out records; free with pgmm_alignments_free. */
PGMM_API int pgmm_align_with_minimap2_lib(int n_blocks, const uint64_t *block_ids, const char *const *consensus,
                                          const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out);
/* split_matches (packages/pangraph/src/pangraph/split_matches.rs:13-24) on one record */
PGMM_API int pgmm_split_matches(const pgmm_alignment_t *aln, const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out);
/* alignment_energy2 (packages/pangraph/src/align/energy.rs:37-54) */
PGMM_API double pgmm_alignment_energy2(const pgmm_alignment_t *aln, const pgmm_alignment_args_t *args);
/* filter_matches (packages/pangraph/src/pangraph/graph_merging.rs:187-216) */
PGMM_API int pgmm_filter_matches(const pgmm_alignment_t *alns, size_t n, const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out);
/* the alignment half of self_merge (graph_merging.rs:95-121): find_matches, drop self hits, split, filter */
PGMM_API int pgmm_find_filtered_matches(int n_blocks, const uint64_t *block_ids, const char *const *consensus,
                                        const pgmm_alignment_args_t *args, pgmm_alignment_t **out, size_t *n_out);
PGMM_API void pgmm_alignments_free(pgmm_alignment_t *alns, size_t n);

/* ===================== Part 3: stage-level entry points (tests, bench) ===================== */

PGMM_API int pgmm_device_count(void);
/* Binds the process to one GPU (one process per GPU; call before anything else, e.g. with LOCAL_RANK).  Without it the
 * device that is current at the first call is used.  0 ok, -1 no such device, -2 already bound to another device. */
PGMM_API int pgmm_set_device(int device);

/* K5 alone: n DP problems over windows of two host code buffers (bases coded 0..4).  flag = KSW_EZ_* of the
 * reference (ksw2.h:8-14) | 0x10000 to read both windows back to front.  out_ez: 11 int32 per problem
 * (max,zdropped,max_q,max_t,mqe,mqe_t,mte,mte_q,score,reach_end,n_cigar); problem i's CIGAR is
 * out_cigar[out_cig_start[i] .. +n_cigar).  Returns 0, -1 if cigar_cap is too small. */
PGMM_API int pgmm_ksw_extd2_batch(int n, const int32_t *qlen, const int32_t *tlen, const uint64_t *q_off,
                                  const uint64_t *t_off, const uint8_t *qcodes, uint64_t q_total,
                                  const uint8_t *tcodes, uint64_t t_total, const int32_t *w, const int32_t *zdrop,
                                  const int32_t *end_bonus, const int32_t *flag, int a, int b, int sc_ambi, int q,
                                  int e, int q2, int e2, int32_t *out_ez, uint32_t *out_cigar, uint64_t cigar_cap,
                                  uint64_t *out_cig_start, double *out_kernel_ms, uint64_t arena_budget_bytes);

/* K1 alone: (w,k)-minimizers of n ASCII sequences in the reference's emission order (sketch.c:77-143):
 * x = hash<<8|span, y = seq<<32 | lastPos<<1 | strand; sequence i's minimizers are out[out_off[i] .. out_off[i+1]).
 * Returns 0, -1 if cap is too small. */
PGMM_API int pgmm_sketch(int n, const char *const *seqs, const int *lens, int w, int k, uint64_t *out_x, uint64_t *out_y,
                         uint64_t cap, uint64_t *out_off);

/* K1+K3 alone: what collect_seed_hits (map.c:168-204) builds for each query: anchors (x,y pairs) in collection order -- or
 * already ordered by target position when no two of them share one (then that order is the reference's sort result) --,
 * the query positions of the seeds used (seed.c:125) and the repeat length (seed.c:113-128).
 * out_n[3*i..3*i+2] = {n_anchors, n_mini_pos, rep_len}. Returns 0, -1 if a capacity is too small. */
PGMM_API int pgmm_collect_seeds(const mm_idx_t *mi, int n, const int *lens, const char *const *seqs, const char *const *names,
                                const mm_mapopt_t *opt, uint64_t *out_anchor, uint64_t anchor_cap, uint64_t *out_mini,
                                uint64_t mini_cap, int64_t *out_n);

/* K4 alone: chains n anchors (2 uint64 each, sorted like radix_sort_128x leaves them) the way mg_lchain_rmq does
 * (lchain.c:250-368): the device fills scores / predecessors / peak scores segment by segment, the host backtracks and
 * compacts.  Segments the device hands back (equal priorities inside an RMQ window, oversized windows) are filled by the
 * host arbiter when host_redo != 0; with host_redo == 0 the call returns -2 if there was any.  Returns the number of
 * chains; u[i] = score<<32 | n_anchors, xy = the kept anchors chain after chain, *n_a_out their number; out_fpv
 * (optional, 3n int32) = f, p, v as filled; seg_stats (optional) = {segments, handed back, their anchors}. */
PGMM_API int64_t pgmm_chain_rmq(uint64_t *xy, int64_t n, int max_dist, int max_dist_inner, int bw, int max_skip, int cap,
                                int min_cnt, int min_sc, float pen_gap, float pen_skip, uint64_t *u, int64_t *n_a_out,
                                int32_t *out_fpv, int64_t *seg_stats, int host_redo);

/* CTA trace of the DP (K5/K5a/K5b) and chaining (K4) kernels: between begin and end every traced CTA appends one record
 * of 40 bytes {uint64 t0, t1, t2 (%globaltimer ns: start, end of the main loop, end); uint32 kernel (1 K5a, 2 K5b first
 * pass, 3 K5b exact pass, 4 K5 generic, 5 K4), block, smid, aux (rows or anchors)}.  begin: 0 ok, -1 already tracing.
 * end: copies up to max_n records to out and returns how many the kernels produced. */
PGMM_API int pgmm_cta_trace_begin(uint64_t capacity);
PGMM_API int64_t pgmm_cta_trace_end(void *out, uint64_t max_n);

/* ===================== Part 4: map_variations ===================== */

/* The Edit the reference's map_variations returns (packages/pangraph/src/align/map_variations.rs:39-80,
 * packages/pangraph/src/pangraph/edits.rs: Sub {pos, alt}, Del {pos, len}, Ins {pos, seq}).  Arrays are malloc blocks owned by the
 * record; pgmm_edits_free releases them.  status: 0 ok; -1 where the reference returns Err (a character outside
 * "TAWCYMHGKRDSBVN", an empty query); -2 / -3 where the reference panics (backtrace dead end / outside the band);
 * -4 band wider than 24 000 columns (the kernel's row state; reported, never guessed). */
typedef struct pgmm_edit_s {
  int32_t status, hit_boundary, attempts, band_width, score;
  int32_t n_sub, n_del, n_ins;
  int32_t *sub_pos; char *sub_chr;          /* substitutions, ascending position */
  int32_t *del_pos, *del_len;               /* inner deletions ascending, then the leading, then the trailing one */
  int32_t *ins_pos, *ins_len; char *ins_seq;/* insertions ascending; pos = position after the insertion; bases concatenated */
} pgmm_edit_t;

/* n problems at once: qry[i] against ref[i] inside the band (mean_shift[i], band_width[i] + extra_band_width), retried with a
 * doubled band up to max_alignment_attempts times while the traceback touches the band boundary (align.rs:52-63).
 * Defaults of the reference: extra_band_width 5, max_alignment_attempts 4 (commands/build/build_args.rs:76-85).
 * *out = malloc array of n records.  stats (optional, 4 doubles): kernel ms, band cells, problem attempts, launches. */
PGMM_API int pgmm_map_variations_batch(int n, const char *const *refs, const int32_t *ref_lens, const char *const *qrys,
                                       const int32_t *qry_lens, const int32_t *mean_shift, const int32_t *band_width,
                                       int extra_band_width, int max_alignment_attempts, pgmm_edit_t **out, double *stats);
PGMM_API void pgmm_edits_free(pgmm_edit_t *edits, int n);

/* ===================== Part 5: the guide tree ===================== */

/* mash_distance (packages/pangraph/src/distance/mash/mash_distance.rs:9-68) over minimizers_sketch
 * (distance/mash/minimizer.rs:49-160): sketches of all sequences, the number of distinct minimizer values each pair shares
 * (K7 on the GPU), dist[i][j] = 1 - shared(i, j) / distinct(min(i, j)).  seqs[i] = raw FASTA bytes of sequence i (any case;
 * everything but ACGTU is an ambiguous base).  The reference always calls it with k = 15, w = 100 (MinimizersParams::default,
 * tree/neighbor_joining.rs:42).  dist = n x n doubles, row-major.
 * -> 0; 1 + i: sequence i has no minimizer (the reference panics: "no minimizer found for a sequence during mash distance
 * evaluation"); -1: k outside 1..31 or w outside 1..255 (the reference's assert!s); -2: n < 1; -3: 2k + bits(n) > 64;
 * -4: a sequence of 2^31 bases or more; -5: the incidence bitmap does not fit the device.
 * stats (optional, 10 doubles): upload ms, sketch ms, sort ms, pair ms, bases, tiles, minimizers, distinct (value, sequence)
 * keys, values shared by two or more sequences, kernel launches. */
PGMM_API int pgmm_mash_distance(int n, const char *const *seqs, const int64_t *lens, int k, int w, double *dist, double *stats);

/* build_tree_using_neighbor_joining (tree/neighbor_joining.rs:16-35, 46-101) on an n x n distance matrix; host code.
 * Leaves are 0..n-1, the node made by the t-th join is n + t with children left[t], right[t] (n - 1 entries each), the root
 * is 2n - 2.  -> 0; -1: n < 2 (the reference indexes nodes[1] and panics); -2: a NaN in the matrix. */
PGMM_API int pgmm_nj_tree(int n, const double *dist, int32_t *left, int32_t *right);

/* parse_newick (tree/newick.rs:43-62): *n_leaves leaves numbered in order of appearance, *names = their labels separated by
 * '\0' (one malloc block), *left / *right = children of the nodes n, n + 1, ... (malloc, n - 1 entries, root last).
 * -> 0, or -1 with the reference's message in err (truncated to err_cap).  Branch lengths and internal labels are read and
 * dropped; only strictly bifurcating trees are accepted. */
PGMM_API int pgmm_newick_parse(const char *text, int32_t *n_leaves, char **names, int64_t *names_bytes, int32_t **left, int32_t **right,
                               char *err, int err_cap);
/* Clade::to_newick (tree/newick.rs:11-38): -> malloc'd string (internal nodes unlabeled) */
PGMM_API char *pgmm_newick_write(int n, const char *const *names, const int32_t *left, const int32_t *right);
/* balance (tree/balance.rs:4-18): the same leaves left to right under a bisected tree */
PGMM_API int pgmm_tree_balance(int n, const int32_t *left, const int32_t *right, int32_t *out_left, int32_t *out_right);
/* postorder (tree/clade.rs:49-71): the 2n - 1 nodes in the order the reference's build loop visits them */
PGMM_API int pgmm_tree_postorder(int n, const int32_t *left, const int32_t *right, int32_t *order);
PGMM_API void pgmm_free(void *p);

/* ===================== Part 6: FASTA input ===================== */

/* FastaRecord (packages/pangraph/src/io/fasta.rs:17-24).  desc == NULL where the reference has None.  seq is upper-cased and
 * NUL-terminated; len = its length. */
typedef struct pgmm_fasta_record_s {
  char *name, *desc, *seq;
  int64_t len, index;
} pgmm_fasta_record_t;

/* FastaReader::from_paths(paths).read_many() (io/fasta.rs:92-129, 131-225): the files one after the other (".gz" inflated;
 * ".bz2" / ".xz" / ".zst" refused: this build links zlib only).  alphabet = accepted characters after upper-casing, NULL = the
 * reference's default "ACGTYRWSKMDVHBN" (no gap).  *recs = malloc array of *n_recs records (pgmm_fasta_free).
 * -> 0, or -1 with the reference's message in err ("FASTA input is incorrectly formatted: ...", "When processing sequence
 * #i: \">name\": FASTA input is incorrect: character \"c\" is not in the alphabet"). */
PGMM_API int pgmm_fasta_read_files(int n_paths, const char *const *paths, const char *alphabet, pgmm_fasta_record_t **recs,
                                   int64_t *n_recs, char *err, int err_cap);
/* the same over one buffer (FastaReader::from_str) */
PGMM_API int pgmm_fasta_read_buffer(const char *data, int64_t n, const char *alphabet, pgmm_fasta_record_t **recs, int64_t *n_recs,
                                    char *err, int err_cap);
PGMM_API void pgmm_fasta_free(pgmm_fasta_record_t *recs, int64_t n_recs);

/* counters since the last reset: [0] total_ms [1] seed_ms [2] dp_kernel_ms [3] index_ms [4] dp_jobs [5] dp_cells
 * [6] dp_waves [7] bases_mapped [8] bases_indexed [9] batches [10] kernel launches [11..16] wall ms of the phases of
 * pgmm_map_batch (encode, seeding, sort+chain+plan, DP waves, stitching between waves, final filters)
 * [17] host->device bytes [18] device->host bytes [19] bases read by the DP kernels [20] cudaMalloc calls
 * [21..32] per DP kernel family (K5 generic, K5a small fills, K5b wide fills): ms on the launch streams, cells, bases read,
 * launches [33..41] chaining: host sort ms, device fill ms (with copies), host rest ms (redo + backtrack + hit skeletons +
 * plan), fill kernel ms, anchors, segments, segments handed back to the host, their anchors, launches 
 * [42] fixed-point iterations of the chain fill summed over its batches, [43] those batches
 * [44..53] host CPU ms (thread CPU clocks) by phase: encode, seeding, anchor sort, chain fill, refill + backtrack + hits +
 * plan, DP waves (round side), DP service workers, stitching, final, index build */
PGMM_API void pgmm_get_stats(double *out, int n, int reset);

#ifdef __cplusplus
}
#endif
#endif

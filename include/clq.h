/*
 * clq.h -- C ABI of libclq: B200-native batched amplicon alignment, a drop-in for the alignment hot
 * path of mckennalab/clique (global affine-gap Gotoh DP + traceback, best-candidate reference selection).
 *
 * The reference has no FFI of its own; the seam is the set of Rust functions its callers use.  Each entry
 * point below names the reference interface it replaces (file:line relative to rust_cmd/src/ of the
 * reference).  INTEGRATION.md shows the Rust `extern "C"` block + shim that binds these symbols.
 *
 * Conventions: plain C, no exceptions across the boundary, every call returns an int32 status (0 = OK,
 * negative = call-level error, see clq_strerror), per-read outcomes are reported in clq_result_t.status.
 * The caller owns all host buffers.  One clq_ctx per GPU; a ctx is not thread-safe, distinct ctxs are fully
 * independent (read-sharded multi-GPU: one ctx per device, no collectives).
 * There is no CPU fallback: without a CUDA device every entry point that needs one fails with CLQ_E_CUDA.
 */
#ifndef CLQ_H
#define CLQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLQ_VERSION 100

/* ---- per-read status (clq_result_t.status) ---- */
#define CLQ_OK 0
#define CLQ_READ_TOO_LONG 1             /* align_reads drops reads >= max_read_size with warn!; alignment_functions.rs:147,240-247 */
#define CLQ_SCORING_NOT_REPRESENTABLE 2 /* scores are not dyadic rationals / gap_open >= 0: no exact integer form */
#define CLQ_TRACEBACK_DIVERGED 3        /* the reference would spin forever on a stale Up(0) cell; alignment/alignment_matrix.rs:977-1051 */
#define CLQ_CIGAR_POOL_FULL 4
#define CLQ_NO_CANDIDATE 5              /* no reference to align to: Option::None; alignment_functions.rs:536-543 */

/* ---- call-level errors (negative return values) ---- */
#define CLQ_E_INVALID (-1)
#define CLQ_E_CUDA (-2)
#define CLQ_E_NOMEM (-3)
#define CLQ_E_LIMIT (-4)
#define CLQ_E_STATE (-5)
#define CLQ_E_UNSUPPORTED (-6)

/* ---- flags for clq_submit / clq_launch ---- */
#define CLQ_BAND_MAXLEN 0u           /* perform_affine_alignment: bandwidth = max(L1,L2); alignment/alignment_matrix.rs:366-372 */
#define CLQ_BAND_READLEN 1u          /* align_two_strings_passed_matrix(.., &read.len()); alignment_functions.rs:737-746,790-799 */
#define CLQ_BAND_K 2u                /* perform_affine_alignment_bandwidth(.., &k) with an explicit bandwidth k = flags >> CLQ_BAND_K_SHIFT
                                        (alignment/alignment_matrix.rs:376-425: row x computes y in [max(1, c - k), min(L2 + 1, c + k)),
                                        c = trunc(f64(x) / (L1 + 1) * (L2 + 1)); skipped cells keep the fresh-matrix state).  No caller in
                                        the reference passes one today (the 1-reference call with &100 is commented out, :586-597);
                                        generic kernels, affine scoring. */
#define CLQ_BAND_K_SHIFT 8
#define CLQ_BAND_MASK 3u
#define CLQ_SEARCH_FIXED (0u << 2)      /* fixed_ref[i] names the reference of read i */
#define CLQ_SEARCH_EXHAUSTIVE (1u << 2) /* exhaustive_alignment_search; alignment_functions.rs:769-827 */
#define CLQ_SEARCH_QUICK (2u << 2)      /* quick_alignment_search (fast_lookup = true); alignment_functions.rs:693-767 */
#define CLQ_SEARCH_MASK (3u << 2)
#define CLQ_SCORE_ONLY (1u << 4)        /* skip traceback: score + ref_index only */
#define CLQ_CONVEX (1u << 5)            /* two-piece affine gaps (scoring given as clq_convex_t) */
#define CLQ_EXTRACT_TAGS (1u << 6)      /* also record the read bytes aligned to the reference's tag columns '0'..'9':
                                           extract_tagged_sequences' digit keys, extractor.rs:271-332 (see clq_tags_download) */

#define CLQ_RUSTBIO (1u << 7)           /* the single-reference branch of align_to_reference_choices (alignment_functions.rs:544-603):
                                           rust-bio `pairwise::Aligner::global(read, reference)` semantics -- unbanded, gap(k) =
                                           open + k*extend on the boundaries without clique's corner rule, a read 'N' matches any
                                           byte, ties: match > insertion > deletion, extension only when strictly better than
                                           opening.  PARITY UNPINNED (un-vendored `bio = "*"`); needs CLQ_SEARCH_FIXED,
                                           scoring from clq_rustbio_scoring.  A read holding a byte the 16 x 8 class table cannot
                                           score exactly gets CLQ_SCORING_NOT_REPRESENTABLE. */

/* CIGAR op encoding in the pool: len << 4 | code, BAM codes (AlignmentTag -> Op, alignment/alignment_matrix.rs:95-107) */
#define CLQ_OP_M 0u /* AlignmentTag::MatchMismatch */
#define CLQ_OP_I 1u /* AlignmentTag::Ins */
#define CLQ_OP_D 2u /* AlignmentTag::Del */

typedef struct clq_ctx clq_ctx;

/* AffineScoring (alignment/scoring_functions.rs:65-73) in exact scaled-integer form: every field is the
 * f64 value times `scale`.  Build it with clq_affine_from_f64. */
typedef struct {
    int32_t scale;         /* power of two, 1..64 */
    int32_t match, mismatch, special;
    int32_t oe_in, e_in;   /* interior cells: gap_open + gap_extend, gap_extend                              */
    int32_t oe_fin, e_fin; /* x == L1 || y == L2: gap_open + gap_extend*mult, gap_extend*mult (:625-627)      */
    int32_t b0, b1;        /* boundary g(k) = b0 + k*b1 = (gap_open + k*gap_extend)*mult (:389-405)           */
    int32_t max_neg;       /* MAX_NEG_SCORE * scale = -100000 * scale (:34)                                   */
} clq_affine_t;

/* two-piece affine ("convex") scoring; semantics defined by this repo, gap(k) = max(o1+k*e1, o2+k*e2) */
typedef struct {
    int32_t match, mismatch, special;
    int32_t o1, e1, o2, e2;
    int32_t max_neg;
} clq_convex_t;

/* capacity of a context; a batch exceeding it is rejected with CLQ_E_LIMIT */
typedef struct {
    uint32_t max_reads;       /* reads per batch (per slot)                                            */
    uint64_t max_read_bytes;  /* total read bytes per batch                                            */
    uint32_t max_read_len;    /* reads >= this are dropped with CLQ_READ_TOO_LONG (align_reads' rule)   */
    uint32_t max_refs;
    uint64_t max_ref_bytes;
    uint64_t cigar_pool_ops;  /* CIGAR pool capacity per slot, in uint32 ops (clamped to 2^32 - 1)       */
    uint32_t n_slots;         /* independent stream slots for double buffering (1..4)                   */
} clq_limits_t;

/* one record per read: AlignmentResult.score / cigar_string + AlignmentWithRef.ref_name
 * (alignment/alignment_matrix.rs:694-706, alignment_functions.rs:451-456) */
typedef struct {
    int32_t score_scaled; /* AlignmentResult.score * scale (exact) */
    uint32_t ref_index;   /* index into the clq_refs_set order; 0xffffffff when no usable reference (CLQ_NO_CANDIDATE) */
    uint32_t cigar_off;   /* first op in the CIGAR pool            */
    uint32_t cigar_len;   /* run-length merged ops                 */
    uint32_t status;
    uint32_t matches;     /* get_reference_alignment_rate's counters (consensus/consensus_builders.rs:288-307): aligned   */
    uint32_t mismatches;  /* columns with ref > 64, ref != 'N', read > 64; the `rm` tag = matches / (matches + mismatches) */
} clq_result_t;

/* timing / accounting of the last clq_launch on a slot */
typedef struct {
    float kernel_ms;       /* CUDA-event time over all kernels of the launch, on the slot's stream */
    float dp_ms;           /* the dominant DP kernel(s) alone                                       */
    uint32_t launches;     /* kernels launched                                                      */
    uint32_t dp_launches;
    uint64_t cells;        /* sum of L1*L2 over every (read, reference) pair filled                 */
    uint64_t h2d_bytes;    /* bytes copied host->device by clq_upload                               */
    uint64_t d2h_bytes;    /* bytes copied device->host by clq_download + clq_wait                  */
    uint32_t variant;      /* kernel family of the last launch: bit0 FAST (PRMT/DPX), bit1 PACK (s16x2, two reads per
                              lane group), bit2 CONVEX, bit3 final-gap multiplier variant, bit4 rust-bio semantics, bit5 PACK with the static
                              row slope (M step off the ALU pipe), bit6 PACK with the adaptive per-row bias + int32 retry pass (long reads);
                              bits 8.. = geometry index */
    uint32_t sub_batches;  /* fill + walk rounds the traceback scratch budget split the batch into  */
    uint32_t pack_retries; /* reads the adaptive s16x2 kernel (variant bit6) handed to the int32 retry pass; set by clq_wait */
    uint32_t reserved0;
} clq_stats_t;

int32_t clq_version(void);
const char* clq_strerror(int32_t code);
int32_t clq_device_count(void);

/* AffineScoring{..} -> exact integer form; CLQ_SCORING_NOT_REPRESENTABLE if no scale <= 64 makes every
 * derived constant integral, if gap_open >= 0 (the direction encoding needs a real opening penalty) or if a
 * value leaves the int32 working range. */
int32_t clq_affine_from_f64(double match_score, double mismatch_score, double special_character_score,
                            double gap_open, double gap_extend, double final_gap_multiplier, clq_affine_t* out);

/* rust_bio_alignment's scoring (alignment_functions.rs:48-61: 1 / -1, gap open -5, gap extend -1 hard-coded) for CLQ_RUSTBIO */
int32_t clq_rustbio_scoring(int32_t match_score, int32_t mismatch_score, int32_t gap_open, int32_t gap_extend, clq_affine_t* out);

/* pinned host memory for the caller's batch buffers (double-buffered H2D/D2H) */
int32_t clq_host_alloc(size_t bytes, void** out);
int32_t clq_host_free(void* p);

/* replaces create_scoring_record_3d + the rayon thread-local matrices: alignment/alignment_matrix.rs:226-233,
 * alignment_functions.rs:115-141 */
int32_t clq_ctx_create(int32_t device, const clq_limits_t* limits, clq_ctx** out);
void clq_ctx_destroy(clq_ctx* ctx);
const char* clq_ctx_last_error(const clq_ctx* ctx);

/* replaces ReferenceManager::from_yaml_input / from_fa_file (the reference set): reference/fasta_reference.rs:90-146.
 * bytes = raw ASCII, case preserved; off has n_refs + 1 entries. */
int32_t clq_refs_set(clq_ctx* ctx, uint32_t n_refs, const uint8_t* bytes, const uint64_t* off);

/* replaces ReferenceManager::unique_kmers: reference/fasta_reference.rs:159-202 (k = 8, skip = 4 in the CLI, main.rs:271) */
int32_t clq_kmer_index_set(clq_ctx* ctx, uint32_t k, uint32_t skip);

/* The batch call: replaces the closure body of align_reads' par_bridge loop up to the alignment result
 * (alignment_functions.rs:135-162) = align_to_reference_choices (:520-631) -> quick/exhaustive search ->
 * align_two_strings_passed_matrix (:383-449) -> perform_affine_alignment_bandwidth + perform_3d_global_traceback
 * (alignment/alignment_matrix.rs:376-425, :941-1086) -> simplify_cigar_string (alignment_manager.rs:386-423).
 * With CLQ_SEARCH_FIXED | CLQ_BAND_MAXLEN it is align_two_strings (alignment_manager.rs:231-273) per read.
 * Asynchronous: enqueues H2D copies, kernels and D2H copies on the slot's stream.  read_off has n_reads + 1
 * entries; fixed_ref may be NULL unless CLQ_SEARCH_FIXED; scoring points at a clq_affine_t (or clq_convex_t with
 * CLQ_CONVEX).  The host buffers must stay valid until clq_wait returns. */
int32_t clq_submit(clq_ctx* ctx, int32_t slot, uint32_t n_reads, const uint8_t* read_bytes, const uint64_t* read_off,
                   const int32_t* fixed_ref, const void* scoring, uint32_t flags, double match_threshold);

/* blocks until the slot's work is done and copies results out; cigar_used receives the ops written */
int32_t clq_wait(clq_ctx* ctx, int32_t slot, clq_result_t* results, uint32_t* cigar_pool, uint64_t cigar_cap,
                 uint64_t* cigar_used);

/* After clq_wait on a batch launched with CLQ_EXTRACT_TAGS: tags[i * tag_stride + k] = the read byte (or '-') aligned to
 * the k-th tag column (reference byte '0'..'9', in reference order) of read i's reference; bytes past that reference's
 * column count are undefined.  The tag string of symbol d is the concatenation of the bytes whose column holds d
 * (extract_tagged_sequences, extractor.rs:271-332; align_reads writes them as the e0..e9 BAM tags,
 * alignment_functions.rs:193-212).  Only defined for reads with status CLQ_OK.  tags == NULL only reports tag_stride. */
int32_t clq_tags_download(clq_ctx* ctx, int32_t slot, uint8_t* tags, uint64_t cap, uint32_t* tag_stride);

/* the three stages of clq_submit, separately (kernel-only timing with device-resident inputs) */
int32_t clq_upload(clq_ctx* ctx, int32_t slot, uint32_t n_reads, const uint8_t* read_bytes, const uint64_t* read_off,
                   const int32_t* fixed_ref);
int32_t clq_launch(clq_ctx* ctx, int32_t slot, const void* scoring, uint32_t flags, double match_threshold);
int32_t clq_download(clq_ctx* ctx, int32_t slot);
int32_t clq_sync(clq_ctx* ctx, int32_t slot);
int32_t clq_slot_stats(clq_ctx* ctx, int32_t slot, clq_stats_t* out);

/* 2-bit packed read ingestion -- the "2-bit packing" of the host layer BASELINE.json's north_star (1) names; the reference keeps
 * reads as raw bytes (Vec<u8>, read_strategies/sequence_layout.rs) and match_mismatch compares them byte for byte, case included
 * (alignment/scoring_functions.rs:100-102), so the packed form carries an exception list and is expanded on the device into
 * the same ASCII buffer clq_upload fills: results are identical to clq_submit's on the same reads.
 *   clq_pack2            host only, no CUDA call: A C G T -> 0 1 2 3, base i of the concatenated batch in bits 2 (i % 16) of
 *                        packed[i / 16] ((n_bytes + 15) / 16 words); every other byte (N, IUPAC, lower case) is stored as 0 and
 *                        listed as (exc_pos, exc_byte) in ascending position order.  *n_exc = number of exceptions; returns
 *                        CLQ_E_LIMIT when exc_cap is too small (*n_exc then tells the capacity needed; packed is complete).
 *   clq_upload_packed2 / clq_submit_packed2   as clq_upload / clq_submit with the batch in that form; read_off are base offsets
 *                        (read_off[0] == 0).  H2D traffic: 0.25 B per base + 9 B per exception instead of 1 B per base. */
int32_t clq_pack2(const uint8_t* bytes, uint64_t n_bytes, uint32_t* packed, uint64_t* exc_pos, uint8_t* exc_byte, uint64_t exc_cap,
                  uint64_t* n_exc);
int32_t clq_upload_packed2(clq_ctx* ctx, int32_t slot, uint32_t n_reads, const uint32_t* packed, const uint64_t* read_off,
                           const uint64_t* exc_pos, const uint8_t* exc_byte, uint64_t n_exc, const int32_t* fixed_ref);
int32_t clq_submit_packed2(clq_ctx* ctx, int32_t slot, uint32_t n_reads, const uint32_t* packed, const uint64_t* read_off,
                           const uint64_t* exc_pos, const uint8_t* exc_byte, uint64_t n_exc, const int32_t* fixed_ref,
                           const void* scoring, uint32_t flags, double match_threshold);

/* Tuning / experiment knobs.  Except "debug_flags" none of them changes a result (every kernel family is bit-exact against the
 * oracle); they pick which family runs, for A/B measurements and the parity tests.  Unknown keys return CLQ_E_INVALID.
 *   "max_scratch_bytes"  traceback scratch per context (direction bits + CIGAR scratch), >= 1 MiB; a batch whose bits exceed it
 *                        runs as several fill + walk rounds (default 96 GiB)
 *   "force_cfg"          wavefront geometry 0..5 = (8,16) (8,24) (8,40) (16,24) (32,16) (32,32); -1 = picked from the read lengths
 *   "force_generic"      1 = generic int32 kernels even when the FAST / s16x2 preconditions hold
 *   "no_pack" / "no_madd" / "no_adapt"   1 = without the s16x2 kernels / their static-row-slope form / the adaptive-bias form
 *   "adapt_guard"        guard band of the adaptive-bias kernel in score units (0 = derived from the scoring)
 *   "no_long8"           1 = long reads keep the geometry their length picks instead of (8,40) with column stripes
 *   "no_overlap"         1 = sub-batches run one after the other (dealt round-robin) instead of two at a time on two streams
 *   "no_group"           1 = multi-reference traceback without the on-device bucketing by reference (int32 kernels)
 *   "serialize_slots"    1 = launches of different stream slots are chained and share one scratch (default), 0 = per-slot scratch
 *   "debug_flags"        measurement only, results incomplete: 1 = skip the traceback walk (no CIGARs), 2 = skip the int32 retry
 *                        pass behind the adaptive-bias kernel */
int32_t clq_set_option(clq_ctx* ctx, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif

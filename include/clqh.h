/*
 * clqh.h -- extern "C" view of the pure host-side functions of the C++ host layer (include/clique_host.hpp), exported by
 * libclq_host.so for the ctypes test-suite.  These are string functions of the path's *output* (SURVEY.md section 8f):
 * they need no GPU.  The drop-in boundary of the alignment path itself is include/clq.h.
 * Citations: file:line relative to rust_cmd/src/ of the reference.
 */
#ifndef CLQH_H
#define CLQH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* extract_tagged_sequences, extractor.rs:271-332.  The BTreeMap is returned as records [key u8][len u32 LE][bytes] in key
 * order; returns bytes written (0 if cap is too small). */
size_t clqh_extract_tagged_sequences(const uint8_t* aligned_read, size_t n_read, const uint8_t* aligned_ref, size_t n_ref, uint8_t* out,
                                     size_t cap);
/* reverse_complement, utils/read_utils.rs:50-72 */
void clqh_reverse_complement(const uint8_t* dna, size_t n, uint8_t* out);
/* f64 Display (score.to_string() in the rm / rs / as tags, alignment/alignment_matrix.rs:747-757) */
size_t clqh_f64_to_string(double v, char* out, size_t cap);
/* get_reference_alignment_rate, consensus/consensus_builders.rs:288-307 */
double clqh_get_reference_alignment_rate(const uint8_t* ref_aligned, const uint8_t* read_aligned, size_t n);
/* simplify_cigar_string, alignment_manager.rs:386-423, on ops encoded len << 4 | code */
size_t clqh_simplify_cigar(const uint32_t* ops, size_t n, uint32_t* out);
/* AlignmentResult rebuilt from CIGAR + sequences (gapped strings + path), alignment/alignment_matrix.rs:1019-1086 */
int32_t clqh_from_cigar(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, const uint32_t* ops, size_t n_ops,
                        uint8_t* ref_aligned, uint8_t* read_aligned, size_t aligned_cap, size_t* aligned_len, uint32_t* path_xy,
                        size_t path_cap, size_t* path_len);
/* AlignmentResult::to_sam_record as one SAM text line, alignment/alignment_matrix.rs:741-771; extra_tags = "k1=v;k2=v" */
size_t clqh_sam_line(const char* ref_name, const char* read_name, const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2,
                     const uint32_t* ops, size_t n_ops, double score, int32_t reference_id, const char* extra_tags, char* out, size_t cap);
/* combine_phred_scores, utils/read_utils.rs:26-38 */
uint8_t clqh_combine_phred_scores(uint8_t a, uint8_t b, int32_t agree);
/* alignment_rate_and_consensus, merger.rs:428-498; (size_t)-1 where the reference panics */
size_t clqh_alignment_rate_and_consensus(const uint8_t* a1, const uint8_t* q1, size_t nq1, const uint8_t* a2, const uint8_t* q2, size_t nq2,
                                         size_t n, uint8_t* out_bases, uint8_t* out_quals);
/* merge_reads_by_alignment (merger.rs:348-396) for n read pairs in one GPU launch per chunk: needs a CUDA device */
int32_t clqh_merge_read_pairs_by_alignment(int32_t device, uint32_t n, const uint8_t* r1, const uint8_t* q1, const uint64_t* off1,
                                           const uint8_t* r2, const uint8_t* q2, const uint64_t* off2, double match_score,
                                           double mismatch_score, double special_score, double gap_open, double gap_extend,
                                           double final_gap_multiplier, uint8_t* out_bases, uint8_t* out_quals, uint64_t cap,
                                           uint64_t* out_off);
/* find_greedy_non_overlapping_segments / orient_by_longest_segment, linked_alignment.rs:24-32, :97-130 (segments as
 * (search_start, ref_start, length) triples) */
size_t clqh_find_greedy_non_overlapping_segments(const uint8_t* search, size_t n, const uint8_t* reference, size_t m, size_t seed_size,
                                                 uint32_t* out_xyz, size_t cap, size_t* start_position);
int32_t clqh_orient_by_longest_segment(const uint8_t* search, size_t n, const uint8_t* reference, size_t m, size_t seed_size);
/* extend_hit, linked_alignment.rs:341-362 */
size_t clqh_extend_hit(const uint8_t* search, size_t n, size_t search_location, const uint8_t* reference, size_t m, size_t reference_location);
/* BamFileAlignmentWriter's wire format (alignment_manager.rs:64-209) for one alignment: BGZF header block + record + EOF */
size_t clqh_bam_file(const char* ref_name, const char* read_name, const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2,
                     const uint32_t* ops, size_t n_ops, double score, const char* extra_tags, uint8_t* out, size_t cap);
/* merge_reads_by_concatenation + orient_sequence, merger.rs:40-126; layout items "1F" "2R" "2C" "S:ACGT"; (size_t)-1 = panic */
size_t clqh_merge_reads_by_concatenation(const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const char* layout, uint8_t* out,
                                         size_t cap);

/* align_reads over an in-memory span through clique::ShardedAligner::align_reads_span (one clq_ctx + host threads per device;
 * the analogue of the par_bridge loop, alignment_functions.rs:135): reads from plain (unpinned) host memory, results into the
 * caller's arrays in input order.  Needs CUDA devices.  stats (nullable, >= 8 + 3 * n_devices doubles): [seconds, setup_seconds,
 * fill_seconds, sink_seconds, reads, aligned, dropped, batches], then per device kernel_ms, reads, cells. */
typedef struct {
    uint32_t max_reads;            /* per batch / stream slot */
    uint64_t max_read_bytes;       /* per batch; 0: max_reads * 512 */
    uint32_t max_read_len;
    uint32_t cigar_ops_per_read;
    uint32_t n_slots;
    int32_t fillers_per_device;
    int32_t fast_lookup;           /* > 1 reference without fixed_ref: quick (1) or exhaustive (0) search */
    int32_t extract_tags;
} clqh_span_options_t;
int32_t clqh_align_reads_span(const int32_t* devices, uint32_t n_devices, const clqh_span_options_t* opt, const uint8_t* ref_bytes,
                              const uint64_t* ref_off, uint32_t n_refs, const uint8_t* read_bytes, const uint64_t* read_off, uint64_t n_reads,
                              const int32_t* fixed_ref, double match_score, double mismatch_score, double special_score, double gap_open,
                              double gap_extend, double final_gap_multiplier, uint32_t passes, void* results /* clq_result_t[n_reads] */,
                              uint32_t* cigar_pool, uint64_t cigar_cap, uint64_t* cigar_used, int32_t* scale, double* stats, char* err,
                              size_t err_cap);


/* The batch cutter of align_reads_span (clique::SpanClaimer, pure host logic: no CUDA device needed): replays the claims a
 * single thread would get over a span with the given read offsets.  order: 0 auto, 1 front to back, 2 longest first, 3 two-ended.
 * claims[4 * k .. 4 * k + 3] = lo, hi, lo2, hi2 of claim k (lo2 == hi2: one range).  Returns the number of claims (those
 * beyond `cap` are counted, not stored); *resolved_order (nullable) = what auto resolved to. */
uint64_t clqh_span_claims(const uint64_t* read_off, uint64_t n_reads, uint32_t n_devices, int32_t claimers_per_device, uint64_t max_reads,
                          uint64_t max_read_bytes, int32_t order, uint64_t* claims, uint64_t cap, int32_t* resolved_order);

#ifdef __cplusplus
}
#endif
#endif

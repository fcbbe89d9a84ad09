// clique_host.hpp -- C++17 host layer above the libclq C ABI (include/clq.h).
//
// The reference (mckennalab/clique) is compiled Rust and no Rust toolchain exists in this image, so the host side of the
// drop-in is C++.  Every type and function here mirrors one of the reference's (same name, argument meaning and error
// behaviour; file:line relative to rust_cmd/src/ of the reference); a Rust maintainer binds the C ABI directly
// (INTEGRATION.md), this layer is what a C++ caller -- and tools/clq_align -- use.  It does no alignment arithmetic:
// scores, CIGARs, selected references, the `rm` counters and the digit tags all come from the CUDA kernels.  There is no
// CPU fallback: constructing an Aligner without a CUDA device throws.
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "clq.h"

namespace clique {

using Bytes = std::vector<uint8_t>;

inline Bytes to_bytes(const std::string& s) { return Bytes(s.begin(), s.end()); }
inline std::string to_string(const Bytes& b) { return std::string(b.begin(), b.end()); }

/// panics / call-level errors of the reference surface become exceptions carrying the libclq status code
class ClqError : public std::runtime_error {
public:
    ClqError(int32_t code, const std::string& what) : std::runtime_error(what), code(code) {}
    int32_t code;
};

// ------------------------------------------------------------------------------------------------ scoring
/// AffineScoring, alignment/scoring_functions.rs:65-113
struct AffineScoring {
    double match_score, mismatch_score, special_character_score, gap_open, gap_extend, final_gap_multiplier;
    static AffineScoring default_dna() { return {5.0, -4.0, 4.0, -10.0, -0.5, 0.5}; }            // :77-86
    static AffineScoring align_reads_default() { return {10.0, -9.0, 9.0, -20.0, -2.0, 1.0}; }   // alignment_functions.rs:104-111
    static AffineScoring merger_default() { return {10.0, -5.0, 8.0, -15.0, -1.0, 0.25}; }       // merger.rs:130-139
    /// match_mismatch, :100-102 (host mirror, used by from_match_segment only; the DP runs on the GPU)
    double match_mismatch(uint8_t a, uint8_t b) const;
    /// exact scaled-integer form for the kernels; throws CLQ_SCORING_NOT_REPRESENTABLE (no CPU fallback exists)
    clq_affine_t to_int() const;
};

/// rust_bio_alignment's hard-coded scoring (alignment_functions.rs:48-61): 1 / -1 (a read 'N' matches anything), gap open -5,
/// extend -1.  Selects CLQ_RUSTBIO, rust-bio `Aligner::global` semantics -- PARITY UNPINNED (un-vendored `bio = "*"`).
struct RustBioScoring {
    int32_t match_score = 1, mismatch_score = -1, gap_open = -5, gap_extend = -1;
    clq_affine_t to_int() const;
};

/// ConvexScoring, alignment/scoring_functions.rs:36-53 -- only a gap *function* in the reference (nothing calls it)
struct ConvexScoring {
    double match_score, mismatch_score, gap_score, gap_open, gap_extend;
    double match_mismatch(uint8_t a, uint8_t b) const { return a == b ? match_score : mismatch_score; }
    double gap(size_t length) const;  // gap_open + log10(length); gap(0) = -inf
};

// ------------------------------------------------------------------------------------------------ results
/// AlignmentTag, alignment/alignment_matrix.rs:58-120 (the op codes are the BAM codes the C ABI uses)
struct AlignmentTag {
    enum Kind : uint8_t { MatchMismatch = CLQ_OP_M, Ins = CLQ_OP_I, Del = CLQ_OP_D, SoftClip = 4, HardClip = 5, InversionOpen = 14, InversionClose = 15 };
    Kind kind;
    size_t len;
    bool operator==(const AlignmentTag& o) const { return kind == o.kind && len == o.len; }
    char op_char() const;
};

std::string cigar_to_string(const std::vector<AlignmentTag>& cigar);
/// simplify_cigar_string, alignment_manager.rs:386-423 (twin: alignment_functions.rs:874-911): merges adjacent M / D / I
/// runs, never soft clips; inversion tags throw when doubled as the reference panics.
std::vector<AlignmentTag> simplify_cigar_string(const std::vector<AlignmentTag>& cigar_tokens);

/// AlignmentLocation, alignment/alignment_matrix.rs:687-690
struct AlignmentLocation {
    size_t x, y;
    bool operator==(const AlignmentLocation& o) const { return x == o.x && y == o.y; }
};

using TagKey = std::array<char, 2>;
using TagMap = std::map<TagKey, std::string>;

/// what AlignmentResult::to_sam_record builds (a noodles RecordBuf in the reference), alignment/alignment_matrix.rs:741-771
struct SamRecord {
    std::string name;
    uint16_t flags = 0;                 // Flags::empty()
    int32_t reference_sequence_id = 0;
    size_t alignment_start = 1;         // reference_start + 1
    std::vector<AlignmentTag> cigar;
    Bytes sequence;                     // read_aligned without '-'
    Bytes quality_scores;               // raw scores: b'H' (72) per base, whatever the input qualities were (:760-763)
    std::vector<std::pair<TagKey, std::string>> data;  // all Z-typed; extra tags in key order, then rm, rs, [ar], as
    /// one SAM text line (no newline); RNAME from `reference_names`, MAPQ 255, RNEXT *, PNEXT 0, TLEN 0, QUAL = score + 33
    std::string to_sam_line(const std::vector<std::string>& reference_names) const;
};

/// AlignmentResult, alignment/alignment_matrix.rs:694-706
struct AlignmentResult {
    std::string reference_name, read_name;
    Bytes reference_aligned, read_aligned;  // gapped, '-' = FASTA_UNSET
    std::optional<Bytes> read_quals;        // passed through
    std::vector<AlignmentTag> cigar_string;
    std::vector<AlignmentLocation> path;
    double score = 0.0;
    size_t reference_start = 0, read_start = 0;
    std::optional<std::pair<AlignmentLocation, AlignmentLocation>> bounding_box;

    /// Rebuilds the record perform_3d_global_traceback returns (alignment/alignment_matrix.rs:941-1086) from the kernels'
    /// output: gapped strings and `path` are pure functions of CIGAR + sequences (one path entry per unit op executed inside
    /// the main loop, i.e. excluding the leading boundary run emitted after it, :1054-1065).
    static AlignmentResult from_cigar(const std::string& reference_name, const std::string& read_name, const uint8_t* reference,
                                      size_t l1, const uint8_t* read, size_t l2, std::optional<Bytes> quals, const uint32_t* cigar_ops,
                                      size_t n_ops, double score);
    /// from_match_segment, :710-733
    static AlignmentResult from_match_segment(const Bytes& str1, const Bytes& str2, const std::string& reference_name,
                                              const std::string& read_name, size_t start_x, size_t start_y, const AffineScoring& af_score);
    /// to_sam_record, :741-771
    SamRecord to_sam_record(int32_t reference_id, const TagMap& extra_tags, const std::optional<std::vector<std::string>>& read_names) const;
};

/// BAM wire format of the records (what BamFileAlignmentWriter hands to noodles, alignment_manager.rs:64-209): the header text
/// ("@HD VN:1.6", one @SQ per reference in index order, "@CO Clique processed"), the binary header and records of the SAM/BAM
/// specification, BGZF framing.  Encoding is split from compression so that several host threads can prepare blocks.
namespace bam {
/// "@HD..@SQ..@CO" text of BamFileAlignmentWriter::new (:76-96)
std::string header_text(const std::vector<std::string>& reference_names, const std::vector<size_t>& reference_lengths);
/// uncompressed BAM header: magic, text, reference dictionary
void append_header(const std::vector<std::string>& reference_names, const std::vector<size_t>& reference_lengths, std::string& out);
/// one uncompressed alignment record (block_size + fields): flags 0, MAPQ 255, mate fields unset, qualities raw, tags as Z
void append_record(const SamRecord& rec, std::string& out);
/// BGZF: `data` cut into blocks of at most 65280 bytes, each a gzip member with the BC extra field; appended to `out`
void bgzf_compress(const char* data, size_t n, std::string& out, int level = 6);
/// the 28-byte end-of-file marker block
void bgzf_eof(std::string& out);
}  // namespace bam

/// f64 `Display` of Rust (what `score.to_string()` writes into the rm / rs / as tags): shortest round-trip digits, never
/// scientific notation, integral values without ".0", "NaN", "inf"
std::string f64_to_string(double v);

// ------------------------------------------------------------------------------------------------ post-alignment strings
/// get_reference_alignment_rate, consensus/consensus_builders.rs:288-307 (the `rm` tag; NaN when nothing is comparable)
double get_reference_alignment_rate(const Bytes& reference_aligned, const Bytes& read_aligned);
/// is_valid_fasta_base, utils/base_utils.rs:17-23
bool is_valid_fasta_base(uint8_t b);
/// extract_tagged_sequences, extractor.rs:271-332: digit keys '0'..'9' collect the read bytes aligned to tag columns,
/// 'A','B',.. / 'a','b',.. the reference / read bytes of each upper-case extractor region.  Zips the inputs.
std::map<uint8_t, std::string> extract_tagged_sequences(const Bytes& aligned_read, const Bytes& aligned_ref);
/// reverse_complement, utils/read_utils.rs:50-72 (upper-cases, IUPAC aware, unknown bytes unchanged)
Bytes reverse_complement(const uint8_t* dna, size_t n);
inline Bytes reverse_complement(const Bytes& dna) { return reverse_complement(dna.data(), dna.size()); }

// ------------------------------------------------------------------------------------------------ input side
/// AlignedReadOrientation / ReadPosition, read_strategies/sequence_layout.rs (the parts merge_reads_by_concatenation uses)
enum class AlignedReadOrientation { Forward, Reverse, ReverseComplement, Unknown };
struct ReadPosition {
    enum Kind { Read1, Read2, Index1, Index2, Spacer } kind;
    AlignedReadOrientation orientation = AlignedReadOrientation::Forward;
    std::string spacer_sequence;
};
struct FastqRecord { std::string id; Bytes seq, qual; };
/// ReadSetContainer, read_strategies/read_set.rs:9-14
struct ReadSetContainer {
    FastqRecord read_one;
    std::optional<FastqRecord> read_two, index_one, index_two;
};
struct MergedSequence { Bytes read_bases, read_quals; };
/// orient_sequence, merger.rs:107-126 (Unknown throws as the reference panics)
Bytes orient_sequence(const uint8_t* sequence, size_t n, AlignedReadOrientation orientation);
/// merge_reads_by_concatenation, merger.rs:40-105
MergedSequence merge_reads_by_concatenation(const ReadSetContainer& reads, const std::vector<ReadPosition>& layout);

/// phred helpers, utils/read_utils.rs:6-38 (the disagreement formula is the reference's own, kept as it is)
double phred_to_prob(uint8_t phred);
uint8_t prob_to_phred(double qual);
uint8_t combine_phred_scores(uint8_t phred_one, uint8_t phred_two, bool agree);
/// alignment_rate_and_consensus, merger.rs:428-498: consensus bases + qualities of two gapped strings; throws std::out_of_range
/// where the reference panics (a gap/gap column consumes qualities).
MergedSequence alignment_rate_and_consensus(const Bytes& alignment_1, const Bytes& qual_scores1, const Bytes& alignment_2,
                                            const Bytes& qual_scores2);

// ------------------------------------------------------------------------------------------------ strand orientation
/// MatchedPosition / SharedSegments, alignment/alignment_matrix.rs (the fields find_greedy_non_overlapping_segments fills)
struct MatchedPosition { size_t search_start, ref_start, length; };
struct SharedSegments { size_t start_position; std::vector<MatchedPosition> alignment_segments; };
/// SuffixTableLookup, reference/fasta_reference.rs:34-38 + ReferenceManager::find_seeds (:155-157): the suffix array of the
/// reference (what the `suffix` crate's SuffixTable holds) and the seed size
struct SuffixTableLookup {
    Bytes text;
    std::vector<uint32_t> table;  // suffix start positions in lexicographic order of the suffixes
    size_t seed_size = 0;
    static SuffixTableLookup find_seeds(const Bytes& reference, size_t seed_size);
    /// SuffixTable::positions: every occurrence of `query`, in suffix-array order
    std::pair<const uint32_t*, const uint32_t*> positions(const uint8_t* query, size_t n) const;
};
/// extend_hit, linked_alignment.rs:341-362 (plain bases equal up to case; anything else ends the hit)
size_t extend_hit(const uint8_t* search, size_t n, size_t search_location, const uint8_t* reference, size_t m, size_t reference_location);
/// find_greedy_non_overlapping_segments, linked_alignment.rs:97-130 (kept as it is, including the cursor that moves inside the
/// candidate loop)
SharedSegments find_greedy_non_overlapping_segments(const Bytes& search_string, const Bytes& reference, const SuffixTableLookup& seeds);
/// bio::alphabets::dna::revcomp (complement table over ACGT + IUPAC in both cases, other bytes unchanged)
Bytes bio_revcomp(const Bytes& text);
/// orient_by_longest_segment, linked_alignment.rs:24-32: (forward shares strictly more bases, forward segments, reverse segments)
struct Orientation { bool forward; SharedSegments fwd, rev; };
Orientation orient_by_longest_segment(const Bytes& search_string, const Bytes& reference, const SuffixTableLookup& seeds);

// ------------------------------------------------------------------------------------------------ references
/// Reference, reference/fasta_reference.rs:41-46
struct Reference { Bytes sequence; Bytes name; };

/// ReferenceManager, reference/fasta_reference.rs:64-146.  `references` keeps insertion order = ascending index, the
/// canonical iteration order of this build (the reference's HashMap order is random per process).
class ReferenceManager {
public:
    ReferenceManager() = default;
    ReferenceManager(std::vector<Reference> refs, size_t kmer_size = 8, size_t kmer_skip = 4);
    static ReferenceManager from_fa_file(const std::string& path, size_t kmer_size = 8, size_t kmer_skip = 4);  // :127-130
    std::vector<Reference> references;
    std::map<Bytes, size_t> reference_name_to_ref;
    size_t kmer_size = 8, kmer_skip = 4, longest_ref = 0;
    std::vector<std::string> names() const;
};

/// AlignmentWithRef, alignment_functions.rs:451-456
struct AlignmentWithRef {
    std::optional<AlignmentResult> alignment;
    Bytes ref_name, ref_sequence;
};

// ------------------------------------------------------------------------------------------------ batches
/// Page-locked staging buffer for one batch of reads (clq_host_alloc): the batch loop fills it in place, the H2D copy is a
/// true asynchronous DMA from it.  read i = bytes[off[i] .. off[i+1]).
class ReadBatch {
public:
    ReadBatch(uint32_t max_reads, uint64_t max_bytes);
    ~ReadBatch();
    ReadBatch(const ReadBatch&) = delete;
    ReadBatch& operator=(const ReadBatch&) = delete;
    void clear();
    /// false when the read does not fit any more (the caller keeps it for the next batch)
    bool push(const std::string& name, const uint8_t* seq, size_t n, const uint8_t* qual = nullptr, int32_t fixed_ref = -1) {
        return push(name.data(), name.size(), seq, n, qual, fixed_ref);
    }
    /// allocation-free form (names live in one flat buffer): the per-read cost is two memcpys
    bool push(const char* name, size_t name_len, const uint8_t* seq, size_t n, const uint8_t* qual = nullptr, int32_t fixed_ref = -1);
    /// replaces the batch by reads [lo, hi) of a caller-owned span (bytes + offsets, any host memory): one memcpy of the bytes,
    /// offsets rebased, unnamed reads, no qualities.  false when the range exceeds the batch's capacity.
    bool assign_span(const uint8_t* bytes, const uint64_t* off, uint64_t lo, uint64_t hi, const int32_t* fixed_ref);
    /// appends reads [lo, hi) of the same span behind the ones assign_span put there (a batch made of two ranges of the input:
    /// reads 0 .. n_first-1 are input reads first_index + i, the others second_index + (i - n_first)).  false when it does not fit.
    bool append_span(const uint8_t* bytes, const uint64_t* off, uint64_t lo, uint64_t hi, const int32_t* fixed_ref);
    uint32_t size() const { return n_; }
    uint32_t capacity() const { return max_reads_; }
    uint64_t byte_capacity() const { return max_bytes_; }
    const uint8_t* read(uint32_t i) const { return bytes_ + off_[i]; }
    uint8_t* read_mut(uint32_t i) { return bytes_ + off_[i]; }  // in-place re-orientation (same length) before the batch is submitted
    size_t read_len(uint32_t i) const { return (size_t)(off_[i + 1] - off_[i]); }
    std::string name(uint32_t i) const { return std::string(name_bytes_.data() + name_off_[i], name_off_[i + 1] - name_off_[i]); }
    const char* name_data(uint32_t i) const { return name_bytes_.data() + name_off_[i]; }
    size_t name_len(uint32_t i) const { return name_off_[i + 1] - name_off_[i]; }
    std::optional<Bytes> quals(uint32_t i) const;
    const uint8_t* bytes() const { return bytes_; }
    const uint64_t* offsets() const { return off_; }
    const int32_t* fixed_ref() const { return fixed_; }
    int32_t* fixed_ref_mut() { return fixed_; }
    /// 2-bit form of the staged bytes (clq_pack2 into a page-locked word buffer, allocated on first use) for
    /// clq_submit_packed2: a quarter of the H2D bytes; the ASCII bytes stay for the sinks (SAM SEQ, tags).  Call it once the
    /// batch is complete (after any read_mut); clear() / push / assign_span drop it.
    void pack2() const;
    bool packed() const { return packed_; }
    const uint32_t* packed_words() const { return words_; }
    const std::vector<uint64_t>& exception_positions() const { return exc_pos_; }
    const std::vector<uint8_t>& exception_bytes() const { return exc_byte_; }
    uint64_t first_index = 0;  // index of read 0 in the whole input
    uint64_t second_index = 0; // index of read n_first in the whole input (batches of two ranges, see append_span)
    uint32_t n_first = 0xffffffffu;  // reads of the first range (0xffffffff: the whole batch is one range)

private:
    uint32_t max_reads_, n_ = 0;
    uint64_t max_bytes_;
    uint8_t* bytes_ = nullptr;
    uint64_t* off_ = nullptr;
    int32_t* fixed_ = nullptr;
    std::vector<char> name_bytes_;
    std::vector<size_t> name_off_;
    Bytes quals_;
    bool have_quals_ = false;
    mutable uint32_t* words_ = nullptr;
    mutable std::vector<uint64_t> exc_pos_;
    mutable std::vector<uint8_t> exc_byte_;
    mutable bool packed_ = false;
};

class Aligner;

/// Results of one batch, valid during the sink callback (views over the aligner's pinned result buffers)
class BatchView {
public:
    uint32_t size() const { return batch->size(); }
    const ReadBatch* batch = nullptr;
    const ReferenceManager* rm = nullptr;
    const clq_result_t* results = nullptr;
    const uint32_t* cigar_pool = nullptr;
    const uint8_t* tags = nullptr;  // CLQ_EXTRACT_TAGS output, tag_stride bytes per read (nullptr when not requested)
    uint32_t tag_stride = 0;
    int32_t scale = 1;
    int device = 0;
    bool rust_bio = false;  // CLQ_RUSTBIO batch: the reference reports score 0.0 and an empty path for these records
    uint64_t cigar_used = 0;  // ops of cigar_pool this batch filled

    uint32_t status(uint32_t i) const { return results[i].status; }
    double score(uint32_t i) const { return (double)results[i].score_scaled / (double)scale; }
    uint32_t ref_index(uint32_t i) const { return results[i].ref_index; }
    const uint32_t* cigar(uint32_t i) const { return cigar_pool + results[i].cigar_off; }
    uint32_t cigar_len(uint32_t i) const { return results[i].cigar_len; }
    std::string cigar_string(uint32_t i) const;
    /// the `rm` tag from the counters the traceback walk kept (get_reference_alignment_rate fused on the GPU)
    double alignment_rate(uint32_t i) const;
    /// extract_tagged_sequences' digit keys, rebuilt from the per-column bytes the walk recorded
    std::map<uint8_t, std::string> digit_tags(uint32_t i) const;
    /// the owned record of the reference surface; nullopt for dropped reads (too long / no candidate / diverged traceback)
    std::optional<AlignmentWithRef> alignment(uint32_t i) const;
    /// the tag set align_reads writes (alignment_functions.rs:193-226): e<symbol> for the symbols in `umi_symbols`,
    /// rc = "1", ar = read name, rm, as
    TagMap align_reads_tags(uint32_t i, const std::string& umi_symbols) const;
    /// Fast path of alignment(i) -> to_sam_record(ref_index, align_reads_tags(i, umi_symbols), None) -> to_sam_line: the same
    /// text (plus '\n') appended to `out` straight from the raw records, without building the intermediate objects.
    /// Returns false (nothing appended) for dropped reads.
    bool append_sam_line(uint32_t i, const std::string& umi_symbols, const std::vector<std::string>& reference_names, std::string& out) const;
    /// the same record as one uncompressed BAM alignment block (== bam::append_record(to_sam_record(..)) of the object path)
    bool append_bam_record(uint32_t i, const std::string& umi_symbols, std::string& out) const;
};

struct AlignerOptions {
    int device = 0;
    uint32_t max_reads = 1u << 18;      // per batch / stream slot
    uint64_t max_read_bytes = 0;        // 0: max_reads * 512
    uint32_t max_read_len = 1u << 16;   // align_reads' drop rule: reads >= this are dropped (alignment_functions.rs:147)
    uint32_t max_refs = 4096;
    uint64_t max_ref_bytes = 1u << 26;
    uint32_t cigar_ops_per_read = 32;
    uint32_t n_slots = 2;               // stream slots: batch k+1 uploads while batch k computes
    bool pack2_upload = false;          // ship batches 2-bit packed (ReadBatch::pack2 + clq_submit_packed2): 0.25 B per base on the
                                        // wire for one more host pass over the staged bytes; results are identical
};

/// the source side of the batch loop: fill `batch` (already cleared), return false once the input is exhausted
using ReadSource = std::function<bool(ReadBatch& batch)>;
using ResultSink = std::function<void(const BatchView& view)>;

struct AlignReadsStats {
    uint64_t reads = 0, aligned = 0, dropped = 0, batches = 0, cells = 0;
    double seconds = 0.0;        // the loop itself: first batch filled -> last batch handed to the sink
    double setup_seconds = 0.0;  // one-time page-locking of the staging buffers (first call on an Aligner)
};

/// One clq_ctx on one GPU (reference set, stream slots, pinned result buffers).  Not thread-safe; one per device / thread.
class Aligner {
public:
    explicit Aligner(const AlignerOptions& opt = AlignerOptions());
    ~Aligner();
    Aligner(const Aligner&) = delete;
    Aligner& operator=(const Aligner&) = delete;

    void set_references(const ReferenceManager& rm, bool build_kmer_index = true);
    const ReferenceManager& references() const { return rm_; }
    const AlignerOptions& options() const { return opt_; }

    // ---- the reference's call surface (single pair / single read) ----
    /// align_two_strings, alignment_manager.rs:231-273: fresh matrix, bandwidth = max(L1, L2).  `local` throws (outside the path).
    AlignmentResult align_two_strings(const Bytes& reference_sequence, const Bytes& read_sequence, std::optional<Bytes> read_qual,
                                      const AffineScoring& scoring_function, bool local, const std::string& ref_name,
                                      const std::string& read_name);
    /// align_two_strings_passed_matrix, alignment_functions.rs:383-449; the caller-owned matrix is gone (scores live in
    /// registers); `max_indel` is the bandwidth: the read length and max(L1, L2) take the fast kernels, any other value the
    /// explicit-band path (CLQ_BAND_K).
    AlignmentResult align_two_strings_passed_matrix(const std::string& ref_name, const std::string& read_name, const Bytes& reference,
                                                    const Bytes& read, std::optional<Bytes> qual, const AffineScoring& scoring,
                                                    size_t max_indel);
    /// exhaustive_alignment_search, alignment_functions.rs:769-827 (ascending index, last maximum wins)
    std::optional<AlignmentWithRef> exhaustive_alignment_search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                                const AffineScoring& scoring);
    /// quick_alignment_search, alignment_functions.rs:693-767
    std::optional<AlignmentWithRef> quick_alignment_search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                           const AffineScoring& scoring, double match_threshold = 0.90);
    /// align_to_reference_choices, alignment_functions.rs:520-631: 0 references -> nullopt; > 1 -> quick (fast_lookup) or
    /// exhaustive search; 1 -> with `rust_bio` what the reference does today (:544-603: rust-bio global, 1/-1/-5/-1, score
    /// reported as 0.0, empty path; PARITY UNPINNED), otherwise clique's own Gotoh with bandwidth = read.len() (the call the
    /// reference has commented out at :586-597; pinned on its goldens).  `known_strand = false` (1 reference only, :549-558):
    /// the read is oriented by orient_by_longest_segment first and reverse-complemented when the reverse strand shares more.
    std::optional<AlignmentWithRef> align_to_reference_choices(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                               bool fast_lookup, const AffineScoring& scoring, bool rust_bio = false,
                                                               bool known_strand = true);

    // ---- paired-read merging (MergeStrategy::Align) ----
    /// merge_reads_by_alignment, merger.rs:348-396: align_two_strings(read1, revcomp(read2)) + alignment_rate_and_consensus
    MergedSequence merge_reads_by_alignment(const FastqRecord& read1, const FastqRecord& read2, const AffineScoring& merge_initial_scoring);
    /// the same for a whole batch in one launch: every pair is its own (reference = read1, read = revcomp(read2)) task; the
    /// caller's reference set is restored afterwards.  Entries whose alignment the reference could not finish are nullopt.
    std::vector<std::optional<MergedSequence>> merge_read_pairs_by_alignment(const std::vector<ReadSetContainer>& pairs,
                                                                             const AffineScoring& merge_initial_scoring);

    // ---- the batch loop ----
    /// align_reads' par_bridge loop (alignment_functions.rs:135-249) up to the writer: drains `source` into pinned batches,
    /// keeps every stream slot busy, hands each finished batch to `sink`.
    /// `rust_bio`: single-reference panels take the rust-bio branch (see align_to_reference_choices); records then carry score 0.
    /// `known_strand = false` (single-reference panels only, alignment_functions.rs:549-558): every read is oriented on the host
    /// by orient_by_longest_segment and reverse-complemented in the pinned batch when the reverse strand shares more bases.
    AlignReadsStats align_reads(const ReadSource& source, const AffineScoring& scoring, bool fast_lookup, const ResultSink& sink,
                                bool extract_tags = true, bool rust_bio = false, bool known_strand = true);

    // ---- raw slot interface (used by ShardedAligner and the bench driver) ----
    uint32_t search_flags(bool fast_lookup) const;
    void submit(int slot, const ReadBatch& batch, const clq_affine_t& scoring, uint32_t flags, double threshold = 0.90);
    BatchView wait(int slot, const ReadBatch& batch);
    clq_stats_t stats(int slot);
    clq_ctx* ctx() { return ctx_; }

private:
    void check(int32_t rc, const char* what) const;
    AlignmentResult single(const Bytes& reference, const Bytes& read, std::optional<Bytes> qual, const AffineScoring& sc, uint32_t band,
                           const std::string& ref_name, const std::string& read_name);
    std::optional<AlignmentWithRef> search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                           const AffineScoring& sc, uint32_t search_mode, double threshold);
    AlignerOptions opt_;
    clq_ctx* ctx_ = nullptr;
    ReferenceManager rm_;
    std::vector<clq_result_t*> res_;
    std::vector<uint32_t*> pool_;
    std::vector<uint8_t*> tags_;
    std::vector<uint32_t> flags_;
    std::vector<int32_t> scale_;
    uint64_t pool_ops_ = 0;
    std::vector<uint64_t> tags_cap_;
    std::unique_ptr<ReadBatch> one_;  // staging for the single-read calls
    std::vector<std::unique_ptr<ReadBatch>> bufs_;  // pinned staging of the batch loop, one per stream slot (allocated once)
};

/// A contiguous span of reads already in host memory -- what a FASTQ block decoder (or the stage before align_reads) hands
/// over: read i = bytes[off[i] .. off[i+1]).  The memory need not be page-locked.  fixed_ref (nullable): per-read reference for
/// multi-reference inputs whose assignment is known; single-reference panels need none.
struct ReadSpan {
    const uint8_t* bytes = nullptr;
    const uint64_t* off = nullptr;
    uint64_t n = 0;
    const int32_t* fixed_ref = nullptr;
    /// how align_reads_span cuts batches from the span.  Auto: front to back, unless the reads at one end are much longer than
    /// at the other (a length-sorted stream); then every batch takes its long reads from that end and fills up with short ones
    /// from the other, so that each launch holds the mix of long and short pairs the persistent grid needs to stay full.
    enum class Order { Auto, Front, LongestFirst, TwoEnded };
    Order order = Order::Auto;
};

/// The batch cutter of ShardedAligner::align_reads_span (thread-safe; pure host logic, see clique_host.cpp).  A claim is one range
/// [lo, hi) of the span plus, for a two-ended claim on a length-sorted stream, a second range [lo2, hi2) (empty otherwise).
class SpanClaimer {
public:
    SpanClaimer(const uint64_t* off, uint64_t n, size_t n_devices, int claimers_per_device, uint64_t max_reads, uint64_t max_read_bytes,
                ReadSpan::Order order = ReadSpan::Order::Auto);
    bool claim(uint64_t& lo, uint64_t& hi, uint64_t& lo2, uint64_t& hi2);
    ReadSpan::Order order() const { return order_; }  // what Auto resolved to

private:
    uint64_t fit_front(uint64_t bytes_budget, uint64_t reads_budget) const;
    uint64_t fit_back(uint64_t bytes_budget, uint64_t reads_budget) const;
    const uint64_t* off_;
    size_t nd_;
    int nf_;
    uint64_t max_reads_, max_bytes_;
    uint64_t front_ = 0, back_, n_claims_ = 0;
    ReadSpan::Order order_;
    bool long_at_back_ = true;
    std::mutex mu_;
};

/// Caller-owned result arrays of ShardedAligner::align_reads_span, in input order: results[i] for read i (cigar_off indexes
/// cigar_pool), tags (nullable) n * tag_stride bytes.
struct SpanOutput {
    clq_result_t* results = nullptr;
    uint32_t* cigar_pool = nullptr;
    uint64_t cigar_cap = 0;
    uint64_t cigar_used = 0;      // out
    uint8_t* tags = nullptr;      // optional, n * tag_stride bytes (tag_stride from the reference set)
    uint32_t tag_stride = 0;      // out
    int32_t scale = 1;            // out: score = score_scaled / scale
};

struct SpanStats {
    AlignReadsStats total;
    std::vector<double> device_kernel_ms;     // sum of the kernels' device time per GPU (CUDA events)
    std::vector<uint64_t> device_reads, device_cells, device_batches;
    double fill_seconds = 0.0;                // sum over filler threads: staging copies into pinned batches
    double sink_seconds = 0.0;                // sum over device threads: results copied out to the caller's arrays
};

/// Read-sharded dispatcher over the GPUs of one box: one Aligner + one host thread per device pull
/// batches from the shared source; no collective, no peer traffic.  The sink is called under a mutex (the reference's
/// writer is behind Arc<Mutex<..>> too) in completion order; BatchView::batch->first_index places a batch in the input.
class ShardedAligner {
public:
    ShardedAligner(const std::vector<int>& devices, AlignerOptions opt = AlignerOptions());
    void set_references(const ReferenceManager& rm, bool build_kmer_index = true);
    AlignReadsStats align_reads(const ReadSource& source, const AffineScoring& scoring, bool fast_lookup, const ResultSink& sink,
                                bool extract_tags = true, bool rust_bio = false, bool known_strand = true);
    /// The same loop fed from an in-memory span (the analogue of `read_iterator.par_bridge().for_each`, alignment_functions.rs:135,
    /// for input that is already decoded): batches are cut dynamically from one shared cursor (shrinking towards the end of
    /// the input so that the GPUs finish together, whatever the read-length order), `fillers_per_device` host threads per GPU
    /// stage them into page-locked batches (the only copy of the read bytes on the host), one thread per GPU submits / waits
    /// and writes the records into `out` in input order.  No collective, no peer traffic.
    SpanStats align_reads_span(const ReadSpan& span, const AffineScoring& scoring, bool fast_lookup, SpanOutput& out,
                               int fillers_per_device = 2, bool extract_tags = false, bool rust_bio = false);
    size_t n_devices() const { return aligners_.size(); }
    Aligner& aligner(size_t k) { return *aligners_[k]; }

private:
    std::vector<std::unique_ptr<Aligner>> aligners_;
    std::vector<std::vector<std::unique_ptr<ReadBatch>>> span_bufs_;  // page-locked staging of align_reads_span, per device (allocated once)
};

}  // namespace clique

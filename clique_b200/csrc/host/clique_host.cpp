// clique_host.cpp -- implementation of include/clique_host.hpp over the libclq C ABI.  No alignment arithmetic here.
#include "../../../include/clique_host.hpp"

#include <algorithm>
#include <charconv>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <thread>

#include <zlib.h>

namespace clique {

namespace {
[[noreturn]] void fail(int32_t code, const std::string& msg) { throw ClqError(code, "libclq error " + std::to_string(code) + ": " + msg); }

template <class T>
T* pinned(size_t n) {
    void* p = nullptr;
    const int32_t rc = clq_host_alloc(std::max<size_t>(n, 1) * sizeof(T), &p);
    if (rc != CLQ_OK) fail(rc, "clq_host_alloc(" + std::to_string(n * sizeof(T)) + " bytes): no CUDA device? libclq has no CPU fallback");
    return static_cast<T*>(p);
}
}  // namespace

// ------------------------------------------------------------------------------------------------ scoring
double AffineScoring::match_mismatch(uint8_t a, uint8_t b) const {
    // alignment/scoring_functions.rs:100-102: N or any byte below ':' on either side is "special"
    if (a == 'N' || b == 'N' || a < 58 || b < 58) return special_character_score;
    return a == b ? match_score : mismatch_score;
}

clq_affine_t AffineScoring::to_int() const {
    clq_affine_t out;
    const int32_t rc = clq_affine_from_f64(match_score, mismatch_score, special_character_score, gap_open, gap_extend, final_gap_multiplier, &out);
    if (rc != CLQ_OK) fail(rc, "scoring has no exact scaled-integer form (or gap_open >= 0)");
    return out;
}

clq_affine_t RustBioScoring::to_int() const {
    clq_affine_t out;
    const int32_t rc = clq_rustbio_scoring(match_score, mismatch_score, gap_open, gap_extend, &out);
    if (rc != CLQ_OK) fail(rc, "rust-bio scoring not representable");
    return out;
}

double ConvexScoring::gap(size_t length) const {
    return gap_open + (length > 0 ? std::log10((double)length) : -std::numeric_limits<double>::infinity());
}

// ------------------------------------------------------------------------------------------------ cigar
char AlignmentTag::op_char() const {
    switch (kind) {
        case MatchMismatch: return 'M';
        case Ins: return 'I';
        case Del: return 'D';
        case SoftClip: return 'S';
        case HardClip: return 'H';
        case InversionOpen: return '<';
        default: return '>';
    }
}

std::string cigar_to_string(const std::vector<AlignmentTag>& cigar) {
    std::string s;
    for (const auto& t : cigar) {
        if (t.kind != AlignmentTag::InversionOpen && t.kind != AlignmentTag::InversionClose) s += std::to_string(t.len);
        s += t.op_char();
    }
    return s;
}

std::vector<AlignmentTag> simplify_cigar_string(const std::vector<AlignmentTag>& cigar_tokens) {
    std::vector<AlignmentTag> out;
    std::optional<AlignmentTag> last;
    for (const auto& tok : cigar_tokens) {
        if (!last) { last = tok; continue; }
        if (last->kind == tok.kind && tok.kind == AlignmentTag::InversionOpen) throw std::logic_error("Cannot have two inversion open tags in a row");
        if (last->kind == tok.kind && tok.kind == AlignmentTag::InversionClose) throw std::logic_error("Cannot have two inversion closed tags in a row");
        const bool mergeable = tok.kind == AlignmentTag::MatchMismatch || tok.kind == AlignmentTag::Del || tok.kind == AlignmentTag::Ins;
        if (mergeable && last->kind == tok.kind) last->len += tok.len;
        else { out.push_back(*last); last = tok; }
    }
    if (last) out.push_back(*last);
    return out;
}

// ------------------------------------------------------------------------------------------------ AlignmentResult
AlignmentResult AlignmentResult::from_cigar(const std::string& reference_name, const std::string& read_name, const uint8_t* reference,
                                            size_t l1, const uint8_t* read, size_t l2, std::optional<Bytes> quals, const uint32_t* ops,
                                            size_t n_ops, double score) {
    AlignmentResult r;
    r.reference_name = reference_name;
    r.read_name = read_name;
    r.read_quals = std::move(quals);
    r.score = score;
    r.reference_aligned.reserve(l1 + l2);
    r.read_aligned.reserve(l1 + l2);
    size_t x = 0, y = 0;
    for (size_t k = 0; k < n_ops; k++) {
        const size_t n = ops[k] >> 4;
        const uint32_t c = ops[k] & 15u;
        r.cigar_string.push_back({(AlignmentTag::Kind)c, n});
        // the leading boundary run (emitted after the main loop of the traceback) has no `path` entries
        const bool boundary = k == 0 && c != CLQ_OP_M;
        if (c == CLQ_OP_M) {
            if (x + n > l1 || y + n > l2) fail(CLQ_E_INVALID, "CIGAR overruns the sequences");
            r.reference_aligned.insert(r.reference_aligned.end(), reference + x, reference + x + n);
            r.read_aligned.insert(r.read_aligned.end(), read + y, read + y + n);
            for (size_t i = 0; i < n; i++) r.path.push_back({x + i + 1, y + i + 1});
            x += n; y += n;
        } else if (c == CLQ_OP_D) {
            if (x + n > l1) fail(CLQ_E_INVALID, "CIGAR overruns the reference");
            r.reference_aligned.insert(r.reference_aligned.end(), reference + x, reference + x + n);
            r.read_aligned.insert(r.read_aligned.end(), n, (uint8_t)'-');
            if (!boundary) for (size_t i = 0; i < n; i++) r.path.push_back({x + i + 1, y});
            x += n;
        } else {
            if (y + n > l2) fail(CLQ_E_INVALID, "CIGAR overruns the read");
            r.reference_aligned.insert(r.reference_aligned.end(), n, (uint8_t)'-');
            r.read_aligned.insert(r.read_aligned.end(), read + y, read + y + n);
            if (!boundary) for (size_t i = 0; i < n; i++) r.path.push_back({x, y + i + 1});
            y += n;
        }
    }
    return r;
}

AlignmentResult AlignmentResult::from_match_segment(const Bytes& str1, const Bytes& str2, const std::string& reference_name,
                                                    const std::string& read_name, size_t start_x, size_t start_y, const AffineScoring& af) {
    AlignmentResult r;
    r.reference_name = reference_name;
    r.read_name = read_name;
    r.reference_aligned = str1;
    r.read_aligned = str2;
    r.cigar_string = {{AlignmentTag::MatchMismatch, str1.size()}};
    for (size_t i = 0; i < str1.size(); i++) r.path.push_back({start_x + i, start_y + i});
    r.score = 0.0;
    for (size_t i = 0; i < std::min(str1.size(), str2.size()); i++) r.score += af.match_mismatch(str1[i], str2[i]);
    r.reference_start = start_x;
    r.read_start = start_y;
    return r;
}

std::string f64_to_string(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[400];
    const auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, res.ptr);
}

SamRecord AlignmentResult::to_sam_record(int32_t reference_id, const TagMap& extra_tags, const std::optional<std::vector<std::string>>& read_names) const {
    SamRecord rec;
    rec.name = read_name;
    rec.reference_sequence_id = reference_id;
    rec.alignment_start = reference_start + 1;
    rec.cigar = cigar_string;
    for (uint8_t b : read_aligned) if (b != '-') rec.sequence.push_back(b);
    rec.quality_scores.assign(rec.sequence.size(), (uint8_t)'H');
    // Data::insert replaces the value of an existing tag in place and appends new ones (extra tags first; the reference
    // iterates a HashMap there, this build takes key order)
    auto put = [&](TagKey k, std::string v) {
        for (auto& kv : rec.data) if (kv.first == k) { kv.second = std::move(v); return; }
        rec.data.emplace_back(k, std::move(v));
    };
    for (const auto& kv : extra_tags) put(kv.first, kv.second);
    put({'r', 'm'}, f64_to_string(get_reference_alignment_rate(reference_aligned, read_aligned)));
    put({'r', 's'}, f64_to_string(score));
    if (read_names) {
        std::string joined;
        for (size_t i = 0; i < read_names->size(); i++) joined += (i ? "," : "") + (*read_names)[i];
        put({'a', 'r'}, joined);
    }
    put({'a', 's'}, f64_to_string(score));
    return rec;
}

std::string SamRecord::to_sam_line(const std::vector<std::string>& reference_names) const {
    std::string s = name.empty() ? "*" : name;
    s += '\t' + std::to_string(flags) + '\t';
    s += (reference_sequence_id >= 0 && (size_t)reference_sequence_id < reference_names.size()) ? reference_names[reference_sequence_id] : "*";
    s += '\t' + std::to_string(alignment_start) + "\t255\t";
    s += cigar.empty() ? "*" : cigar_to_string(cigar);
    s += "\t*\t0\t0\t";
    s += sequence.empty() ? "*" : std::string(sequence.begin(), sequence.end());
    s += '\t';
    if (quality_scores.empty()) s += '*';
    else for (uint8_t q : quality_scores) s += (char)(q + 33);
    for (const auto& kv : data) {
        s += '\t';
        s += kv.first[0]; s += kv.first[1];
        s += ":Z:" + kv.second;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------ strings
double get_reference_alignment_rate(const Bytes& reference, const Bytes& read) {
    uint64_t m = 0, mm = 0;
    for (size_t i = 0; i < reference.size(); i++) {
        const uint8_t a = reference[i], b = read.at(i);  // the reference unwraps read.get(index): shorter read panics
        if (a > 64 && a != 'N' && b > 64) { if (a == b) m++; else mm++; }
    }
    return (double)m / (double)(m + mm);
}

bool is_valid_fasta_base(uint8_t b) {
    if (b >= 'a' && b <= 'z') b -= 32;
    switch (b) {
        case 'A': case 'C': case 'G': case 'T': case 'U': case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M':
        case 'B': case 'D': case 'H': case 'V': case 'N': return true;
        default: return false;
    }
}

std::map<uint8_t, std::string> extract_tagged_sequences(const Bytes& aligned_read, const Bytes& aligned_ref) {
    std::map<uint8_t, std::string> out;
    bool in_extractor = false;
    uint8_t next_read = 'a', next_ref = 'A';
    const size_t n = std::min(aligned_read.size(), aligned_ref.size());
    for (size_t i = 0; i < n; i++) {
        const uint8_t rb = aligned_ref[i], qb = aligned_read[i];
        const bool valid = is_valid_fasta_base(rb);
        const bool upper = (rb >= 'A' && rb <= 'Z') || (rb == '-' && in_extractor);
        const bool special = rb >= '0' && rb <= '9';
        if (upper) {
            in_extractor = true;
            out[next_ref] += (char)rb;
            out[next_read] += (char)qb;
        } else if (!valid && !in_extractor && special) {
            out[rb] += (char)qb;
        } else if (!valid && in_extractor && special) {
            out[next_ref] += (char)rb;
            out[next_read] += (char)qb;
            out[rb] += (char)qb;
        } else {
            if (in_extractor) { next_read++; next_ref++; }
            in_extractor = false;
        }
    }
    return out;
}

Bytes reverse_complement(const uint8_t* dna, size_t n) {
    Bytes out(n);
    for (size_t i = 0; i < n; i++) {
        uint8_t b = dna[n - 1 - i];
        if (b >= 'a' && b <= 'z') b -= 32;
        switch (b) {
            case 'A': b = 'T'; break; case 'T': b = 'A'; break; case 'G': b = 'C'; break; case 'C': b = 'G'; break;
            case 'R': b = 'Y'; break; case 'Y': b = 'R'; break; case 'K': b = 'M'; break; case 'M': b = 'K'; break;
            case 'B': b = 'V'; break; case 'V': b = 'B'; break; case 'D': b = 'H'; break; case 'H': b = 'D'; break;
            default: break;
        }
        out[i] = b;
    }
    return out;
}

Bytes orient_sequence(const uint8_t* sequence, size_t n, AlignedReadOrientation o) {
    switch (o) {
        case AlignedReadOrientation::Forward: return Bytes(sequence, sequence + n);
        case AlignedReadOrientation::Reverse: { Bytes b(sequence, sequence + n); std::reverse(b.begin(), b.end()); return b; }
        case AlignedReadOrientation::ReverseComplement: return reverse_complement(sequence, n);
        default: throw std::logic_error("We can't merge reads when the orientation is marked 'Unknown' in the yaml specification file");
    }
}

MergedSequence merge_reads_by_concatenation(const ReadSetContainer& reads, const std::vector<ReadPosition>& layout) {
    MergedSequence m;
    auto add = [&](const FastqRecord& r, AlignedReadOrientation o) {
        const Bytes s = orient_sequence(r.seq.data(), r.seq.size(), o);
        m.read_bases.insert(m.read_bases.end(), s.begin(), s.end());
        m.read_quals.insert(m.read_quals.end(), r.qual.begin(), r.qual.end());  // qualities are appended as they are (merger.rs:55)
    };
    auto need = [](const std::optional<FastqRecord>& r, const char* what) -> const FastqRecord& {
        if (!r) throw std::logic_error(std::string("read layout names ") + what + " but the read set has none");
        return *r;
    };
    for (const auto& pos : layout) {
        switch (pos.kind) {
            case ReadPosition::Read1: add(reads.read_one, pos.orientation); break;
            case ReadPosition::Read2: add(need(reads.read_two, "read2"), pos.orientation); break;
            case ReadPosition::Index1: add(need(reads.index_one, "index1"), pos.orientation); break;
            case ReadPosition::Index2: add(need(reads.index_two, "index2"), pos.orientation); break;
            case ReadPosition::Spacer:
                m.read_bases.insert(m.read_bases.end(), pos.spacer_sequence.begin(), pos.spacer_sequence.end());
                m.read_quals.insert(m.read_quals.end(), pos.spacer_sequence.size(), (uint8_t)'H');
                break;
        }
    }
    return m;
}

double phred_to_prob(uint8_t phred) {
    const double phred_f64 = (double)((size_t)phred - 33);
    return std::pow(10.0, (-1.0 * phred_f64) / 10.0);
}

uint8_t prob_to_phred(double qual) {
    const double v = ((-10.0) * std::log10(qual)) + 33.0;  // `as u8` saturates, NaN -> 0
    if (!(v > 0.0)) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

uint8_t combine_phred_scores(uint8_t phred_one, uint8_t phred_two, bool agree) {
    const double prob1 = phred_to_prob(phred_one), prob2 = phred_to_prob(phred_two);
    return agree ? prob_to_phred(prob1 * prob2) : prob_to_phred(1.0 - ((1.0 - prob2) * (1.0 * prob1)));
}

MergedSequence alignment_rate_and_consensus(const Bytes& a1, const Bytes& q1, const Bytes& a2, const Bytes& q2) {
    if (a1.size() != a2.size()) throw std::logic_error("alignment_rate_and_consensus: the two alignment strings differ in length");
    MergedSequence m;
    m.read_bases.reserve(a1.size());
    m.read_quals.reserve(a1.size());
    size_t p1 = 0, p2 = 0;
    for (size_t i = 0; i < a1.size(); i++) {
        const uint8_t a = a1[i], b = a2[i];
        if (a == b) {  // includes gap/gap, as in the reference
            m.read_bases.push_back(a);
            m.read_quals.push_back(combine_phred_scores(q1.at(p1), q2.at(p2), true));
            p1++; p2++;
        } else if (a == '-') {
            m.read_bases.push_back(b);
            m.read_quals.push_back(q2.at(p2++));
        } else if (b == '-') {
            m.read_bases.push_back(a);
            m.read_quals.push_back(q1.at(p1++));
        } else {
            m.read_bases.push_back(q1.at(p1) >= q2.at(p2) ? a : b);  // bases disagree: the higher quality base
            m.read_quals.push_back(combine_phred_scores(q1.at(p1), q2.at(p2), false));
            p1++; p2++;
        }
    }
    return m;
}

// ------------------------------------------------------------------------------------------------ strand orientation
namespace {
inline int acgt(uint8_t b) {
    switch (b) { case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 3; case 'T': case 't': return 4; default: return 0; }
}
}  // namespace

SuffixTableLookup SuffixTableLookup::find_seeds(const Bytes& reference, size_t seed_size) {
    SuffixTableLookup s;
    s.text = reference;
    s.seed_size = seed_size;
    s.table.resize(reference.size());
    for (size_t i = 0; i < reference.size(); i++) s.table[i] = (uint32_t)i;
    const uint8_t* t = s.text.data();
    const size_t m = s.text.size();
    std::sort(s.table.begin(), s.table.end(), [t, m](uint32_t a, uint32_t b) {
        const size_t la = m - a, lb = m - b;
        const int c = std::memcmp(t + a, t + b, std::min(la, lb));
        return c ? c < 0 : la < lb;
    });
    return s;
}

std::pair<const uint32_t*, const uint32_t*> SuffixTableLookup::positions(const uint8_t* query, size_t n) const {
    const uint8_t* t = text.data();
    const size_t m = text.size();
    // first suffix >= query, then first suffix that does not start with query
    auto lo = std::partition_point(table.begin(), table.end(), [&](uint32_t p) {
        const size_t l = std::min(n, m - p);
        const int c = std::memcmp(t + p, query, l);
        return c ? c < 0 : (m - p) < n;
    });
    auto hi = std::partition_point(lo, table.end(), [&](uint32_t p) { return m - p >= n && std::memcmp(t + p, query, n) == 0; });
    const uint32_t* base = table.data();
    return {base + (lo - table.begin()), base + (hi - table.begin())};
}

size_t extend_hit(const uint8_t* search, size_t n, size_t sloc, const uint8_t* reference, size_t m, size_t rloc) {
    size_t len = 0;
    while (len + sloc < n && len + rloc < m) {
        const int a = acgt(search[sloc + len]), b = acgt(reference[rloc + len]);
        if (!a || a != b) return len;
        len++;
    }
    return len;
}

SharedSegments find_greedy_non_overlapping_segments(const Bytes& search, const Bytes& reference, const SuffixTableLookup& seeds) {
    SharedSegments out{reference.size(), {}};
    size_t position = 0, greatest_ref_pos = 0;
    while ((int64_t)position <= (int64_t)search.size() - (int64_t)seeds.seed_size) {
        const auto range = seeds.seed_size ? seeds.positions(search.data() + position, seeds.seed_size)
                                           : std::pair<const uint32_t*, const uint32_t*>{nullptr, nullptr};
        size_t longest_hit = 0;
        for (const uint32_t* it = range.first; it != range.second; ++it) {
            const size_t ref_position = *it;
            if (ref_position >= greatest_ref_pos) {
                const size_t ext = extend_hit(search.data(), search.size(), position, reference.data(), reference.size(), ref_position);
                if (ext > longest_hit) {
                    out.alignment_segments.push_back({position, ref_position, ext});
                    position += ext;
                    out.start_position = std::min(out.start_position, ref_position);
                    greatest_ref_pos = std::max(greatest_ref_pos, ref_position + ext);
                    longest_hit = ext;
                }
            }
        }
        position += 1;
    }
    return out;
}

Bytes bio_revcomp(const Bytes& text) {
    static const char from[] = "AGCTYRWSKMDVHBNagctyrwskmdvhbn";
    static const char to[] = "TCGARYWSMKHBDVNtcgarywsmkhbdvn";
    uint8_t comp[256];
    for (int i = 0; i < 256; i++) comp[i] = (uint8_t)i;
    for (int i = 0; from[i]; i++) comp[(uint8_t)from[i]] = (uint8_t)to[i];
    Bytes out(text.size());
    for (size_t i = 0; i < text.size(); i++) out[i] = comp[text[text.size() - 1 - i]];
    return out;
}

Orientation orient_by_longest_segment(const Bytes& search, const Bytes& reference, const SuffixTableLookup& seeds) {
    Orientation o;
    o.fwd = find_greedy_non_overlapping_segments(search, reference, seeds);
    o.rev = find_greedy_non_overlapping_segments(bio_revcomp(search), reference, seeds);
    size_t f = 0, r = 0;
    for (const auto& p : o.fwd.alignment_segments) f += p.length;
    for (const auto& p : o.rev.alignment_segments) r += p.length;
    o.forward = f > r;
    return o;
}

// ------------------------------------------------------------------------------------------------ references
ReferenceManager::ReferenceManager(std::vector<Reference> refs, size_t ks, size_t kskip) : references(std::move(refs)), kmer_size(ks), kmer_skip(kskip) {
    for (size_t i = 0; i < references.size(); i++) {
        reference_name_to_ref[references[i].name] = i;
        longest_ref = std::max(longest_ref, references[i].sequence.size());
    }
}

ReferenceManager ReferenceManager::from_fa_file(const std::string& path, size_t ks, size_t kskip) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Unable to open reference file " + path);
    std::vector<Reference> refs;
    std::string line;
    while (std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            const size_t e = line.find_first_of(" \t", 1);
            refs.push_back({Bytes(), to_bytes(line.substr(1, e == std::string::npos ? std::string::npos : e - 1))});
        } else if (!refs.empty()) {
            refs.back().sequence.insert(refs.back().sequence.end(), line.begin(), line.end());
        }
    }
    return ReferenceManager(std::move(refs), ks, kskip);
}

std::vector<std::string> ReferenceManager::names() const {
    std::vector<std::string> n;
    for (const auto& r : references) n.push_back(to_string(r.name));
    return n;
}

// ------------------------------------------------------------------------------------------------ ReadBatch
ReadBatch::ReadBatch(uint32_t max_reads, uint64_t max_bytes) : max_reads_(max_reads), max_bytes_(max_bytes) {
    bytes_ = pinned<uint8_t>(max_bytes + 16);
    off_ = pinned<uint64_t>((size_t)max_reads + 1);
    fixed_ = pinned<int32_t>(max_reads);
    name_off_.reserve((size_t)max_reads + 1);
    name_off_.push_back(0);
    off_[0] = 0;
}

ReadBatch::~ReadBatch() {
    if (words_) clq_host_free(words_);
    clq_host_free(bytes_);
    clq_host_free(off_);
    clq_host_free(fixed_);
}

void ReadBatch::pack2() const {
    if (packed_) return;
    if (!words_) words_ = pinned<uint32_t>((size_t)((max_bytes_ + 15) / 16) + 4);
    const uint64_t nb = off_[n_];
    uint64_t ne = 0;
    if (exc_pos_.size() < 1024) { exc_pos_.resize(1024); exc_byte_.resize(1024); }
    int32_t rc = clq_pack2(bytes_, nb, words_, exc_pos_.data(), exc_byte_.data(), exc_pos_.size(), &ne);
    if (rc == CLQ_E_LIMIT) {  // ne = the capacity the list needs
        exc_pos_.resize((size_t)ne); exc_byte_.resize((size_t)ne);
        rc = clq_pack2(bytes_, nb, words_, exc_pos_.data(), exc_byte_.data(), exc_pos_.size(), &ne);
    }
    if (rc != CLQ_OK) fail(rc, "clq_pack2");
    exc_pos_.resize((size_t)ne); exc_byte_.resize((size_t)ne);
    packed_ = true;
}

void ReadBatch::clear() {
    n_ = 0;
    packed_ = false;
    n_first = 0xffffffffu;
    off_[0] = 0;
    name_bytes_.clear();
    name_off_.assign(1, 0);
    quals_.clear();
    have_quals_ = false;
}

bool ReadBatch::push(const char* name, size_t name_len, const uint8_t* seq, size_t n, const uint8_t* qual, int32_t fixed_ref) {
    if (n_ >= max_reads_ || off_[n_] + n > max_bytes_) return false;
    packed_ = false;
    if (n) std::memcpy(bytes_ + off_[n_], seq, n);
    if (qual) {
        if (!have_quals_) { quals_.assign((size_t)off_[n_], (uint8_t)'H'); have_quals_ = true; }
        quals_.insert(quals_.end(), qual, qual + n);
    } else if (have_quals_) {
        quals_.insert(quals_.end(), n, (uint8_t)'H');
    }
    fixed_[n_] = fixed_ref;
    off_[n_ + 1] = off_[n_] + n;
    name_bytes_.insert(name_bytes_.end(), name, name + name_len);
    name_off_.push_back(name_bytes_.size());
    n_++;
    return true;
}

bool ReadBatch::assign_span(const uint8_t* bytes, const uint64_t* off, uint64_t lo, uint64_t hi, const int32_t* fixed_ref) {
    const uint64_t n = hi - lo, nb = off[hi] - off[lo];
    if (n > max_reads_ || nb > max_bytes_) return false;
    clear();
    if (nb) std::memcpy(bytes_, bytes + off[lo], nb);
    const uint64_t base = off[lo];
    for (uint64_t i = 0; i <= n; i++) off_[i] = off[lo + i] - base;
    if (fixed_ref) std::memcpy(fixed_, fixed_ref + lo, n * sizeof(int32_t));
    else std::fill(fixed_, fixed_ + n, -1);
    name_off_.assign((size_t)n + 1, 0);
    n_ = (uint32_t)n;
    return true;
}

bool ReadBatch::append_span(const uint8_t* bytes, const uint64_t* off, uint64_t lo, uint64_t hi, const int32_t* fixed_ref) {
    const uint64_t n = hi - lo, nb = off[hi] - off[lo], have = off_[n_];
    if (n_ + n > max_reads_ || have + nb > max_bytes_) return false;
    packed_ = false;
    if (nb) std::memcpy(bytes_ + have, bytes + off[lo], nb);
    const uint64_t base = off[lo];
    for (uint64_t i = 1; i <= n; i++) off_[n_ + i] = have + (off[lo + i] - base);
    if (fixed_ref) std::memcpy(fixed_ + n_, fixed_ref + lo, n * sizeof(int32_t));
    else std::fill(fixed_ + n_, fixed_ + n_ + n, -1);
    name_off_.resize((size_t)(n_ + n) + 1, 0);
    if (n_first == 0xffffffffu) { n_first = n_; second_index = lo; }
    n_ += (uint32_t)n;
    return true;
}

std::optional<Bytes> ReadBatch::quals(uint32_t i) const {
    if (!have_quals_) return std::nullopt;
    return Bytes(quals_.begin() + (size_t)off_[i], quals_.begin() + (size_t)off_[i + 1]);
}

// ------------------------------------------------------------------------------------------------ BatchView
std::string BatchView::cigar_string(uint32_t i) const {
    std::string s;
    const uint32_t* c = cigar(i);
    for (uint32_t k = 0; k < cigar_len(i); k++) {
        s += std::to_string(c[k] >> 4);
        s += "MID"[c[k] & 3u];
    }
    return s;
}

// (score() of a CLQ_RUSTBIO batch is rust-bio's own score; the record the reference builds carries 0.0, see alignment())
double BatchView::alignment_rate(uint32_t i) const {
    const double m = results[i].matches, mm = results[i].mismatches;
    return m / (m + mm);
}

std::map<uint8_t, std::string> BatchView::digit_tags(uint32_t i) const {
    std::map<uint8_t, std::string> out;
    if (!tags || status(i) != CLQ_OK) return out;
    const Bytes& ref = rm->references[ref_index(i)].sequence;
    const uint8_t* row = tags + (size_t)i * tag_stride;
    uint32_t k = 0;
    for (uint8_t b : ref)
        if (b >= '0' && b <= '9' && k < tag_stride) out[b] += (char)row[k++];
    return out;
}

std::optional<AlignmentWithRef> BatchView::alignment(uint32_t i) const {
    if (status(i) != CLQ_OK) return std::nullopt;  // the reference warns / logs and moves on (alignment_functions.rs:165-176, :240-247)
    const Reference& r = rm->references[ref_index(i)];
    AlignmentWithRef a;
    a.alignment = AlignmentResult::from_cigar(to_string(r.name), batch->name(i), r.sequence.data(), r.sequence.size(), batch->read(i),
                                              batch->read_len(i), batch->quals(i), cigar(i), cigar_len(i), rust_bio ? 0.0 : score(i));
    if (rust_bio) a.alignment->path.clear();  // alignment_functions.rs:571-583: path: vec!(), score: 0.0
    a.ref_name = r.name;
    a.ref_sequence = r.sequence;
    return a;
}

TagMap BatchView::align_reads_tags(uint32_t i, const std::string& umi_symbols) const {
    TagMap t;
    const auto dt = digit_tags(i);
    for (char sym : umi_symbols) {
        const auto it = dt.find((uint8_t)sym);
        if (it != dt.end()) t[{'e', sym}] = it->second;
    }
    t[{'r', 'c'}] = "1";
    t[{'a', 'r'}] = batch->name(i);
    t[{'r', 'm'}] = f64_to_string(alignment_rate(i));
    t[{'a', 's'}] = f64_to_string(rust_bio ? 0.0 : score(i));
    return t;
}

namespace {
inline void append_uint(std::string& s, uint64_t v) {
    char buf[24];
    const auto r = std::to_chars(buf, buf + sizeof(buf), v);
    s.append(buf, r.ptr);
}
inline void append_f64(std::string& s, double v) {
    if (std::isnan(v)) { s += "NaN"; return; }
    if (std::isinf(v)) { s += v > 0 ? "inf" : "-inf"; return; }
    char buf[400];
    const auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    s.append(buf, r.ptr);
}
}  // namespace

bool BatchView::append_sam_line(uint32_t i, const std::string& umi_symbols, const std::vector<std::string>& reference_names,
                                std::string& out) const {
    if (status(i) != CLQ_OK) return false;
    const uint32_t ri = ref_index(i);
    const size_t l2 = batch->read_len(i);
    // QNAME FLAG RNAME POS MAPQ
    if (batch->name_len(i)) out.append(batch->name_data(i), batch->name_len(i)); else out += '*';
    out += "\t0\t";
    out += ri < reference_names.size() ? reference_names[ri] : std::string("*");
    out += "\t1\t255\t";
    // CIGAR
    const uint32_t* c = cigar(i);
    const uint32_t nc = cigar_len(i);
    if (!nc) out += '*';
    for (uint32_t k = 0; k < nc; k++) { append_uint(out, c[k] >> 4); out += "MID"[c[k] & 3u]; }
    out += "\t*\t0\t0\t";
    // SEQ = the read (read_aligned without gaps), QUAL = raw 'H' per base -> 'i' in SAM text
    if (l2) { out.append(reinterpret_cast<const char*>(batch->read(i)), l2); out += '\t'; out.append(l2, 'i'); }
    else out += "*\t*";
    // the extra set of align_reads in key order (ar, as, e<sym>.., rc, rm); to_sam_record's own rm / as replace those in place, rs is appended
    const double sc = rust_bio ? 0.0 : score(i);
    out += "\tar:Z:";
    out.append(batch->name_data(i), batch->name_len(i));
    out += "\tas:Z:"; append_f64(out, sc);
    if (tags && !umi_symbols.empty()) {
        const Bytes& ref = rm->references[ri].sequence;
        const uint8_t* row = tags + (size_t)i * tag_stride;
        bool seen[10] = {};
        for (uint8_t b : ref) if (b >= '0' && b <= '9') seen[b - '0'] = true;
        for (int d = 0; d < 10; d++) {   // key order: e0 < e1 < ..
            if (!seen[d] || umi_symbols.find((char)('0' + d)) == std::string::npos) continue;
            out += "\te"; out += (char)('0' + d); out += ":Z:";
            uint32_t k = 0;
            for (uint8_t b : ref)
                if (b >= '0' && b <= '9') { if (k < tag_stride && b == '0' + d) out += (char)row[k]; k++; }
        }
    }
    out += "\trc:Z:1\trm:Z:";
    append_f64(out, alignment_rate(i));
    out += "\trs:Z:"; append_f64(out, sc);
    out += '\n';
    return true;
}

// ------------------------------------------------------------------------------------------------ BAM
namespace bam {
namespace {
template <class T>
inline void put(std::string& s, T v) { s.append(reinterpret_cast<const char*>(&v), sizeof(T)); }  // little-endian host

inline int reg2bin(int64_t beg, int64_t end) {  // SAM specification, section 5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

inline uint8_t base_code(uint8_t b) {
    static const char codes[] = "=ACMGRSVTWYHKDBN";
    if (b >= 'a' && b <= 'z') b -= 32;
    for (int i = 0; i < 16; i++) if ((uint8_t)codes[i] == b) return (uint8_t)i;
    return 15;
}

// shared by the object path (append_record) and the raw path (BatchView::append_bam_record)
void record_core(std::string& out, int32_t ref_id, int32_t pos0, const char* name, size_t name_len, const uint32_t* cigar, size_t n_cigar,
                 const uint8_t* seq, size_t l_seq, uint8_t qual_byte, const std::string& aux) {
    const size_t start = out.size();
    put<int32_t>(out, 0);  // block_size, patched below
    int64_t ref_len = 0;
    for (size_t k = 0; k < n_cigar; k++) { const uint32_t op = cigar[k] & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += cigar[k] >> 4; }
    put<int32_t>(out, ref_id);
    put<int32_t>(out, pos0);
    const size_t nl = name_len ? name_len : 1;
    put<uint8_t>(out, (uint8_t)(nl + 1));
    put<uint8_t>(out, 255);  // mapping quality: missing
    put<uint16_t>(out, (uint16_t)reg2bin(pos0, pos0 + (ref_len > 0 ? ref_len : 1)));
    put<uint16_t>(out, (uint16_t)n_cigar);
    put<uint16_t>(out, 0);  // Flags::empty()
    put<uint32_t>(out, (uint32_t)l_seq);
    put<int32_t>(out, -1);  // mate reference
    put<int32_t>(out, -1);  // mate position
    put<int32_t>(out, 0);   // template length
    if (name_len) out.append(name, name_len); else out += '*';
    out += '\0';
    out.append(reinterpret_cast<const char*>(cigar), n_cigar * sizeof(uint32_t));  // len << 4 | op: the encoding the C ABI already uses
    for (size_t i = 0; i < l_seq; i += 2) {
        const uint8_t hi = base_code(seq[i]), lo = i + 1 < l_seq ? base_code(seq[i + 1]) : 0;
        out += (char)((hi << 4) | lo);
    }
    out.append(l_seq, (char)qual_byte);
    out += aux;
    const int32_t block = (int32_t)(out.size() - start - 4);
    std::memcpy(&out[start], &block, 4);
}

inline void aux_z(std::string& aux, char a, char b, const std::string& v) { aux += a; aux += b; aux += 'Z'; aux += v; aux += '\0'; }
}  // namespace

std::string header_text(const std::vector<std::string>& names, const std::vector<size_t>& lengths) {
    std::string t = "@HD\tVN:1.6\n";
    for (size_t i = 0; i < names.size(); i++) t += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lengths[i]) + "\n";
    t += "@CO\tClique processed\n";
    return t;
}

void append_header(const std::vector<std::string>& names, const std::vector<size_t>& lengths, std::string& out) {
    const std::string text = header_text(names, lengths);
    out += "BAM\1";
    put<int32_t>(out, (int32_t)text.size());
    out += text;
    put<int32_t>(out, (int32_t)names.size());
    for (size_t i = 0; i < names.size(); i++) {
        put<int32_t>(out, (int32_t)names[i].size() + 1);
        out += names[i];
        out += '\0';
        put<int32_t>(out, (int32_t)lengths[i]);
    }
}

void append_record(const SamRecord& rec, std::string& out) {
    std::vector<uint32_t> cig;
    for (const auto& t : rec.cigar) cig.push_back((uint32_t)(t.len << 4) | (uint32_t)t.kind);
    std::string aux;
    for (const auto& kv : rec.data) aux_z(aux, kv.first[0], kv.first[1], kv.second);
    record_core(out, rec.reference_sequence_id, (int32_t)rec.alignment_start - 1, rec.name.data(), rec.name.size(), cig.data(), cig.size(),
                rec.sequence.data(), rec.sequence.size(), rec.quality_scores.empty() ? 0xff : rec.quality_scores[0], aux);
}

void bgzf_compress(const char* data, size_t n, std::string& out, int level) {
    constexpr size_t kBlock = 65280;  // htslib's input block size: the compressed member always fits 64 KiB
    std::vector<uint8_t> buf(compressBound(kBlock) + 64);
    for (size_t off = 0; off < n || (n == 0 && off == 0); off += kBlock) {
        const size_t len = std::min(kBlock, n - off);
        z_stream zs{};
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2 failed");
        zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(data + off));
        zs.avail_in = (uInt)len;
        zs.next_out = buf.data();
        zs.avail_out = (uInt)buf.size();
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) throw std::runtime_error("deflate failed");
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef*>(data + off), (uInt)len);
        const uint16_t bsize = (uint16_t)(clen + 25);  // total block size - 1
        static const uint8_t head[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
        out.append(reinterpret_cast<const char*>(head), 12);
        out += 'B'; out += 'C';
        put<uint16_t>(out, 2);
        put<uint16_t>(out, bsize);
        out.append(reinterpret_cast<const char*>(buf.data()), clen);
        put<uint32_t>(out, crc);
        put<uint32_t>(out, (uint32_t)len);
        if (n == 0) break;
    }
}

void bgzf_eof(std::string& out) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    out.append(reinterpret_cast<const char*>(eof), 28);
}
}  // namespace bam

bool BatchView::append_bam_record(uint32_t i, const std::string& umi_symbols, std::string& out) const {
    if (status(i) != CLQ_OK) return false;
    const uint32_t ri = ref_index(i);
    const double sc = rust_bio ? 0.0 : score(i);
    std::string aux, num;
    bam::aux_z(aux, 'a', 'r', batch->name(i));
    num.clear(); append_f64(num, sc);
    bam::aux_z(aux, 'a', 's', num);
    if (tags && !umi_symbols.empty()) {
        for (const auto& kv : digit_tags(i))
            if (umi_symbols.find((char)kv.first) != std::string::npos) bam::aux_z(aux, 'e', (char)kv.first, kv.second);
    }
    bam::aux_z(aux, 'r', 'c', "1");
    std::string rate; append_f64(rate, alignment_rate(i));
    bam::aux_z(aux, 'r', 'm', rate);
    bam::aux_z(aux, 'r', 's', num);
    bam::record_core(out, (int32_t)ri, 0, batch->name_data(i), batch->name_len(i), cigar(i), cigar_len(i), batch->read(i), batch->read_len(i),
                     (uint8_t)'H', aux);
    return true;
}

// ------------------------------------------------------------------------------------------------ Aligner
Aligner::Aligner(const AlignerOptions& opt) : opt_(opt) {
    if (clq_device_count() <= opt.device) fail(CLQ_E_CUDA, "CUDA device " + std::to_string(opt.device) + " not available; libclq has no CPU fallback");
    if (!opt_.max_read_bytes) opt_.max_read_bytes = (uint64_t)opt_.max_reads * 512;
    clq_limits_t lim = {};
    lim.max_reads = opt_.max_reads;
    lim.max_read_bytes = opt_.max_read_bytes;
    lim.max_read_len = opt_.max_read_len;
    lim.max_refs = opt_.max_refs;
    lim.max_ref_bytes = opt_.max_ref_bytes;
    pool_ops_ = std::max<uint64_t>(1024, (uint64_t)opt_.max_reads * opt_.cigar_ops_per_read);
    lim.cigar_pool_ops = pool_ops_;
    lim.n_slots = std::max<uint32_t>(1, std::min<uint32_t>(4, opt_.n_slots));
    opt_.n_slots = lim.n_slots;
    const int32_t rc = clq_ctx_create(opt_.device, &lim, &ctx_);
    if (rc != CLQ_OK) fail(rc, clq_strerror(rc));
    for (uint32_t s = 0; s < opt_.n_slots; s++) {
        res_.push_back(pinned<clq_result_t>(opt_.max_reads));
        pool_.push_back(pinned<uint32_t>(pool_ops_));
        tags_.push_back(nullptr);
        tags_cap_.push_back(0);
    }
    flags_.assign(opt_.n_slots, 0);
    scale_.assign(opt_.n_slots, 1);
}

Aligner::~Aligner() {
    if (ctx_) clq_ctx_destroy(ctx_);
    for (auto* p : res_) clq_host_free(p);
    for (auto* p : pool_) clq_host_free(p);
    for (auto* p : tags_) if (p) clq_host_free(p);
}

void Aligner::check(int32_t rc, const char* what) const {
    if (rc != CLQ_OK) fail(rc, std::string(what) + ": " + clq_strerror(rc) + ": " + clq_ctx_last_error(ctx_));
}

void Aligner::set_references(const ReferenceManager& rm, bool build_kmer_index) {
    Bytes flat;
    std::vector<uint64_t> off(1, 0);
    for (const auto& r : rm.references) {
        flat.insert(flat.end(), r.sequence.begin(), r.sequence.end());
        off.push_back(flat.size());
    }
    if (flat.empty()) flat.push_back(0);
    check(clq_refs_set(ctx_, (uint32_t)rm.references.size(), flat.data(), off.data()), "clq_refs_set");
    if (build_kmer_index && !rm.references.empty())
        check(clq_kmer_index_set(ctx_, (uint32_t)rm.kmer_size, (uint32_t)rm.kmer_skip), "clq_kmer_index_set");
    rm_ = rm;
}

uint32_t Aligner::search_flags(bool fast_lookup) const {
    if (rm_.references.size() == 1) return CLQ_SEARCH_FIXED | CLQ_BAND_READLEN;
    return (fast_lookup ? CLQ_SEARCH_QUICK : CLQ_SEARCH_EXHAUSTIVE) | CLQ_BAND_READLEN;
}

void Aligner::submit(int slot, const ReadBatch& b, const clq_affine_t& sc, uint32_t flags, double threshold) {
    const bool fixed = (flags & CLQ_SEARCH_MASK) == CLQ_SEARCH_FIXED;
    if (opt_.pack2_upload) {
        b.pack2();  // no-op when a filler thread already packed the batch
        const auto& ep = b.exception_positions();
        check(clq_submit_packed2(ctx_, slot, b.size(), b.packed_words(), b.offsets(), ep.empty() ? nullptr : ep.data(),
                                 ep.empty() ? nullptr : b.exception_bytes().data(), ep.size(), fixed ? b.fixed_ref() : nullptr, &sc, flags, threshold),
              "clq_submit_packed2");
    } else {
        check(clq_submit(ctx_, slot, b.size(), b.bytes(), b.offsets(), fixed ? b.fixed_ref() : nullptr, &sc, flags, threshold), "clq_submit");
    }
    flags_[slot] = flags;
    scale_[slot] = sc.scale;
}

BatchView Aligner::wait(int slot, const ReadBatch& b) {
    uint64_t used = 0;
    check(clq_wait(ctx_, slot, res_[slot], pool_[slot], pool_ops_, &used), "clq_wait");
    BatchView v;
    v.batch = &b;
    v.rm = &rm_;
    v.results = res_[slot];
    v.cigar_pool = pool_[slot];
    v.scale = scale_[slot];
    v.device = opt_.device;
    v.rust_bio = (flags_[slot] & CLQ_RUSTBIO) != 0;
    v.cigar_used = used;
    if (flags_[slot] & CLQ_EXTRACT_TAGS) {
        uint32_t stride = 0;
        check(clq_tags_download(ctx_, slot, nullptr, 0, &stride), "clq_tags_download");
        const uint64_t need = (uint64_t)opt_.max_reads * stride;
        if (need > tags_cap_[slot]) {  // per slot: a BatchView of another slot may still point into its buffer
            if (tags_[slot]) clq_host_free(tags_[slot]);
            tags_[slot] = pinned<uint8_t>(need);
            tags_cap_[slot] = need;
        }
        if (stride && b.size()) check(clq_tags_download(ctx_, slot, tags_[slot], tags_cap_[slot], &stride), "clq_tags_download");
        v.tags = stride ? tags_[slot] : nullptr;
        v.tag_stride = stride;
    }
    return v;
}

clq_stats_t Aligner::stats(int slot) {
    clq_stats_t st;
    check(clq_slot_stats(ctx_, slot, &st), "clq_slot_stats");
    return st;
}

AlignmentResult Aligner::single(const Bytes& reference, const Bytes& read, std::optional<Bytes> qual, const AffineScoring& sc, uint32_t band,
                                const std::string& ref_name, const std::string& read_name) {
    // a private one-reference set for this call; the caller's set is restored afterwards
    const ReferenceManager saved = rm_;
    set_references(ReferenceManager({Reference{reference, to_bytes(ref_name)}}), false);
    if (!one_ || one_->capacity() < 1) one_ = std::make_unique<ReadBatch>(1, opt_.max_read_bytes);
    one_->clear();
    if (!one_->push(read_name, read.data(), read.size(), qual ? qual->data() : nullptr, 0)) fail(CLQ_E_LIMIT, "read exceeds max_read_bytes");
    std::optional<AlignmentWithRef> out;
    uint32_t st = CLQ_OK;
    try {
        submit(0, *one_, sc.to_int(), CLQ_SEARCH_FIXED | band);
        const BatchView v = wait(0, *one_);
        st = v.status(0);
        out = v.alignment(0);
    } catch (...) {
        set_references(saved);  // also when the aligner had no references: the temporary one must not stay installed
        throw;
    }
    set_references(saved);
    if (st == CLQ_TRACEBACK_DIVERGED) fail((int32_t)st, "the reference's traceback does not terminate for this pair (stale band cell)");
    if (st != CLQ_OK || !out) fail((int32_t)st, clq_strerror((int32_t)st));
    return std::move(*out->alignment);
}

AlignmentResult Aligner::align_two_strings(const Bytes& reference_sequence, const Bytes& read_sequence, std::optional<Bytes> read_qual,
                                           const AffineScoring& scoring_function, bool local, const std::string& ref_name,
                                           const std::string& read_name) {
    if (local) fail(CLQ_E_UNSUPPORTED, "local alignment is outside the hot path (SURVEY.md section 2)");
    return single(reference_sequence, read_sequence, std::move(read_qual), scoring_function, CLQ_BAND_MAXLEN, ref_name, read_name);
}

AlignmentResult Aligner::align_two_strings_passed_matrix(const std::string& ref_name, const std::string& read_name, const Bytes& reference,
                                                         const Bytes& read, std::optional<Bytes> qual, const AffineScoring& scoring,
                                                         size_t max_indel) {
    uint32_t band;
    if (max_indel == read.size()) band = CLQ_BAND_READLEN;
    else if (max_indel >= std::max(reference.size(), read.size())) band = CLQ_BAND_MAXLEN;  // the band covers the whole matrix
    else if (max_indel < (1u << 24)) band = CLQ_BAND_K | ((uint32_t)max_indel << CLQ_BAND_K_SHIFT);  // explicit bandwidth
    else fail(CLQ_E_UNSUPPORTED, "bandwidth out of range");
    return single(reference, read, std::move(qual), scoring, band, ref_name, read_name);
}

std::optional<AlignmentWithRef> Aligner::search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                const AffineScoring& sc, uint32_t mode, double threshold) {
    if (rm_.references.empty()) return std::nullopt;
    if (!one_) one_ = std::make_unique<ReadBatch>(1, opt_.max_read_bytes);
    one_->clear();
    if (!one_->push(read_name, read.data(), read.size(), qual ? qual->data() : nullptr, 0)) fail(CLQ_E_LIMIT, "read exceeds max_read_bytes");
    submit(0, *one_, sc.to_int(), mode | CLQ_BAND_READLEN, threshold);
    const BatchView v = wait(0, *one_);
    const uint32_t st = v.status(0);
    if (st == CLQ_NO_CANDIDATE) return std::nullopt;
    if (st == CLQ_TRACEBACK_DIVERGED) fail((int32_t)st, "the reference's traceback does not terminate for this pair (stale band cell)");
    if (st != CLQ_OK) fail((int32_t)st, clq_strerror((int32_t)st));
    return v.alignment(0);
}

std::optional<AlignmentWithRef> Aligner::exhaustive_alignment_search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                                     const AffineScoring& scoring) {
    return search(read_name, read, std::move(qual), scoring, CLQ_SEARCH_EXHAUSTIVE, 0.90);
}

std::optional<AlignmentWithRef> Aligner::quick_alignment_search(const std::string& read_name, const Bytes& read, std::optional<Bytes> qual,
                                                                const AffineScoring& scoring, double match_threshold) {
    return search(read_name, read, std::move(qual), scoring, CLQ_SEARCH_QUICK, match_threshold);
}

std::optional<AlignmentWithRef> Aligner::align_to_reference_choices(const std::string& read_name, const Bytes& read_in, std::optional<Bytes> qual,
                                                                    bool fast_lookup, const AffineScoring& scoring, bool rust_bio,
                                                                    bool known_strand) {
    if (rm_.references.empty()) return std::nullopt;
    Bytes oriented;
    if (rm_.references.size() == 1 && !known_strand) {  // alignment_functions.rs:549-558 (seed size = the reference manager's k-mer size, :97)
        const Bytes& ref = rm_.references[0].sequence;
        const Orientation o = orient_by_longest_segment(read_in, ref, SuffixTableLookup::find_seeds(ref, rm_.kmer_size));
        if (!o.forward) oriented = reverse_complement(read_in);
    }
    const Bytes& read = oriented.empty() ? read_in : oriented;
    if (rm_.references.size() == 1 && rust_bio) {
        if (!one_) one_ = std::make_unique<ReadBatch>(1, opt_.max_read_bytes);
        one_->clear();
        if (!one_->push(read_name, read.data(), read.size(), qual ? qual->data() : nullptr, 0)) fail(CLQ_E_LIMIT, "read exceeds max_read_bytes");
        submit(0, *one_, RustBioScoring().to_int(), CLQ_SEARCH_FIXED | CLQ_BAND_MAXLEN | CLQ_RUSTBIO);
        const BatchView v = wait(0, *one_);
        if (v.status(0) != CLQ_OK) fail((int32_t)v.status(0), clq_strerror((int32_t)v.status(0)));
        return v.alignment(0);
    }
    if (rm_.references.size() == 1) return search(read_name, read, std::move(qual), scoring, CLQ_SEARCH_FIXED, 0.90);
    return search(read_name, read, std::move(qual), scoring, fast_lookup ? CLQ_SEARCH_QUICK : CLQ_SEARCH_EXHAUSTIVE, 0.90);
}

MergedSequence Aligner::merge_reads_by_alignment(const FastqRecord& read1, const FastqRecord& read2, const AffineScoring& sc) {
    const Bytes rc2 = reverse_complement(read2.seq);
    Bytes q2(read2.qual.rbegin(), read2.qual.rend());
    const AlignmentResult r = align_two_strings(read1.seq, rc2, std::nullopt, sc, false, read1.id, read2.id);
    return alignment_rate_and_consensus(r.reference_aligned, read1.qual, r.read_aligned, q2);
}

std::vector<std::optional<MergedSequence>> Aligner::merge_read_pairs_by_alignment(const std::vector<ReadSetContainer>& pairs,
                                                                                  const AffineScoring& sc) {
    std::vector<std::optional<MergedSequence>> out(pairs.size());
    if (pairs.empty()) return out;
    const ReferenceManager saved = rm_;
    const clq_affine_t sci = sc.to_int();
    const uint32_t per = std::min<uint32_t>(opt_.max_reads, opt_.max_refs);
    ReadBatch batch(per, opt_.max_read_bytes);
    try {
        for (size_t lo = 0; lo < pairs.size();) {
            // one launch per chunk: reference k = read1 of pair lo + k, read k = revcomp(read2), fixed_ref[k] = k
            std::vector<Reference> refs;
            batch.clear();
            size_t hi = lo;
            for (; hi < pairs.size() && hi - lo < per; hi++) {
                const ReadSetContainer& p = pairs[hi];
                if (!p.read_two) throw std::logic_error("merge_read_pairs_by_alignment: pair without read2");
                const Bytes rc2 = reverse_complement(p.read_two->seq);
                if (!batch.push(p.read_one.id, rc2.data(), rc2.size(), nullptr, (int32_t)(hi - lo))) break;
                refs.push_back({p.read_one.seq, to_bytes(p.read_one.id)});
            }
            if (hi == lo) fail(CLQ_E_LIMIT, "a read pair does not fit an empty batch");
            set_references(ReferenceManager(std::move(refs)), false);
            submit(0, batch, sci, CLQ_SEARCH_FIXED | CLQ_BAND_MAXLEN);
            const BatchView v = wait(0, batch);
            for (size_t k = 0; k < hi - lo; k++) {
                const auto al = v.alignment((uint32_t)k);
                if (!al) continue;
                const ReadSetContainer& p = pairs[lo + k];
                const Bytes q2(p.read_two->qual.rbegin(), p.read_two->qual.rend());
                try {
                    out[lo + k] = alignment_rate_and_consensus(al->alignment->reference_aligned, p.read_one.qual, al->alignment->read_aligned, q2);
                } catch (const std::out_of_range&) {
                    // qualities shorter than the bases: the reference panics here; a batch reports the pair as unmerged
                }
            }
            lo = hi;
        }
    } catch (...) {
        set_references(saved);
        throw;
    }
    set_references(saved);
    return out;
}

namespace {
void account(AlignReadsStats& st, const BatchView& v) {
    st.batches++;
    st.reads += v.size();
    for (uint32_t i = 0; i < v.size(); i++) (v.status(i) == CLQ_OK ? st.aligned : st.dropped)++;
}

void prepare_fixed(ReadBatch& b, uint32_t flags) {
    // single-reference panels: every read aligns to reference 0 (alignment_functions.rs:544-548)
    if ((flags & CLQ_SEARCH_MASK) == CLQ_SEARCH_FIXED)
        for (uint32_t i = 0; i < b.size(); i++) if (b.fixed_ref()[i] < 0) b.fixed_ref_mut()[i] = 0;
}
}  // namespace

AlignReadsStats Aligner::align_reads(const ReadSource& source, const AffineScoring& scoring, bool fast_lookup, const ResultSink& sink,
                                     bool extract_tags, bool rust_bio, bool known_strand) {
    const auto tsetup = std::chrono::steady_clock::now();
    AlignReadsStats st;
    if (rm_.references.empty()) return st;
    const bool rb = rust_bio && rm_.references.size() == 1;
    const clq_affine_t sc = rb ? RustBioScoring().to_int() : scoring.to_int();
    const uint32_t flags = search_flags(fast_lookup) | (extract_tags ? CLQ_EXTRACT_TAGS : 0u) | (rb ? CLQ_RUSTBIO : 0u);
    // known_strand = false with one reference (alignment_functions.rs:549-558): orient on the host before the batch goes out
    const bool orient = !known_strand && rm_.references.size() == 1;
    const SuffixTableLookup seeds = orient ? SuffixTableLookup::find_seeds(rm_.references[0].sequence, rm_.kmer_size) : SuffixTableLookup();
    const uint32_t ns = opt_.n_slots;
    while (bufs_.size() < ns) bufs_.push_back(std::make_unique<ReadBatch>(opt_.max_reads, opt_.max_read_bytes));  // page-locked once
    auto& bufs = bufs_;
    const auto t0 = std::chrono::steady_clock::now();
    st.setup_seconds = std::chrono::duration<double>(t0 - tsetup).count();
    std::vector<bool> busy(ns, false);
    uint64_t next_index = 0;
    bool more = true;
    uint32_t slot = 0;
    auto drain = [&](uint32_t s) {
        const BatchView v = wait((int)s, *bufs[s]);
        st.cells += stats((int)s).cells;
        account(st, v);
        sink(v);
        busy[s] = false;
    };
    while (more) {
        if (busy[slot]) drain(slot);
        ReadBatch& b = *bufs[slot];
        b.clear();
        b.first_index = next_index;  // a sharded source overrides this with the batch's place in the whole input
        more = source(b);
        if (b.size()) {
            prepare_fixed(b, flags);
            if (orient) {
                Bytes tmp;
                for (uint32_t i = 0; i < b.size(); i++) {
                    tmp.assign(b.read(i), b.read(i) + b.read_len(i));
                    if (!orient_by_longest_segment(tmp, rm_.references[0].sequence, seeds).forward) {
                        const Bytes rc = reverse_complement(tmp);
                        std::memcpy(b.read_mut(i), rc.data(), rc.size());
                    }
                }
            }
            submit((int)slot, b, sc, flags);
            busy[slot] = true;
            next_index += b.size();
            slot = (slot + 1) % ns;
        }
    }
    for (uint32_t k = 0; k < ns; k++) {
        const uint32_t s = (slot + k) % ns;
        if (busy[s]) drain(s);
    }
    st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return st;
}

// ------------------------------------------------------------------------------------------------ ShardedAligner
ShardedAligner::ShardedAligner(const std::vector<int>& devices, AlignerOptions opt) {
    for (int d : devices) {
        opt.device = d;
        aligners_.push_back(std::make_unique<Aligner>(opt));
    }
}

void ShardedAligner::set_references(const ReferenceManager& rm, bool build_kmer_index) {
    for (auto& a : aligners_) a->set_references(rm, build_kmer_index);
}

AlignReadsStats ShardedAligner::align_reads(const ReadSource& source, const AffineScoring& scoring, bool fast_lookup, const ResultSink& sink,
                                            bool extract_tags, bool rust_bio, bool known_strand) {
    const auto t0 = std::chrono::steady_clock::now();
    AlignReadsStats total;
    std::mutex src_mu, sink_mu;
    bool more = true;
    uint64_t next_index = 0;
    std::exception_ptr err;
    // every device thread runs the single-GPU loop over a shared, mutex-guarded source: batches go to whichever GPU is
    // free next (dynamic read sharding, no collective)
    const ReadSource shared_source = [&](ReadBatch& b) -> bool {
        std::lock_guard<std::mutex> g(src_mu);
        if (!more) return false;
        b.first_index = next_index;
        more = source(b);
        next_index += b.size();
        return more;
    };
    const ResultSink locked_sink = [&](const BatchView& v) {
        std::lock_guard<std::mutex> g(sink_mu);
        sink(v);
    };
    auto work = [&](Aligner* a) {
        try {
            const AlignReadsStats st = a->align_reads(shared_source, scoring, fast_lookup, locked_sink, extract_tags, rust_bio, known_strand);
            std::lock_guard<std::mutex> g(sink_mu);
            total.reads += st.reads; total.aligned += st.aligned; total.dropped += st.dropped; total.batches += st.batches; total.cells += st.cells;
            total.setup_seconds = std::max(total.setup_seconds, st.setup_seconds);
        } catch (...) {
            std::lock_guard<std::mutex> g(sink_mu);
            if (!err) err = std::current_exception();
        }
    };
    std::vector<std::thread> th;
    for (auto& a : aligners_) th.emplace_back(work, a.get());
    for (auto& t : th) t.join();
    if (err) std::rethrow_exception(err);
    total.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - total.setup_seconds;
    return total;
}

// ------------------------------------------------------------------------------------------------ span feed
namespace {
// a tiny blocking queue of batch buffers (indices into the device's buffer pool)
struct IdxQueue {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> q;
    bool closed = false;
    void push(int v) { { std::lock_guard<std::mutex> g(mu); q.push_back(v); } cv.notify_one(); }
    void close() { { std::lock_guard<std::mutex> g(mu); closed = true; } cv.notify_all(); }
    bool pop(int& v) {  // false once closed and drained
        std::unique_lock<std::mutex> g(mu);
        cv.wait(g, [&] { return !q.empty() || closed; });
        if (q.empty()) return false;
        v = q.front(); q.pop_front();
        return true;
    }
};
}  // namespace

// ------------------------------------------------------------------------------------------------ SpanClaimer
// Claims of align_reads_span.  One pair of cursors over the whole span ([front, back) is what is left): a claim takes up to a full
// batch, but with several devices never more than 1 / (2 * devices) of the bytes that are left (guided self-scheduling), so the
// last batches are small and the devices drain together.  On a length-sorted stream (the reads at one end much longer than at
// the other) a claim takes half of its bytes from the LONG end and fills up from the short end: a launch of only-long pairs runs
// in whole waves of the persistent grid (1184 pairs of 5 kb, ~50 ms each, whatever the batch holds), a mixed one back-fills the
// last wave with its short pairs, and the cheap short reads are what is left for the small claims at the end.
SpanClaimer::SpanClaimer(const uint64_t* off, uint64_t n, size_t n_devices, int claimers_per_device, uint64_t max_reads,
                         uint64_t max_read_bytes, ReadSpan::Order order)
    : off_(off), nd_(std::max<size_t>(n_devices, 1)), nf_(std::max(claimers_per_device, 1)), max_reads_(std::max<uint64_t>(max_reads, 1)),
      max_bytes_(std::max<uint64_t>(max_read_bytes, 1)), back_(n), order_(order) {
    const uint64_t m = std::min<uint64_t>(4096, n / 8);
    double head = 0.0, tail = 0.0;
    if (m) {
        head = (double)(off[m] - off[0]) / (double)m;
        tail = (double)(off[n] - off[n - m]) / (double)m;
    }
    const bool skewed = m && (tail > 1.5 * head || head > 1.5 * tail);
    long_at_back_ = tail >= head;
    if (order_ == ReadSpan::Order::Auto) order_ = skewed ? ReadSpan::Order::TwoEnded : ReadSpan::Order::Front;
}

// reads taken from one end of [front, back) within a byte and a read budget
uint64_t SpanClaimer::fit_front(uint64_t bytes_budget, uint64_t reads_budget) const {
    const uint64_t* first = off_ + front_ + 1;
    const uint64_t* last = off_ + back_ + 1;
    const uint64_t n = (uint64_t)(std::upper_bound(first, last, off_[front_] + bytes_budget) - first);
    return std::min<uint64_t>(n, std::min<uint64_t>(reads_budget, back_ - front_));
}

uint64_t SpanClaimer::fit_back(uint64_t bytes_budget, uint64_t reads_budget) const {
    const uint64_t* first = off_ + front_;
    const uint64_t* last = off_ + back_ + 1;
    const uint64_t floor_off = off_[back_] > bytes_budget ? off_[back_] - bytes_budget : 0;
    const uint64_t i = front_ + (uint64_t)(std::lower_bound(first, last, floor_off) - first);  // first read boundary >= floor_off
    return std::min<uint64_t>(back_ - std::min(i, back_), std::min<uint64_t>(reads_budget, back_ - front_));
}

bool SpanClaimer::claim(uint64_t& lo, uint64_t& hi, uint64_t& lo2, uint64_t& hi2) {
    std::lock_guard<std::mutex> g(mu_);
    lo2 = hi2 = 0;
    if (front_ >= back_) return false;
    const uint64_t min_claim = std::min<uint64_t>(max_reads_, 4096);
    const uint64_t left_bytes = off_[back_] - off_[front_];
    // one device: nothing to balance, full batches.  Several: a claim is at most half of an equal share of what is left
    const uint64_t share = nd_ > 1 ? left_bytes / (2 * nd_) : left_bytes;
    const uint64_t want_bytes = std::min<uint64_t>(max_bytes_, std::max<uint64_t>(share, 1));
    // ramp-up: the first claims are small (32 Ki reads, doubling per round of claims) so that every GPU starts computing
    // after about a millisecond of staging instead of after a whole batch
    const uint64_t round = n_claims_++ / (uint64_t)(nd_ * (size_t)nf_);
    const uint64_t max_n = std::min<uint64_t>(max_reads_, 32768ull << std::min<uint64_t>(round, 10));
    const bool from_back = order_ != ReadSpan::Order::Front && long_at_back_;
    auto fit_long = [&](uint64_t bb, uint64_t rb) { return from_back ? fit_back(bb, rb) : fit_front(bb, rb); };
    auto fit_short = [&](uint64_t bb, uint64_t rb) { return from_back ? fit_front(bb, rb) : fit_back(bb, rb); };
    auto take_long = [&](uint64_t n, uint64_t& a, uint64_t& b) { if (from_back) { a = back_ - n; b = back_; back_ -= n; } else { a = front_; b = front_ + n; front_ += n; } };
    auto take_short = [&](uint64_t n, uint64_t& a, uint64_t& b) { if (from_back) { a = front_; b = front_ + n; front_ += n; } else { a = back_ - n; b = back_; back_ -= n; } };
    if (order_ == ReadSpan::Order::TwoEnded) {
        uint64_t nl = fit_long(std::max<uint64_t>(want_bytes / 2, 1), max_n);
        if (nl == 0) nl = 1;  // at least the longest read (alone if it is larger than a whole batch: the caller reports CLQ_READ_TOO_LONG)
        take_long(nl, lo, hi);
        const uint64_t long_bytes = off_[hi] - off_[lo];
        if (front_ < back_ && long_bytes <= max_bytes_ && nl < max_n) {
            // the rest of the byte budget from the short end; a tiny claim (the end of the stream) is topped up to min_claim reads,
            // within 1 MiB (by reads alone the top-up would eat the short end and leave the expensive reads for the last claims)
            uint64_t ns = fit_short(want_bytes > long_bytes ? want_bytes - long_bytes : 0, max_n - nl);
            const uint64_t tiny = std::min<uint64_t>(max_bytes_, 1ull << 20);
            if (nl + ns < min_claim && nl < std::min<uint64_t>(min_claim, max_n) && long_bytes < tiny)
                ns = std::max(ns, fit_short(tiny - long_bytes, std::min<uint64_t>(min_claim, max_n) - nl));
            if (ns) take_short(ns, lo2, hi2);
        }
        return true;
    }
    // one-ended claims (front to back, or longest first): at least min_claim reads when they fit
    const uint64_t by_bytes = fit_long(want_bytes, max_n);
    const uint64_t cap_bytes = fit_long(max_bytes_, max_n);
    uint64_t n = std::max<uint64_t>(by_bytes, std::min<uint64_t>(min_claim, cap_bytes));
    if (n == 0) n = 1;  // a read larger than a whole batch: handed over alone, the caller reports CLQ_READ_TOO_LONG
    take_long(n, lo, hi);
    return true;
}

SpanStats ShardedAligner::align_reads_span(const ReadSpan& span, const AffineScoring& scoring, bool fast_lookup, SpanOutput& out,
                                           int fillers_per_device, bool extract_tags, bool rust_bio) {
    SpanStats stats;
    const size_t nd = aligners_.size();
    stats.device_kernel_ms.assign(nd, 0.0);
    stats.device_reads.assign(nd, 0); stats.device_cells.assign(nd, 0); stats.device_batches.assign(nd, 0);
    out.cigar_used = 0;
    if (!nd || !span.n) return stats;
    const ReferenceManager& rm = aligners_[0]->references();
    if (rm.references.empty()) return stats;
    if (!span.bytes || !span.off || !out.results) fail(CLQ_E_INVALID, "align_reads_span: null span or result array");
    const AlignerOptions& opt = aligners_[0]->options();
    const bool rb = rust_bio && rm.references.size() == 1;
    const clq_affine_t sc = rb ? RustBioScoring().to_int() : scoring.to_int();
    out.scale = sc.scale;
    uint32_t flags = aligners_[0]->search_flags(fast_lookup) | (extract_tags ? CLQ_EXTRACT_TAGS : 0u) | (rb ? CLQ_RUSTBIO : 0u);
    if (span.fixed_ref && rm.references.size() > 1) flags = (flags & ~CLQ_SEARCH_MASK) | CLQ_SEARCH_FIXED;  // the caller knows the reference of every read
    const int nf = std::max(1, fillers_per_device);
    const uint32_t ns = opt.n_slots;
    const int nbuf = (int)ns + nf;  // per device: one batch per stream slot in flight + one per filler being staged

    const auto tsetup = std::chrono::steady_clock::now();
    struct Dev {
        std::vector<std::unique_ptr<ReadBatch>>* pool = nullptr;
        IdxQueue free_q, ready_q;
        std::atomic<int> fillers_left{0};
    };
    std::vector<std::unique_ptr<Dev>> devs;
    span_bufs_.resize(nd);
    for (size_t d = 0; d < nd; d++) {
        devs.push_back(std::make_unique<Dev>());
        while ((int)span_bufs_[d].size() < nbuf) span_bufs_[d].push_back(std::make_unique<ReadBatch>(opt.max_reads, opt.max_read_bytes));  // page-locked once
        devs[d]->pool = &span_bufs_[d];
        for (int b = 0; b < nbuf; b++) devs[d]->free_q.push(b);
        devs[d]->fillers_left = nf;
    }
    const auto t0 = std::chrono::steady_clock::now();
    stats.total.setup_seconds = std::chrono::duration<double>(t0 - tsetup).count();

    SpanClaimer claimer(span.off, span.n, nd, nf, opt.max_reads, opt.max_read_bytes, span.order);
    auto claim = [&](uint64_t& lo, uint64_t& hi, uint64_t& lo2, uint64_t& hi2) { return claimer.claim(lo, hi, lo2, hi2); };

    std::mutex stat_mu, pool_mu;
    std::exception_ptr err;
    std::atomic<bool> failed{false};
    auto fail_with = [&](std::exception_ptr e) {
        std::lock_guard<std::mutex> g(stat_mu);
        if (!err) err = e;
        failed = true;
    };

    auto filler = [&](size_t d) {
        Dev& D = *devs[d];
        double secs = 0.0;
        try {
            uint64_t lo, hi, lo2, hi2;
            while (!failed && claim(lo, hi, lo2, hi2)) {
                int b;
                if (!D.free_q.pop(b)) break;
                const auto f0 = std::chrono::steady_clock::now();
                ReadBatch& rbuf = *(*D.pool)[b];
                if (!rbuf.assign_span(span.bytes, span.off, lo, hi, span.fixed_ref)) {
                    // does not fit even an empty batch (one read beyond max_read_bytes): the reference drops such reads (:240-247)
                    for (uint64_t i = lo; i < hi; i++) { clq_result_t r = {}; r.status = CLQ_READ_TOO_LONG; r.ref_index = 0xffffffffu; out.results[i] = r; }
                    D.free_q.push(b);
                    std::lock_guard<std::mutex> g(stat_mu);
                    stats.total.reads += hi - lo; stats.total.dropped += hi - lo;
                    continue;   // (such a claim is a single read and has no second range)
                }
                rbuf.first_index = lo;
                if (hi2 > lo2 && !rbuf.append_span(span.bytes, span.off, lo2, hi2, span.fixed_ref)) fail(CLQ_E_STATE, "align_reads_span: a two-ended claim does not fit its batch");
                if (opt.pack2_upload) rbuf.pack2();  // on the filler thread, so that the packing pass scales with fillers_per_device
                secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - f0).count();
                D.ready_q.push(b);
            }
        } catch (...) { fail_with(std::current_exception()); }
        if (--D.fillers_left == 0) D.ready_q.close();
        std::lock_guard<std::mutex> g(stat_mu);
        stats.fill_seconds += secs;
    };

    auto device = [&](size_t d) {
        Dev& D = *devs[d];
        Aligner& a = *aligners_[d];
        double sink_secs = 0.0, kernel_ms = 0.0;
        uint64_t reads = 0, cells = 0, batches = 0, aligned = 0, dropped = 0;
        std::vector<int> in_slot(ns, -1);
        auto drain = [&](uint32_t s) {
            const int b = in_slot[s];
            const BatchView v = a.wait((int)s, *(*D.pool)[b]);
            const clq_stats_t st = a.stats((int)s);
            kernel_ms += st.kernel_ms; cells += st.cells;
            const auto s0 = std::chrono::steady_clock::now();
            const uint64_t first = v.batch->first_index, n = v.size();
            uint64_t base;
            {
                std::lock_guard<std::mutex> g(pool_mu);
                base = out.cigar_used; out.cigar_used += v.cigar_used;
                if (v.tags && v.tag_stride) out.tag_stride = v.tag_stride;
            }
            if (base + v.cigar_used > out.cigar_cap || base + v.cigar_used > 0xffffffffull) fail(CLQ_E_LIMIT, "align_reads_span: the caller's CIGAR pool is too small");
            if (v.cigar_used) std::memcpy(out.cigar_pool + base, v.cigar_pool, v.cigar_used * sizeof(uint32_t));
            // a batch is one range of the input, or two (a two-ended claim): reads 0 .. n1-1 and n1 .. n-1
            const uint64_t n1 = std::min<uint64_t>(v.batch->n_first, n), second = v.batch->second_index;
            for (uint64_t i = 0; i < n; i++) {
                clq_result_t r = v.results[i];
                r.cigar_off += (uint32_t)base;
                out.results[i < n1 ? first + i : second + (i - n1)] = r;
                (r.status == CLQ_OK ? aligned : dropped)++;
            }
            if (out.tags && v.tags && v.tag_stride) {
                std::memcpy(out.tags + first * v.tag_stride, v.tags, n1 * v.tag_stride);
                if (n > n1) std::memcpy(out.tags + second * v.tag_stride, v.tags + n1 * v.tag_stride, (n - n1) * v.tag_stride);
            }
            sink_secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
            reads += n; batches++;
            in_slot[s] = -1;
            D.free_q.push(b);
        };
        try {
            uint32_t slot = 0;
            int b;
            while (D.ready_q.pop(b)) {
                if (failed) { D.free_q.push(b); continue; }
                if (in_slot[slot] >= 0) drain(slot);
                ReadBatch& rbuf = *(*D.pool)[b];
                prepare_fixed(rbuf, flags);
                a.submit((int)slot, rbuf, sc, flags);
                in_slot[slot] = b;
                slot = (slot + 1) % ns;
            }
            for (uint32_t k = 0; k < ns; k++) {
                const uint32_t s = (slot + k) % ns;
                if (in_slot[s] >= 0) drain(s);
            }
        } catch (...) {
            fail_with(std::current_exception());
            D.free_q.close();  // unblock this device's fillers
            int b;
            while (D.ready_q.pop(b)) {}
        }
        std::lock_guard<std::mutex> g(stat_mu);
        stats.device_kernel_ms[d] = kernel_ms; stats.device_reads[d] = reads; stats.device_cells[d] = cells; stats.device_batches[d] = batches;
        stats.total.reads += reads; stats.total.cells += cells; stats.total.batches += batches; stats.total.aligned += aligned; stats.total.dropped += dropped;
        stats.sink_seconds += sink_secs;
    };

    std::vector<std::thread> th;
    for (size_t d = 0; d < nd; d++) {
        th.emplace_back(device, d);
        for (int f = 0; f < nf; f++) th.emplace_back(filler, d);
    }
    for (auto& t : th) t.join();
    if (err) std::rethrow_exception(err);
    stats.total.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return stats;
}

}  // namespace clique

// clqh_c.cpp -- a thin extern "C" view of the pure host functions of include/clique_host.hpp, so that the CPU test suite
// (ctypes) can check them against the oracle and the reference's golden vectors without a GPU.  Not part of the drop-in
// boundary (that is include/clq.h); declared in include/clqh.h.
#include <cstdlib>
#include <cctype>
#include <cstring>

#include "../../../include/clique_host.hpp"
#include "../../../include/clqh.h"

using namespace clique;

namespace {
size_t put_records(const std::map<uint8_t, std::string>& m, uint8_t* out, size_t cap) {
    size_t w = 0;
    for (const auto& kv : m) {
        if (w + 5 + kv.second.size() > cap) return 0;
        out[w] = kv.first;
        const uint32_t len = (uint32_t)kv.second.size();
        std::memcpy(out + w + 1, &len, 4);
        std::memcpy(out + w + 5, kv.second.data(), len);
        w += 5 + len;
    }
    return w;
}

size_t put_string(const std::string& s, char* out, size_t cap) {
    if (s.size() + 1 > cap) return 0;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return s.size();
}

std::vector<AlignmentTag> decode(const uint32_t* ops, size_t n) {
    std::vector<AlignmentTag> v;
    for (size_t i = 0; i < n; i++) v.push_back({(AlignmentTag::Kind)(ops[i] & 15u), ops[i] >> 4});
    return v;
}
}  // namespace

extern "C" {

size_t clqh_extract_tagged_sequences(const uint8_t* aligned_read, size_t n_read, const uint8_t* aligned_ref, size_t n_ref, uint8_t* out,
                                     size_t cap) {
    return put_records(extract_tagged_sequences(Bytes(aligned_read, aligned_read + n_read), Bytes(aligned_ref, aligned_ref + n_ref)), out, cap);
}

void clqh_reverse_complement(const uint8_t* dna, size_t n, uint8_t* out) {
    const Bytes r = reverse_complement(dna, n);
    if (n) std::memcpy(out, r.data(), n);
}

size_t clqh_f64_to_string(double v, char* out, size_t cap) { return put_string(f64_to_string(v), out, cap); }

double clqh_get_reference_alignment_rate(const uint8_t* ref_aligned, const uint8_t* read_aligned, size_t n) {
    return get_reference_alignment_rate(Bytes(ref_aligned, ref_aligned + n), Bytes(read_aligned, read_aligned + n));
}

size_t clqh_simplify_cigar(const uint32_t* ops, size_t n, uint32_t* out) {
    const auto v = simplify_cigar_string(decode(ops, n));
    for (size_t i = 0; i < v.size(); i++) out[i] = (uint32_t)(v[i].len << 4) | (uint32_t)v[i].kind;
    return v.size();
}

int32_t clqh_from_cigar(const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2, const uint32_t* ops, size_t n_ops,
                        uint8_t* ref_aligned, uint8_t* read_aligned, size_t aligned_cap, size_t* aligned_len, uint32_t* path_xy,
                        size_t path_cap, size_t* path_len) {
    try {
        const AlignmentResult r = AlignmentResult::from_cigar("ref", "read", ref, l1, read, l2, std::nullopt, ops, n_ops, 0.0);
        if (r.reference_aligned.size() > aligned_cap || r.path.size() > path_cap) return CLQ_E_LIMIT;
        std::memcpy(ref_aligned, r.reference_aligned.data(), r.reference_aligned.size());
        std::memcpy(read_aligned, r.read_aligned.data(), r.read_aligned.size());
        *aligned_len = r.reference_aligned.size();
        for (size_t i = 0; i < r.path.size(); i++) { path_xy[2 * i] = (uint32_t)r.path[i].x; path_xy[2 * i + 1] = (uint32_t)r.path[i].y; }
        *path_len = r.path.size();
        return CLQ_OK;
    } catch (const ClqError& e) {
        return e.code;
    }
}

size_t clqh_sam_line(const char* ref_name, const char* read_name, const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2,
                     const uint32_t* ops, size_t n_ops, double score, int32_t reference_id, const char* extra_tags, char* out, size_t cap) {
    try {
        const AlignmentResult r = AlignmentResult::from_cigar(ref_name, read_name, ref, l1, read, l2, std::nullopt, ops, n_ops, score);
        TagMap extra;  // "k1=value;k2=value"
        std::string s = extra_tags ? extra_tags : "";
        size_t pos = 0;
        while (pos < s.size()) {
            const size_t e = s.find(';', pos);
            const std::string item = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            if (item.size() >= 3 && item[2] == '=') extra[{item[0], item[1]}] = item.substr(3);
            if (e == std::string::npos) break;
            pos = e + 1;
        }
        std::vector<std::string> names((size_t)std::max(reference_id, 0) + 1, "*");
        if (reference_id >= 0) names[reference_id] = ref_name;
        return put_string(r.to_sam_record(reference_id, extra, std::nullopt).to_sam_line(names), out, cap);
    } catch (const std::exception&) {
        return 0;
    }
}

uint8_t clqh_combine_phred_scores(uint8_t a, uint8_t b, int32_t agree) { return combine_phred_scores(a, b, agree != 0); }

size_t clqh_alignment_rate_and_consensus(const uint8_t* a1, const uint8_t* q1, size_t nq1, const uint8_t* a2, const uint8_t* q2, size_t nq2,
                                         size_t n, uint8_t* out_bases, uint8_t* out_quals) {
    try {
        const MergedSequence m = alignment_rate_and_consensus(Bytes(a1, a1 + n), Bytes(q1, q1 + nq1), Bytes(a2, a2 + n), Bytes(q2, q2 + nq2));
        if (n) { std::memcpy(out_bases, m.read_bases.data(), n); std::memcpy(out_quals, m.read_quals.data(), n); }
        return n;
    } catch (const std::exception&) {
        return (size_t)-1;
    }
}

/* GPU: Aligner::merge_read_pairs_by_alignment over n pairs given as packed arrays (read1 / read2 bytes + qualities with
 * shared offsets per side); merged bases / qualities are written back to back, out_off gets n + 1 offsets, a pair the
 * reference could not finish has an empty record.  Returns 0 or a negative libclq code. */
int32_t clqh_merge_read_pairs_by_alignment(int32_t device, uint32_t n, const uint8_t* r1, const uint8_t* q1, const uint64_t* off1,
                                           const uint8_t* r2, const uint8_t* q2, const uint64_t* off2, double match_score,
                                           double mismatch_score, double special_score, double gap_open, double gap_extend,
                                           double final_gap_multiplier, uint8_t* out_bases, uint8_t* out_quals, uint64_t cap,
                                           uint64_t* out_off) {
    try {
        AlignerOptions opt;
        opt.device = device;
        opt.max_reads = 1u << 16;
        opt.max_refs = 1u << 16;
        opt.max_ref_bytes = 1u << 26;
        opt.cigar_ops_per_read = 128;
        opt.n_slots = 1;
        Aligner al(opt);
        std::vector<ReadSetContainer> pairs(n);
        for (uint32_t i = 0; i < n; i++) {
            pairs[i].read_one = {"r" + std::to_string(i), Bytes(r1 + off1[i], r1 + off1[i + 1]), Bytes(q1 + off1[i], q1 + off1[i + 1])};
            pairs[i].read_two = FastqRecord{"r" + std::to_string(i), Bytes(r2 + off2[i], r2 + off2[i + 1]), Bytes(q2 + off2[i], q2 + off2[i + 1])};
        }
        const auto merged = al.merge_read_pairs_by_alignment(pairs, {match_score, mismatch_score, special_score, gap_open, gap_extend, final_gap_multiplier});
        uint64_t w = 0;
        out_off[0] = 0;
        for (uint32_t i = 0; i < n; i++) {
            if (merged[i]) {
                const size_t len = merged[i]->read_bases.size();
                if (w + len > cap) return CLQ_E_LIMIT;
                if (len) { std::memcpy(out_bases + w, merged[i]->read_bases.data(), len); std::memcpy(out_quals + w, merged[i]->read_quals.data(), len); }
                w += len;
            }
            out_off[i + 1] = w;
        }
        return CLQ_OK;
    } catch (const ClqError& e) {
        return e.code < 0 ? e.code : CLQ_E_INVALID;
    } catch (const std::exception&) {
        return CLQ_E_INVALID;
    }
}

size_t clqh_find_greedy_non_overlapping_segments(const uint8_t* search, size_t n, const uint8_t* reference, size_t m, size_t seed_size,
                                                 uint32_t* out_xyz, size_t cap, size_t* start_position) {
    const Bytes ref(reference, reference + m);
    const SharedSegments sg = find_greedy_non_overlapping_segments(Bytes(search, search + n), ref, SuffixTableLookup::find_seeds(ref, seed_size));
    if (start_position) *start_position = sg.start_position;
    const size_t k = std::min(cap, sg.alignment_segments.size());
    for (size_t i = 0; i < k; i++) {
        out_xyz[3 * i] = (uint32_t)sg.alignment_segments[i].search_start;
        out_xyz[3 * i + 1] = (uint32_t)sg.alignment_segments[i].ref_start;
        out_xyz[3 * i + 2] = (uint32_t)sg.alignment_segments[i].length;
    }
    return k;
}

int32_t clqh_orient_by_longest_segment(const uint8_t* search, size_t n, const uint8_t* reference, size_t m, size_t seed_size) {
    const Bytes ref(reference, reference + m);
    return orient_by_longest_segment(Bytes(search, search + n), ref, SuffixTableLookup::find_seeds(ref, seed_size)).forward ? 1 : 0;
}

size_t clqh_extend_hit(const uint8_t* search, size_t n, size_t search_location, const uint8_t* reference, size_t m, size_t reference_location) {
    return extend_hit(search, n, search_location, reference, m, reference_location);
}

size_t clqh_bam_file(const char* ref_name, const char* read_name, const uint8_t* ref, size_t l1, const uint8_t* read, size_t l2,
                     const uint32_t* ops, size_t n_ops, double score, const char* extra_tags, uint8_t* out, size_t cap) {
    // a complete BAM file (BGZF: header block, one record, EOF marker) for one alignment, through the object path
    try {
        const AlignmentResult r = AlignmentResult::from_cigar(ref_name, read_name, ref, l1, read, l2, std::nullopt, ops, n_ops, score);
        TagMap extra;
        std::string s = extra_tags ? extra_tags : "";
        size_t pos = 0;
        while (pos < s.size()) {
            const size_t e = s.find(';', pos);
            const std::string item = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            if (item.size() >= 3 && item[2] == '=') extra[{item[0], item[1]}] = item.substr(3);
            if (e == std::string::npos) break;
            pos = e + 1;
        }
        std::string raw, z;
        bam::append_header({ref_name}, {l1}, raw);
        bam::bgzf_compress(raw.data(), raw.size(), z);
        raw.clear();
        bam::append_record(r.to_sam_record(0, extra, std::nullopt), raw);
        bam::bgzf_compress(raw.data(), raw.size(), z);
        bam::bgzf_eof(z);
        if (z.size() > cap) return 0;
        std::memcpy(out, z.data(), z.size());
        return z.size();
    } catch (const std::exception&) {
        return 0;
    }
}

size_t clqh_merge_reads_by_concatenation(const uint8_t* r1, size_t n1, const uint8_t* r2, size_t n2, const char* layout, uint8_t* out,
                                         size_t cap) {
    // layout: comma-separated items "1F" / "2R" / "2C" (read number + Forward / Reverse / reverse-Complement) or "S:ACGT" (spacer)
    try {
        ReadSetContainer rs;
        rs.read_one = {"r", Bytes(r1, r1 + n1), Bytes(n1, (uint8_t)'I')};
        if (r2) rs.read_two = FastqRecord{"r", Bytes(r2, r2 + n2), Bytes(n2, (uint8_t)'I')};
        std::vector<ReadPosition> lay;
        std::string s = layout ? layout : "";
        size_t pos = 0;
        while (pos < s.size()) {
            const size_t e = s.find(',', pos);
            const std::string item = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
            ReadPosition p;
            if (item.rfind("S:", 0) == 0) { p.kind = ReadPosition::Spacer; p.spacer_sequence = item.substr(2); }
            else if (item.size() == 2) {
                p.kind = item[0] == '1' ? ReadPosition::Read1 : ReadPosition::Read2;
                p.orientation = item[1] == 'F' ? AlignedReadOrientation::Forward
                              : item[1] == 'R' ? AlignedReadOrientation::Reverse
                              : item[1] == 'C' ? AlignedReadOrientation::ReverseComplement : AlignedReadOrientation::Unknown;
            } else return 0;
            lay.push_back(p);
            if (e == std::string::npos) break;
            pos = e + 1;
        }
        const MergedSequence m = merge_reads_by_concatenation(rs, lay);
        if (m.read_bases.size() > cap) return 0;
        if (!m.read_bases.empty()) std::memcpy(out, m.read_bases.data(), m.read_bases.size());
        return m.read_bases.size();
    } catch (const std::exception&) {
        return (size_t)-1;
    }
}


int32_t clqh_align_reads_span(const int32_t* devices, uint32_t n_devices, const clqh_span_options_t* o, const uint8_t* ref_bytes,
                              const uint64_t* ref_off, uint32_t n_refs, const uint8_t* read_bytes, const uint64_t* read_off, uint64_t n_reads,
                              const int32_t* fixed_ref, double match_score, double mismatch_score, double special_score, double gap_open,
                              double gap_extend, double final_gap_multiplier, uint32_t passes, void* results, uint32_t* cigar_pool,
                              uint64_t cigar_cap, uint64_t* cigar_used, int32_t* scale, double* stats, char* err, size_t err_cap) {
    try {
        if (!devices || !n_devices || !o) throw ClqError(CLQ_E_INVALID, "clqh_align_reads_span: no devices / options");
        AlignerOptions opt;
        opt.max_reads = o->max_reads;
        opt.max_read_bytes = o->max_read_bytes;
        opt.max_read_len = o->max_read_len;
        opt.cigar_ops_per_read = o->cigar_ops_per_read;
        opt.n_slots = o->n_slots;
        opt.max_refs = std::max<uint32_t>(64, n_refs);
        if (const char* e = std::getenv("CLQ_SPAN_PACK2")) opt.pack2_upload = std::atoi(e) != 0;  // bench / test view of AlignerOptions::pack2_upload
        ShardedAligner sh(std::vector<int>(devices, devices + n_devices), opt);
        std::vector<Reference> refs;
        for (uint32_t r = 0; r < n_refs; r++)
            refs.push_back({Bytes(ref_bytes + ref_off[r], ref_bytes + ref_off[r + 1]), to_bytes("ref" + std::to_string(r))});
        sh.set_references(ReferenceManager(std::move(refs)));
        ReadSpan span;
        span.bytes = read_bytes; span.off = read_off; span.n = n_reads; span.fixed_ref = fixed_ref;
        // experiment knobs of this test / bench view (the C++ API takes them as ReadSpan::order and Aligner::ctx() options)
        if (const char* e = std::getenv("CLQ_SPAN_ORDER")) {
            const std::string v(e);
            span.order = v == "front" ? ReadSpan::Order::Front : v == "longest" ? ReadSpan::Order::LongestFirst : v == "two" ? ReadSpan::Order::TwoEnded : ReadSpan::Order::Auto;
        }
        for (const char* key : {"serialize_slots", "max_scratch_bytes", "no_overlap"}) {
            std::string envn = "CLQ_SPAN_";
            for (const char* c = key; *c; c++) envn += (char)std::toupper((unsigned char)*c);
            if (const char* e = std::getenv(envn.c_str()))
                for (size_t k = 0; k < sh.n_devices(); k++)
                    if (clq_set_option(sh.aligner(k).ctx(), key, std::atoll(e)) != CLQ_OK) throw ClqError(CLQ_E_INVALID, std::string("bad ") + envn);
        }
        SpanOutput out;
        out.results = static_cast<clq_result_t*>(results);
        out.cigar_pool = cigar_pool;
        out.cigar_cap = cigar_cap;
        const AffineScoring sc{match_score, mismatch_score, special_score, gap_open, gap_extend, final_gap_multiplier};
        SpanStats st;
        // pass 0 warms up (page-locks the staging buffers, sizes the device scratch); the last pass is the one reported
        for (uint32_t k = 0; k < std::max<uint32_t>(1, passes); k++) st = sh.align_reads_span(span, sc, o->fast_lookup != 0, out, o->fillers_per_device, o->extract_tags != 0);
        if (cigar_used) *cigar_used = out.cigar_used;
        if (scale) *scale = out.scale;
        if (stats) {
            stats[0] = st.total.seconds; stats[1] = st.total.setup_seconds; stats[2] = st.fill_seconds; stats[3] = st.sink_seconds;
            stats[4] = (double)st.total.reads; stats[5] = (double)st.total.aligned; stats[6] = (double)st.total.dropped; stats[7] = (double)st.total.batches;
            for (uint32_t d = 0; d < n_devices; d++) {
                stats[8 + 3 * d] = st.device_kernel_ms[d]; stats[9 + 3 * d] = (double)st.device_reads[d]; stats[10 + 3 * d] = (double)st.device_cells[d];
            }
        }
        return CLQ_OK;
    } catch (const ClqError& e) {
        if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return e.code ? e.code : CLQ_E_INVALID;
    } catch (const std::exception& e) {
        if (err && err_cap) { std::strncpy(err, e.what(), err_cap - 1); err[err_cap - 1] = 0; }
        return CLQ_E_INVALID;
    }
}

uint64_t clqh_span_claims(const uint64_t* read_off, uint64_t n_reads, uint32_t n_devices, int32_t claimers_per_device, uint64_t max_reads,
                          uint64_t max_read_bytes, int32_t order, uint64_t* claims, uint64_t cap, int32_t* resolved_order) {
    const ReadSpan::Order ord = order == 1 ? ReadSpan::Order::Front : order == 2 ? ReadSpan::Order::LongestFirst : order == 3 ? ReadSpan::Order::TwoEnded : ReadSpan::Order::Auto;
    SpanClaimer cl(read_off, n_reads, n_devices, claimers_per_device, max_reads, max_read_bytes, ord);
    if (resolved_order) *resolved_order = cl.order() == ReadSpan::Order::Front ? 1 : cl.order() == ReadSpan::Order::LongestFirst ? 2 : cl.order() == ReadSpan::Order::TwoEnded ? 3 : 0;
    uint64_t k = 0, lo, hi, lo2, hi2;
    while (cl.claim(lo, hi, lo2, hi2)) {
        if (claims && k < cap) { claims[4 * k] = lo; claims[4 * k + 1] = hi; claims[4 * k + 2] = lo2; claims[4 * k + 3] = hi2; }
        k++;
    }
    return k;
}

}  // extern "C"

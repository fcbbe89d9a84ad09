// clq_align -- a small driver over the C++ host layer (include/clique_host.hpp): FASTA references + FASTQ / one-read-per-
// line input -> SAM text, through ShardedAligner::align_reads (the batch loop that replaces align_reads' par_bridge closure,
// alignment_functions.rs:135-249).  It exists to exercise and time the host layer end to end; it is not a port of the
// reference's CLI (main.rs), YAML layouts or BAM writer.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

#include "../../../include/clique_host.hpp"

using namespace clique;

namespace {

struct Args {
    std::string refs, reads, reads2, out = "-", stats_json, umi_symbols = "0123456789", layout;
    std::vector<int> gpus = {0};
    uint32_t batch = 1u << 18, max_read_len = 1u << 16, cigar_ops_per_read = 32, max_reference_multiplier = 2, threads = 1;
    int bam_level = 6;
    uint64_t synthetic = 0;  // --synthetic N: N in-memory noisy copies of the references instead of a read file
    bool exhaustive = false, tags = true, quiet_sam = false, rust_bio = false, slow_sam = false, unknown_strand = false;
    AffineScoring scoring = AffineScoring::align_reads_default();
};

[[noreturn]] void usage(const char* msg) {
    if (msg) std::fprintf(stderr, "clq_align: %s\n", msg);
    std::fprintf(stderr,
                 "usage: clq_align --refs refs.fa --reads reads.fastq|reads.txt [--reads2 r2.fastq --layout 1F,2C] [--out out.sam|out.bam|-] [--bam-level 0..9]\n"
                 "                 [--gpus 0,1,..] [--batch N] [--exhaustive] [--scoring match,mismatch,special,open,extend,final_mult]\n"
                 "                 [--umi-symbols 012] [--no-tags] [--no-sam] [--slow-sam] [--rust-bio] [--unknown-strand] [--max-read-len N] [--max-reference-multiplier N]\n"
                 "                 [--cigar-ops-per-read N] [--stats-json path] [--threads T (SAM text built by T host threads)]\n"
                 "                 [--synthetic N (N in-memory noisy copies of the references, 0.3%% substitutions, instead of --reads)]\n");
    std::exit(2);
}

Args parse(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        const std::string k = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) usage(("missing value for " + k).c_str()); return argv[++i]; };
        if (k == "--refs") a.refs = val();
        else if (k == "--reads") a.reads = val();
        else if (k == "--reads2") a.reads2 = val();
        else if (k == "--layout") a.layout = val();
        else if (k == "--out") a.out = val();
        else if (k == "--stats-json") a.stats_json = val();
        else if (k == "--umi-symbols") a.umi_symbols = val();
        else if (k == "--batch") a.batch = (uint32_t)std::stoul(val());
        else if (k == "--max-read-len") a.max_read_len = (uint32_t)std::stoul(val());
        else if (k == "--cigar-ops-per-read") a.cigar_ops_per_read = (uint32_t)std::stoul(val());
        else if (k == "--max-reference-multiplier") a.max_reference_multiplier = (uint32_t)std::stoul(val());
        else if (k == "--threads") a.threads = std::max<uint32_t>(1, (uint32_t)std::stoul(val()));
        else if (k == "--synthetic") a.synthetic = std::stoull(val());
        else if (k == "--bam-level") a.bam_level = std::stoi(val());
        else if (k == "--exhaustive") a.exhaustive = true;
        else if (k == "--no-tags") a.tags = false;
        else if (k == "--rust-bio") a.rust_bio = true;
        else if (k == "--slow-sam") a.slow_sam = true;
        else if (k == "--unknown-strand") a.unknown_strand = true;
        else if (k == "--no-sam") a.quiet_sam = true;
        else if (k == "--gpus") {
            a.gpus.clear();
            std::stringstream ss(val());
            std::string t;
            while (std::getline(ss, t, ',')) a.gpus.push_back(std::stoi(t));
        } else if (k == "--scoring") {
            std::stringstream ss(val());
            std::string t;
            double v[6];
            int n = 0;
            while (n < 6 && std::getline(ss, t, ',')) v[n++] = std::stod(t);
            if (n != 6) usage("--scoring needs six comma-separated numbers");
            a.scoring = {v[0], v[1], v[2], v[3], v[4], v[5]};
        } else usage(("unknown option " + k).c_str());
    }
    if (a.refs.empty() || (a.reads.empty() && !a.synthetic)) usage("--refs and --reads (or --synthetic N) are required");
    return a;
}

// FASTQ (4-line records, '@' header) or plain text (one read per line)
class ReadFile {
public:
    explicit ReadFile(const std::string& path) : in_(path), n_(0) {
        if (!in_) throw std::runtime_error("Unable to open input file " + path);
        const int c = in_.peek();
        fastq_ = c == '@';
    }
    bool next(FastqRecord& r) {
        std::string l1, l2, l3, l4;
        if (fastq_) {
            if (!std::getline(in_, l1) || !std::getline(in_, l2) || !std::getline(in_, l3) || !std::getline(in_, l4)) return false;
            strip(l1); strip(l2); strip(l4);
            const size_t e = l1.find_first_of(" \t");
            r.id = l1.substr(1, e == std::string::npos ? std::string::npos : e - 1);
            r.seq.assign(l2.begin(), l2.end());
            r.qual.assign(l4.begin(), l4.end());
        } else {
            do { if (!std::getline(in_, l2)) return false; strip(l2); } while (l2.empty());
            r.id = "read" + std::to_string(n_);
            r.seq.assign(l2.begin(), l2.end());
            r.qual.assign(l2.size(), (uint8_t)'H');
        }
        n_++;
        return true;
    }

private:
    static void strip(std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); }
    std::ifstream in_;
    bool fastq_;
    uint64_t n_;
};

std::vector<ReadPosition> parse_layout(const std::string& s) {
    std::vector<ReadPosition> lay;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) {
        ReadPosition p;
        if (item.rfind("S:", 0) == 0) { p.kind = ReadPosition::Spacer; p.spacer_sequence = item.substr(2); }
        else if (item.size() == 2 && (item[0] == '1' || item[0] == '2')) {
            p.kind = item[0] == '1' ? ReadPosition::Read1 : ReadPosition::Read2;
            p.orientation = item[1] == 'F' ? AlignedReadOrientation::Forward
                          : item[1] == 'R' ? AlignedReadOrientation::Reverse
                          : item[1] == 'C' ? AlignedReadOrientation::ReverseComplement : AlignedReadOrientation::Unknown;
        } else usage("bad --layout item (want 1F / 2R / 2C / S:ACGT)");
        lay.push_back(p);
    }
    return lay;
}

}  // namespace

int main(int argc, char** argv) {
    const Args a = parse(argc, argv);
    try {
        const ReferenceManager rm = ReferenceManager::from_fa_file(a.refs);
        if (rm.references.empty()) throw std::runtime_error("no references in " + a.refs);
        AlignerOptions opt;
        opt.max_reads = a.batch;
        // align_reads' drop rule: reads of (longest_ref + 1) * max_reference_multiplier bases or more are dropped with a
        // warning (alignment_functions.rs:122,147,240-247)
        const uint64_t max_read_size = (uint64_t)(rm.longest_ref + 1) * a.max_reference_multiplier;
        opt.max_read_len = (uint32_t)std::min<uint64_t>(a.max_read_len, max_read_size);
        opt.max_read_bytes = (uint64_t)a.batch * opt.max_read_len;
        opt.cigar_ops_per_read = a.cigar_ops_per_read;
        ShardedAligner sh(a.gpus, opt);
        sh.set_references(rm);

        // --synthetic: reads generated up front in host memory (xorshift, 0.3 % substitutions, tag symbols and N filled with bases)
        std::vector<Bytes> synth;
        if (a.synthetic) {
            uint64_t st = 0x9E3779B97F4A7C15ull;
            auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
            const size_t pool = (size_t)std::min<uint64_t>(a.synthetic, 1u << 16);
            for (size_t i = 0; i < pool; i++) {
                const Bytes& ref = rm.references[rnd() % rm.references.size()].sequence;
                Bytes r = ref;
                for (auto& b : r) if (b < 58 || b == 'N' || rnd() % 1000 < 3) b = "ACGT"[rnd() & 3];
                synth.push_back(std::move(r));
            }
        }
        uint64_t synth_next = 0;
        const ReadSource synth_source = [&](ReadBatch& b) -> bool {
            while (synth_next < a.synthetic) {
                const Bytes& r = synth[synth_next % synth.size()];
                char nm[24];
                int nl = 23;
                uint64_t v = synth_next;
                do { nm[nl--] = (char)('0' + v % 10); v /= 10; } while (v);
                nm[nl] = 's';
                if (!b.push(nm + nl, (size_t)(24 - nl), r.data(), r.size())) return true;
                synth_next++;
            }
            return false;
        };
        std::unique_ptr<ReadFile> f1p;
        if (!a.synthetic) f1p = std::make_unique<ReadFile>(a.reads);
        ReadFile* f1ptr = f1p.get();
        std::unique_ptr<ReadFile> f2;
        if (!a.reads2.empty() && !a.synthetic) f2 = std::make_unique<ReadFile>(a.reads2);
        const std::vector<ReadPosition> layout = a.layout.empty() ? std::vector<ReadPosition>{} : parse_layout(a.layout);
        ReadSetContainer pending;
        bool have_pending = false;
        uint64_t too_big = 0;
        // the input side of the loop (ReadIterator + MergedReadSequence, read_strategies/read_set.rs:60-125, merger.rs:196-330):
        // records are concatenated per the layout straight into the pinned batch
        const ReadSource source = [&](ReadBatch& b) -> bool {
            for (;;) {
                if (!have_pending) {
                    if (!f1ptr->next(pending.read_one)) return false;
                    if (f2) {
                        FastqRecord r2;
                        if (!f2->next(r2)) return false;
                        pending.read_two = std::move(r2);
                    }
                    have_pending = true;
                }
                bool ok;
                if (layout.empty()) ok = b.push(pending.read_one.id, pending.read_one.seq.data(), pending.read_one.seq.size(), pending.read_one.qual.data());
                else {
                    const MergedSequence m = merge_reads_by_concatenation(pending, layout);
                    ok = b.push(pending.read_one.id, m.read_bases.data(), m.read_bases.size(), m.read_quals.size() == m.read_bases.size() ? m.read_quals.data() : nullptr);
                }
                if (!ok) {
                    if (b.size() == 0) { too_big++; have_pending = false; continue; }  // does not fit an empty batch: skip it
                    return true;  // batch full: keep the record for the next one
                }
                have_pending = false;
            }
        };

        std::ofstream fout;
        std::ostream* out = &std::cout;
        if (a.out != "-") { fout.open(a.out); if (!fout) throw std::runtime_error("Unable to open " + a.out); out = &fout; }
        const std::vector<std::string> names = rm.names();
        // --out x.bam: the BAM container (BamFileAlignmentWriter, alignment_manager.rs:64-209): same records, binary, BGZF
        const bool as_bam = a.out.size() > 4 && a.out.compare(a.out.size() - 4, 4, ".bam") == 0 && !a.quiet_sam;
        if (as_bam) {
            std::vector<size_t> lens;
            for (const auto& r : rm.references) lens.push_back(r.sequence.size());
            std::string raw, z;
            bam::append_header(names, lens, raw);
            bam::bgzf_compress(raw.data(), raw.size(), z);
            out->write(z.data(), (std::streamsize)z.size());
        } else if (!a.quiet_sam) {
            *out << "@HD\tVN:1.6\n";
            for (const auto& r : rm.references) *out << "@SQ\tSN:" << to_string(r.name) << "\tLN:" << r.sequence.size() << "\n";
            *out << "@CO\tClique processed\n";
        }
        // batches complete in any order across GPUs; a small reorder buffer restores input order.  The SAM text of a batch is
        // built by `threads` host threads into per-thread buffers (contiguous read ranges) that are written in order and reused.
        std::map<uint64_t, std::pair<uint64_t, std::vector<std::string>>> done;  // first_index -> (reads, parts)
        uint64_t next_out = 0;
        std::vector<std::string> part;
        double sink_seconds = 0.0;
        const std::string syms = a.tags ? a.umi_symbols : std::string();
        const ResultSink sink = [&](const BatchView& v) {
            const auto t0 = std::chrono::steady_clock::now();
            const uint32_t nt = a.quiet_sam ? 0 : std::min<uint32_t>(a.threads, std::max<uint32_t>(1, v.size()));
            part.resize(nt);
            auto work = [&](uint32_t t) {
                const uint32_t lo = (uint32_t)((uint64_t)v.size() * t / nt), hi = (uint32_t)((uint64_t)v.size() * (t + 1) / nt);
                std::string& s = part[t];
                s.clear();
                if (as_bam) {  // records -> uncompressed BAM blocks -> BGZF members, all on this thread
                    std::string raw;
                    raw.reserve((size_t)(hi - lo) * 500);
                    for (uint32_t i = lo; i < hi; i++) {
                        if (!a.slow_sam) { v.append_bam_record(i, syms, raw); continue; }
                        const auto al = v.alignment(i);
                        if (al) bam::append_record(al->alignment->to_sam_record((int32_t)v.ref_index(i), v.align_reads_tags(i, syms), std::nullopt), raw);
                    }
                    if (!raw.empty()) bam::bgzf_compress(raw.data(), raw.size(), s, a.bam_level);
                    return;
                }
                for (uint32_t i = lo; i < hi; i++) {
                    if (!a.slow_sam) { v.append_sam_line(i, syms, names, s); continue; }
                    const auto al = v.alignment(i);  // --slow-sam: through the owned AlignmentResult / to_sam_record objects
                    if (!al) continue;
                    const TagMap tags = v.align_reads_tags(i, syms);
                    s += al->alignment->to_sam_record((int32_t)v.ref_index(i), tags, std::nullopt).to_sam_line(names);
                    s += '\n';
                }
            };
            std::vector<std::thread> th;
            for (uint32_t t = 1; t < nt; t++) th.emplace_back(work, t);
            if (nt) work(0);
            for (auto& t : th) t.join();
            if (v.batch->first_index == next_out) {
                for (const auto& s : part) out->write(s.data(), (std::streamsize)s.size());
                next_out += v.size();
            } else {
                done[v.batch->first_index] = {v.size(), std::move(part)};
                part.clear();
            }
            while (!done.empty() && done.begin()->first == next_out) {
                for (const auto& s : done.begin()->second.second) out->write(s.data(), (std::streamsize)s.size());
                next_out += done.begin()->second.first;
                done.erase(done.begin());
            }
            sink_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        };
        const AlignReadsStats st = sh.align_reads(a.synthetic ? synth_source : source, a.scoring, !a.exhaustive, sink, a.tags, a.rust_bio,
                                                  /*known_strand=*/!a.unknown_strand);  // the library orients the reads (alignment_functions.rs:549-558)
        if (as_bam) { std::string z; bam::bgzf_eof(z); out->write(z.data(), (std::streamsize)z.size()); }
        out->flush();
        char js[640];
        std::snprintf(js, sizeof(js),
                      "{\"reads\": %llu, \"aligned\": %llu, \"dropped\": %llu, \"batches\": %llu, \"cells\": %llu, \"seconds\": %.6f, "
                      "\"reads_per_s\": %.1f, \"gcups\": %.3f, \"gpus\": %zu, \"skipped_oversize\": %llu, \"sink_seconds\": %.6f, \"threads\": %u, \"setup_seconds\": %.6f}",
                      (unsigned long long)st.reads, (unsigned long long)st.aligned, (unsigned long long)st.dropped,
                      (unsigned long long)st.batches, (unsigned long long)st.cells, st.seconds, st.seconds > 0 ? st.reads / st.seconds : 0.0,
                      st.seconds > 0 ? st.cells / st.seconds / 1e9 : 0.0, sh.n_devices(), (unsigned long long)too_big, sink_seconds, a.threads, st.setup_seconds);
        std::fprintf(stderr, "%s\n", js);
        if (!a.stats_json.empty()) { std::ofstream sj(a.stats_json); sj << js << "\n"; }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "clq_align: %s\n", e.what());
        return 1;
    }
}

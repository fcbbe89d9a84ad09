//! Replacement for the per-read closure of `align_reads` (rust_cmd/src/alignment_functions.rs:135-249), to be added to the
//! reference crate as `rust_cmd/src/gpu_align_reads.rs` next to a `clq = { path = ".." }` dependency.
//!
//! Source only: written against clique's own types, it cannot be compiled outside that crate (and this image has no Rust
//! toolchain).  The compiled twin of this loop is `clique::Aligner::align_reads` (clique_b200/csrc/host/clique_host.cpp).
//!
//! What changes for the caller: nothing but the loop.  `read_iterator`, `rm`, `read_structure`, the `AffineScoring` literal
//! and the `BamFileAlignmentWriter` stay as they are; instead of one `align_to_reference_choices` call per read on a rayon
//! worker with a thread-local `Alignment<Ix3>`, reads are drained into two pinned batches that take turns on the GPU
//! (batch k+1 is copied and aligned while batch k is written out).  Output order is batch order; the reference's own BAM
//! order is already non-deterministic (rayon), so nothing downstream depends on it.
use std::collections::HashMap;

use clq::sys::{CLQ_NO_CANDIDATE, CLQ_OK, CLQ_READ_TOO_LONG, CLQ_TRACEBACK_DIVERGED};
use clq::{affine_scoring, rustbio_scoring, BatchResults, CigarOp, Context, Mode, ReadBatch};
// `ctx` comes from Context::new(device, clq_limits_t { max_reads: 1 << 19, max_read_bytes: 1 << 28, max_read_len: max_read_size as u32,
//                                                      max_refs, max_ref_bytes, cigar_pool_ops: 1 << 26, n_slots: 2 })

use crate::alignment::alignment_matrix::{AlignmentLocation, AlignmentResult, AlignmentTag};
use crate::alignment::scoring_functions::AffineScoring;
use crate::alignment_manager::{BamFileAlignmentWriter, OutputAlignmentWriter};
use crate::linked_alignment::orient_by_longest_segment;
use crate::utils::read_utils::reverse_complement;
use crate::read_strategies::read_disk_sorter::SortingReadSetContainer;
use crate::read_strategies::sequence_layout::SequenceLayout;
use crate::reference::fasta_reference::ReferenceManager;

/// The gapped strings, `path` and CIGAR of `perform_3d_global_traceback` (alignment/alignment_matrix.rs:941-1086) are a
/// deterministic function of the merged CIGAR and the two sequences: one `path` entry per unit step of the main loop, none
/// for the leading boundary run that the traceback emits after the loop.
fn rebuild(reference_name: &str, read_name: &str, reference: &[u8], read: &[u8], quals: Option<Vec<u8>>,
           cigar: impl Iterator<Item = CigarOp>, score: f64) -> AlignmentResult {
    let (mut ra, mut qa) = (Vec::with_capacity(reference.len() + read.len()), Vec::with_capacity(reference.len() + read.len()));
    let (mut tags, mut path) = (Vec::new(), Vec::new());
    let (mut x, mut y) = (0usize, 0usize);
    for (k, op) in cigar.enumerate() {
        match op {
            CigarOp::MatchMismatch(n) => {
                ra.extend_from_slice(&reference[x..x + n]);
                qa.extend_from_slice(&read[y..y + n]);
                path.extend((1..=n).map(|i| AlignmentLocation { x: x + i, y: y + i }));
                x += n;
                y += n;
                tags.push(AlignmentTag::MatchMismatch(n));
            }
            CigarOp::Del(n) => {
                ra.extend_from_slice(&reference[x..x + n]);
                qa.extend(std::iter::repeat(b'-').take(n));
                if k > 0 {
                    path.extend((1..=n).map(|i| AlignmentLocation { x: x + i, y }));
                }
                x += n;
                tags.push(AlignmentTag::Del(n));
            }
            CigarOp::Ins(n) => {
                ra.extend(std::iter::repeat(b'-').take(n));
                qa.extend_from_slice(&read[y..y + n]);
                if k > 0 {
                    path.extend((1..=n).map(|i| AlignmentLocation { x, y: y + i }));
                }
                y += n;
                tags.push(AlignmentTag::Ins(n));
            }
        }
    }
    AlignmentResult {
        reference_name: reference_name.to_string(),
        read_name: read_name.to_string(),
        reference_aligned: ra,
        read_aligned: qa,
        read_quals: quals,
        cigar_string: tags,
        path,
        score,
        reference_start: 0,
        read_start: 0,
        bounding_box: None,
    }
}

/// Names and qualities of the reads of a batch in flight (the sequences live in the pinned `ReadBatch`).
struct InFlight {
    names: Vec<String>,
    quals: Vec<Option<Vec<u8>>>,
}

pub fn align_reads_gpu<I>(ctx: &mut Context, read_iterator: I, rm: &ReferenceManager, read_structure: &SequenceLayout,
                          my_aff_score: &AffineScoring, max_reference_multiplier: usize, writer: &mut BamFileAlignmentWriter,
                          use_rust_bio_branch: bool)
where
    I: Iterator<Item = crate::merger::UnifiedRead>,
{
    // reference set in ascending index order: `ref_index` of a result is the position in this list
    let mut order: Vec<usize> = rm.references.keys().copied().collect();
    order.sort_unstable();
    ctx.set_references(order.iter().map(|i| rm.references[i].sequence.as_slice())).expect("clq_refs_set");
    ctx.build_kmer_index(rm.kmer_size as u32, rm.kmer_skip as u32).expect("clq_kmer_index_set");
    let ref_names: Vec<String> = order.iter().map(|i| String::from_utf8(rm.references[i].name.clone()).unwrap()).collect();

    // align_to_reference_choices' dispatch on the reference count (alignment_functions.rs:535-630), once per run:
    // one reference -> that reference (clique's own Gotoh, or the rust-bio semantics the reference runs today);
    // several -> quick_alignment_search with fast_lookup = true and the 0.90 vote share (:153, :604-628)
    let single = rm.references.len() == 1;
    let (scoring, mode) = if single && use_rust_bio_branch {
        let mut m = Mode::fixed_full().with_tags();
        m.flags |= clq::sys::CLQ_RUSTBIO;
        (rustbio_scoring(1, -1, -5, -1).expect("rust-bio scoring"), m)
    } else {
        let sc = affine_scoring(my_aff_score.match_score, my_aff_score.mismatch_score, my_aff_score.special_character_score,
                                my_aff_score.gap_open, my_aff_score.gap_extend, my_aff_score.final_gap_multiplier)
            .expect("scores must be dyadic rationals with gap_open < 0 (there is no CPU fallback)");
        (sc, if single { Mode::fixed_readlen().with_tags() } else { Mode::quick(0.90).with_tags() })
    };

    // The length filter of :147 runs on the device: the context was created with limits.max_read_len =
    // (longest_ref + 1) * multiplier, such reads come back with CLQ_READ_TOO_LONG and are dropped with the same warning.
    let max_read_size = (rm.longest_ref + 1) * max_reference_multiplier;
    let n_slots = ctx.n_slots().min(2).max(1);
    let mut meta: Vec<InFlight> = (0..n_slots).map(|_| InFlight { names: Vec::new(), quals: Vec::new() }).collect();
    let mut reads = read_iterator.peekable();
    let mut slot = 0usize;
    loop {
        // 1. the batch that used this slot one round ago: wait, rebuild, write -- while the other slot computes
        if ctx.in_flight(slot) {
            let res = ctx.wait(slot).expect("clq_wait");
            write_batch(&res, ctx.batch(slot), &meta[slot], rm, &order, &ref_names, read_structure, writer, max_read_size,
                        single && use_rust_bio_branch);
        }
        if reads.peek().is_none() {
            if (0..n_slots).all(|s| !ctx.in_flight(s)) {
                break;
            }
            slot = (slot + 1) % n_slots;
            continue;
        }
        // 2. refill the slot's pinned batch in place and submit it
        let m = &mut meta[slot];
        m.names.clear();
        m.quals.clear();
        let batch = ctx.batch_mut(slot).expect("slot is idle");
        batch.clear();
        while let Some(r) = reads.peek() {
            // align_to_reference_choices, :549-558: with one reference and an unknown strand the read is oriented on the host
            // first (orient_by_longest_segment against the reference's suffix table) and reverse-complemented when the
            // reverse strand shares more bases; the GPU then aligns the oriented bytes
            let oriented: Option<Vec<u8>> = if single && !read_structure.known_strand {
                let reference = &rm.references[&order[0]];
                let read_vec: Vec<u8> = r.seq().to_vec(); // the reference's helpers take &Vec<u8>
                let (forward, _, _) = orient_by_longest_segment(&read_vec, &reference.sequence, &reference.suffix_table);
                if forward { None } else { Some(reverse_complement(&read_vec)) }
            } else {
                None
            };
            let seq: &[u8] = oriented.as_deref().unwrap_or_else(|| r.seq());
            if !batch.push(seq, if single { Some(0) } else { None }) {
                if batch.len() == 0 {
                    // a read that does not fit an EMPTY batch never will (longer than the slot's max_read_bytes): drop it with
                    // the reference's own message (:240-247) instead of submitting empty batches forever
                    let r = reads.next().unwrap();
                    warn!("Dropped read {} is it's length {} exceeds 2x the reference length {}",
                          String::from_utf8_lossy(r.name()), r.seq().len(), max_read_size);
                    continue;
                }
                break;
            }
            let r = reads.next().unwrap();
            m.names.push(String::from_utf8(r.name().clone()).unwrap());
            m.quals.push(r.quals.clone());
        }
        if batch.len() > 0 {
            ctx.submit(slot, &scoring, mode).expect("clq_submit");
        }
        slot = (slot + 1) % n_slots;
    }
}

#[allow(clippy::too_many_arguments)]
fn write_batch(res: &BatchResults, batch: &ReadBatch, meta: &InFlight, rm: &ReferenceManager, order: &[usize], ref_names: &[String],
               read_structure: &SequenceLayout, writer: &mut BamFileAlignmentWriter, max_read_size: usize, rust_bio: bool) {
    for i in 0..batch.len() {
        let name = &meta.names[i];
        match res.status(i) {
            CLQ_OK => {}
            CLQ_READ_TOO_LONG => {
                warn!("Dropped read {} is it's length {} exceeds 2x the reference length {}", name, batch.read(i).len(), max_read_size);
                continue;
            }
            CLQ_NO_CANDIDATE => {
                debug!("Unable to create alignment for read {}", name);
                continue;
            }
            CLQ_TRACEBACK_DIVERGED => {
                // perform_3d_global_traceback would never return for this pair (it walks onto a band-skipped Up(0) cell)
                warn!("Dropped read {}: its traceback does not terminate in the reference implementation", name);
                continue;
            }
            other => panic!("libclq status {} for read {}", other, name),
        }
        let ri = res.ref_index(i);
        let reference = &rm.references[&order[ri]];
        let score = if rust_bio { 0.0 } else { res.score(i) }; // the rust-bio branch reports score 0.0 and an empty path (:571-583)
        let mut aln = rebuild(&ref_names[ri], name, &reference.sequence, batch.read(i), meta.quals[i].clone(), res.cigar(i), score);
        if rust_bio {
            aln.path.clear();
        }
        assert_eq!(aln.reference_aligned.len(), aln.read_aligned.len());

        // the e0..e9 tags: the walk already collected the read bytes under every '0'..'9' column of the reference, in
        // reference order; group them by the column's symbol (extract_tagged_sequences' digit keys, extractor.rs:271-332)
        let mut by_symbol: HashMap<u8, String> = HashMap::new();
        let mut k = 0usize;
        for b in reference.sequence.iter().filter(|b| b.is_ascii_digit()) {
            by_symbol.entry(*b).or_default().push(res.tag_bytes(i)[k] as char);
            k += 1;
        }
        let mut added_tags: HashMap<[u8; 2], String> = HashMap::new();
        let structure = read_structure.references.get(&aln.reference_name).unwrap();
        for (_, cfg) in structure.umi_configurations.iter() {
            if let Some(v) = by_symbol.get(&(cfg.symbol as u8)) {
                added_tags.insert([b'e', cfg.symbol as u8], v.clone());
            }
        }
        added_tags.insert([b'r', b'c'], 1.to_string());
        added_tags.insert([b'a', b'r'], aln.read_name.clone());
        added_tags.insert([b'r', b'm'], res.alignment_rate(i).to_string()); // get_reference_alignment_rate, counted in the walk
        added_tags.insert([b'a', b's'], aln.score.to_string());

        let read = SortingReadSetContainer::empty_tags(aln);
        writer.write_read(&read, &added_tags).expect("Unable to write a read to the BAM writer");
    }
}

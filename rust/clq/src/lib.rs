//! Safe wrapper over `clq-sys`: what the Rust host layer of clique needs to drive libclq -- pinned batch buffers, one
//! `Context` per GPU, asynchronous `submit` on a stream slot and `wait` for its results.  No alignment arithmetic lives here
//! and there is no CPU fallback: every constructor fails when libclq reports no CUDA device.
//!
//! Source only (no Rust toolchain in the build image); the same sequence of C calls is compiled and tested in
//! clique_b200/csrc/host/clique_host.cpp (`clique::Aligner`) and clique_b200/aligner.py.
use std::ffi::{c_void, CStr, CString};
use std::fmt;
use std::marker::PhantomData;
use std::ptr;

pub use clq_sys as sys;
use clq_sys::*;

/// A call-level error (negative return code of the C ABI) with the context's message when there is one.
#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}

impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        write!(f, "libclq error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}

pub type Result<T> = std::result::Result<T, Error>;

fn strerror(code: i32) -> String {
    unsafe { CStr::from_ptr(clq_strerror(code)) }.to_string_lossy().into_owned()
}

fn check(code: i32, ctx: *const clq_ctx) -> Result<()> {
    if code == 0 {
        return Ok(());
    }
    let mut message = strerror(code);
    if !ctx.is_null() {
        let detail = unsafe { CStr::from_ptr(clq_ctx_last_error(ctx)) }.to_string_lossy().into_owned();
        if !detail.is_empty() {
            message = format!("{message}: {detail}");
        }
    }
    Err(Error { code, message })
}

/// `AffineScoring` (rust_cmd/src/alignment/scoring_functions.rs:65-73) in the exact scaled-integer form the kernels use.
/// Fails with `CLQ_SCORING_NOT_REPRESENTABLE` (code 2) for non-dyadic scores or `gap_open >= 0`.
pub fn affine_scoring(match_score: f64, mismatch_score: f64, special_character_score: f64, gap_open: f64, gap_extend: f64,
                      final_gap_multiplier: f64) -> Result<clq_affine_t> {
    let mut out = clq_affine_t::default();
    let rc = unsafe {
        clq_affine_from_f64(match_score, mismatch_score, special_character_score, gap_open, gap_extend, final_gap_multiplier, &mut out)
    };
    if rc != 0 {
        return Err(Error { code: rc, message: strerror(rc) });
    }
    Ok(out)
}

/// The scoring `rust_bio_alignment` hard-codes (alignment_functions.rs:48-61) for the `CLQ_RUSTBIO` branch.
pub fn rustbio_scoring(match_score: i32, mismatch_score: i32, gap_open: i32, gap_extend: i32) -> Result<clq_affine_t> {
    let mut out = clq_affine_t::default();
    let rc = unsafe { clq_rustbio_scoring(match_score, mismatch_score, gap_open, gap_extend, &mut out) };
    if rc != 0 {
        return Err(Error { code: rc, message: strerror(rc) });
    }
    Ok(out)
}

pub fn device_count() -> i32 {
    unsafe { clq_device_count() }
}

/// Page-locked host memory from `clq_host_alloc`: H2D / D2H copies of a batch are asynchronous DMA only from pinned buffers.
pub struct PinnedBuf<T: Copy> {
    ptr: *mut T,
    cap: usize,
    _own: PhantomData<T>,
}

unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}

impl<T: Copy> PinnedBuf<T> {
    pub fn new(cap: usize) -> Result<Self> {
        let mut p: *mut c_void = ptr::null_mut();
        let bytes = cap.max(1) * std::mem::size_of::<T>();
        check(unsafe { clq_host_alloc(bytes, &mut p) }, ptr::null())?;
        unsafe { ptr::write_bytes(p as *mut u8, 0, bytes) }; // the slices below hand out initialised memory
        Ok(PinnedBuf { ptr: p as *mut T, cap, _own: PhantomData })
    }
    pub fn capacity(&self) -> usize {
        self.cap
    }
    pub fn as_ptr(&self) -> *const T {
        self.ptr
    }
    pub fn as_mut_ptr(&mut self) -> *mut T {
        self.ptr
    }
    /// The first `n` elements (the caller tracks how many it has written).
    pub fn slice(&self, n: usize) -> &[T] {
        assert!(n <= self.cap);
        unsafe { std::slice::from_raw_parts(self.ptr, n) }
    }
    pub fn slice_mut(&mut self, n: usize) -> &mut [T] {
        assert!(n <= self.cap);
        unsafe { std::slice::from_raw_parts_mut(self.ptr, n) }
    }
}

impl<T: Copy> Drop for PinnedBuf<T> {
    fn drop(&mut self) {
        unsafe { clq_host_free(self.ptr as *mut c_void) };
    }
}

/// One batch of reads in the layout `clq_submit` takes: raw ASCII bytes back to back, `n + 1` byte offsets, optionally the
/// fixed reference of every read.  Lives in pinned memory and is reused across batches (`clear`).
pub struct ReadBatch {
    bytes: PinnedBuf<u8>,
    off: PinnedBuf<u64>,
    fixed: PinnedBuf<i32>,
    n: usize,
    used: usize,
    has_fixed: bool,
}

impl ReadBatch {
    pub fn new(max_reads: usize, max_bytes: usize) -> Result<Self> {
        let mut b = ReadBatch {
            bytes: PinnedBuf::new(max_bytes)?,
            off: PinnedBuf::new(max_reads + 1)?,
            fixed: PinnedBuf::new(max_reads)?,
            n: 0,
            used: 0,
            has_fixed: false,
        };
        b.off.slice_mut(1)[0] = 0;
        Ok(b)
    }
    pub fn clear(&mut self) {
        self.n = 0;
        self.used = 0;
        self.has_fixed = false;
    }
    pub fn len(&self) -> usize {
        self.n
    }
    pub fn is_empty(&self) -> bool {
        self.n == 0
    }
    /// Appends a read; `false` when the batch is full (submit it and start the next one).
    pub fn push(&mut self, seq: &[u8], fixed_ref: Option<i32>) -> bool {
        if self.n + 1 > self.fixed.capacity() || self.used + seq.len() > self.bytes.capacity() {
            return false;
        }
        let start = self.used;
        self.bytes.slice_mut(start + seq.len())[start..].copy_from_slice(seq);
        self.used += seq.len();
        self.n += 1;
        let n = self.n;
        self.off.slice_mut(n + 1)[n] = self.used as u64;
        if let Some(r) = fixed_ref {
            self.fixed.slice_mut(n)[n - 1] = r;
            self.has_fixed = true;
        }
        true
    }
    pub fn read(&self, i: usize) -> &[u8] {
        let off = self.off.slice(self.n + 1);
        &self.bytes.slice(self.used)[off[i] as usize..off[i + 1] as usize]
    }
}

/// Exception list of a 2-bit packed batch (`Context::submit_packed2`): positions and bytes of everything that is not an
/// upper-case `ACGT`; reused across batches, must outlive the batch's `wait`.
#[derive(Default)]
pub struct Packed2Exceptions {
    pub pos: Vec<u64>,
    pub byte: Vec<u8>,
}

/// What to run on a batch: the `flags` word of `clq_submit`.
#[derive(Clone, Copy, Debug)]
pub struct Mode {
    pub flags: u32,
    /// `quick_alignment_search`'s vote share threshold (0.90 in the CLI, alignment_functions.rs:153)
    pub match_threshold: f64,
}

impl Mode {
    /// `align_two_strings_passed_matrix(.., &read.len())` against `fixed_ref[i]` (single-reference panels).
    pub fn fixed_readlen() -> Mode {
        Mode { flags: CLQ_SEARCH_FIXED | CLQ_BAND_READLEN, match_threshold: 0.0 }
    }
    /// `align_two_strings` (unbanded: bandwidth = max(L1, L2)).
    pub fn fixed_full() -> Mode {
        Mode { flags: CLQ_SEARCH_FIXED | CLQ_BAND_MAXLEN, match_threshold: 0.0 }
    }
    /// `quick_alignment_search` then `exhaustive_alignment_search` on the voted references (`fast_lookup = true`).
    pub fn quick(match_threshold: f64) -> Mode {
        Mode { flags: CLQ_SEARCH_QUICK | CLQ_BAND_READLEN, match_threshold }
    }
    pub fn exhaustive() -> Mode {
        Mode { flags: CLQ_SEARCH_EXHAUSTIVE | CLQ_BAND_READLEN, match_threshold: 0.0 }
    }
    /// Explicit bandwidth `k` of `perform_affine_alignment_bandwidth`.
    pub fn with_bandwidth(mut self, k: u32) -> Mode {
        self.flags = (self.flags & !CLQ_BAND_MASK & ((1 << CLQ_BAND_K_SHIFT) - 1)) | CLQ_BAND_K | (k << CLQ_BAND_K_SHIFT);
        self
    }
    /// Also return the read bytes aligned to the reference's `'0'..'9'` columns (the `e0..e9` tags of `align_reads`).
    pub fn with_tags(mut self) -> Mode {
        self.flags |= CLQ_EXTRACT_TAGS;
        self
    }
    pub fn score_only(mut self) -> Mode {
        self.flags |= CLQ_SCORE_ONLY;
        self
    }
}

/// Results of one batch, in the order the reads were pushed.
pub struct BatchResults {
    pub records: Vec<clq_result_t>,
    pub cigar_pool: Vec<u32>,
    pub tags: Vec<u8>,
    pub tag_stride: usize,
    pub scale: i32,
}

/// One run-length CIGAR element as the pool holds it.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum CigarOp {
    MatchMismatch(usize),
    Ins(usize),
    Del(usize),
}

impl BatchResults {
    pub fn status(&self, i: usize) -> u32 {
        self.records[i].status
    }
    /// `AlignmentResult.score` (exact: the kernels compute in integers scaled by a power of two).
    pub fn score(&self, i: usize) -> f64 {
        self.records[i].score_scaled as f64 / self.scale as f64
    }
    pub fn ref_index(&self, i: usize) -> usize {
        self.records[i].ref_index as usize
    }
    pub fn cigar(&self, i: usize) -> impl Iterator<Item = CigarOp> + '_ {
        let r = &self.records[i];
        self.cigar_pool[r.cigar_off as usize..r.cigar_off as usize + r.cigar_len as usize].iter().map(|w| {
            let n = (w >> 4) as usize;
            match w & 15 {
                CLQ_OP_M => CigarOp::MatchMismatch(n),
                CLQ_OP_I => CigarOp::Ins(n),
                _ => CigarOp::Del(n),
            }
        })
    }
    /// `get_reference_alignment_rate` (consensus/consensus_builders.rs:288-307): the value of the `rm` tag.
    pub fn alignment_rate(&self, i: usize) -> f64 {
        let r = &self.records[i];
        r.matches as f64 / (r.matches + r.mismatches) as f64
    }
    /// Bytes aligned to the tag columns of read `i`'s reference (one per `'0'..'9'` column, in reference order).
    pub fn tag_bytes(&self, i: usize) -> &[u8] {
        &self.tags[i * self.tag_stride..(i + 1) * self.tag_stride]
    }
}

/// One context per GPU.  Not `Sync`: a context is driven by one host thread; distinct contexts are independent
/// (read-sharded multi-GPU: one context and one thread per device, no collective).  The context owns one pinned `ReadBatch`
/// per stream slot; a slot's batch cannot be refilled while it is in flight (`batch_mut` fails with `CLQ_E_STATE`).
pub struct Context {
    raw: *mut clq_ctx,
    limits: clq_limits_t,
    pool: Vec<u32>, // landing buffer of clq_wait (cigar_pool_ops entries, allocated on first use)
    slots: Vec<Slot>,
    _not_sync: PhantomData<*mut ()>,
}

struct Slot {
    batch: ReadBatch,
    in_flight: Option<Submitted>,
}

#[derive(Clone, Copy)]
struct Submitted {
    n: usize,
    scale: i32,
    tags: bool,
}

unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32, limits: clq_limits_t) -> Result<Context> {
        let mut raw: *mut clq_ctx = ptr::null_mut();
        check(unsafe { clq_ctx_create(device, &limits, &mut raw) }, ptr::null())?;
        let mut ctx = Context { raw, limits, pool: Vec::new(), slots: Vec::new(), _not_sync: PhantomData };
        for _ in 0..limits.n_slots.clamp(1, 4) {
            let batch = ReadBatch::new(limits.max_reads as usize, limits.max_read_bytes as usize)?;
            ctx.slots.push(Slot { batch, in_flight: None });
        }
        Ok(ctx)
    }

    pub fn n_slots(&self) -> usize {
        self.slots.len()
    }

    /// The reference set, in the index order results refer to (`ReferenceManager.references` by ascending key).
    pub fn set_references<'a, I: IntoIterator<Item = &'a [u8]>>(&mut self, refs: I) -> Result<()> {
        let mut bytes: Vec<u8> = Vec::new();
        let mut off: Vec<u64> = vec![0];
        for r in refs {
            bytes.extend_from_slice(r);
            off.push(bytes.len() as u64);
        }
        check(unsafe { clq_refs_set(self.raw, (off.len() - 1) as u32, bytes.as_ptr(), off.as_ptr()) }, self.raw)
    }

    /// `ReferenceManager::unique_kmers` (reference/fasta_reference.rs:159-202); the CLI uses k = 8, skip = 4 (main.rs:271).
    pub fn build_kmer_index(&mut self, k: u32, skip: u32) -> Result<()> {
        check(unsafe { clq_kmer_index_set(self.raw, k, skip) }, self.raw)
    }

    pub fn in_flight(&self, slot: usize) -> bool {
        self.slots[slot].in_flight.is_some()
    }

    /// The slot's batch, to be filled for the next `submit`.  Fails while the slot is in flight: the device is still
    /// reading these buffers.
    pub fn batch_mut(&mut self, slot: usize) -> Result<&mut ReadBatch> {
        if self.slots[slot].in_flight.is_some() {
            return Err(Error { code: CLQ_E_STATE, message: strerror(CLQ_E_STATE) });
        }
        Ok(&mut self.slots[slot].batch)
    }

    /// The slot's batch, read-only (valid at any time: the reads of the results `wait` returned).
    pub fn batch(&self, slot: usize) -> &ReadBatch {
        &self.slots[slot].batch
    }

    /// Enqueues the copies and kernels for the slot's batch on its stream and returns at once.
    pub fn submit(&mut self, slot: usize, scoring: &clq_affine_t, mode: Mode) -> Result<()> {
        if self.slots[slot].in_flight.is_some() {
            return Err(Error { code: CLQ_E_STATE, message: strerror(CLQ_E_STATE) });
        }
        let b = &self.slots[slot].batch;
        let fixed = if b.has_fixed { b.fixed.as_ptr() } else { ptr::null() };
        check(
            unsafe {
                clq_submit(self.raw, slot as i32, b.n as u32, b.bytes.as_ptr(), b.off.as_ptr(), fixed,
                           scoring as *const clq_affine_t as *const c_void, mode.flags, mode.match_threshold)
            },
            self.raw,
        )?;
        let sub = Submitted { n: b.n, scale: scoring.scale, tags: mode.flags & CLQ_EXTRACT_TAGS != 0 };
        self.slots[slot].in_flight = Some(sub);
        Ok(())
    }

    /// `submit` with the batch shipped 2-bit packed (`clq_pack2` + `clq_submit_packed2`): the host packs the staged bytes into
    /// `words` (a pinned buffer of at least `(bytes + 15) / 16` u32, owned by the caller and kept alive until `wait`), bytes
    /// outside `ACGT` travel in an exception list.  A quarter of the H2D bytes for one more host pass; identical results.
    pub fn submit_packed2(&mut self, slot: usize, scoring: &clq_affine_t, mode: Mode, words: &mut [u32], exc: &mut Packed2Exceptions) -> Result<()> {
        if self.slots[slot].in_flight.is_some() {
            return Err(Error { code: CLQ_E_STATE, message: strerror(CLQ_E_STATE) });
        }
        let b = &self.slots[slot].batch;
        if words.len() < (b.used + 15) / 16 {
            return Err(Error { code: CLQ_E_INVALID, message: strerror(CLQ_E_INVALID) });
        }
        let mut n_exc: u64 = 0;
        loop {
            let rc = unsafe {
                clq_pack2(b.bytes.as_ptr(), b.used as u64, words.as_mut_ptr(), exc.pos.as_mut_ptr(), exc.byte.as_mut_ptr(),
                          exc.pos.len() as u64, &mut n_exc)
            };
            if rc == CLQ_E_LIMIT {
                // n_exc = the capacity the list needs
                exc.pos.resize(n_exc as usize, 0);
                exc.byte.resize(n_exc as usize, 0);
                continue;
            }
            check(rc, self.raw)?;
            break;
        }
        let fixed = if b.has_fixed { b.fixed.as_ptr() } else { ptr::null() };
        let (ep, eb) = if n_exc > 0 { (exc.pos.as_ptr(), exc.byte.as_ptr()) } else { (ptr::null(), ptr::null()) };
        check(
            unsafe {
                clq_submit_packed2(self.raw, slot as i32, b.n as u32, words.as_ptr(), b.off.as_ptr(), ep, eb, n_exc, fixed,
                                   scoring as *const clq_affine_t as *const c_void, mode.flags, mode.match_threshold)
            },
            self.raw,
        )?;
        let sub = Submitted { n: b.n, scale: scoring.scale, tags: mode.flags & CLQ_EXTRACT_TAGS != 0 };
        self.slots[slot].in_flight = Some(sub);
        Ok(())
    }

    /// Blocks until the slot's work is done and returns its records; the slot's batch may be refilled afterwards.
    pub fn wait(&mut self, slot: usize) -> Result<BatchResults> {
        let sub = match self.slots[slot].in_flight.take() {
            Some(s) => s,
            None => return Err(Error { code: CLQ_E_STATE, message: strerror(CLQ_E_STATE) }),
        };
        let mut records = vec![clq_result_t::default(); sub.n];
        let cap = self.limits.cigar_pool_ops as usize;
        if self.pool.len() < cap {
            self.pool.resize(cap, 0);
        }
        let mut used: u64 = 0;
        check(unsafe { clq_wait(self.raw, slot as i32, records.as_mut_ptr(), self.pool.as_mut_ptr(), cap as u64, &mut used) }, self.raw)?;
        let pool = self.pool[..used as usize].to_vec();
        let (mut tags, mut stride) = (Vec::new(), 0u32);
        if sub.tags {
            check(unsafe { clq_tags_download(self.raw, slot as i32, ptr::null_mut(), 0, &mut stride) }, self.raw)?;
            tags = vec![0u8; sub.n * stride as usize];
            check(unsafe { clq_tags_download(self.raw, slot as i32, tags.as_mut_ptr(), tags.len() as u64, &mut stride) }, self.raw)?;
        }
        Ok(BatchResults { records, cigar_pool: pool, tags, tag_stride: stride as usize, scale: sub.scale })
    }

    pub fn stats(&mut self, slot: usize) -> Result<clq_stats_t> {
        let mut s = clq_stats_t::default();
        check(unsafe { clq_slot_stats(self.raw, slot as i32, &mut s) }, self.raw)?;
        Ok(s)
    }

    pub fn set_option(&mut self, key: &str, value: i64) -> Result<()> {
        let k = CString::new(key).expect("option names hold no NUL");
        check(unsafe { clq_set_option(self.raw, k.as_ptr(), value) }, self.raw)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        // the context goes first: clq_ctx_destroy synchronises the streams that may still read the pinned batches
        unsafe { clq_ctx_destroy(self.raw) };
    }
}

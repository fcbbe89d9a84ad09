//! Raw bindings of `include/clq.h`, one item per C item, same order.  Every entry point names the reference interface it
//! replaces in the header; nothing here does arithmetic.  Not compiled in the build image (no Rust toolchain): kept in step
//! with the header by tests/test_abi.py::test_rust_sys_crate_matches_header.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_void};

pub const CLQ_VERSION: i32 = 100;

// ---- per-read status (clq_result_t.status) ----
pub const CLQ_OK: u32 = 0;
pub const CLQ_READ_TOO_LONG: u32 = 1;
pub const CLQ_SCORING_NOT_REPRESENTABLE: u32 = 2;
pub const CLQ_TRACEBACK_DIVERGED: u32 = 3;
pub const CLQ_CIGAR_POOL_FULL: u32 = 4;
pub const CLQ_NO_CANDIDATE: u32 = 5;

// ---- call-level errors (negative return values) ----
pub const CLQ_E_INVALID: i32 = -1;
pub const CLQ_E_CUDA: i32 = -2;
pub const CLQ_E_NOMEM: i32 = -3;
pub const CLQ_E_LIMIT: i32 = -4;
pub const CLQ_E_STATE: i32 = -5;
pub const CLQ_E_UNSUPPORTED: i32 = -6;

// ---- flags for clq_submit / clq_launch ----
pub const CLQ_BAND_MAXLEN: u32 = 0;
pub const CLQ_BAND_READLEN: u32 = 1;
pub const CLQ_BAND_K: u32 = 2;
pub const CLQ_BAND_K_SHIFT: u32 = 8;
pub const CLQ_BAND_MASK: u32 = 3;
pub const CLQ_SEARCH_FIXED: u32 = 0 << 2;
pub const CLQ_SEARCH_EXHAUSTIVE: u32 = 1 << 2;
pub const CLQ_SEARCH_QUICK: u32 = 2 << 2;
pub const CLQ_SEARCH_MASK: u32 = 3 << 2;
pub const CLQ_SCORE_ONLY: u32 = 1 << 4;
pub const CLQ_CONVEX: u32 = 1 << 5;
pub const CLQ_EXTRACT_TAGS: u32 = 1 << 6;
pub const CLQ_RUSTBIO: u32 = 1 << 7;

// ---- CIGAR op codes in the pool: len << 4 | code ----
pub const CLQ_OP_M: u32 = 0;
pub const CLQ_OP_I: u32 = 1;
pub const CLQ_OP_D: u32 = 2;

/// Opaque context (one per GPU, not thread-safe).
#[repr(C)]
pub struct clq_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct clq_affine_t {
    pub scale: i32,
    pub match_: i32, // C field `match`
    pub mismatch: i32,
    pub special: i32,
    pub oe_in: i32,
    pub e_in: i32,
    pub oe_fin: i32,
    pub e_fin: i32,
    pub b0: i32,
    pub b1: i32,
    pub max_neg: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct clq_convex_t {
    pub match_: i32, // C field `match`
    pub mismatch: i32,
    pub special: i32,
    pub o1: i32,
    pub e1: i32,
    pub o2: i32,
    pub e2: i32,
    pub max_neg: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct clq_limits_t {
    pub max_reads: u32,
    pub max_read_bytes: u64,
    pub max_read_len: u32,
    pub max_refs: u32,
    pub max_ref_bytes: u64,
    pub cigar_pool_ops: u64,
    pub n_slots: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct clq_result_t {
    pub score_scaled: i32,
    pub ref_index: u32,
    pub cigar_off: u32,
    pub cigar_len: u32,
    pub status: u32,
    pub matches: u32,
    pub mismatches: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct clq_stats_t {
    pub kernel_ms: f32,
    pub dp_ms: f32,
    pub launches: u32,
    pub dp_launches: u32,
    pub cells: u64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub variant: u32,
    pub sub_batches: u32,
    pub pack_retries: u32,
    pub reserved0: u32,
}

extern "C" {
    pub fn clq_version() -> i32;
    pub fn clq_strerror(code: i32) -> *const c_char;
    pub fn clq_device_count() -> i32;
    pub fn clq_affine_from_f64(match_score: f64, mismatch_score: f64, special_character_score: f64, gap_open: f64,
                               gap_extend: f64, final_gap_multiplier: f64, out: *mut clq_affine_t) -> i32;
    pub fn clq_rustbio_scoring(match_score: i32, mismatch_score: i32, gap_open: i32, gap_extend: i32, out: *mut clq_affine_t) -> i32;
    pub fn clq_host_alloc(bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn clq_host_free(p: *mut c_void) -> i32;
    pub fn clq_ctx_create(device: i32, limits: *const clq_limits_t, out: *mut *mut clq_ctx) -> i32;
    pub fn clq_ctx_destroy(ctx: *mut clq_ctx);
    pub fn clq_ctx_last_error(ctx: *const clq_ctx) -> *const c_char;
    pub fn clq_refs_set(ctx: *mut clq_ctx, n_refs: u32, bytes: *const u8, off: *const u64) -> i32;
    pub fn clq_kmer_index_set(ctx: *mut clq_ctx, k: u32, skip: u32) -> i32;
    pub fn clq_submit(ctx: *mut clq_ctx, slot: i32, n_reads: u32, read_bytes: *const u8, read_off: *const u64,
                      fixed_ref: *const i32, scoring: *const c_void, flags: u32, match_threshold: f64) -> i32;
    pub fn clq_wait(ctx: *mut clq_ctx, slot: i32, results: *mut clq_result_t, cigar_pool: *mut u32, cigar_cap: u64,
                    cigar_used: *mut u64) -> i32;
    pub fn clq_tags_download(ctx: *mut clq_ctx, slot: i32, tags: *mut u8, cap: u64, tag_stride: *mut u32) -> i32;
    pub fn clq_upload(ctx: *mut clq_ctx, slot: i32, n_reads: u32, read_bytes: *const u8, read_off: *const u64,
                      fixed_ref: *const i32) -> i32;
    pub fn clq_launch(ctx: *mut clq_ctx, slot: i32, scoring: *const c_void, flags: u32, match_threshold: f64) -> i32;
    pub fn clq_download(ctx: *mut clq_ctx, slot: i32) -> i32;
    pub fn clq_sync(ctx: *mut clq_ctx, slot: i32) -> i32;
    pub fn clq_slot_stats(ctx: *mut clq_ctx, slot: i32, out: *mut clq_stats_t) -> i32;
    pub fn clq_set_option(ctx: *mut clq_ctx, key: *const c_char, value: i64) -> i32;
    // 2-bit packed read ingestion (host packer + the packed forms of clq_upload / clq_submit)
    pub fn clq_pack2(bytes: *const u8, n_bytes: u64, packed: *mut u32, exc_pos: *mut u64, exc_byte: *mut u8, exc_cap: u64,
                     n_exc: *mut u64) -> i32;
    pub fn clq_upload_packed2(ctx: *mut clq_ctx, slot: i32, n_reads: u32, packed: *const u32, read_off: *const u64,
                              exc_pos: *const u64, exc_byte: *const u8, n_exc: u64, fixed_ref: *const i32) -> i32;
    pub fn clq_submit_packed2(ctx: *mut clq_ctx, slot: i32, n_reads: u32, packed: *const u32, read_off: *const u64,
                              exc_pos: *const u64, exc_byte: *const u8, n_exc: u64, fixed_ref: *const i32,
                              scoring: *const c_void, flags: u32, match_threshold: f64) -> i32;
}

// clq_kernels.cuh -- sm_100a kernels of libclq: batched global affine-gap Gotoh alignment.
//
// One G-lane sub-warp group aligns one (read, reference) pair as an anti-diagonal wavefront:
// lane l owns C consecutive read columns in registers, the reference streams past row by row from shared
// memory, lane l works on row x = t - l at step t, and the right-edge values of a lane's stripe travel to the
// next lane with __shfl_up_sync.  Reads longer than G*C columns are processed in column stripes whose boundary
// column is parked in a small global scratch.  With TB the kernel also records 4 direction bits per cell
// (coalesced, one G*C/8-word row per step) into the task's slot of the bits scratch; walk_kernel (one thread per
// pair) then walks them back and emits the run-length CIGAR.
//
// Arithmetic restated from the reference (file:line relative to rust_cmd/src/ of mckennalab/clique):
//   update_3d_score                alignment/alignment_matrix.rs:618-665
//   three_way_max_and_direction    alignment/alignment_matrix.rs:671-683   (ties: Diag > Left > Up)
//   boundary init + f64 band       alignment/alignment_matrix.rs:385-424
//   perform_3d_global_traceback    alignment/alignment_matrix.rs:941-1086  (start layer = last max)
//   simplify_cigar_string          alignment_manager.rs:386-423
//   match_mismatch                 alignment/scoring_functions.rs:100-102
//
// Value recurrence used here (identical values because gap_open < 0 makes E+le >= E+x1):
//   M = B[x-1,y-1] + m,  E = max(E[x-1,y] + le, B[x-1,y] + x1),  F = max(F[x,y-1] + le, B[x,y-1] + x1),
//   B = max(M, E, F).
// Direction bits per cell (x,y), enough to replay all three traceback layers exactly:
//   ext1 (bit 3) = E[x,y] extended E[x-1,y] (strictly better than opening)          -> T1 = Up
//   ext2 (bit 2) = F[x,y] extended F[x,y-1] (>= the E-open, > the M-open)            -> T2 = Left
//   eP   (bit 1) = E > max(M,F),  fM (bit 0) = F > M   =>  A = argmax(M,E,F) with priority M > F > E = eP ? E : (fM ? F : M)
//   T0[x,y] = A[x-1,y-1];  T1[x,y] = ext1 ? Up : (A[x-1,y]==F ? Left : Diag);  T2[x,y] = ext2 ? Left : (A[x,y-1]==E ? Up : Diag)
// Band-skipped ("stale") cells are (x <= K, y == L2); the walker recognises them by coordinates.
//
// FAST variant (uniform gap constants, <= 8 byte classes, int8 scores): substitution score by one PRMT from an 8-byte
// profile row, E and F kept shifted by -x1 so that every max-plus step is one DPX instruction
//   Eh = viaddmax(Eh_up, le, B_up); Fh = viaddmax(Fh_left, le, B_left); P = viaddmax(Fh, x1, M); B = viaddmax(Eh, x1, P)
// and the four direction bits are the sign bits of four differences, shifted into the row word with SHF.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/clq.h"

namespace clq {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kThreads = 128;  // 4 warps per CTA

struct TbRec;

struct KParams {
    const uint8_t* ref_bytes;
    const uint64_t* ref_off;
    uint32_t n_refs;
    const uint8_t* read_bytes;
    const uint64_t* read_off;
    uint32_t n_reads;
    const uint32_t* order;       // processing order over reads (nullptr = identity)
    const int32_t* ref_of_read;  // per-read reference (pair mode); < 0 => no candidate
    uint32_t n_tasks;
    uint32_t all_pairs;          // 1: task = (read, ref) over every reference, output -> scores
    const uint32_t* cand_mask;   // all_pairs: optional candidate bitmask, mask_words per read
    uint32_t mask_words;
    int32_t* scores;             // [n_reads * n_refs]
    clq_result_t* results;       // [n_reads]
    clq_affine_t sc;
    uint32_t band_mode;
    uint32_t band_k;             // CLQ_BAND_K: explicit bandwidth (generic kernels only)
    uint32_t max_read_len;
    uint32_t ref_sm_stride;      // bytes of shared memory per group
    uint32_t* bits;              // traceback bits scratch
    uint64_t bits_stride;        // words per task slot (uniform slots)
    const uint64_t* bits_off;    // or: word offset of every processing position's slot (variable-size slots), nullptr = uniform
    uint32_t* cig_scratch;
    uint32_t cig_stride;         // ops per group
    int32_t* col_scratch;        // stripe boundary column: 4 arrays of col_stride per group
    uint32_t col_stride;
    uint32_t* cigar_pool;
    uint64_t cigar_cap;
    unsigned long long* cigar_cursor;
    unsigned int* task_counter;
    unsigned long long* cells;
    const uint8_t* cls_lut;      // FAST: byte -> bits 0..2 read class (column of the profile), bits 3..6 reference class (row),
                                 //       bit 7 = a READ holding this byte cannot be scored exactly (rust-bio mode only)
    uint32_t tab[32];            // FAST: 16 profile rows (reference class) x 8 int8 scores (read class)
    uint32_t rustbio;            // CLQ_RUSTBIO: rust-bio global semantics (boundary init, tie order), see gotoh_kernel<.., RB>
    uint32_t debug_flags;        // experiments only: 1 = skip the traceback walk
    uint32_t task_base;          // first read (processing position) of this sub-batch
    uint32_t task_end;           // one past its last read (PACK kernel: two reads per task)
    struct TbRec* tb_rec;        // TB: one record per task of the sub-batch for walk_kernel
    const uint16_t* tag_slot;    // CLQ_EXTRACT_TAGS: per reference byte, rank among the reference's tag columns (0xffff = none)
    uint8_t* tags;               // [n_reads * tag_stride] read bytes aligned to the tag columns ('-' = deleted)
    uint32_t tag_stride;
    // retry pass after pack_adapt_kernel (clq_pack_adapt.cuh): the tasks are the sub-batch positions listed in retry_list, their
    // number is read from device memory (the host does not know it when it queues the launch)
    const uint32_t* retry_list;
    const unsigned int* n_tasks_dev;
};

// what the fill kernel leaves for the walker: one record per task of the sub-batch
struct TbRec {
    uint32_t ridx;   // read index (results slot)
    int32_t L1, L2;  // L1 < 0: nothing to walk (dropped read / no candidate)
    int32_t zK;      // start layer | stale rows K << 2 | (columns per lane of the narrow last stripe) / 8 << 20 | its index << 24
};

// Narrow last stripe (G >= 16 geometries of the FAST / PACK kernels): the last column stripe of a task holds
// R = L2 - (NS-1)*W <= W columns; instead of C columns per lane (most of them padding) every lane takes
// Cs = ceil(R / (8 G)) * 8 columns, so a 1100-column read on the (32,32) geometry costs 1024 + 256 columns, not 2048.
// The walker gets (Cs / 8, stripe index) in TbRec.zK.
template <int G>
__host__ __device__ __forceinline__ int narrow_cols(int R, int C) {
    const int cs = (R + G * 8 - 1) / (G * 8) * 8;
    return cs < C ? cs : C;
}
constexpr int kMaxNarrowStripe = 255;

__device__ __forceinline__ bool is_special(int c) { return c == 'N' || c < 58; }

// direction-bit slot of the read at slot index `slot` of the current sub-batch (processing position task_base + slot)
__device__ __forceinline__ size_t bits_slot(const uint64_t* bits_off, uint64_t bits_stride, uint32_t task_base, uint32_t slot) {
    return bits_off ? (size_t)(bits_off[task_base + slot] - bits_off[task_base]) : (size_t)slot * bits_stride;
}

// prmt.b32 in its default mode: selector nibble bits [2:0] pick a byte of {hi,lo}, bit 3 replicates that byte's sign.
// (__byte_perm masks bit 3 away, so the raw instruction is needed for the sign-extending byte lookup.)
__device__ __forceinline__ int prmt_s8(uint32_t lo, uint32_t hi, uint32_t sel) {
    int d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(sel));
    return d;
}

// rows 1..K whose band skips the last column (f64 centre, alignment/alignment_matrix.rs:413-417)
__device__ inline int stale_rows(int L1, int L2, uint32_t band_mode) {
    if (band_mode > CLQ_BAND_READLEN) return 0;  // unbanded (rust-bio mode)
    long long bw = band_mode == CLQ_BAND_READLEN ? L2 : (L1 > L2 ? L1 : L2);
    long long lim = (long long)L2 - bw;
    if (lim < 0) return 0;
    int K = 0;
    for (int x = 1; x <= L1; x++) {
        long long yc = (long long)(((double)x / (double)(L1 + 1)) * (double)(L2 + 1));
        if (yc <= lim) K = x; else break;
    }
    return K;
}

// perform_affine_alignment_bandwidth's band of row x (alignment/alignment_matrix.rs:413-424): cells y in [lo, hi] are
// computed, every other cell of the row keeps the fresh-matrix state (0,0,0) / Up(0).  The centre is computed in f64.
__device__ __forceinline__ void band_of_row(int x, int L1, int L2, long long bw, int& lo, int& hi) {
    const long long yc = (long long)(((double)x / (double)(L1 + 1)) * (double)(L2 + 1));
    const long long a = yc - bw, b = yc + bw;
    lo = (int)(a > 1 ? a : 1);
    hi = (int)((b < (long long)L2 + 1 ? b : (long long)L2 + 1) - 1);
}

// Time-transposed direction-bit store.  The walker follows a path that moves up one row per step, so it wants the
// words of consecutive steps side by side; the fill produces one row per step.  Each lane therefore parks its WPL words of
// the last 8 steps in shared memory ([word][step & 7][lane], conflict-free, private to the lane so no barrier) and
// writes them out as full 32-byte sectors: bits[block of 8 steps][lane][word][step & 7].
template <int G, int WPL>
__device__ __forceinline__ void tt_store(uint32_t* tt, uint32_t* bits_g, const uint32_t (&w)[WPL], int s, int Tb, int t, int lane,
                                         int gl, bool last_row) {
    const int ph = (t - 1) & 7;
#pragma unroll
    for (int k = 0; k < WPL; k++) tt[(k * 8 + ph) * 32 + lane] = w[k];
    if (ph == 7 || last_row) {
        uint32_t* dst = bits_g + ((size_t)(s * Tb + ((t - 1) >> 3)) * (G * WPL) + gl * WPL) * 8;
#pragma unroll
        for (int k = 0; k < WPL; k++) {
            const uint32_t* src = tt + (k * 8) * 32 + lane;
            *reinterpret_cast<uint4*>(dst + k * 8) = make_uint4(src[0], src[32], src[64], src[96]);
            *reinterpret_cast<uint4*>(dst + k * 8 + 4) = make_uint4(src[128], src[160], src[192], src[224]);
        }
    }
}

// word index of cell (x, y)'s nibble in the layout tt_store writes (lane ln owns column y, word k of its row)
template <int G, int WPL>
__device__ __forceinline__ size_t tt_index(int s, int Tb, int t, int ln, int k) {
    return ((size_t)(s * Tb + ((t - 1) >> 3)) * (G * WPL) + ln * WPL + k) * 8 + ((t - 1) & 7);
}

// Row-per-step layout (long-read geometries, G >= 16): one row of G*WPL words per (stripe, step), 128-bit stores of whole
// sectors [G x 4 words][G x (WPL-4) words].  The transposed layout only pays off when the walk is a visible share of the
// step (short reads); for long reads the extra shared-memory traffic costs more than the walk gains.
template <int G, int WPL>
__device__ __forceinline__ void row_store(uint32_t* bits_g, const uint32_t (&w)[WPL], int s, int T, int t, int gl, int nb) {
    uint32_t* row = bits_g + (size_t)(s * T + (t - 1)) * (G * WPL);
    if constexpr (WPL >= 4) {
        *reinterpret_cast<uint4*>(row + gl * 4) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
        for (int k = 4; k < WPL; k++)
            if (k < nb) row[G * 4 + gl * (WPL - 4) + (k - 4)] = w[k];
    } else if constexpr (WPL == 2) {
        *reinterpret_cast<uint2*>(row + gl * 2) = make_uint2(w[0], w[1]);
    } else {
#pragma unroll
        for (int k = 0; k < WPL; k++) row[gl * WPL + k] = w[k];
    }
}

template <int G, int WPL>
__device__ __forceinline__ size_t row_index(int s, int T, int t, int ln, int k) {
    const int in_row = (WPL >= 4) ? (k < 4 ? ln * 4 + k : G * 4 + ln * (WPL - 4) + (k - 4)) : ln * WPL + k;
    return (size_t)(s * T + (t - 1)) * (G * WPL) + in_row;
}

// which layout a geometry uses (host: clq_api.cu::bits_words_per_pair must agree)
#ifndef CLQ_TRANSPOSE_MAX_G
#define CLQ_TRANSPOSE_MAX_G 8
#endif
constexpr int kTransposeMaxG = CLQ_TRANSPOSE_MAX_G;
template <int G>
struct BitsLayout { static constexpr bool transposed = (G <= kTransposeMaxG); };

template <int G, int WPL>
__device__ __forceinline__ void bits_store(uint32_t* tt, uint32_t* bits_g, const uint32_t (&w)[WPL], int s, int T, int t, int lane, int gl,
                                           bool last_row, int nb = WPL) {
    if constexpr (BitsLayout<G>::transposed) tt_store<G, WPL>(tt, bits_g, w, s, (T + 7) >> 3, t, lane, gl, last_row);
    else row_store<G, WPL>(bits_g, w, s, T, t, gl, nb);
}

template <int G, int WPL>
__device__ __forceinline__ size_t bits_index(int s, int T, int t, int ln, int k) {
    if constexpr (BitsLayout<G>::transposed) return tt_index<G, WPL>(s, (T + 7) >> 3, t, ln, k);
    else return row_index<G, WPL>(s, T, t, ln, k);
}

// One wavefront step of one lane: C cells of row x.
// BAND: columns outside [jlo, jhi] (the explicit band of this row) are not computed: they hold the fresh-matrix state.
template <int C, bool TB, bool FIN, bool LAST, bool BAND = false>
__device__ __forceinline__ void row_step(int (&E)[C], int (&B)[C], const int (&bq)[C], uint32_t (&w)[C / 8],
                                         int& Fl, int& El, int& Ml, int& Bl, int diag, int rcode, int mt, int mm,
                                         const clq_affine_t& sc, bool own_last, int jL, int& capM, int& capE, int& capF,
                                         int jlo = 0, int jhi = C) {
    const int x1row = LAST ? sc.oe_fin : sc.oe_in;
    const int lerow = LAST ? sc.e_fin : sc.e_in;
#pragma unroll
    for (int k = 0; k < C / 8; k++) w[k] = 0;
#pragma unroll
    for (int j = 0; j < C; j++) {
        int m = (bq[j] == rcode) ? mt : mm;
        if (bq[j] & 0x100) m = sc.special;
        int x1c = x1row, lec = lerow;
        if (FIN && !LAST) {
            if (own_last && j == jL) { x1c = sc.oe_fin; lec = sc.e_fin; }
        }
        int Mv = diag + m;
        const int Eext = E[j] + lec, Eopen = B[j] + x1c;
        int Ev = max(Eext, Eopen);
        const int Fext = Fl + lec, Fopen = Bl + x1c;
        int Fv = max(Fext, Fopen);
        const int Pv = max(Mv, Fv);
        int Bv = max(Pv, Ev);
        if (BAND) {
            if (j < jlo || j > jhi) { Mv = 0; Ev = 0; Fv = 0; Bv = 0; }
        }
        if (TB) {
            const bool ext1 = Eext > Eopen;
            const bool ext2 = (Fext >= El + x1c) && (Fext > Ml + x1c);
            const uint32_t nib = (ext1 ? 8u : 0u) | (ext2 ? 4u : 0u) | ((Ev > Pv) ? 2u : 0u) | ((Fv > Mv) ? 1u : 0u);
            w[j >> 3] |= nib << (28 - 4 * (j & 7));
        }
        diag = B[j];
        E[j] = Ev;
        B[j] = Bv;
        Fl = Fv; Bl = Bv; El = Ev; Ml = Mv;
        if (LAST) {
            if (own_last && j == jL) { capM = Mv; capE = Ev; capF = Fv; }
        }
    }
}

constexpr int kLutBytes = 256, kTabBytes = 128;  // FAST kernels: class LUT + 16-row profile table at the start of smem

// FAST wavefront step: Eh/Fh are E/F shifted by -x1 (uniform gap constants), m by PRMT, max-plus by DPX, bits by SHF.
// RB (rust-bio global, oracle/clq_oracle.h::orc_rustbio_global): same values, but a gap extends only when strictly better
// than opening from the best state S = max(M, I, D) of the neighbour: ext2 <=> F_left + e > B_left + o + e.
template <int C, bool TB, bool LAST, bool RB, int JB>
__device__ __forceinline__ void row_blocks_fast(int (&Eh)[C], int (&B)[C], const int (&sel)[C], uint32_t (&w)[C / 8], int& Fh, int& Ehl, int& Ml,
                                                int& Bl, int diag, uint32_t tlo, uint32_t thi, int le, int x1, bool own_last, int jL, int& capM,
                                                int& capE, int& capF, int nb) {
    // blocks of 8 columns, nested (block JB + 1 inside block JB's `if`): a narrow last stripe (only nb blocks per lane are real)
    // leaves through one forward branch instead of a reconvergence region per block
    if constexpr (JB < C / 8) {
        if (JB < nb) {
            constexpr int jb = JB;
            const int x1m1 = x1 - 1;
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int j = jb * 8 + jj;
                const int m = prmt_s8(tlo, thi, (uint32_t)sel[j]);
                const int Mv = diag + m;
                const int EhU = Eh[j], BU = B[j];
                const int Ehn = __viaddmax_s32(EhU, le, BU);
                int d2 = 0;
                if (TB) d2 = RB ? (Bl - Fh - le) : (__viaddmax_s32(Ehl, x1m1, Ml) - Fh - le);  // < 0  <=>  F extends (>= E-open, > M-open)
                const int Fhn = __viaddmax_s32(Fh, le, Bl);
                const int Pv = __viaddmax_s32(Fhn, x1, Mv);
                const int Bn = __viaddmax_s32(Ehn, x1, Pv);
                if (TB) {
                    uint32_t acc = w[jb];
                    acc = __funnelshift_l((uint32_t)(BU - Ehn), acc, 1);  // ext1: Eh_up + le > B_up
                    acc = __funnelshift_l((uint32_t)d2, acc, 1);          // ext2
                    acc = __funnelshift_l((uint32_t)(Pv - Bn), acc, 1);   // eP: E > max(M,F)
                    acc = __funnelshift_l((uint32_t)(Mv - Pv), acc, 1);   // fM: F > M
                    w[jb] = acc;
                }
                diag = BU;
                Eh[j] = Ehn;
                B[j] = Bn;
                Fh = Fhn; Ehl = Ehn; Ml = Mv; Bl = Bn;
                if (LAST) {
                    if (own_last && j == jL) { capM = Mv; capE = Ehn + x1; capF = Fhn + x1; }
                }
            }
            row_blocks_fast<C, TB, LAST, RB, JB + 1>(Eh, B, sel, w, Fh, Ehl, Ml, Bl, diag, tlo, thi, le, x1, own_last, jL, capM, capE, capF, nb);
        }
    }
}

template <int C, bool TB, bool LAST, bool RB = false>
__device__ __forceinline__ void row_step_fast(int (&Eh)[C], int (&B)[C], const int (&sel)[C], uint32_t (&w)[C / 8], int& Fh,
                                              int& Ehl, int& Ml, int& Bl, int diag, uint32_t tlo, uint32_t thi, int le, int x1,
                                              bool own_last, int jL, int& capM, int& capE, int& capF, int nb) {
    row_blocks_fast<C, TB, LAST, RB, 0>(Eh, B, sel, w, Fh, Ehl, Ml, Bl, diag, tlo, thi, le, x1, own_last, jL, capM, capE, capF, nb);
}

// predicated global accesses for the stripe-boundary column (a branch around four loads / stores costs more issue slots in the
// step loop than executing the address arithmetic on every lane)
// the four values of one boundary-column row (F, E, M, B) sit side by side: one predicated 128-bit access per row
__device__ __forceinline__ void ldg4_if_s32(const int32_t* ptr, bool cond, int& a, int& b, int& c, int& d) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.v4.s32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "l"(ptr), "r"((uint32_t)cond));
}
__device__ __forceinline__ void stg4_if_s32(int32_t* ptr, bool cond, int a, int b, int c, int d) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.global.v4.s32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"((uint32_t)cond) : "memory");
}

template <int G, int C, bool TB, bool FIN, bool FAST, bool RB = false>
__global__ void __launch_bounds__(kThreads, (FAST && TB && C >= 32) ? 3 : 1) gotoh_kernel(const KParams p) {
    static_assert(C % 8 == 0, "C must be a multiple of 8 (4 direction bits per cell, whole words per lane)");
    static_assert(!(FAST && FIN), "the FAST variant needs uniform gap constants");
    static_assert(!RB || FAST, "rust-bio mode runs on the FAST kernels");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // FAST: [0,256) class LUT, [256,384) profile table, then the per-group reference rows
    uint8_t* smem = smem_raw + (FAST ? kLutBytes + kTabBytes : 0);
    if (FAST) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) smem_raw[i] = p.cls_lut[i];
        if (threadIdx.x < 32) ((uint32_t*)(smem_raw + kLutBytes))[threadIdx.x] = p.tab[threadIdx.x];
        __syncthreads();
    }
    const uint8_t* lut_sm = smem_raw;
    const uint8_t* tab_sm = smem_raw + kLutBytes;
    constexpr int GPW = 32 / G;
    constexpr int W = G * C;
    constexpr int WPL = C / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gw = lane / G;
    const int wpb = blockDim.x >> 5;
    const uint32_t ggid = (blockIdx.x * wpb + warp) * GPW + gw;
    uint8_t* ref_sm = smem + (size_t)(warp * GPW + gw) * p.ref_sm_stride;
    // TB: per-warp transposition buffer for the direction bits, after the reference rows
    uint32_t* tt_sm = reinterpret_cast<uint32_t*>(smem + (size_t)wpb * GPW * p.ref_sm_stride) + (size_t)warp * (WPL * 256);
    int32_t* col_g = p.col_scratch + (size_t)ggid * 4 * p.col_stride;
    const clq_affine_t sc = p.sc;
    int staged_ref = -1;

    const uint32_t n_tasks = p.n_tasks_dev ? *p.n_tasks_dev : p.n_tasks;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.task_counter, (unsigned)GPW);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n_tasks) break;
        uint32_t task = base + gw;
        bool valid = task < n_tasks;
        if (valid && p.retry_list) task = p.retry_list[task];  // retry pass: the task is a sub-batch position some earlier kernel gave up on
        uint32_t ridx = 0;
        int ref = -1;
        uint32_t* bits_g = (TB && valid) ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, task) : nullptr;
        if (valid) {
            if (p.all_pairs) {
                const uint32_t q = task / p.n_refs;
                ref = (int)(task - q * p.n_refs);
                ridx = p.order ? p.order[q] : q;
                if (p.cand_mask && !((p.cand_mask[(size_t)ridx * p.mask_words + (ref >> 5)] >> (ref & 31)) & 1u)) valid = false;
            } else {
                ridx = p.order ? p.order[p.task_base + task] : p.task_base + task;
                if (ridx == 0xffffffffu) {  // padding position of a reference group (host order / ref_scatter_kernel): nothing to align
                    if (TB && gl == 0) { TbRec rec; rec.ridx = 0; rec.L1 = -1; rec.L2 = 0; rec.zK = 0; p.tb_rec[task] = rec; }
                    valid = false;
                } else
                    ref = p.ref_of_read[ridx];
            }
        }
        int L1 = 0, L2 = 0;
        const uint8_t* refp = nullptr;
        const uint8_t* readp = nullptr;
        uint32_t status = CLQ_OK;
        if (valid) {
            const uint64_t r0 = p.read_off[ridx];
            L2 = (int)(p.read_off[ridx + 1] - r0);
            readp = p.read_bytes + r0;
            if ((uint32_t)L2 >= p.max_read_len) status = CLQ_READ_TOO_LONG;
            else if (ref < 0 || (uint32_t)ref >= p.n_refs) status = CLQ_NO_CANDIDATE;
            else {
                const uint64_t f0 = p.ref_off[ref];
                L1 = (int)(p.ref_off[ref + 1] - f0);
                refp = p.ref_bytes + f0;
            }
        }
        const bool ok = valid && status == CLQ_OK;
        const bool run = ok && L1 > 0 && L2 > 0;

        if (run && ref != staged_ref) {
            for (int i = gl; i < L1; i += G) ref_sm[i] = FAST ? ((lut_sm[refp[i]] >> 3) & 15) : refp[i];
            staged_ref = ref;
        }
        __syncwarp();

        const int K = run ? stale_rows(L1, L2, p.band_mode) : 0;
        const int NS = run ? (L2 + W - 1) / W : 0;
        const int NSmax = __reduce_max_sync(FULL, NS);
        const int T = run ? L1 + G - 1 : 0;
        const int Tmax = __reduce_max_sync(FULL, T);
        // the lane / register that own column L2 in the last stripe (narrow: CsL <= C columns per lane there)
        const int cL = run ? (L2 - 1) - (NS - 1) * W : 0;
        const bool narrow = FAST && G >= 16 && NS - 1 <= kMaxNarrowStripe;
        const int CsL = (run && narrow) ? narrow_cols<G>(cL + 1, C) : C;
        const int lL = cL / CsL, jL = cL - lL * CsL;
        int capM = 0, capE = 0, capF = 0;
        int bad = 0;  // RB: the read holds a byte the class table cannot score exactly

        for (int s = 0; s < NSmax; s++) {
            const bool act_s = run && s < NS;
            const bool own_last = act_s && s == NS - 1 && gl == lL;
            const int Cs = (G >= 16 && s == NS - 1) ? CsL : C;  // compile-time C for the short-read geometries
            const int nb = Cs >> 3;
            const int y0 = s * W + gl * Cs;  // columns y0+1 .. y0+Cs
            int E[C], B[C], bq[C];
            uint32_t w[WPL];
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int y = y0 + j + 1;
                int code = FAST ? 1 : 0x400;  // padding column: never equal, not special
                if (act_s && y <= L2 && j < Cs) {
                    const int c = readp[y - 1];
                    code = FAST ? (int)lut_sm[c] : (is_special(c) ? (c | 0x100) : c);
                    if (FAST) { bad |= code >> 7; code &= 7; }
                }
                bq[j] = FAST ? (code * 0x1111 | 0x8880) : code;  // FAST: PRMT selector (byte + sign replication)
                B[j] = sc.b0 + y * sc.b1;  // row 0: S[0,y] = (MAXNEG, g(y), g(y))
                E[j] = RB ? sc.max_neg : B[j] - (FAST ? sc.oe_in : 0);  // rust-bio: D[0][j] = MIN_SCORE
            }
            int prevBl = (y0 == 0) ? 0 : sc.b0 + y0 * sc.b1;  // B[0, y0]
            int oF = 0, oE = 0, oM = 0, oB = 0;
            // stripe boundary column (left edge of lane 0 when s > 0), prefetched one row ahead
            int nF = 0, nE = 0, nM = 0, nB = 0;
            if (s > 0 && gl == 0 && act_s) { const int4 v = *(const int4*)(col_g + 4); nF = v.x; nE = v.y; nM = v.z; nB = v.w; }  // row 1 of the boundary column
            int rnext = act_s ? ref_sm[0] : 0;  // every lane starts at row 1 (lane gl at step t = 1 + gl)
            // per-step conditions as single compares / loop-invariant flags (a compound condition is re-evaluated from its parts on
            // every step: predicates do not survive the row step)
            const uint32_t L1act = act_s ? (uint32_t)L1 : 0u;                      // act  <=>  (unsigned)(x - 1) < L1act
            const int xlast = FAST ? (own_last ? L1 : -1) : L1;                     // FAST: the capture variant only on the lane owning column L2
            const bool first_col = (gl == 0) && (s == 0);
            const bool ld_col = (gl == 0) && (s > 0), st_col = (gl == G - 1) && (s < NS - 1);
            const int k_own = own_last ? K : 0;

            for (int t = 1; t <= Tmax; t++) {
                const int x = t - gl;
                int Fl = __shfl_up_sync(FULL, oF, 1, G);
                int Bl = __shfl_up_sync(FULL, oB, 1, G);
                int El = 0, Ml = 0;
                if (TB && !RB) {
                    El = __shfl_up_sync(FULL, oE, 1, G);
                    Ml = __shfl_up_sync(FULL, oM, 1, G);
                }
                if ((uint32_t)(x - 1) < L1act) {
                    if (first_col) {  // S[x,0] = (MAXNEG, g(x), g(x)); selects
                        Bl = sc.b0 + x * sc.b1;
                        Fl = El = Bl - (FAST ? sc.oe_in : 0);
                        Ml = sc.max_neg;
                        if (RB) Fl = sc.max_neg;  // rust-bio: I[i][0] = MIN_SCORE on the empty-read boundary
                    }
                    if (ld_col) { Fl = nF; El = nE; Ml = nM; Bl = nB; }
                    {
                        const bool nx = ld_col && x < L1;  // the next row's boundary values, one step ahead
                        ldg4_if_s32(col_g + 4 * (x + 1), nx, nF, nE, nM, nB);
                    }
                    const int r = rnext;
                    rnext = ref_sm[x < L1 ? x : L1 - 1];
                    const int BlIn = Bl;
                    if (FAST) {
                        const uint2 tr = *(const uint2*)(tab_sm + r * 8);
                        if (x == xlast)
                            row_step_fast<C, TB, true, RB>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, sc.e_in, sc.oe_in, own_last, jL, capM, capE, capF, nb);
                        else
                            row_step_fast<C, TB, false, RB>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, sc.e_in, sc.oe_in, own_last, jL, capM, capE, capF, nb);
                    } else {
                        const bool rsp = is_special(r);
                        const int rcode = rsp ? 0x200 : r;
                        const int mt = rsp ? sc.special : sc.match;
                        const int mm = rsp ? sc.special : sc.mismatch;
                        if (p.band_mode == CLQ_BAND_K) {  // explicit bandwidth: per-row column window of this lane
                            int lo, hi;
                            band_of_row(x, L1, L2, (long long)p.band_k, lo, hi);
                            const int jlo = lo - (y0 + 1), jhi = hi - (y0 + 1);
                            if (x == L1)
                                row_step<C, TB, FIN, true, true>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, rcode, mt, mm, sc, own_last, jL, capM, capE, capF, jlo, jhi);
                            else
                                row_step<C, TB, FIN, false, true>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, rcode, mt, mm, sc, own_last, jL, capM, capE, capF, jlo, jhi);
                        } else if (x == L1)
                            row_step<C, TB, FIN, true>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, rcode, mt, mm, sc, own_last, jL, capM, capE, capF);
                        else
                            row_step<C, TB, FIN, false>(E, B, bq, w, Fl, El, Ml, Bl, prevBl, rcode, mt, mm, sc, own_last, jL, capM, capE, capF);
                    }
                    prevBl = BlIn;
                    oF = Fl; oE = El; oM = Ml; oB = Bl;
                    if (x <= k_own) {
                        // band-skipped cell (x, L2): fresh-matrix state (0,0,0) / Up(0)
                        const int e0 = FAST ? -sc.oe_in : 0;
#pragma unroll
                        for (int j = 0; j < C; j++)
                            if (j == jL) { E[j] = e0; B[j] = 0; }
                        if (jL == Cs - 1) { oF = e0; oE = e0; oM = 0; oB = 0; }
                        if (x == L1) { capM = 0; capE = 0; capF = 0; }
                    }
                    if (TB) bits_store<G, WPL>(tt_sm, bits_g, w, s, T, t, lane, gl, x == L1, nb);
                    stg4_if_s32(col_g + 4 * x, st_col, oF, oE, oM, oB);
                }
            }
            __syncwarp();
        }

        // ---- final cell: score + start layer = LAST maximum of (M, E, F)  (alignment/alignment_matrix.rs:963-972) ----
        const int src = gw * G + lL;
        capM = __shfl_sync(FULL, capM, src);
        capE = __shfl_sync(FULL, capE, src);
        capF = __shfl_sync(FULL, capF, src);
        int score = 0, z = 0;
        if (run) {
            score = capM; z = 0;
            if (RB) {  // rust-bio: S takes the match first, then the insertion, then the deletion, each only when strictly better
                if (capF > score) { score = capF; z = 2; }
                if (capE > score) { score = capE; z = 1; }
            } else {
                if (capE >= score) { score = capE; z = 1; }
                if (capF >= score) { score = capF; z = 2; }
            }
        } else if (ok) {
            const int n = L1 > L2 ? L1 : L2;
            if (n > 0) { score = sc.b0 + n * sc.b1; z = 2; }
        }
        if (RB) {
            const unsigned bm = __ballot_sync(FULL, bad != 0);
            const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (gw * G));
            if (ok && (bm & gm)) status = CLQ_SCORING_NOT_REPRESENTABLE;
        }

        if (!TB) {
            if (valid && gl == 0) {
                if (p.all_pairs) p.scores[(size_t)ridx * p.n_refs + ref] = ok ? score : INT32_MIN;
                else {
                    clq_result_t r;
                    r.score_scaled = score; r.ref_index = (status != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu; r.cigar_off = 0; r.cigar_len = 0; r.status = status; r.matches = 0; r.mismatches = 0;
                    p.results[ridx] = r;
                }
                if (run) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2);
            }
            continue;
        }

        // ---- leave the start layer for walk_kernel ----
        if (valid && gl == 0) {
            clq_result_t r;
            r.score_scaled = score; r.ref_index = (status != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu; r.cigar_off = 0; r.cigar_len = 0; r.status = status; r.matches = 0; r.mismatches = 0;
            p.results[ridx] = r;
            TbRec rec;
            rec.ridx = ridx; rec.L1 = (ok && status == CLQ_OK) ? L1 : -1; rec.L2 = L2;
            rec.zK = z | (K << 2) | ((CsL >> 3) << 20) | ((narrow && run ? NS - 1 : 0) << 24);
            p.tb_rec[task] = rec;
            if (run) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2);
        }
    }
}

#ifndef CLQ_WALK_PREFETCH
#define CLQ_WALK_PREFETCH 12
#endif
constexpr int kWalkPrefetch = CLQ_WALK_PREFETCH;  // steps ahead along the diagonal (0 = off)

// The walker's load of a direction-bit word.  CLQ_WALK_L2_HINT = 64 / 128 / 256 adds the L2 prefetch-size hint: a miss then
// brings the aligned 64 / 128 / 256 bytes around the sector into L2 in one DRAM access.  Within one block of 8 steps a diagonal
// path reads two neighbouring sectors (words k and k - 1 of its lane), so the second one is then an L2 hit.
#ifndef CLQ_WALK_L2_HINT
#define CLQ_WALK_L2_HINT 0
#endif
__device__ __forceinline__ uint32_t ld_bits(const uint32_t* p) {
#if CLQ_WALK_L2_HINT == 64
    uint32_t v;
    asm("ld.global.nc.L2::64B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#elif CLQ_WALK_L2_HINT == 128
    uint32_t v;
    asm("ld.global.nc.L2::128B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#elif CLQ_WALK_L2_HINT == 256
    uint32_t v;
    asm("ld.global.nc.L2::256B.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// perform_3d_global_traceback (alignment/alignment_matrix.rs:941-1086) + simplify_cigar_string (alignment_manager.rs:386-423)
// over the direction bits gotoh_kernel<.., TB=true, ..> stored.  One thread per pair: the walk is a chain of dependent
// loads (one 32-byte sector per step), so it is spread over as many threads as there are pairs in the sub-batch.
// WARP (the long-read geometries, G >= 16): one WARP per pair instead of one thread.  A sub-batch of long reads holds only a few
// thousand pairs with paths of ~10^4 steps, and with one thread per pair the walk is that many dependent DRAM round trips with
// the GPU nearly idle (C5: 14 % of the step).  All 32 lanes carry the same walk state; when the word of the current cell is not
// in the warp's window, lane i loads the word of the cell i steps further up the diagonal, and the following steps take
// their words from the lanes by shuffle as long as the path stays inside those words: one memory round trip per ~20-30
// steps instead of one per step.  Lane 0 alone writes.
template <int G, int C, bool WARP = false>
__global__ void __launch_bounds__(128) walk_kernel(const TbRec* recs, uint32_t n_tasks, const uint32_t* bits, uint64_t bits_stride,
                                                   const uint64_t* bits_off, uint32_t task_base,
                                                   uint32_t* cig_scratch, uint32_t cig_stride, uint32_t* cigar_pool, uint64_t cigar_cap,
                                                   unsigned long long* cigar_cursor, clq_result_t* results, const uint8_t* ref_bytes,
                                                   const uint64_t* ref_off, const uint8_t* read_bytes, const uint64_t* read_off,
                                                   const uint16_t* tag_slot, uint8_t* tags, uint32_t tag_stride, uint32_t rb,
                                                   uint32_t band_mode, uint32_t band_k) {
    constexpr int W = G * C;
    constexpr int WPL = C / 8;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = WARP ? (gtid >> 5) : gtid;
    const int wlane = threadIdx.x & 31;
    const bool writer = !WARP || wlane == 0;
    if (q >= n_tasks) return;  // warp-uniform with WARP
    const TbRec rec = recs[q];
    if (rec.L1 < 0) return;
    const int L1 = rec.L1, L2 = rec.L2, K = (rec.zK >> 2) & 0x3ffff;
    const int CsL = ((rec.zK >> 20) & 15) * 8, sL = (rec.zK >> 24) & 255;  // the task's narrow last stripe (G >= 16 only)
    int z = rec.zK & 3;
    const int T = L1 + G - 1;
    const uint32_t* bits_g = bits + bits_slot(bits_off, bits_stride, task_base, q);
    uint32_t* cig_g = cig_scratch + (size_t)q * cig_stride;
    uint32_t status = CLQ_OK;
    int x = L1, y = L2;
    // get_reference_alignment_rate (consensus/consensus_builders.rs:288-307) fused into the walk: only M columns can count
    const uint64_t ref0 = ref_off[results[rec.ridx].ref_index];
    const uint8_t* refp = ref_bytes + ref0;
    const uint8_t* readp = read_bytes + read_off[rec.ridx];
    uint32_t n_match = 0, n_mismatch = 0;
    // extract_tagged_sequences' digit tags (extractor.rs:271-332, the (false,_,false) arm) fused into the walk: the read byte
    // (or '-') aligned to every reference column that holds a tag symbol '0'..'9', at that column's rank
    const uint16_t* slotp = tag_slot ? tag_slot + ref0 : nullptr;
    uint8_t* tagp = tags ? tags + (size_t)rec.ridx * tag_stride : nullptr;
    auto tag = [&](int xx, uint8_t b) {
        if (slotp) {
            const uint16_t sl = __ldg(slotp + xx - 1);
            if (sl != 0xffffu && writer) tagp[sl] = b;
        }
    };
    auto count = [&](int xx, int yy) {
        const uint8_t rb = __ldg(refp + xx - 1), qb = __ldg(readp + yy - 1);
        if (rb > 64 && rb != 'N' && qb > 64) { if (rb == qb) n_match++; else n_mismatch++; }
        if (rb < 58) tag(xx, qb);
    };
    int cpos = (int)cig_stride;
    uint32_t cur_op = 3, cur_len = 0;
    auto emit = [&](uint32_t op, uint32_t n) {
        if (op == cur_op) cur_len += n;
        else {
            if (cur_len) { --cpos; if (writer) cig_g[cpos] = (cur_len << 4) | cur_op; }
            cur_op = op; cur_len = n;
        }
    };
    // word index of cell (xx, yy) in the task's slot and the shift of its nibble
    auto locate = [&](int xx, int yy, int& shift) -> size_t {
        int c = yy - 1;
        const int s = c / W;
        c -= s * W;
        const int cs = (s == sL && CsL > 0) ? CsL : C;  // the narrow last stripe of a multi-stripe task (fill kernels leave CsL = C, sL = 0 otherwise)
        const int ln = c / cs, j = c - ln * cs;
        shift = 28 - 4 * (j & 7);
        return bits_index<G, WPL>(s, T, xx + ln, ln, j >> 3);
    };
    int win_x = -1;                 // WARP: lane i holds the word of cell (win_x - i, win_y - i) and its index
    uint32_t win_word = 0, win_idx = 0xffffffffu;
    auto nibble = [&](int xx, int yy) -> uint32_t {
        int shift;
        const size_t idx = locate(xx, yy, shift);
        if (!WARP) return (ld_bits(bits_g + idx) >> shift) & 15u;
        const int d = win_x - xx;   // every lane walks the same path: all of this is warp-uniform
        uint32_t v = 0;
        bool hit = false;
        if (d >= 0 && d < 32) {
            v = __shfl_sync(FULL, win_word, d);
            hit = __shfl_sync(FULL, win_idx, d) == (uint32_t)idx;
        }
        if (!hit) {
            win_x = xx;
            const int px = xx - wlane, py = yy - wlane;
            win_idx = 0xffffffffu;
            if (px >= 1 && py >= 1) {
                int sh;
                const size_t pi = locate(px, py, sh);
                win_idx = (uint32_t)pi;
                win_word = ld_bits(bits_g + pi);
            }
            v = __shfl_sync(FULL, win_word, 0);
        }
        return (v >> shift) & 15u;
    };
    // The walk is a chain of dependent sector loads (ncu: long-scoreboard stall 15 per issue, 5.4 KB of DRAM reads per C2 read
    // = 168 sectors for a 515-step path).  Alignment paths run along diagonals, so every step also prefetches the word of the cell
    // kWalkPrefetch steps further up the diagonal into L1: the demand load then finds its sector on chip and the misses of
    // consecutive sectors overlap instead of queueing behind one another.  A wrong guess (indels move the path off the diagonal)
    // costs one extra sector, never a wrong result.
    auto prefetch = [&](int xx, int yy) {
        int c = yy - 1;
        const int s = c / W;
        c -= s * W;
        const int cs = (s == sL && CsL > 0) ? CsL : C;  // the narrow last stripe of a multi-stripe task (fill kernels leave CsL = C, sL = 0 otherwise)
        const int ln = c / cs, j = c - ln * cs;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(bits_g + bits_index<G, WPL>(s, T, xx + ln, ln, j >> 3)));
    };
    auto argmax = [](uint32_t nb) -> int { return (nb & 2u) ? 1 : ((nb & 1u) ? 2 : 0); };
    // cells the fill skipped keep the fresh-matrix state: (x <= K, y == L2) for the two implicit bands, everything outside the
    // row's window for an explicit bandwidth
    auto stale = [&](int xx, int yy) -> bool {
        if (band_mode == CLQ_BAND_K) {
            int lo, hi;
            band_of_row(xx, L1, L2, (long long)band_k, lo, hi);
            return yy < lo || yy > hi;
        }
        return yy == L2 && xx <= K;
    };
    uint32_t nib = (x > 0 && y > 0) ? nibble(x, y) : 0;
    bool cur_stale = (x > 0 && y > 0) ? stale(x, y) : false;
    while (x > 0 && y > 0) {
        if (cur_stale) { status = CLQ_TRACEBACK_DIVERGED; break; }  // stale Up(0) cell: the reference spins here
        const uint32_t old = nib;
        if (z == 0) { emit(CLQ_OP_M, 1); count(x, y); x--; y--; }
        else if (z == 1) { emit(CLQ_OP_D, 1); tag(x, '-'); x--; }
        else { emit(CLQ_OP_I, 1); y--; }
        if (x == 0 || y == 0) break;
        if (!WARP && G >= 16 && kWalkPrefetch > 0 && x > kWalkPrefetch && y > kWalkPrefetch) prefetch(x - kWalkPrefetch, y - kWalkPrefetch);
        nib = nibble(x, y);
        cur_stale = stale(x, y);
        const int a = cur_stale ? 0 : argmax(nib);  // a stale source cell holds (0,0,0): Diag
        if (z == 0) z = a;
        else if (rb) z = (z == 1) ? ((old & 8u) ? 1 : a) : ((old & 4u) ? 2 : a);  // rust-bio: an opened gap resumes in S = argmax
        else if (z == 1) z = (old & 8u) ? 1 : (a == 2 ? 2 : 0);
        else z = (old & 4u) ? 2 : (a == 1 ? 1 : 0);
    }
    if (status == CLQ_OK) {
        if (x > 0) {
            emit(CLQ_OP_D, (uint32_t)x);
            if (slotp) for (int xx = x; xx >= 1; xx--) tag(xx, '-');
        }
        if (y > 0) emit(CLQ_OP_I, (uint32_t)y);
    }
    if (cur_len) { --cpos; if (writer) cig_g[cpos] = (cur_len << 4) | cur_op; }
    int nops = (int)cig_stride - cpos;
    unsigned long long off = 0;
    if (status != CLQ_OK) nops = 0;
    if (nops > 0) {
        if (writer) off = atomicAdd(cigar_cursor, (unsigned long long)nops);
        if (WARP) {
            off = __shfl_sync(FULL, off, 0);
            __syncwarp();  // lane 0's scratch writes, read by every lane below
        }
        if (off + (unsigned long long)nops > cigar_cap) { status = CLQ_CIGAR_POOL_FULL; nops = 0; }
    }
    for (int i = WARP ? wlane : 0; i < nops; i += WARP ? 32 : 1) cigar_pool[off + i] = cig_g[cpos + i];
    if (writer) {
        clq_result_t* r = results + rec.ridx;
        r->cigar_off = (uint32_t)off;
        r->cigar_len = (uint32_t)nops;
        r->status = status;
        r->matches = n_match;
        r->mismatches = n_mismatch;
    }
}

// Grouping of a multi-reference batch for the PACK traceback stage: the s16x2 kernels align two reads against ONE reference,
// so the reads are bucketed by their (selected or fixed) reference; every bucket is padded to an even number of processing
// positions (padding = 0xffffffff, skipped by the kernel).  Bucket n_refs collects reads without a usable reference.
__global__ void ref_hist_kernel(const int32_t* ref_of_read, uint32_t n_reads, uint32_t n_refs, uint32_t* hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const int32_t r = ref_of_read[i];
    atomicAdd(hist + ((r >= 0 && (uint32_t)r < n_refs) ? (uint32_t)r : n_refs), 1u);
}

__global__ void ref_scan_kernel(const uint32_t* hist, uint32_t n_buckets, uint32_t* cursor) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t at = 0;
    for (uint32_t b = 0; b < n_buckets; b++) {
        cursor[b] = at;
        at += (hist[b] + 1u) & ~1u;
    }
}

__global__ void ref_scatter_kernel(const int32_t* ref_of_read, uint32_t n_reads, uint32_t n_refs, uint32_t* cursor, uint32_t* order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const int32_t r = ref_of_read[i];
    order[atomicAdd(cursor + ((r >= 0 && (uint32_t)r < n_refs) ? (uint32_t)r : n_refs), 1u)] = i;
}

// exhaustive_alignment_search's arg-max: ascending reference index, LAST maximum wins
// (max_by(partial_cmp), alignment_functions.rs:809-813).  One thread per read.
__global__ void select_best_kernel(const int32_t* scores, uint32_t n_reads, uint32_t n_refs, const uint32_t* cand_mask,
                                   uint32_t mask_words, int32_t* ref_of_read, const int32_t* single_ref) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    if (single_ref && single_ref[i] >= 0) { ref_of_read[i] = single_ref[i]; return; }
    int best = -1;
    int32_t bs = INT32_MIN;
    for (uint32_t r = 0; r < n_refs; r++) {
        if (cand_mask && !((cand_mask[(size_t)i * mask_words + (r >> 5)] >> (r & 31)) & 1u)) continue;
        const int32_t v = scores[(size_t)i * n_refs + r];
        if (v == INT32_MIN) continue;
        if (best < 0 || v >= bs) { bs = v; best = (int)r; }
    }
    ref_of_read[i] = best;
}

// quick_alignment_search's k-mer vote (alignment_functions.rs:693-767) against the unique-k-mer table built by
// clq_kmer_index_set (reference/fasta_reference.rs:159-202).  One thread per read:
//   votes per reference over the read's sampled k-mer runs; best share > threshold -> single_ref[i] = that reference;
//   otherwise the candidate mask = references with >= 1 vote, or every reference when nothing voted.
__global__ void kmer_vote_kernel(const uint8_t* read_bytes, const uint64_t* read_off, uint32_t n_reads, const uint8_t* keys,
                                 const uint32_t* owner, uint32_t n_keys, uint32_t k, uint32_t skip, uint32_t n_refs,
                                 double threshold, uint32_t* votes_scratch, uint32_t* cand_mask, uint32_t mask_words,
                                 int32_t* single_ref) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const uint8_t* rd = read_bytes + read_off[i];
    const uint32_t L2 = (uint32_t)(read_off[i + 1] - read_off[i]);
    uint32_t* votes = votes_scratch + (size_t)i * n_refs;
    for (uint32_t r = 0; r < n_refs; r++) votes[r] = 0;
    uint32_t total = 0;
    long long prev = -1;  // start of the previous window (consecutive dedup_with_count)
    for (uint32_t pos = 0; k > 0 && pos + k <= L2; pos += skip) {
        bool same = prev >= 0;
        if (same) {
            for (uint32_t c = 0; c < k; c++) {
                uint8_t a = rd[prev + c], b = rd[pos + c];
                a = (a >= 'a' && a <= 'z') ? a - 32 : a;
                b = (b >= 'a' && b <= 'z') ? b - 32 : b;
                if (a != b) { same = false; break; }
            }
        }
        prev = pos;
        if (same) continue;  // same run: one vote per run
        uint32_t lo = 0, hi = n_keys;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            int cmp = 0;
            for (uint32_t c = 0; c < k; c++) {
                uint8_t b = rd[pos + c];
                b = (b >= 'a' && b <= 'z') ? b - 32 : b;
                const uint8_t a = keys[(size_t)mid * k + c];
                if (a != b) { cmp = a < b ? -1 : 1; break; }
            }
            if (cmp == 0) { votes[owner[mid]]++; total++; break; }
            if (cmp < 0) lo = mid + 1; else hi = mid;
        }
    }
    uint32_t* mask = cand_mask + (size_t)i * mask_words;
    for (uint32_t wd = 0; wd < mask_words; wd++) mask[wd] = 0;
    int single = -1;
    if (total == 0) {
        for (uint32_t r = 0; r < n_refs; r++) mask[r >> 5] |= 1u << (r & 31);
    } else {
        const double count = (double)total;
        double bestp = -1.0;
        int bi = -1;
        for (uint32_t r = 0; r < n_refs; r++) {
            if (!votes[r]) continue;
            mask[r >> 5] |= 1u << (r & 31);
            const double pr = (double)votes[r] / count;
            if (pr >= bestp) { bestp = pr; bi = (int)r; }
        }
        if (bestp > threshold) single = bi;
    }
    single_ref[i] = single;
    if (single >= 0) {
        for (uint32_t wd = 0; wd < mask_words; wd++) mask[wd] = 0;  // no exhaustive fill for this read
    }
}

// Fast form of kmer_vote_kernel for k <= 8 (clique's CLI uses k = 8, skip = 4, main.rs:271): the unique k-mers are packed
// big-endian into uint64 (numeric order == the byte-wise order of the host's sort), the table and the per-thread vote counters
// live in shared memory, the read's window is a rolling 64-bit register and the run-length de-duplication one compare.  Same
// votes, same candidate masks, same single_ref as the generic kernel (the host falls back to that one when the table or the
// counters do not fit).  Dynamic shared memory: n_keys * 12 bytes (keys, owners) + n_refs * blockDim.x * 2 (u16 counters).
__global__ void __launch_bounds__(128) kmer_vote_packed_kernel(const uint8_t* read_bytes, const uint64_t* read_off, uint32_t n_reads,
                                                                const uint64_t* keys64, const uint32_t* owner, uint32_t n_keys, uint32_t k,
                                                                uint32_t skip, uint32_t n_refs, double threshold, uint32_t* cand_mask,
                                                                uint32_t mask_words, int32_t* single_ref) {
    extern __shared__ __align__(16) uint8_t vote_sm[];
    uint64_t* keys_sm = reinterpret_cast<uint64_t*>(vote_sm);
    uint32_t* own_sm = reinterpret_cast<uint32_t*>(vote_sm + (size_t)n_keys * 8);
    uint16_t* cnt_sm = reinterpret_cast<uint16_t*>(vote_sm + (size_t)n_keys * 12);
    for (uint32_t i = threadIdx.x; i < n_keys; i += blockDim.x) { keys_sm[i] = keys64[i]; own_sm[i] = owner[i]; }
    for (uint32_t r = 0; r < n_refs; r++) cnt_sm[r * blockDim.x + threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_reads) return;
    const uint8_t* rd = read_bytes + read_off[i];
    const uint32_t L2 = (uint32_t)(read_off[i + 1] - read_off[i]);
    uint16_t* votes = cnt_sm + threadIdx.x;  // votes[r * blockDim.x]
    const uint64_t kmask = k >= 8 ? ~0ull : ((1ull << (8 * k)) - 1ull);
    uint32_t total = 0;
    uint64_t win = 0, prev = 0;
    bool have_prev = false;
    uint32_t filled = 0;  // bytes of the read already shifted into the window
    for (uint32_t pos = 0; pos + k <= L2; pos += skip) {
        // advance the window to cover read[pos .. pos + k)
        for (; filled < pos + k; filled++) {
            uint8_t b = __ldg(rd + filled);
            b = (b >= 'a' && b <= 'z') ? b - 32 : b;
            win = (win << 8) | b;
        }
        const uint64_t key = win & kmask;
        const bool same = have_prev && key == prev;  // consecutive dedup_with_count: one vote per run
        prev = key; have_prev = true;
        if (same) continue;
        uint32_t lo = 0, hi = n_keys;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint64_t a = keys_sm[mid];
            if (a == key) { votes[own_sm[mid] * blockDim.x]++; total++; break; }
            if (a < key) lo = mid + 1; else hi = mid;
        }
    }
    uint32_t* mask = cand_mask + (size_t)i * mask_words;
    for (uint32_t wd = 0; wd < mask_words; wd++) mask[wd] = 0;
    int single = -1;
    if (total == 0) {
        for (uint32_t r = 0; r < n_refs; r++) mask[r >> 5] |= 1u << (r & 31);
    } else {
        const double count = (double)total;
        double bestp = -1.0;
        int bi = -1;
        for (uint32_t r = 0; r < n_refs; r++) {
            const uint32_t v = votes[r * blockDim.x];
            if (!v) continue;
            mask[r >> 5] |= 1u << (r & 31);
            const double pr = (double)v / count;
            if (pr >= bestp) { bestp = pr; bi = (int)r; }
        }
        if (bestp > threshold) single = bi;
    }
    single_ref[i] = single;
    if (single >= 0) {
        for (uint32_t wd = 0; wd < mask_words; wd++) mask[wd] = 0;  // no exhaustive fill for this read
    }
}

}  // namespace clq

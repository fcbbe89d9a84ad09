// clq_pack.cuh -- s16x2 ("PACK") variant of the FAST affine kernel: one lane group aligns TWO reads against the same
// reference at once, read A in the low and read B in the high 16-bit half of every register, so each DPX instruction
// (VIADDMNMX.S16x2) advances two cells.  Same recurrence, same direction bits, same bit layout and the same walker as
// gotoh_kernel<.., FAST> (clq_kernels.cuh); only the arithmetic width differs.
//
// Exactness: the host proves per batch that every value a cell can take lies in a window narrower than 2^15
// (clq_api.cu::pack_bias), values are stored with a common positive bias so that both halves are always in [64, 32767]:
// signed and unsigned order coincide, 32-bit subtractions of a >= b never borrow across halves, and the boundary
// sentinel MAX_NEG (which only ever loses a max, alignment/alignment_matrix.rs:385-406) is represented by 0.
// Direction bits: with x >= y per half, "x > y" is min_u16x2(x - y, 1); four such bit pairs per cell pair are shifted
// into packed accumulators (low half = read A, high half = read B) and unzipped with PRMT into each read's row word.
#pragma once

#include <cstdio>
#include <type_traits>

#include "clq_kernels.cuh"

namespace clq {

__device__ __forceinline__ uint32_t dup16(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
__device__ __forceinline__ uint32_t set_lo(uint32_t w, int v) { return (w & 0xffff0000u) | ((uint32_t)v & 0xffffu); }
__device__ __forceinline__ uint32_t set_hi(uint32_t w, int v) { return (w & 0x0000ffffu) | ((uint32_t)v << 16); }
__device__ __forceinline__ int get_lo(uint32_t w) { return (int)(w & 0xffffu); }
__device__ __forceinline__ int get_hi(uint32_t w) { return (int)(w >> 16); }

// shared-memory accessors on 32-bit shared addresses (the generic-pointer form re-derives the CTA's shared window -- S2UR
// SR_CgaCtaId + ULEA -- at every use inside the step loop)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// Direction-bit store of both reads of a PACK task: one flush branch per step for the pair (bits_store per read costs two).
// The words are parked unconditionally (the buffer is private to the lane); only the global stores depend on the read.
template <int G, int WPL>
__device__ __forceinline__ void bits_store2(uint32_t* tt, uint32_t* bitsA, uint32_t* bitsB, const uint32_t (&wA)[WPL], const uint32_t (&wB)[WPL],
                                            bool stA, bool stB, int s, int T, int t, int lane, int gl, bool last_row, int nb) {
    if constexpr (BitsLayout<G>::transposed) {
        const int Tb = (T + 7) >> 3;
        const int ph = (t - 1) & 7;
        uint32_t* ttB = tt + WPL * 256;
#pragma unroll
        for (int k = 0; k < WPL; k++) { tt[(k * 8 + ph) * 32 + lane] = wA[k]; ttB[(k * 8 + ph) * 32 + lane] = wB[k]; }
        if (ph == 7 || last_row) {
            const size_t at = ((size_t)(s * Tb + ((t - 1) >> 3)) * (G * WPL) + gl * WPL) * 8;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (h ? stB : stA) {
                    uint32_t* dst = (h ? bitsB : bitsA) + at;
                    const uint32_t* buf = h ? ttB : tt;
#pragma unroll
                    for (int k = 0; k < WPL; k++) {
                        const uint32_t* src = buf + (k * 8) * 32 + lane;
                        *reinterpret_cast<uint4*>(dst + k * 8) = make_uint4(src[0], src[32], src[64], src[96]);
                        *reinterpret_cast<uint4*>(dst + k * 8 + 4) = make_uint4(src[128], src[160], src[192], src[224]);
                    }
                }
            }
        }
    } else {
        if (stA) row_store<G, WPL>(bitsA, wA, s, T, t, gl, nb);
        if (stB) row_store<G, WPL>(bitsB, wB, s, T, t, gl, nb);
    }
}

// predicated global accesses for the stripe-boundary column: a branch around four loads / stores costs more issue slots on this
// kernel than executing the address arithmetic on every lane
__device__ __forceinline__ uint32_t ldg_if(const uint32_t* ptr, bool cond, uint32_t keep) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.global.u32 %0, [%1];\n\t}" : "+r"(keep) : "l"(ptr), "r"((uint32_t)cond));
    return keep;
}
__device__ __forceinline__ void stg_if(uint32_t* ptr, bool cond, uint32_t v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u32 [%0], %1;\n\t}" ::"l"(ptr), "r"(v), "r"((uint32_t)cond) : "memory");
}
// the four values of one boundary-column row (F, E, M, B) sit side by side: one 128-bit access per row
__device__ __forceinline__ void ldg4_if(const uint32_t* ptr, bool cond, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "l"(ptr), "r"((uint32_t)cond));
}
__device__ __forceinline__ void stg4_if(uint32_t* ptr, bool cond, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.global.v4.u32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"((uint32_t)cond) : "memory");
}

// -DCLQ_PACK_CANARY=1 (debug builds, tools/canary_gpu.sh): every value a PACK kernel stores (M, Eh, Fh, B of every cell) is
// tracked per task; a task whose values left [64, 32767] WITHOUT being handed to the retry pass is counted here and reported by
// clq_debug_canary().  Must stay 0: it is the run-time check of the window proofs (static, sloped, adaptive guard band).
#ifndef CLQ_PACK_CANARY
#define CLQ_PACK_CANARY 0
#endif
__device__ unsigned long long g_pack_canary[2];  // [0] tasks checked, [1] violations
struct Canary {
    uint32_t mn = 0xffffffffu, mx = 0u;
    int nA = 1 << 30, nB = 1 << 30;  // columns j < nA / nB of this lane hold real cells of read A / B (static PACK: the window proof
                                     // covers padding columns too; the adaptive kernel only guards real cells, padding may wrap)
    __device__ __forceinline__ void see(uint32_t v, int j) {
#if CLQ_PACK_CANARY
        const uint32_t w = (j < nA ? (v & 0xffffu) : 0x4000u) | (j < nB ? (v & 0xffff0000u) : 0x40000000u);
        mn = __vminu2(mn, w); mx = __vmaxu2(mx, w);
#endif
    }
    // call with full-warp convergence; `skip`: the task is redone elsewhere (its values do not matter)
    __device__ __forceinline__ void report(bool active, bool skip, int a = 0, int b = 0, int c = 0, int d = 0) {
#if CLQ_PACK_CANARY
        const bool bad = active && !skip && ((mn & 0xffffu) < 64u || (mn >> 16) < 64u || (mx & 0xffffu) > 32767u || (mx >> 16) > 32767u);
        if (bad) printf("canary: lane %d block %d min %u/%u max %u/%u ctx %d %d %d %d\n", (int)(threadIdx.x & 31), (int)blockIdx.x, mn & 0xffffu, mn >> 16, mx & 0xffffu, mx >> 16, a, b, c, d);
        const unsigned nbad = __popc(__ballot_sync(0xffffffffu, bad)), nact = __popc(__ballot_sync(0xffffffffu, active && !skip));
        if ((threadIdx.x & 31) == 0) { atomicAdd(&g_pack_canary[0], (unsigned long long)nact); if (nbad) atomicAdd(&g_pack_canary[1], (unsigned long long)nbad); }
#endif
    }
};

#ifndef CLQ_PACK_PIN_NIBBLE
#define CLQ_PACK_PIN_NIBBLE 1
#endif
__device__ __forceinline__ uint32_t mad2(uint32_t a, uint32_t b) {
#if CLQ_PACK_PIN_NIBBLE
    uint32_t d;
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return a * 2u + b;
#endif
}

struct PackParams {
    int32_t bias;   // added to every stored score
    int32_t slope;  // MADD kernels: row x is stored with x * slope more (slope = -min(substitution score, 0) >= 0)
};

// MADD (static row slope, see pack_kernel): the profile bytes are >= 0, so M = diag + m is a plain 32-bit add of packed halves
// (no carry can cross them) and leaves the ALU pipe; LEe / X1b are the E-step extend and the B-step open constants, which then
// differ from the F-step / P-step ones (LE, X1) by the slope.
template <int C, bool TB, bool LAST, bool RB, bool MADD, int JB>
__device__ __forceinline__ void pack_row_blocks(uint32_t (&Eh)[C], uint32_t (&B)[C], const uint32_t (&sel)[C], uint32_t (&wA)[C / 8],
                                                uint32_t (&wB)[C / 8], uint32_t& Fh, uint32_t& Ehl, uint32_t& Ml, uint32_t& Bl, uint32_t diag,
                                                uint32_t tlo, uint32_t thi, uint32_t LE, uint32_t X1, uint32_t X1M1, uint32_t LEe, uint32_t X1b,
                                                bool ownA, int jA, bool ownB, int jB, uint32_t (&cap)[3], int nb, Canary& cy) {
    // Blocks of 8 columns, nested: block JB + 1 only runs inside block JB's `if`, so a narrow last stripe (only nb blocks per lane
    // are real, G >= 16 geometries) leaves through ONE forward branch instead of a reconvergence region per block.
    if constexpr (JB < C / 8) {
        if (JB < nb) {
            constexpr int jb = JB;
            const uint32_t ONE = 0x00010001u;
            uint32_t acc0 = 0, acc1 = 0;
            // M of the next cell is issued inside the current one (it needs B[j] of the previous row, which the current cell
            // overwrites): B[j]'s old value then dies within the cell and no register copy is needed to carry it as `diag`
            uint32_t Mnext = MADD ? diag + (uint32_t)prmt_s8(tlo, thi, sel[jb * 8])
                                  : __viaddmax_s16x2(diag, (uint32_t)prmt_s8(tlo, thi, sel[jb * 8]), X1);  // per-half add: biased values are > 0 > x1
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int j = jb * 8 + jj;
                const uint32_t Mv = Mnext;                               // [mA, mB] profile bytes added as s16x2 (X1: a live register; a literal 0 costs a PRMT)
                const uint32_t EhU = Eh[j], BU = B[j];
                if (jj < 7) Mnext = MADD ? BU + (uint32_t)prmt_s8(tlo, thi, sel[j + 1]) : __viaddmax_s16x2(BU, (uint32_t)prmt_s8(tlo, thi, sel[j + 1]), X1);
                const uint32_t Ehn = __viaddmax_s16x2(EhU, LEe, BU);
                uint32_t t2 = 0, u2 = 0, f2 = 0;
                if (TB && !RB) {
                    // F extends  <=>  F_left + le > max(E_left + x1 - 1, M_left + x1)  (>= the E-open, > the M-open).  LE - t2 is the
                    // per-half value le - t2 (no borrow crosses the halves: t2 - le is in (0, 65536) per half), so one DPX with
                    // relu clamps Fh + le - t2 to {0, 1}: the bit itself, without a second max and a subtract + min
                    t2 = __viaddmax_s16x2(Ehl, X1M1, Ml);
                    f2 = __viaddmin_s16x2_relu(Fh, LE - t2, ONE);
                }
                const uint32_t Fhn = __viaddmax_s16x2(Fh, LE, Bl);
                if (TB && RB) { u2 = Fhn; t2 = Bl; }  // rust-bio: F extends <=> F_left + e > B_left + o + e
                const uint32_t Pv = __viaddmax_s16x2(Fhn, X1, Mv);
                const uint32_t Bn = __viaddmax_s16x2(Ehn, X1b, Pv);
                if (TB) {
                    // nibble pair [ext1 ext2 eP fM] of this cell pair, then 4 cells per 16-bit half
                    // every flag is shifted into the accumulator by one multiply-add (mad2: an opaque a * 2 + b, so that the compiler
                    // cannot re-associate the chain into a tree of ~4.75 adds and multiply-adds per cell for the sake of its depth)
                    uint32_t a = ((jj & 4) ? acc1 : acc0);
                    const uint32_t e1 = __vminu2(Ehn - BU, ONE);           // ext1
                    a = ((jj & 3) == 0) ? e1 : mad2(a, e1);
                    a = mad2(a, (TB && !RB) ? f2 : __vminu2(u2 - t2, ONE));  // ext2
                    a = mad2(a, __vminu2(Bn - Pv, ONE));                     // eP: E > max(M,F)
                    a = mad2(a, __vminu2(Pv - Mv, ONE));                     // fM: F > M
                    if (jj & 4) acc1 = a; else acc0 = a;
                    if (jj == 7) {
                        wA[jb] = __byte_perm(acc1, acc0, 0x5410);  // low halves: read A's 8 nibbles
                        wB[jb] = __byte_perm(acc1, acc0, 0x7632);  // high halves: read B's
                    }
                }
                if (jj == 7) diag = BU;  // carried to the next block of 8 columns
                Eh[j] = Ehn;
                B[j] = Bn;
                Fh = Fhn; Ehl = Ehn; Ml = Mv; Bl = Bn;
                if (CLQ_PACK_CANARY) { cy.see(Mv, j); cy.see(Ehn, j); cy.see(Fhn, j); cy.see(Bn, j); }
                if (LAST) {
                    if (ownA && j == jA) { cap[0] = set_lo(cap[0], get_lo(Mv)); cap[1] = set_lo(cap[1], get_lo(Ehn)); cap[2] = set_lo(cap[2], get_lo(Fhn)); }
                    if (ownB && j == jB) { cap[0] = set_hi(cap[0], get_hi(Mv)); cap[1] = set_hi(cap[1], get_hi(Ehn)); cap[2] = set_hi(cap[2], get_hi(Fhn)); }
                }
            }
            pack_row_blocks<C, TB, LAST, RB, MADD, JB + 1>(Eh, B, sel, wA, wB, Fh, Ehl, Ml, Bl, diag, tlo, thi, LE, X1, X1M1, LEe, X1b, ownA, jA, ownB, jB, cap, nb, cy);
        }
    }
}

template <int C, bool TB, bool LAST, bool RB = false, bool MADD = false>
__device__ __forceinline__ void pack_row_step(uint32_t (&Eh)[C], uint32_t (&B)[C], const uint32_t (&sel)[C], uint32_t (&wA)[C / 8],
                                              uint32_t (&wB)[C / 8], uint32_t& Fh, uint32_t& Ehl, uint32_t& Ml, uint32_t& Bl, uint32_t diag,
                                              uint32_t tlo, uint32_t thi, uint32_t LE, uint32_t X1, uint32_t X1M1, uint32_t LEe, uint32_t X1b,
                                              bool ownA, int jA, bool ownB, int jB, uint32_t (&cap)[3], int nb, Canary& cy) {
    pack_row_blocks<C, TB, LAST, RB, MADD, 0>(Eh, B, sel, wA, wB, Fh, Ehl, Ml, Bl, diag, tlo, thi, LE, X1, X1M1, LEe, X1b, ownA, jA, ownB, jB, cap, nb, cy);
}

// Tasks are read PAIRS.  Pair mode (all_pairs == 0): pair t = reads at processing positions task_base + 2t, +1, all against
// reference ref_of_read[.] (the host only takes this kernel when the batch has a single reference).  All-pairs mode: pair
// t = (read pair t / n_refs, reference t % n_refs).
// occupancy: the (8,40) short-read geometry runs best with all 255 registers and no spills (2 CTAs/SM: C2 24.1 M reads/s
// against 23.2 M at 3 CTAs/SM with 168 registers and spills, 23.0 M at 200 registers without spills whatever the CTA size), the
// long-read (32,32) geometry at 3 CTAs/SM (C3 1447 vs 1341 GCUPS)
template <int G, int C, bool TB>
constexpr int pack_min_blocks() { return (TB && C >= 32) ? (G <= 8 ? 2 : 3) : 1; }

//
// MADD ("static row slope"): row x is stored as v + bias + x * s with s = -min(substitution score) >= 0.  Only constants of
// the recurrence change (derivation with B' = B + s x, M' = M + s x, Fh' = Fh + s x, Eh'(x,.) = Eh(x,.) + s (x - 1)):
//   M'  = B'[x-1,y-1] + (m + s)              profile bytes m + s >= 0: a plain packed add, off the ALU pipe
//   Eh' = max(Eh'_up + (le + s), B'_up)      Fh' = max(Fh'_left + le, B'_left)
//   P'  = max(Fh' + x1, M')                  B'  = max(Eh' + (x1 + s), P')
//   ext2: t2' = max(Eh'_left + (x1 - 1 + s), M'_left); every direction bit compares two values of the same row: unchanged.
// The host takes these kernels when the 15-bit window still holds with L1 * s more head-room (clq_api.cu) and hands over the
// profile table with s already added.
template <int G, int C, bool TB, bool RB = false, bool MADD = false>
__global__ void __launch_bounds__(kThreads, pack_min_blocks<G, C, TB>()) pack_kernel(const KParams p, const PackParams pp) {
    static_assert(!(RB && MADD), "the rust-bio variant keeps the unsloped form");
    static_assert(C % 8 == 0, "C must be a multiple of 8");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + kLutBytes + kTabBytes;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) smem_raw[i] = p.cls_lut[i];
    if (threadIdx.x < 32) ((uint32_t*)(smem_raw + kLutBytes))[threadIdx.x] = p.tab[threadIdx.x];
    __syncthreads();
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
    // TB: per-warp transposition buffers (read A, read B) for the direction bits, after the reference rows
    uint32_t* tt_sm = reinterpret_cast<uint32_t*>(smem + (size_t)wpb * GPW * p.ref_sm_stride) + (size_t)warp * (2 * WPL * 256);
    uint32_t* col_g = (uint32_t*)p.col_scratch + (size_t)ggid * 4 * p.col_stride;
    const clq_affine_t sc = p.sc;
    const int bias = pp.bias;
    const int x1 = sc.oe_in, le = sc.e_in;
    const int slope = MADD ? pp.slope : 0;
    const uint32_t LE = dup16(le), X1 = dup16(x1), X1M1 = dup16(x1 - 1 + slope);
    const uint32_t LEe = dup16(le + slope), X1b = dup16(x1 + slope);  // E-step extend / B-step open of a sloped row
    int staged_ref = -1;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.task_counter, (unsigned)GPW);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.n_tasks) break;
        const uint32_t task = base + gw;
        const bool tvalid = task < p.n_tasks;
        // ---- decode the two reads of this task ----
        uint32_t ridx[2] = {0, 0};
        bool valid[2] = {false, false};
        int ref = -1;
        if (tvalid) {
            if (p.all_pairs) {
                const uint32_t q = task / p.n_refs;
                ref = (int)(task - q * p.n_refs);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t pos = 2 * q + h;
                    if (pos < p.n_reads) {
                        ridx[h] = p.order ? p.order[pos] : pos;
                        valid[h] = !(p.cand_mask && !((p.cand_mask[(size_t)ridx[h] * p.mask_words + (ref >> 5)] >> (ref & 31)) & 1u));
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t pos = p.task_base + 2 * task + h;
                    if (pos < p.task_end) {
                        ridx[h] = p.order ? p.order[pos] : pos;
                        valid[h] = ridx[h] != 0xffffffffu;  // padding position of a reference bucket (ref_scatter_kernel)
                    }
                }
                // the reference of the pair = the first usable per-read reference (the host takes this kernel only for
                // single-reference batches, so two usable references are always equal)
                for (int h = 1; h >= 0; h--)
                    if (valid[h]) {
                        const int rh = p.ref_of_read[ridx[h]];
                        if (rh >= 0 && (uint32_t)rh < p.n_refs) ref = rh;
                    }
            }
        }
        int L1 = 0, L2[2] = {0, 0};
        const uint8_t* refp = nullptr;
        const uint8_t* readp[2] = {nullptr, nullptr};
        uint32_t status[2] = {CLQ_OK, CLQ_OK};
        bool ok[2], run[2];
        const bool ref_ok = ref >= 0 && (uint32_t)ref < p.n_refs;
        if (ref_ok) {
            const uint64_t f0 = p.ref_off[ref];
            L1 = (int)(p.ref_off[ref + 1] - f0);
            refp = p.ref_bytes + f0;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (valid[h]) {
                const uint64_t r0 = p.read_off[ridx[h]];
                L2[h] = (int)(p.read_off[ridx[h] + 1] - r0);
                readp[h] = p.read_bytes + r0;
                int rh = ref;
                if (!p.all_pairs) rh = p.ref_of_read[ridx[h]];
                if ((uint32_t)L2[h] >= p.max_read_len) status[h] = CLQ_READ_TOO_LONG;
                else if (rh < 0 || (uint32_t)rh >= p.n_refs || rh != ref) status[h] = CLQ_NO_CANDIDATE;
            }
            ok[h] = valid[h] && status[h] == CLQ_OK;
            run[h] = ok[h] && L1 > 0 && L2[h] > 0;
        }
        const bool anyrun = run[0] || run[1];
        if (anyrun && ref != staged_ref) {
            for (int i = gl; i < L1; i += G) ref_sm[i] = (lut_sm[refp[i]] >> 3) & 15;
            staged_ref = ref;
        }
        __syncwarp();

        const int L2m = max(run[0] ? L2[0] : 0, run[1] ? L2[1] : 0);
        const int NS = anyrun ? (L2m + W - 1) / W : 0;
        // narrow last stripe of the PAIR: both reads share the lane mapping, so the longer one sets the width
        // (short-read geometries: only when the task spans several stripes -- long reads on (8,40) -- so that the single-stripe fast
        // path keeps its compile-time column count)
        const bool narrow = (G >= 16 || NS > 1) && NS - 1 <= kMaxNarrowStripe;
        const int CsL = (anyrun && narrow) ? narrow_cols<G>(L2m - (NS - 1) * W, C) : C;
        int K[2], NSh[2], lLh[2], jLh[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            K[h] = run[h] ? stale_rows(L1, L2[h], p.band_mode) : 0;
            NSh[h] = run[h] ? (L2[h] + W - 1) / W : 0;
            const int cL = run[h] ? (L2[h] - 1) - (NSh[h] - 1) * W : 0;
            const int csh = (NSh[h] == NS) ? CsL : C;  // its last column lies in the pair's last stripe?
            lLh[h] = cL / csh;
            jLh[h] = cL - lLh[h] * csh;
        }
        const int NSmax = __reduce_max_sync(FULL, NS);
        const int T = anyrun ? L1 + G - 1 : 0;
        const int Tmax = __reduce_max_sync(FULL, T);
        Canary cy;
        uint32_t cap[3] = {0, 0, 0};
        uint32_t bad = 0;  // RB: bit h = read h holds a byte the class table cannot score exactly
        uint32_t* bitsA = TB ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task) : nullptr;
        uint32_t* bitsB = (TB && valid[1]) ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task + 1) : nullptr;

        // Warp-uniform fast path: one column stripe and no band-skipped cells anywhere in the warp.  The step loop is compiled
        // twice (steps<SIMPLE>): per-step branches cost ~12 issue slots each on a 2-warps-per-scheduler kernel (ncu source page,
        // profiles/ncu_r02_s1_pack_and_walk.txt: 40 % of the warp samples sat in the 17 % of instructions around the row step),
        // so the common case keeps only four: the loop, `act`, the last-row capture and the 8-step bit flush.
        // (short-read geometries only: a long-read warp would alternate between the two copies from task to task and the second
        // copy costs more instruction-cache misses than the branches it removes -- C3 1510 -> 1383 GCUPS when both were compiled)
        const bool simple = G <= 8 && NSmax <= 1 && __reduce_max_sync(FULL, (unsigned)(K[0] | K[1])) == 0u;
        // boundary column g(x) + s x: plain packed arithmetic; (b1 + s) * 0x10001 mod 2^32 adds b1 + s to both halves, whatever its sign
        const uint32_t GD = (uint32_t)((sc.b1 + slope) * 0x10001), NX1 = dup16(-x1), SL = dup16(slope);

        for (int s = 0; s < NSmax; s++) {
            const bool act_s = anyrun && s < NS;
            const bool ownA = run[0] && s == NSh[0] - 1 && gl == lLh[0];
            const bool ownB = run[1] && s == NSh[1] - 1 && gl == lLh[1];
            const int Cs = (s == NS - 1) ? CsL : C;  // CsL == C unless the last stripe is narrow
            const int nb = Cs >> 3;
            const int y0 = s * W + gl * Cs;
            uint32_t Eh[C], B[C], sel[C];
            uint32_t wA[WPL], wB[WPL];
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int y = y0 + j + 1;
                uint32_t ca = 1, cb = 1;  // padding column: class "other"
                if (run[0] && y <= L2[0] && j < Cs) ca = lut_sm[readp[0][y - 1]];
                if (run[1] && y <= L2[1] && j < Cs) cb = lut_sm[readp[1][y - 1]];
                bad |= (ca >> 7) | ((cb >> 7) << 1);
                ca &= 7; cb &= 7;
                sel[j] = (ca * 0x11u | 0x80u) | ((cb * 0x11u | 0x80u) << 8);  // bytes [mA, sign(mA), mB, sign(mB)]
                const int g = sc.b0 + y * sc.b1 + bias;                        // row 0: S[0,y] = (MAXNEG, g(y), g(y))
                B[j] = dup16(g);
                Eh[j] = RB ? 0u : dup16(g - x1 - slope);                       // rust-bio: D[0][j] = MIN_SCORE (the sentinel 0); Eh' of row x carries s (x - 1)
            }
            uint32_t prevBl = dup16(((y0 == 0) ? 0 : sc.b0 + y0 * sc.b1) + bias);
            uint32_t oF = 0, oE = 0, oM = 0, oB = 0;
            uint32_t nF = 0, nE = 0, nM = 0, nB = 0;
            if (s > 0 && gl == 0 && act_s) { const uint4 v = *(const uint4*)(col_g + 4); nF = v.x; nE = v.y; nM = v.z; nB = v.w; }  // row 1 of the boundary column
            // per-step conditions as single integer compares (predicates do not survive the 900-instruction row step, so a
            // compound condition is re-evaluated from its parts every step)
            const uint32_t L1act = act_s ? (uint32_t)L1 : 0u;          // act  <=>  (unsigned)(x - 1) < L1act
            const int xcap = (ownA || ownB) ? L1 : -1;                 // the capture variant: only the lane(s) owning column L2, at row L1
            const bool first_col = (gl == 0) && (s == 0);
            const bool ld_col = (gl == 0) && (s > 0), st_col = (gl == G - 1) && (s < NS - 1);  // stripe-boundary column: consumer / producer lane
            const int k_own = max(ownA ? K[0] : 0, ownB ? K[1] : 0);  // band-skipped cell (x <= K, L2) of this lane's reads, if it owns column L2
            const bool stA = TB && run[0] && s < NSh[0], stB = TB && run[1] && s < NSh[1];
            // the profile row of the current reference class comes from a lane shuffle (lane k holds row k of the 16-row table):
            // no shared-memory address arithmetic in the loop; the class of the next row is read one step ahead
            // (the long-read geometries run at 3 CTAs/SM on 168 registers and cannot spare the two table registers: they load
            // the row from shared memory instead)
            constexpr bool TAB_SHFL = G <= 8;
            const uint2 tab_row = TAB_SHFL ? *(const uint2*)(tab_sm + (lane & 15) * 8) : make_uint2(0u, 0u);
            uint32_t rcur = act_s ? (uint32_t)ref_sm[0] : 0u;
            auto steps = [&](auto simple_tag) {
                constexpr bool SIMPLE = decltype(simple_tag)::value;
                const int nbx = SIMPLE ? C / 8 : nb;  // the single-stripe fast path runs all C / 8 blocks: no per-block branch
                uint32_t gB = dup16(sc.b0 + bias);  // lane 0 of stripe 0: g(x) + bias, advanced by one row per step (x = t there)
                int t = 1;
                if (Tmax >= 1) do {
                    const int x = t - gl;
                    uint32_t Fl = __shfl_up_sync(FULL, oF, 1, G);
                    uint32_t Bl = __shfl_up_sync(FULL, oB, 1, G);
                    uint32_t El = 0, Ml = 0;
                    if (TB && !RB) {
                        El = __shfl_up_sync(FULL, oE, 1, G);
                        Ml = __shfl_up_sync(FULL, oM, 1, G);
                    }
                    uint2 tr;
                    if (TAB_SHFL) {
                        tr.x = __shfl_sync(FULL, tab_row.x, (int)rcur);
                        tr.y = __shfl_sync(FULL, tab_row.y, (int)rcur);
                    } else {
                        tr = *(const uint2*)(tab_sm + rcur * 8);
                    }
                    gB += GD;
                    if ((uint32_t)(x - 1) < L1act) {
                        if (first_col) {  // S[x,0] = (MAXNEG, g(x), g(x)); selects, not a branch
                            Bl = gB;
                            Fl = gB + NX1;
                            El = Fl - SL;  // Eh' of row x carries s (x - 1)
                            Ml = 0;        // the sentinel: below every biased value
                            if (RB) Fl = 0;  // rust-bio: I[i][0] = MIN_SCORE on the empty-read boundary
                        }
                        if (!SIMPLE) {
                            if (ld_col) { Fl = nF; El = nE; Ml = nM; Bl = nB; }  // selects; the next row's values are fetched one step ahead
                            const bool nx = ld_col && x < L1;
                            ldg4_if(col_g + 4 * (x + 1), nx, nF, nE, nM, nB);
                        }
                        rcur = ref_sm[x < L1 ? x : L1 - 1];  // class of row x + 1 (clamped on the last row: unused)
                        const uint32_t BlIn = Bl;
                        // the capture variant only runs on the lane(s) that own column L2, at their last row: once per task
                        if (x == xcap)
                            pack_row_step<C, TB, true, RB, MADD>(Eh, B, sel, wA, wB, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, LE, X1, X1M1, LEe, X1b, ownA, jLh[0], ownB, jLh[1], cap, nbx, cy);
                        else
                            pack_row_step<C, TB, false, RB, MADD>(Eh, B, sel, wA, wB, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, LE, X1, X1M1, LEe, X1b, ownA, jLh[0], ownB, jLh[1], cap, nbx, cy);
                        prevBl = BlIn;
                        oF = Fl; oE = El; oM = Ml; oB = Bl;
                        if (!SIMPLE) {
                            // band-skipped cells (x <= K, y == L2): fresh-matrix state (0,0,0), per read
                            if (x <= k_own) {
                                const bool sa = ownA && x <= K[0], sb = ownB && x <= K[1];
                                // true values (0,0,0): stored with this row's slope terms (Eh' carries s (x - 1), the rest s x)
                                const int b0v = bias + slope * x, f0 = -x1 + b0v, e0 = f0 - slope;
#pragma unroll
                                for (int j = 0; j < C; j++) {
                                    if (sa && j == jLh[0]) { Eh[j] = set_lo(Eh[j], e0); B[j] = set_lo(B[j], b0v); }
                                    if (sb && j == jLh[1]) { Eh[j] = set_hi(Eh[j], e0); B[j] = set_hi(B[j], b0v); }
                                }
                                if (sa && jLh[0] == Cs - 1) { oF = set_lo(oF, f0); oE = set_lo(oE, e0); oM = set_lo(oM, b0v); oB = set_lo(oB, b0v); }
                                if (sb && jLh[1] == Cs - 1) { oF = set_hi(oF, f0); oE = set_hi(oE, e0); oM = set_hi(oM, b0v); oB = set_hi(oB, b0v); }
                                if (x == L1) {
                                    if (sa) { cap[0] = set_lo(cap[0], b0v); cap[1] = set_lo(cap[1], e0); cap[2] = set_lo(cap[2], f0); }
                                    if (sb) { cap[0] = set_hi(cap[0], b0v); cap[1] = set_hi(cap[1], e0); cap[2] = set_hi(cap[2], f0); }
                                }
                            }
                        }
                        if (TB) bits_store2<G, WPL>(tt_sm, bitsA, bitsB, wA, wB, stA, stB, s, T, t, lane, gl, x == L1, nb);
                        if (!SIMPLE) {
                            stg4_if(col_g + 4 * x, st_col, oF, oE, oM, oB);
                        }
                    }
                } while (++t <= Tmax);
            };
            if (G <= 8 && simple) steps(std::true_type{}); else steps(std::false_type{});
            __syncwarp();
        }

        cy.report(anyrun, false, L1, L2[0], L2[1], (int)(TB ? 1 : 0) | (RB ? 2 : 0) | (MADD ? 4 : 0) | (G << 8) | (slope << 16));
        // ---- final cells: score + start layer = LAST maximum of (M, E, F) per read ----
        if (RB) {
            const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (gw * G));
            const unsigned b0 = __ballot_sync(FULL, (bad & 1u) != 0), b1 = __ballot_sync(FULL, (bad & 2u) != 0);
            if (ok[0] && (b0 & gm)) { status[0] = CLQ_SCORING_NOT_REPRESENTABLE; }
            if (ok[1] && (b1 & gm)) { status[1] = CLQ_SCORING_NOT_REPRESENTABLE; }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int src = gw * G + lLh[h];
            const uint32_t c0 = __shfl_sync(FULL, cap[0], src), c1 = __shfl_sync(FULL, cap[1], src), c2 = __shfl_sync(FULL, cap[2], src);
            int score = 0, z = 0;
            if (run[h]) {
                const int cM = (h ? get_hi(c0) : get_lo(c0)) - bias - slope * L1;
                const int cE = (h ? get_hi(c1) : get_lo(c1)) - bias - slope * (L1 - 1) + x1;
                const int cF = (h ? get_hi(c2) : get_lo(c2)) - bias - slope * L1 + x1;
                score = cM; z = 0;
                if (RB) {  // rust-bio: match first, then insertion, then deletion, each only when strictly better
                    if (cF > score) { score = cF; z = 2; }
                    if (cE > score) { score = cE; z = 1; }
                } else {
                    if (cE >= score) { score = cE; z = 1; }
                    if (cF >= score) { score = cF; z = 2; }
                }
            } else if (ok[h]) {
                const int n = L1 > L2[h] ? L1 : L2[h];
                if (n > 0) { score = sc.b0 + n * sc.b1; z = 2; }
            }
            if (valid[h] && gl == 0) {
                if (!TB && p.all_pairs) p.scores[(size_t)ridx[h] * p.n_refs + ref] = ok[h] ? score : INT32_MIN;
                else {
                    clq_result_t r;
                    r.score_scaled = score; r.ref_index = (status[h] != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu; r.cigar_off = 0; r.cigar_len = 0; r.status = status[h]; r.matches = 0; r.mismatches = 0;
                    p.results[ridx[h]] = r;
                }
                if (run[h]) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2[h]);
            }
            if (TB && tvalid && gl == 0) {
                TbRec rec;
                rec.ridx = ridx[h]; rec.L1 = (ok[h] && status[h] == CLQ_OK) ? L1 : -1; rec.L2 = L2[h];
                rec.zK = z | (K[h] << 2) | ((CsL >> 3) << 20) | ((narrow && anyrun ? NS - 1 : 0) << 24);
                p.tb_rec[2 * task + h] = rec;
            }
        }
    }
}

}  // namespace clq

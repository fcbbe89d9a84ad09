// clq_pack_adapt.cuh -- s16x2 PACK fill for pairs whose scores do NOT fit a static 15-bit window (long reads: a 5 kb pair
// spans 70 000 score units), with an adaptive per-row bias and an exact fallback.
//
// Idea.  Inside one column stripe the values of a row differ by at most ~(match - e) * W, far less than 2^15; what overflows a
// static window is the drift of the whole row as x grows (+match per row above the diagonal, e below it).  So row x of stripe s
// is stored as  v' = v + beta_s(x),  beta_s(x) = beta_s(x - 1) + sigma_x,  with a small integer slope sigma_x in [-16, 15] that is
// constant over blocks of kAdaptBlock rows and is chosen at run time, one block ahead, from the centre of the stored values the
// lane group sees (a proportional controller that steers the centre of the row to the middle of the window).  Only constants of
// the recurrence change (same derivation as the static slope of clq_pack.cuh, with s -> sigma_x):
//   M'  = B'[x-1,y-1] + (m + sigma_x)                 profile rows for every sigma live in shared memory (32 copies, 4 KB)
//   Eh' = max(Eh'_up + (le + sigma_{x-1}), B'_up)      Fh' = max(Fh'_left + le, B'_left)          [Eh' of row x carries beta(x-1)]
//   P'  = max(Fh' + x1, M')                            B'  = max(Eh' + (x1 + sigma_x), P')
//   ext2: t2' = max(Eh'_left + (x1 - 1 + sigma_x), M'_left);  every direction bit compares two values of one row: unchanged.
// The schedule is part of the arithmetic only through beta, which cancels in every comparison, so ANY schedule gives the exact
// result as long as no stored half leaves [64, 32767].  That is not provable for arbitrary reads, so it is CHECKED: every lane
// tracks min / max of the first stored B of each of its rows (two ops per row), the boundary values a stripe inherits are
// checked after re-biasing, and all other cells of a lane-row lie within `guard` of the tracked one (clq_api.cu derives the
// guard from the scoring: adjacent cells differ by at most match - 2 * x1).  A task that leaves the guarded window -- or two
// reads that drift apart, since both halves share one bias -- is handed to the int32 kernel through the retry list: its
// records are written by that kernel, bit-exact either way.
//
// Pair mode only (two reads against one reference; the host orders reads so that pairs share theirs), traceback only.  Same
// direction bits, same slots, same walker as pack_kernel.  Runs on (8,40) with as many column stripes as the read needs: the
// short-read geometry keeps 8 of 8 lanes busy and is 24 % faster on 1 kb reads than (32,32) even with four stripes (measured,
// profiles/geometry_sweep_r02_s13.txt).
#pragma once

#include "clq_pack.cuh"

namespace clq {

#ifndef CLQ_ADAPT_MIN_BLOCKS
#define CLQ_ADAPT_MIN_BLOCKS 2  // 255 registers: at 168 (3 CTAs/SM) the compiler re-derives ~230 instructions of loop-invariant state per step (C5 335 vs 322 ms)
#endif
constexpr int kAdaptBlock = 64;     // rows per slope block (>= G: every lane has switched to a block's slope before the next decision)
constexpr int kAdaptSigmaMin = -16, kAdaptSigmaMax = 15, kAdaptTabs = 32;
constexpr int kAdaptCentre = 16384;
constexpr int kAdaptMaxRebias = 30000;  // |beta_s - beta_{s-1}| beyond this is treated as an overflow: with |d| <= 30000 a re-biased value that wraps lands outside [64, 32767] as s16 and is caught

struct AdaptParams {
    int32_t guard;              // every cell of a lane-row lies within `guard` of the lane's first column in that row
    uint32_t* retry_list;       // sub-batch positions (2 * task, 2 * task + 1) of the pairs the int32 kernel must redo
    unsigned int* retry_count;  // reset per sub-batch
    unsigned int* retry_total;  // per launch (statistics)
    uint32_t ref_stride;        // bytes of shared memory per staged reference row: TWO reference classes per byte (a 5 kb
                                // amplicon is 2.5 KB per lane group, 40 KB per CTA on (8,40): two CTAs per SM instead of one)
};

__device__ __forceinline__ uint32_t vmin_s16x2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
__device__ __forceinline__ uint32_t vmax_s16x2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ int lo16s(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
__device__ __forceinline__ int hi16s(uint32_t w) { return (int)(int16_t)(w >> 16); }

template <int G, int C>
__global__ void __launch_bounds__(kThreads, CLQ_ADAPT_MIN_BLOCKS) pack_adapt_kernel(const KParams p, const AdaptParams ap) {
    static_assert(C % 8 == 0, "C must be a multiple of 8");
    static_assert(kAdaptBlock >= G, "a block must be at least as long as the wavefront skew");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // [0,256) class LUT, [256, 256 + 32 * 128) one profile table per slope, then the reference rows
    constexpr int kAdaptTabBytes = kAdaptTabs * kTabBytes;
    uint8_t* smem = smem_raw + kLutBytes + kAdaptTabBytes;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) smem_raw[i] = p.cls_lut[i];
    for (int i = threadIdx.x; i < kAdaptTabs * 128; i += blockDim.x) {
        const int sg = i / 128 + kAdaptSigmaMin, e = i % 128;
        smem_raw[kLutBytes + i] = (uint8_t)(int8_t)(((const int8_t*)p.tab)[e] + sg);  // the host checked that every sum fits int8
    }
    __syncthreads();
    const uint8_t* lut_sm = smem_raw;
    const uint8_t* tab_sm = smem_raw + kLutBytes;
    constexpr int W = G * C;
    constexpr int WPL = C / 8;
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gw = lane / G;
    const int wpb = blockDim.x >> 5;
    const uint32_t ggid = (blockIdx.x * wpb + warp) * GPW + gw;
    uint8_t* ref_sm = smem + (size_t)(warp * GPW + gw) * ap.ref_stride;
    // per-warp transposition buffers (read A, read B) of the direction bits for the time-transposed layout (G <= 8), after the reference rows
    uint32_t* tt_sm = reinterpret_cast<uint32_t*>(smem + (size_t)wpb * GPW * ap.ref_stride) + (size_t)warp * (2 * WPL * 256);
    uint32_t* col_g = (uint32_t*)p.col_scratch + (size_t)ggid * 5 * p.col_stride;  // F, E, M, B of the boundary column + beta per row
    const clq_affine_t sc = p.sc;
    const int x1 = sc.oe_in, le = sc.e_in;
    const uint32_t LE = dup16(le), X1 = dup16(x1), NX1 = dup16(-x1);
    const int lo_lim = 64 + ap.guard, hi_lim = 32767 - ap.guard;
    int staged_ref = -1;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.task_counter, (unsigned)GPW);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.n_tasks) break;
        const uint32_t task = base + gw;
        const bool tvalid = task < p.n_tasks;
        uint32_t ridx[2] = {0, 0};
        bool valid[2] = {false, false};
        int ref = -1;
        if (tvalid) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t pos = p.task_base + 2 * task + h;
                if (pos < p.task_end) {
                    ridx[h] = p.order ? p.order[pos] : pos;
                    valid[h] = ridx[h] != 0xffffffffu;  // padding position (the host pads every reference group to an even count)
                }
            }
            for (int h = 1; h >= 0; h--)
                if (valid[h]) {
                    const int rh = p.ref_of_read[ridx[h]];
                    if (rh >= 0 && (uint32_t)rh < p.n_refs) ref = rh;
                }
        }
        int L1 = 0, L2[2] = {0, 0};
        const uint8_t* refp = nullptr;
        const uint8_t* readp[2] = {nullptr, nullptr};
        uint32_t status[2] = {CLQ_OK, CLQ_OK};
        bool ok[2], run[2];
        if (ref >= 0 && (uint32_t)ref < p.n_refs) {
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
                const int rh = p.ref_of_read[ridx[h]];
                if ((uint32_t)L2[h] >= p.max_read_len) status[h] = CLQ_READ_TOO_LONG;
                else if (rh < 0 || (uint32_t)rh >= p.n_refs || rh != ref) status[h] = CLQ_NO_CANDIDATE;
            }
            ok[h] = valid[h] && status[h] == CLQ_OK;
            run[h] = ok[h] && L1 > 0 && L2[h] > 0;
        }
        const bool anyrun = run[0] || run[1];
        if (anyrun && ref != staged_ref) {
            for (int b = gl; 2 * b < L1; b += G) {  // two classes per byte, low nibble first
                const uint32_t c0 = (lut_sm[refp[2 * b]] >> 3) & 15, c1 = (2 * b + 1 < L1) ? (lut_sm[refp[2 * b + 1]] >> 3) & 15 : 0u;
                ref_sm[b] = (uint8_t)(c0 | (c1 << 4));
            }
            staged_ref = ref;
        }
        __syncwarp();

        const int L2m = max(run[0] ? L2[0] : 0, run[1] ? L2[1] : 0);
        const int NS = anyrun ? (L2m + W - 1) / W : 0;
        const bool narrow = NS - 1 <= kMaxNarrowStripe;
        const int CsL = (anyrun && narrow) ? narrow_cols<G>(L2m - (NS - 1) * W, C) : C;
        int K[2], NSh[2], lLh[2], jLh[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            K[h] = run[h] ? stale_rows(L1, L2[h], p.band_mode) : 0;
            NSh[h] = run[h] ? (L2[h] + W - 1) / W : 0;
            const int cL = run[h] ? (L2[h] - 1) - (NSh[h] - 1) * W : 0;
            const int csh = (NSh[h] == NS) ? CsL : C;
            lLh[h] = cL / csh;
            jLh[h] = cL - lLh[h] * csh;
        }
        const int NSmax = __reduce_max_sync(FULL, NS);
        const int T = anyrun ? L1 + G - 1 : 0;
        const int Tmax = __reduce_max_sync(FULL, T);
        Canary cy;
        uint32_t cap[3] = {0, 0, 0};
        int cap_beta[2] = {0, 0}, cap_beta_prev[2] = {0, 0};  // beta(L1), beta(L1 - 1) of the stripe that holds column L2 of read h
        uint32_t* bitsA = p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task);
        uint32_t* bitsB = valid[1] ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task + 1) : nullptr;
        // the stored values this lane has seen: first column of every row (per half) + re-biased boundary values; and hard faults
        uint32_t seen_mn = dup16(kAdaptCentre), seen_mx = dup16(kAdaptCentre);
        uint32_t fault = 0;  // reason bits: 1 boundary row, 2 first column, 4 re-bias too large, 8 stale cell, 16 watched values left the window

        for (int s = 0; s < NSmax; s++) {
            const bool act_s = anyrun && s < NS;
            const bool ownA = run[0] && s == NSh[0] - 1 && gl == lLh[0];
            const bool ownB = run[1] && s == NSh[1] - 1 && gl == lLh[1];
            const int Cs = (s == NS - 1) ? CsL : C;
            const int nb = Cs >> 3;
            const int y0 = s * W + gl * Cs;
            cy.nA = run[0] ? max(0, min(Cs, L2[0] - y0)) : 0;  // canary builds: only real cells are guarded
            cy.nB = run[1] ? max(0, min(Cs, L2[1] - y0)) : 0;
            // beta_s(0): the boundary row g(y) of this stripe is centred on the window
            const int ymid = s * W + min(W, max(L2m - s * W, 1)) / 2;
            int beta = kAdaptCentre - (sc.b0 + ymid * sc.b1);
            const int beta0 = beta;
            uint32_t Eh[C], B[C], sel[C];
            uint32_t wA[WPL], wB[WPL];
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int y = y0 + j + 1;
                uint32_t ca = 1, cb = 1;
                if (run[0] && y <= L2[0] && j < Cs) ca = lut_sm[readp[0][y - 1]];
                if (run[1] && y <= L2[1] && j < Cs) cb = lut_sm[readp[1][y - 1]];
                ca &= 7; cb &= 7;
                sel[j] = (ca * 0x11u | 0x80u) | ((cb * 0x11u | 0x80u) << 8);
                const int g = sc.b0 + y * sc.b1 + beta0;   // row 0: S[0,y] = (MAXNEG, g(y), g(y)); beta(-1) = beta(0)
                B[j] = dup16(g);
                Eh[j] = dup16(g - x1);
            }
            {   // the boundary row itself must sit inside the guarded window
                const int g_first = sc.b0 + (y0 + 1) * sc.b1 + beta0, g_last = sc.b0 + (y0 + C) * sc.b1 + beta0;
                if (act_s && (min(g_first, g_last) < lo_lim || max(g_first, g_last) > hi_lim)) fault |= 1u;
            }
            uint32_t prevBl = dup16(((y0 == 0) ? 0 : sc.b0 + y0 * sc.b1) + beta0);
            uint32_t oF = 0, oE = 0, oM = 0, oB = 0;
            uint32_t nF = 0, nE = 0, nM = 0, nB = 0;
            int nBeta = 0, pBeta = 0;  // beta_{s-1}(x + 1) prefetched, beta_{s-1}(x - 1)
            const bool first_col = (gl == 0) && (s == 0);
            const bool ld_col = (gl == 0) && (s > 0) && act_s, st_col = (gl == G - 1) && (s < NS - 1) && act_s;
            if (ld_col) {
                { const uint4 v = *(const uint4*)(col_g + 4); nF = v.x; nE = v.y; nM = v.z; nB = v.w; }
                nBeta = (int)col_g[4 * p.col_stride + 1];
                pBeta = (int)col_g[4 * p.col_stride + 0];  // beta_{s-1}(0), stored by the producer with row 1
            }
            const uint32_t L1act = act_s ? (uint32_t)L1 : 0u;
            const int xcap = (ownA || ownB) ? L1 : -1;
            const int k_own = max(ownA ? K[0] : 0, ownB ? K[1] : 0);
            const bool stA = run[0] && s < NSh[0], stB = run[1] && s < NSh[1];
            // which halves of this lane's FIRST column are real cells (padding columns may hold anything)
            const bool realA = run[0] && y0 + 1 <= L2[0], realB = run[1] && y0 + 1 <= L2[1];
            const uint32_t real_mask = (realA ? 0x0000ffffu : 0u) | (realB ? 0xffff0000u : 0u);
            const uint32_t neutral = dup16(kAdaptCentre) & ~real_mask;
            uint32_t rcur = act_s ? (uint32_t)ref_sm[0] & 15u : 0u;
            // slope schedule: rows <= blk_end take sig_cur, later rows sig_next (decided at step blk_end, one block ahead)
            int sig_cur = 0, sig_next = 0, sig_prev_row = 0, blk_end = kAdaptBlock;
            int ctr_prev = kAdaptCentre, sig_last = 0;  // controller state (uniform over the group)
            {   // first block: cancel the drift the scoring predicts (rows above the diagonal gain match - e, rows below lose e)
                const int drift = (s * W >= L1) ? -(sc.match - le) : -(sc.match - le) * 3 / 4;
                sig_cur = max(kAdaptSigmaMin, min(kAdaptSigmaMax, drift));
                sig_last = sig_cur;
            }

            for (int t = 1; t <= Tmax; t++) {
                const int x = t - gl;
                uint32_t Fl = __shfl_up_sync(FULL, oF, 1, G);
                uint32_t Bl = __shfl_up_sync(FULL, oB, 1, G);
                uint32_t El = __shfl_up_sync(FULL, oE, 1, G);
                uint32_t Ml = __shfl_up_sync(FULL, oM, 1, G);
                if ((uint32_t)(x - 1) < L1act) {
                    // this row's slope and bias
                    if (x > blk_end) { sig_cur = sig_next; blk_end += kAdaptBlock; }
                    const int sg = sig_cur, sgp = sig_prev_row;
                    sig_prev_row = sg;
                    const int beta_prev = beta;
                    beta += sg;
                    const uint32_t X1b = dup16(x1 + sg), X1M1 = dup16(x1 - 1 + sg), LEe = dup16(le + sgp);
                    if (first_col) {  // S[x,0] = (MAXNEG, g(x), g(x))
                        const int g = sc.b0 + x * sc.b1 + beta;
                        if (g < lo_lim || g > hi_lim) fault |= 2u;
                        Bl = dup16(g);
                        Fl = Bl + NX1;
                        El = Fl - (uint32_t)(sg * 0x10001);  // Eh' of row x carries beta(x - 1); (sg * 0x10001) mod 2^32 subtracts sg from both halves
                        Ml = 0;               // the sentinel: below every stored value
                    }
                    if (ld_col) {
                        // inherit the boundary column of stripe s - 1, stored under ITS bias: re-bias to this stripe's
                        const int d = beta - nBeta, dE = beta_prev - pBeta;
                        if (max(abs(d), abs(dE)) > kAdaptMaxRebias) fault |= 4u;
                        const uint32_t D = (uint32_t)(d * 0x10001), DE = (uint32_t)(dE * 0x10001);
                        Fl = nF + D; El = nE + DE; Ml = nM + D; Bl = nB + D;
                        // (only the halves that hold real cells in this stripe: the padded partner of an odd pair drifts freely)
                        seen_mn = vmin_s16x2(seen_mn, (vmin_s16x2(Bl, vmin_s16x2(El, Fl)) & real_mask) | neutral);
                        seen_mx = vmax_s16x2(seen_mx, (vmax_s16x2(Bl, vmax_s16x2(Ml, Fl)) & real_mask) | neutral);
                        pBeta = nBeta;
                    }
                    {
                        const bool nx = ld_col && x < L1;
                        ldg4_if(col_g + 4 * (x + 1), nx, nF, nE, nM, nB);
                        nBeta = (int)ldg_if(col_g + 4 * p.col_stride + x + 1, nx, (uint32_t)nBeta);
                    }
                    const uint2 tr = *(const uint2*)(tab_sm + ((sg - kAdaptSigmaMin) * 16 + (int)rcur) * 8);
                    { const int xi = x < L1 ? x : L1 - 1; rcur = ((uint32_t)ref_sm[xi >> 1] >> ((xi & 1) << 2)) & 15u; }
                    const uint32_t BlIn = Bl;
                    if (x == xcap) {
                        pack_row_step<C, true, true, false, false>(Eh, B, sel, wA, wB, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, LE, X1, X1M1, LEe, X1b, ownA, jLh[0], ownB, jLh[1], cap, nb, cy);
                        if (ownA) { cap_beta[0] = beta; cap_beta_prev[0] = beta_prev; }
                        if (ownB) { cap_beta[1] = beta; cap_beta_prev[1] = beta_prev; }
                    } else
                        pack_row_step<C, true, false, false, false>(Eh, B, sel, wA, wB, Fl, El, Ml, Bl, prevBl, tr.x, tr.y, LE, X1, X1M1, LEe, X1b, ownA, jLh[0], ownB, jLh[1], cap, nb, cy);
                    prevBl = BlIn;
                    oF = Fl; oE = El; oM = Ml; oB = Bl;
                    if (x <= k_own) {
                        // band-skipped cells (x <= K, y == L2): fresh-matrix state (0,0,0), stored under this row's bias
                        const bool sa = ownA && x <= K[0], sb = ownB && x <= K[1];
                        const int b0v = beta, f0 = -x1 + beta, e0 = -x1 + beta_prev;
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
                        if (b0v < lo_lim || b0v > hi_lim) fault |= 8u;
                    }
                    // overflow watch: the first stored B of this lane-row (padding halves neutralised)
                    {
                        const uint32_t b0w = (B[0] & real_mask) | neutral;
                        seen_mn = vmin_s16x2(seen_mn, b0w);
                        seen_mx = vmax_s16x2(seen_mx, b0w);
                    }
                    bits_store2<G, WPL>(tt_sm, bitsA, bitsB, wA, wB, stA, stB, s, T, t, lane, gl, x == L1, nb);
                    stg4_if(col_g + 4 * x, st_col, oF, oE, oM, oB);
                    stg_if(col_g + 4 * p.col_stride + x, st_col, (uint32_t)beta);
                    if (x == 1) stg_if(col_g + 4 * p.col_stride, st_col, (uint32_t)beta0);
                }
                // ---- slope decision for the next block (every lane of the group computes the same numbers) ----
                if ((t % kAdaptBlock) == 0) {
                    const bool in_rows = (uint32_t)(x - 1) < L1act;
                    uint32_t vmx = in_rows ? ((B[0] & real_mask) | (dup16(-32768) & ~real_mask)) : dup16(-32768);
                    uint32_t vmn = in_rows ? ((B[0] & real_mask) | (dup16(32767) & ~real_mask)) : dup16(32767);
#pragma unroll
                    for (int o = G / 2; o >= 1; o >>= 1) {
                        vmx = vmax_s16x2(vmx, __shfl_xor_sync(FULL, vmx, o, G));
                        vmn = vmin_s16x2(vmn, __shfl_xor_sync(FULL, vmn, o, G));
                    }
                    const int hi = max(lo16s(vmx), hi16s(vmx)), lo = min(lo16s(vmn), hi16s(vmn));
                    int sg_new = sig_last;
                    if (hi >= lo) {  // some lane holds real cells
                        const int ctr = (hi + lo) / 2;
                        // centre_{b+1} = centre_b + R (g + sigma_b): cancel the drift seen over the last block, close half of the offset
                        sg_new = sig_last - (ctr - ctr_prev) / kAdaptBlock + (kAdaptCentre - ctr) / (2 * kAdaptBlock);
                        sg_new = max(kAdaptSigmaMin, min(kAdaptSigmaMax, sg_new));
                        ctr_prev = ctr;
                    }
                    sig_last = sg_new;
                    sig_next = sg_new;
                }
            }
            __syncwarp();
        }

        // ---- did every stored value stay inside the guarded window? (per group; padding halves were neutralised) ----
        {
            const int mn = min(lo16s(seen_mn), hi16s(seen_mn)), mx = max(lo16s(seen_mx), hi16s(seen_mx));
            if (mn < lo_lim || mx > hi_lim) fault |= 16u;
#if CLQ_PACK_CANARY
            if (fault && anyrun) printf("adapt retry: reason %u lane %d L1 %d L2 %d %d seen %d..%d limits %d..%d\n", fault, gl, L1, L2[0], L2[1], mn, mx, lo_lim, hi_lim);
#endif
        }
        const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (gw * G));
        const unsigned faulted = __ballot_sync(FULL, fault != 0u);  // every lane votes (an idle group must not skip the warp-wide vote)
        const bool redo = anyrun && (faulted & gm) != 0u;
        cy.report(anyrun, redo, L1, L2[0], L2[1], 0x1000 | G);  // a pair that is NOT redone must never have left the window: the guard band's soundness, checked

        // ---- final cells: score + start layer = LAST maximum of (M, E, F) per read ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int src = gw * G + lLh[h];
            const uint32_t c0 = __shfl_sync(FULL, cap[0], src), c1 = __shfl_sync(FULL, cap[1], src), c2 = __shfl_sync(FULL, cap[2], src);
            const int bL = __shfl_sync(FULL, cap_beta[h], src), bP = __shfl_sync(FULL, cap_beta_prev[h], src);
            int score = 0, z = 0;
            if (run[h]) {
                const int cM = (h ? get_hi(c0) : get_lo(c0)) - bL;
                const int cE = (h ? get_hi(c1) : get_lo(c1)) - bP + x1;
                const int cF = (h ? get_hi(c2) : get_lo(c2)) - bL + x1;
                score = cM; z = 0;
                if (cE >= score) { score = cE; z = 1; }
                if (cF >= score) { score = cF; z = 2; }
            } else if (ok[h]) {
                const int n = L1 > L2[h] ? L1 : L2[h];
                if (n > 0) { score = sc.b0 + n * sc.b1; z = 2; }
            }
            if (gl == 0 && tvalid) {
                if (redo) {
                    // the int32 kernel redoes this pair: it writes the result record, the walker record and the direction bits
                    if (valid[h]) ap.retry_list[atomicAdd(ap.retry_count, 1u)] = 2 * task + h;
                    TbRec rec;
                    rec.ridx = ridx[h]; rec.L1 = -1; rec.L2 = 0; rec.zK = 0;
                    p.tb_rec[2 * task + h] = rec;
                    if (valid[h]) atomicAdd(ap.retry_total, 1u);
                } else {
                    if (valid[h]) {
                        clq_result_t r;
                        r.score_scaled = score;
                        r.ref_index = (status[h] != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu;
                        r.cigar_off = 0; r.cigar_len = 0; r.status = status[h]; r.matches = 0; r.mismatches = 0;
                        p.results[ridx[h]] = r;
                        if (run[h]) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2[h]);
                    }
                    TbRec rec;
                    rec.ridx = ridx[h]; rec.L1 = (ok[h] && status[h] == CLQ_OK) ? L1 : -1; rec.L2 = L2[h];
                    rec.zK = z | (K[h] << 2) | ((CsL >> 3) << 20) | ((narrow && anyrun ? NS - 1 : 0) << 24);
                    p.tb_rec[2 * task + h] = rec;
                }
            }
        }
    }
}

}  // namespace clq

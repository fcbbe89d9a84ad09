// clq_convex_pack.cuh -- s16x2 ("PACK") variant of the two-piece affine ("convex") kernel: one lane group aligns TWO reads
// against the same reference, read A in the low and read B in the high 16-bit half of every register.  Same recurrence,
// same 8 direction bits per cell, same row layout and the same walker as convex_kernel (clq_convex.cuh) -- PARITY UNPINNED
// like that kernel (the reference has no convex DP; semantics in oracle/clq_oracle.c::orc_convex_align_pair).
//
// Exactness: as in clq_pack.cuh the host proves per batch that every stored value fits a 15-bit window
// (clq_api.cu: B >= gap(x) + gap(y) >= 2*o1 + (L1+L2+W)*e1, gap states are stored minus their o_i + e_i so they sit within
// one open of B, nothing exceeds max(match, special) * min(L1, L2)); values carry a common bias so both halves stay in
// [64, 32767], the boundary sentinel of the M layer is 0, and every direction bit is min_u16x2(x - y, 1) of two values
// with x >= y per half.  Pair mode only (tasks = read pairs against one reference; multi-reference batches are bucketed by
// reference first, ref_scatter_kernel).
#pragma once

#include "clq_convex.cuh"
#include "clq_pack.cuh"

namespace clq {

template <int C, bool TB, bool LAST>
__device__ __forceinline__ void convex_pack_row_step(uint32_t (&E1)[C], uint32_t (&E2)[C], uint32_t (&B)[C], const uint32_t (&sel)[C],
                                                     uint32_t (&wA)[C / 4], uint32_t (&wB)[C / 4], uint32_t& F1, uint32_t& F2, uint32_t& Bl,
                                                     uint32_t diag, uint32_t tlo, uint32_t thi, uint32_t LE1, uint32_t LE2, uint32_t X1,
                                                     uint32_t X2, bool ownA, int jA, bool ownB, int jB, uint32_t (&cap)[5], int nb) {
    const uint32_t ONE = 0x00010001u;
    uint32_t acc0 = 0, acc1 = 0;
#pragma unroll
    for (int jb = 0; jb < C / 16; jb++) {
      if (jb < nb) {  // narrow last stripe: only nb blocks of 16 columns per lane are real
#pragma unroll
        for (int jj = 0; jj < 16; jj++) {
        const int j = jb * 16 + jj;
        const uint32_t m = (uint32_t)prmt_s8(tlo, thi, sel[j]);
        const uint32_t Mv = __viaddmax_s16x2(diag, m, X1);  // per-half add (biased values are > 0 > x1)
        const uint32_t BU = B[j];
        const uint32_t E1n = __viaddmax_s16x2(E1[j], LE1, BU);
        const uint32_t E2n = __viaddmax_s16x2(E2[j], LE2, BU);
        const uint32_t F1n = __viaddmax_s16x2(F1, LE1, Bl);
        const uint32_t F2n = __viaddmax_s16x2(F2, LE2, Bl);
        const uint32_t P1 = __viaddmax_s16x2(F1n, X1, Mv);
        const uint32_t P2 = __viaddmax_s16x2(F2n, X2, P1);
        const uint32_t B1 = __viaddmax_s16x2(E1n, X1, P2);
        const uint32_t Bn = __viaddmax_s16x2(E2n, X2, B1);
        if (TB) {
            // byte pair [extE1 extE2 extF1 extF2 | F1>M  F2>P1  E1>P2  E2>B1] of this cell pair (low half read A, high half read B)
            uint32_t by = __vminu2(E1n - BU, ONE);
            by = by * 2u + __vminu2(E2n - BU, ONE);
            by = by * 2u + __vminu2(F1n - Bl, ONE);
            by = by * 2u + __vminu2(F2n - Bl, ONE);
            by = by * 2u + __vminu2(P1 - Mv, ONE);
            by = by * 2u + __vminu2(P2 - P1, ONE);
            by = by * 2u + __vminu2(B1 - P2, ONE);
            by = by * 2u + __vminu2(Bn - B1, ONE);
            if ((j & 3) == 0) acc0 = by;
            else if ((j & 3) == 1) acc0 = acc0 * 256u + by;
            else if ((j & 3) == 2) acc1 = by;
            else {
                acc1 = acc1 * 256u + by;
                wA[j >> 2] = __byte_perm(acc1, acc0, 0x5410);  // low halves: read A's 4 cells, cell 0 in the top byte
                wB[j >> 2] = __byte_perm(acc1, acc0, 0x7632);  // high halves: read B's
            }
        }
        if (LAST) {
            if (ownA && j == jA) {
                cap[0] = set_lo(cap[0], get_lo(Mv)); cap[1] = set_lo(cap[1], get_lo(P1)); cap[2] = set_lo(cap[2], get_lo(P2));
                cap[3] = set_lo(cap[3], get_lo(B1)); cap[4] = set_lo(cap[4], get_lo(Bn));
            }
            if (ownB && j == jB) {
                cap[0] = set_hi(cap[0], get_hi(Mv)); cap[1] = set_hi(cap[1], get_hi(P1)); cap[2] = set_hi(cap[2], get_hi(P2));
                cap[3] = set_hi(cap[3], get_hi(B1)); cap[4] = set_hi(cap[4], get_hi(Bn));
            }
        }
        diag = BU;
        E1[j] = E1n; E2[j] = E2n; B[j] = Bn;
        F1 = F1n; F2 = F2n; Bl = Bn;
        }
      }
    }
}

template <int G, int C, bool TB>
__global__ void __launch_bounds__(kThreads) convex_pack_kernel(const KParams p, const ConvexParams cp, const PackParams pp) {
    static_assert(C % 16 == 0, "C must be a multiple of 16 (8 direction bits per cell, 128-bit row stores)");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + kLutBytes + kTabBytes;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) smem_raw[i] = p.cls_lut[i];
    if (threadIdx.x < 32) ((uint32_t*)(smem_raw + kLutBytes))[threadIdx.x] = p.tab[threadIdx.x];
    __syncthreads();
    const uint8_t* lut_sm = smem_raw;
    const uint8_t* tab_sm = smem_raw + kLutBytes;
    constexpr int GPW = 32 / G;
    constexpr int W = G * C;
    constexpr int WPL = C / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gw = lane / G;
    const int wpb = blockDim.x >> 5;
    const uint32_t ggid = (blockIdx.x * wpb + warp) * GPW + gw;
    uint8_t* ref_sm = smem + (size_t)(warp * GPW + gw) * p.ref_sm_stride;
    uint32_t* col_g = (uint32_t*)p.col_scratch + (size_t)ggid * 4 * p.col_stride;
    const clq_convex_t cv = cp.cv;
    const int bias = pp.bias;
    const int x1 = cv.o1 + cv.e1, x2 = cv.o2 + cv.e2;
    const uint32_t LE1 = dup16(cv.e1), LE2 = dup16(cv.e2), X1 = dup16(x1), X2 = dup16(x2);
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
                    valid[h] = ridx[h] != 0xffffffffu;
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
        if (ref >= 0) {
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
            for (int i = gl; i < L1; i += G) ref_sm[i] = (lut_sm[refp[i]] >> 3) & 15;
            staged_ref = ref;
        }
        __syncwarp();

        const int L2m = max(run[0] ? L2[0] : 0, run[1] ? L2[1] : 0);
        const int NS = anyrun ? (L2m + W - 1) / W : 0;
        // narrow last stripe of the pair (clq_kernels.cuh::narrow_cols), in blocks of 16 columns (one 128-bit store of bits)
        const bool narrow = G >= 16 && NS - 1 <= kMaxNarrowStripe;
        int CsL = C;
        if (anyrun && narrow) {
            const int R = L2m - (NS - 1) * W;
            CsL = min(C, (R + G * 16 - 1) / (G * 16) * 16);
        }
        int NSh[2], lLh[2], jLh[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            NSh[h] = run[h] ? (L2[h] + W - 1) / W : 0;
            const int cL = run[h] ? (L2[h] - 1) - (NSh[h] - 1) * W : 0;
            const int csh = (NSh[h] == NS) ? CsL : C;
            lLh[h] = cL / csh;
            jLh[h] = cL - lLh[h] * csh;
        }
        const int NSmax = __reduce_max_sync(FULL, NS);
        const int T = anyrun ? L1 + G - 1 : 0;
        const int Tmax = __reduce_max_sync(FULL, T);
        uint32_t cap[5] = {0, 0, 0, 0, 0};
        uint32_t* bitsA = TB ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task) : nullptr;
        uint32_t* bitsB = (TB && valid[1]) ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, 2 * task + 1) : nullptr;

        for (int s = 0; s < NSmax; s++) {
            const bool act_s = anyrun && s < NS;
            const bool ownA = run[0] && s == NSh[0] - 1 && gl == lLh[0];
            const bool ownB = run[1] && s == NSh[1] - 1 && gl == lLh[1];
            const int Cs = (G >= 16 && s == NS - 1) ? CsL : C;
            const int nb = Cs >> 4;
            const int y0 = s * W + gl * Cs;
            uint32_t E1[C], E2[C], B[C], sel[C];
            uint32_t wA[WPL], wB[WPL];
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int y = y0 + j + 1;
                uint32_t ca = 1, cb = 1;
                if (run[0] && y <= L2[0] && j < Cs) ca = lut_sm[readp[0][y - 1]] & 7;
                if (run[1] && y <= L2[1] && j < Cs) cb = lut_sm[readp[1][y - 1]] & 7;
                sel[j] = (ca * 0x11u | 0x80u) | ((cb * 0x11u | 0x80u) << 8);
                const int g1 = cv.o1 + y * cv.e1, g2 = cv.o2 + y * cv.e2;  // row 0: E_i = F_i = o_i + y*e_i, M = NEG
                B[j] = dup16(max(g1, g2) + bias);
                E1[j] = dup16(g1 - x1 + bias);
                E2[j] = dup16(g2 - x2 + bias);
            }
            uint32_t prevBl = dup16(((y0 == 0) ? 0 : max(cv.o1 + y0 * cv.e1, cv.o2 + y0 * cv.e2)) + bias);
            uint32_t oF1 = 0, oF2 = 0, oB = 0;
            uint32_t nF1 = 0, nF2 = 0, nB = 0;
            if (s > 0 && gl == 0 && act_s) { nF1 = col_g[1]; nF2 = col_g[p.col_stride + 1]; nB = col_g[3 * p.col_stride + 1]; }
            int rnext = act_s ? ref_sm[0] : 0;

            for (int t = 1; t <= Tmax; t++) {
                const int x = t - gl;
                uint32_t F1 = __shfl_up_sync(FULL, oF1, 1, G);
                uint32_t F2 = __shfl_up_sync(FULL, oF2, 1, G);
                uint32_t Bl = __shfl_up_sync(FULL, oB, 1, G);
                const bool act = act_s && x >= 1 && x <= L1;
                if (act) {
                    if (gl == 0) {
                        if (s == 0) {
                            const int g1 = cv.o1 + x * cv.e1, g2 = cv.o2 + x * cv.e2;
                            F1 = dup16(g1 - x1 + bias); F2 = dup16(g2 - x2 + bias); Bl = dup16(max(g1, g2) + bias);
                        } else {
                            F1 = nF1; F2 = nF2; Bl = nB;
                            if (x < L1) { nF1 = col_g[x + 1]; nF2 = col_g[p.col_stride + x + 1]; nB = col_g[3 * p.col_stride + x + 1]; }
                        }
                    }
                    const int r = rnext;
                    if (x < L1) rnext = ref_sm[x];
                    const uint32_t BlIn = Bl;
                    const uint2 tr = *(const uint2*)(tab_sm + r * 8);
                    if (x == L1)
                        convex_pack_row_step<C, TB, true>(E1, E2, B, sel, wA, wB, F1, F2, Bl, prevBl, tr.x, tr.y, LE1, LE2, X1, X2, ownA, jLh[0], ownB, jLh[1], cap, nb);
                    else
                        convex_pack_row_step<C, TB, false>(E1, E2, B, sel, wA, wB, F1, F2, Bl, prevBl, tr.x, tr.y, LE1, LE2, X1, X2, ownA, jLh[0], ownB, jLh[1], cap, nb);
                    prevBl = BlIn;
                    oF1 = F1; oF2 = F2; oB = Bl;
                    if (TB) {
                        const size_t rowoff = (size_t)(s * T + (t - 1)) * (G * WPL);
                        if (run[0] && s < NSh[0]) {
#pragma unroll
                            for (int k = 0; k < WPL; k += 4)
                                if (k < nb * 4) *reinterpret_cast<uint4*>(bitsA + rowoff + (k / 4) * (G * 4) + gl * 4) = make_uint4(wA[k], wA[k + 1], wA[k + 2], wA[k + 3]);
                        }
                        if (run[1] && s < NSh[1]) {
#pragma unroll
                            for (int k = 0; k < WPL; k += 4)
                                if (k < nb * 4) *reinterpret_cast<uint4*>(bitsB + rowoff + (k / 4) * (G * 4) + gl * 4) = make_uint4(wB[k], wB[k + 1], wB[k + 2], wB[k + 3]);
                        }
                    }
                    if (gl == G - 1 && s < NS - 1) { col_g[x] = oF1; col_g[p.col_stride + x] = oF2; col_g[3 * p.col_stride + x] = oB; }
                }
            }
            __syncwarp();
        }

#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int src = gw * G + lLh[h];
            int v[5];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const uint32_t c = __shfl_sync(FULL, cap[k], src);
                v[k] = h ? get_hi(c) : get_lo(c);
            }
            int score = 0, z = 0;
            if (run[h]) {
                const int Mv = v[0], P1 = v[1], P2 = v[2], B1 = v[3], Bn = v[4];
                score = Bn - bias;
                z = (Bn > B1) ? 2 : ((B1 > P2) ? 1 : ((P2 > P1) ? 4 : ((P1 > Mv) ? 3 : 0)));
            } else if (ok[h]) {
                const int n = L1 > L2[h] ? L1 : L2[h];
                if (n > 0) {
                    const int g1 = cv.o1 + n * cv.e1, g2 = cv.o2 + n * cv.e2;
                    score = max(g1, g2); z = g2 > g1 ? 4 : 3;
                }
            }
            if (valid[h] && gl == 0) {
                clq_result_t r;
                r.score_scaled = score; r.ref_index = (status[h] != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu; r.cigar_off = 0; r.cigar_len = 0; r.status = status[h]; r.matches = 0; r.mismatches = 0;
                p.results[ridx[h]] = r;
                if (run[h]) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2[h]);
            }
            if (TB && tvalid && gl == 0) {
                TbRec rec;
                rec.ridx = ridx[h]; rec.L1 = ok[h] ? L1 : -1; rec.L2 = L2[h];
                rec.zK = z | ((CsL >> 4) << 20) | ((narrow && anyrun ? NS - 1 : 0) << 24);
                p.tb_rec[2 * task + h] = rec;
            }
        }
    }
}

}  // namespace clq

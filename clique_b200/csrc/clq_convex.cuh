// clq_convex.cuh -- two-piece affine ("convex") global alignment, the K4 kernel of SURVEY.md section 2a.
//
// PARITY UNPINNED: the reference has no convex DP (only the never-called ConvexScoring::gap,
// alignment/scoring_functions.rs:36-53).  The semantics are defined by this repository and restated on the CPU in
// oracle/clq_oracle.c::orc_convex_align_pair, which is what the parity tests compare against:
//   gap of length k costs max(o1 + k*e1, o2 + k*e2);  states M, E1, E2 (Del), F1, F2 (Ins);
//   A(cell) = argmax over (M, F1, F2, E1, E2) scanned in that order with strict '>' (earlier wins ties), B = that max;
//   M  = B(x-1,y-1) + m;   Ei = max(Ei_up + ei, B_up + oi + ei);   Fi = max(Fi_left + ei, B_left + oi + ei);
//   a gap state extends iff strictly greater than opening, otherwise it came from A of the source cell;
//   score = B(L1,L2), start state = A(L1,L2); boundary / sentinel / trailing run as in the affine path.
// Same wavefront as gotoh_kernel (G lanes x C register columns, shuffles to the right neighbour, column stripes), with the
// gap states kept shifted by -(oi+ei) so every max-plus step is one VIADDMNMX; 8 direction bits per cell
//   [extE1 extE2 extF1 extF2 | F1>M  F2>P1  E1>P2  E2>B1]   (sign bits of eight differences, shifted in with SHF)
// stored as one row of G*C/4 words per step; convex_walk_kernel (one thread per pair) walks them back.
#pragma once

#include "clq_kernels.cuh"

namespace clq {

template <int C, bool TB, bool LAST>
__device__ __forceinline__ void convex_row_step(int (&E1)[C], int (&E2)[C], int (&B)[C], const int (&sel)[C], uint32_t (&w)[C / 4],
                                                int& F1, int& F2, int& Bl, int diag, uint32_t tlo, uint32_t thi, int e1, int e2, int x1,
                                                int x2, bool own_last, int jL, int& capB, int& capZ) {
#pragma unroll
    for (int j = 0; j < C; j++) {
        const int m = prmt_s8(tlo, thi, (uint32_t)sel[j]);
        const int Mv = diag + m;
        const int BU = B[j];
        const int E1n = __viaddmax_s32(E1[j], e1, BU);
        const int E2n = __viaddmax_s32(E2[j], e2, BU);
        const int F1n = __viaddmax_s32(F1, e1, Bl);
        const int F2n = __viaddmax_s32(F2, e2, Bl);
        const int P1 = __viaddmax_s32(F1n, x1, Mv);
        const int P2 = __viaddmax_s32(F2n, x2, P1);
        const int B1 = __viaddmax_s32(E1n, x1, P2);
        const int Bn = __viaddmax_s32(E2n, x2, B1);
        if (TB) {
            uint32_t acc = w[j >> 2];
            acc = __funnelshift_l((uint32_t)(BU - E1n), acc, 1);  // E1 extends
            acc = __funnelshift_l((uint32_t)(BU - E2n), acc, 1);  // E2 extends
            acc = __funnelshift_l((uint32_t)(Bl - F1n), acc, 1);  // F1 extends
            acc = __funnelshift_l((uint32_t)(Bl - F2n), acc, 1);  // F2 extends
            acc = __funnelshift_l((uint32_t)(Mv - P1), acc, 1);   // F1 > M
            acc = __funnelshift_l((uint32_t)(P1 - P2), acc, 1);   // F2 > max(M,F1)
            acc = __funnelshift_l((uint32_t)(P2 - B1), acc, 1);   // E1 > max(M,F1,F2)
            acc = __funnelshift_l((uint32_t)(B1 - Bn), acc, 1);   // E2 > max(M,F1,F2,E1)
            w[j >> 2] = acc;
        }
        if (LAST) {
            if (own_last && j == jL) {
                capB = Bn;
                capZ = (Bn > B1) ? 2 : ((B1 > P2) ? 1 : ((P2 > P1) ? 4 : ((P1 > Mv) ? 3 : 0)));
            }
        }
        diag = BU;
        E1[j] = E1n; E2[j] = E2n; B[j] = Bn;
        F1 = F1n; F2 = F2n; Bl = Bn;
    }
}

struct ConvexParams {
    clq_convex_t cv;
};

template <int G, int C, bool TB>
__global__ void __launch_bounds__(kThreads) convex_kernel(const KParams p, const ConvexParams cp) {
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
    constexpr int WPL = C / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane % G, gw = lane / G;
    const int wpb = blockDim.x >> 5;
    const uint32_t ggid = (blockIdx.x * wpb + warp) * GPW + gw;
    uint8_t* ref_sm = smem + (size_t)(warp * GPW + gw) * p.ref_sm_stride;
    int32_t* col_g = p.col_scratch + (size_t)ggid * 4 * p.col_stride;
    const clq_convex_t cv = cp.cv;
    const int x1 = cv.o1 + cv.e1, x2 = cv.o2 + cv.e2;
    int staged_ref = -1;

    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(p.task_counter, (unsigned)GPW);
        base = __shfl_sync(FULL, base, 0);
        if (base >= p.n_tasks) break;
        const uint32_t task = base + gw;
        bool valid = task < p.n_tasks;
        uint32_t ridx = 0;
        int ref = -1;
        uint32_t* bits_g = TB ? p.bits + bits_slot(p.bits_off, p.bits_stride, p.task_base, task) : nullptr;
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
            for (int i = gl; i < L1; i += G) ref_sm[i] = (lut_sm[refp[i]] >> 3) & 15;
            staged_ref = ref;
        }
        __syncwarp();

        const int NS = run ? (L2 + W - 1) / W : 0;
        const int NSmax = __reduce_max_sync(FULL, NS);
        const int T = run ? L1 + G - 1 : 0;
        const int Tmax = __reduce_max_sync(FULL, T);
        const int cL = run ? (L2 - 1) - (NS - 1) * W : 0;
        const int lL = cL / C, jL = cL - lL * C;
        int capB = 0, capZ = 0;

        for (int s = 0; s < NSmax; s++) {
            const bool act_s = run && s < NS;
            const bool own_last = act_s && s == NS - 1 && gl == lL;
            const int y0 = s * W + gl * C;
            int E1[C], E2[C], B[C], sel[C];
            uint32_t w[WPL];
#pragma unroll
            for (int j = 0; j < C; j++) {
                const int y = y0 + j + 1;
                int code = 1;
                if (act_s && y <= L2) code = (int)lut_sm[readp[y - 1]] & 7;
                sel[j] = code * 0x1111 | 0x8880;
                const int g1 = cv.o1 + y * cv.e1, g2 = cv.o2 + y * cv.e2;  // row 0: E_i = F_i = o_i + y*e_i, M = NEG
                B[j] = max(g1, g2);
                E1[j] = g1 - x1;
                E2[j] = g2 - x2;
            }
            int prevBl = (y0 == 0) ? 0 : max(cv.o1 + y0 * cv.e1, cv.o2 + y0 * cv.e2);
            int oF1 = 0, oF2 = 0, oB = 0;
            int nF1 = 0, nF2 = 0, nB = 0;
            if (s > 0 && gl == 0 && act_s) { nF1 = col_g[1]; nF2 = col_g[p.col_stride + 1]; nB = col_g[3 * p.col_stride + 1]; }
            int rnext = act_s ? ref_sm[0] : 0;

            for (int t = 1; t <= Tmax; t++) {
                const int x = t - gl;
                int F1 = __shfl_up_sync(FULL, oF1, 1, G);
                int F2 = __shfl_up_sync(FULL, oF2, 1, G);
                int Bl = __shfl_up_sync(FULL, oB, 1, G);
                const bool act = act_s && x >= 1 && x <= L1;
                if (act) {
                    if (gl == 0) {
                        if (s == 0) {
                            const int g1 = cv.o1 + x * cv.e1, g2 = cv.o2 + x * cv.e2;
                            F1 = g1 - x1; F2 = g2 - x2; Bl = max(g1, g2);
                        } else {
                            F1 = nF1; F2 = nF2; Bl = nB;
                            if (x < L1) { nF1 = col_g[x + 1]; nF2 = col_g[p.col_stride + x + 1]; nB = col_g[3 * p.col_stride + x + 1]; }
                        }
                    }
                    const int r = rnext;
                    if (x < L1) rnext = ref_sm[x];
                    const int BlIn = Bl;
                    const uint2 tr = *(const uint2*)(tab_sm + r * 8);
                    if (x == L1)
                        convex_row_step<C, TB, true>(E1, E2, B, sel, w, F1, F2, Bl, prevBl, tr.x, tr.y, cv.e1, cv.e2, x1, x2, own_last, jL, capB, capZ);
                    else
                        convex_row_step<C, TB, false>(E1, E2, B, sel, w, F1, F2, Bl, prevBl, tr.x, tr.y, cv.e1, cv.e2, x1, x2, own_last, jL, capB, capZ);
                    prevBl = BlIn;
                    oF1 = F1; oF2 = F2; oB = Bl;
                    if (TB) {
                        uint32_t* row = bits_g + (size_t)(s * T + (t - 1)) * (G * WPL);
#pragma unroll
                        for (int k = 0; k < WPL; k += 4)
                            *reinterpret_cast<uint4*>(row + (k / 4) * (G * 4) + gl * 4) = make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]);
                    }
                    if (gl == G - 1 && s < NS - 1) { col_g[x] = oF1; col_g[p.col_stride + x] = oF2; col_g[3 * p.col_stride + x] = oB; }
                }
            }
            __syncwarp();
        }

        const int src = gw * G + lL;
        capB = __shfl_sync(FULL, capB, src);
        capZ = __shfl_sync(FULL, capZ, src);
        int score = 0, z = 0;
        if (run) { score = capB; z = capZ; }
        else if (ok) {
            const int n = L1 > L2 ? L1 : L2;
            if (n > 0) {  // S[n,0] / S[0,n]: argmax over (M=NEG, F1, F2, E1, E2) -> F1 unless piece 2 is strictly better
                const int g1 = cv.o1 + n * cv.e1, g2 = cv.o2 + n * cv.e2;
                score = max(g1, g2); z = g2 > g1 ? 4 : 3;
            }
        }
        if (valid && gl == 0) {
            if (!TB && p.all_pairs) p.scores[(size_t)ridx * p.n_refs + ref] = ok ? score : INT32_MIN;
            else {
                clq_result_t r;
                r.score_scaled = score; r.ref_index = (status != CLQ_NO_CANDIDATE && ref >= 0 && (uint32_t)ref < p.n_refs) ? (uint32_t)ref : 0xffffffffu; r.cigar_off = 0; r.cigar_len = 0; r.status = status; r.matches = 0; r.mismatches = 0;
                p.results[ridx] = r;
                if (TB) {
                    TbRec rec;
                    rec.ridx = ridx; rec.L1 = ok ? L1 : -1; rec.L2 = L2; rec.zK = z;
                    p.tb_rec[task] = rec;
                }
            }
            if (run) atomicAdd(p.cells, (unsigned long long)L1 * (unsigned long long)L2);
        }
    }
}

// traceback over the 8-bit direction records of convex_kernel; states: 0 M, 1 E1, 2 E2, 3 F1, 4 F2
template <int G, int C>
__global__ void __launch_bounds__(128) convex_walk_kernel(const TbRec* recs, uint32_t n_tasks, const uint32_t* bits, uint64_t bits_stride,
                                                          const uint64_t* bits_off, uint32_t task_base,
                                                          uint32_t* cig_scratch, uint32_t cig_stride, uint32_t* cigar_pool, uint64_t cigar_cap,
                                                          unsigned long long* cigar_cursor, clq_result_t* results, const uint8_t* ref_bytes,
        const uint64_t* ref_off, const uint8_t* read_bytes, const uint64_t* read_off) {
    constexpr int W = G * C;
    constexpr int WPL = C / 4;
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_tasks) return;
    const TbRec rec = recs[q];
    if (rec.L1 < 0) return;
    const int L1 = rec.L1, L2 = rec.L2;
    int z = rec.zK & 7;
    const int CsL = ((rec.zK >> 20) & 15) * 16, sL = (rec.zK >> 24) & 255;  // narrow last stripe (convex PACK kernel, G >= 16)
    const int T = L1 + G - 1;
    const uint32_t* bits_g = bits + bits_slot(bits_off, bits_stride, task_base, q);
    uint32_t* cig_g = cig_scratch + (size_t)q * cig_stride;
    uint32_t status = CLQ_OK;
    int x = L1, y = L2;
    // get_reference_alignment_rate (consensus/consensus_builders.rs:288-307) fused into the walk: only M columns can count
    const uint8_t* refp = ref_bytes + ref_off[results[rec.ridx].ref_index];
    const uint8_t* readp = read_bytes + read_off[rec.ridx];
    uint32_t n_match = 0, n_mismatch = 0;
    auto count = [&](int xx, int yy) {
        const uint8_t rb = __ldg(refp + xx - 1), qb = __ldg(readp + yy - 1);
        if (rb > 64 && rb != 'N' && qb > 64) { if (rb == qb) n_match++; else n_mismatch++; }
    };
    int cpos = (int)cig_stride;
    uint32_t cur_op = 3, cur_len = 0;
    auto emit = [&](uint32_t op, uint32_t n) {
        if (op == cur_op) cur_len += n;
        else {
            if (cur_len) cig_g[--cpos] = (cur_len << 4) | cur_op;
            cur_op = op; cur_len = n;
        }
    };
    auto rec8 = [&](int xx, int yy) -> uint32_t {
        int c = yy - 1;
        const int s = c / W;
        c -= s * W;
        const int cs = (G >= 16 && s == sL && CsL > 0) ? CsL : C;
        const int ln = c / cs, j = c - ln * cs;
        const int k = j >> 2;
        const size_t idx = (size_t)(s * T + (xx + ln - 1)) * (G * WPL) + (k / 4) * (G * 4) + ln * 4 + (k & 3);
        return (__ldg(bits_g + idx) >> (24 - 8 * (j & 3))) & 255u;
    };
    auto argmax = [](uint32_t b) -> int { return (b & 1u) ? 2 : ((b & 2u) ? 1 : ((b & 4u) ? 4 : ((b & 8u) ? 3 : 0))); };
    uint32_t cur = (x > 0 && y > 0) ? rec8(x, y) : 0;
    while (x > 0 && y > 0) {
        const uint32_t old = cur;
        bool ext;
        if (z == 0) { emit(CLQ_OP_M, 1); count(x, y); x--; y--; ext = false; }
        else if (z <= 2) { emit(CLQ_OP_D, 1); x--; ext = (old >> (z == 1 ? 7 : 6)) & 1u; }
        else { emit(CLQ_OP_I, 1); y--; ext = (old >> (z == 3 ? 5 : 4)) & 1u; }
        if (x == 0 || y == 0) break;
        cur = rec8(x, y);
        if (!ext) z = argmax(cur);
    }
    if (x > 0) emit(CLQ_OP_D, (uint32_t)x);
    if (y > 0) emit(CLQ_OP_I, (uint32_t)y);
    if (cur_len) cig_g[--cpos] = (cur_len << 4) | cur_op;
    int nops = (int)cig_stride - cpos;
    unsigned long long off = 0;
    if (nops > 0) {
        off = atomicAdd(cigar_cursor, (unsigned long long)nops);
        if (off + (unsigned long long)nops > cigar_cap) { status = CLQ_CIGAR_POOL_FULL; nops = 0; }
    }
    for (int i = 0; i < nops; i++) cigar_pool[off + i] = cig_g[cpos + i];
    clq_result_t* r = results + rec.ridx;
    r->cigar_off = (uint32_t)off;
    r->cigar_len = (uint32_t)nops;
    r->status = status;
    r->matches = n_match;
    r->mismatches = n_mismatch;
}

}  // namespace clq

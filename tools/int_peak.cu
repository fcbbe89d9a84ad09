// INT32 / DPX issue-rate microbenchmark for the B200 roofline denominator.
// Each thread runs ITER iterations over NACC independent accumulator chains of one op.
// Reports warp-lane ops per clock per SM (from clock64) and Tops/s (from CUDA events).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NACC 8
#define ITER 4096

enum Op { ADD, MAX, VIADDMAX, VIMAX3, VIBMAX, PRMT, ADDMAX16, MAX3_16, ADDMAXU16, IMAD, LOP3, SHF, SEL, SHFL,
          MIX_DPX_IMAD, MIX_CELL, MIX_CELL16, UMULHI, IADD3, VMINU2, LDS_RAND, MIX_ALU2_FMA2, NOPS };
static const char* names[] = {"iadd", "imax", "viaddmax_s32", "vimax3_s32", "vibmax_s32", "prmt", "viaddmax_s16x2",
          "vimax3_s16x2", "viaddmax_u16x2", "imad", "lop3", "shf", "isetp+sel", "shfl_up",
          "mix(viaddmax+imad)", "cell6(int32)", "cell6(s16x2)",
          "umulhi(imad.hi)", "iadd3", "vminu2", "lds_rand_u32", "mix(2dpx+2imad)"};
// ops counted per accumulator per iteration
static const int opcount[] = {1,1,1,1,1,1,1,1,1,1,1,1,2,1,2,6,6,1,1,1,1,4};

template <int OP>
__global__ void __launch_bounds__(256) k(int* out, int b, int c, int d, unsigned long long* cyc) {
    __shared__ int lds_tab[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) lds_tab[i] = i * 2654435761u >> 23;
    __syncthreads();
    int a[NACC];
    int e[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { a[i] = threadIdx.x * 7 + i * b; e[i] = threadIdx.x + i * c; }
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (OP == ADD) a[i] = a[i] + b;
            else if (OP == MAX) a[i] = max(a[i] ^ b, c);   // xor folded? keep dependent: see LOP3 count
            else if (OP == VIADDMAX) a[i] = __viaddmax_s32(a[i], b, c + i);
            else if (OP == VIMAX3) a[i] = __vimax3_s32(a[i], e[i], c) ;
            else if (OP == VIBMAX) { bool p; a[i] = __vibmax_s32(a[i], e[i], &p); e[i] = p ? e[i] : d; }
            else if (OP == PRMT) a[i] = __byte_perm(a[i], b, e[i]);
            else if (OP == ADDMAX16) a[i] = __viaddmax_s16x2(a[i], b, c + i);
            else if (OP == MAX3_16) a[i] = __vimax3_s16x2(a[i], e[i], c);
            else if (OP == ADDMAXU16) a[i] = __viaddmax_u16x2(a[i], b, c + i);
            else if (OP == IMAD) a[i] = a[i] * b + c;
            else if (OP == LOP3) a[i] = (a[i] & b) ^ c;
            else if (OP == SHF) a[i] = __funnelshift_l(a[i], b, 3);
            else if (OP == SEL) a[i] = (a[i] > e[i]) ? b : (a[i] + c);
            else if (OP == SHFL) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
            else if (OP == MIX_DPX_IMAD) { a[i] = __viaddmax_s32(a[i], b, c); e[i] = e[i] * b + d; }
            else if (OP == MIX_CELL) {
                // one Gotoh cell, score-only, shifted-E/F form: prmt, add, 2x viaddmax, max, viaddmax
                int m = __byte_perm(b, c, e[i]);            // substitution lookup
                int M = a[i] + m;                           // diag + m
                int E = __viaddmax_s32(e[i], d, a[i]);      // Ehat
                int F = __viaddmax_s32(E, d, M);            // Fhat (fake dependency)
                int t = max(E, F);
                a[i] = __viaddmax_s32(t, c, M);
                e[i] = E;
            } else if (OP == UMULHI) a[i] = __umulhi((unsigned)a[i], (unsigned)b << 16) + i;
            else if (OP == IADD3) a[i] = a[i] + e[i] + b;
            else if (OP == VMINU2) a[i] = __vminu2(a[i], e[i]) + 0;
            else if (OP == LDS_RAND) a[i] = lds_tab[(a[i] + i) & 511];
            else if (OP == MIX_ALU2_FMA2) { a[i] = __viaddmax_s16x2(a[i], b, c); e[i] = e[i] * b + d; a[i] = __viaddmax_s16x2(a[i], d, e[i]); e[i] = e[i] * c + a[i]; }
            else if (OP == MIX_CELL16) {
                int m = __byte_perm(b, c, e[i]);
                int M = a[i] + m;
                int E = __viaddmax_s16x2(e[i], d, a[i]);
                int F = __viaddmax_s16x2(E, d, M);
                int t = __vimax3_s16x2(E, F, E);
                a[i] = __viaddmax_s16x2(t, c, M);
                e[i] = E;
            }
        }
    }
    unsigned long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += a[i] + e[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int nsm, int* out, unsigned long long* cyc, int b, int c, int d) {
    const int blocks = nsm * 8, threads = 256;   // 8 CTAs x 8 warps = 64 warps/SM
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, b, c, d, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, b, c, d, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = (unsigned long long*)malloc(blocks * 8);
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double lane_ops = (double)blocks * threads * ITER * NACC * opcount[OP];
    // all 8 CTAs of an SM run concurrently for ~avg cycles
    double per_sm_clk = (double)8 * threads * ITER * NACC * opcount[OP] / avg;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"tops\": %.3f, \"lanes_per_clk_per_sm\": %.2f, \"mhz_eff\": %.0f}\n",
           names[OP], ms, lane_ops / ms / 1e9, per_sm_clk, avg / ms / 1e3);
    free(h);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
    int* out; unsigned long long* cyc;
    cudaMalloc(&out, nsm * 8 * 256 * 4); cudaMalloc(&cyc, nsm * 8 * 8);
    int b = 3, c = 5, d = -2;
    if (getenv("NEVER")) { b = 17; c = 1; d = 9; }
    run<ADD>(nsm, out, cyc, b, c, d);
    run<MAX>(nsm, out, cyc, b, c, d);
    run<VIADDMAX>(nsm, out, cyc, b, c, d);
    run<VIMAX3>(nsm, out, cyc, b, c, d);
    run<VIBMAX>(nsm, out, cyc, b, c, d);
    run<PRMT>(nsm, out, cyc, b, c, d);
    run<ADDMAX16>(nsm, out, cyc, b, c, d);
    run<MAX3_16>(nsm, out, cyc, b, c, d);
    run<ADDMAXU16>(nsm, out, cyc, b, c, d);
    run<IMAD>(nsm, out, cyc, b, c, d);
    run<LOP3>(nsm, out, cyc, b, c, d);
    run<SHF>(nsm, out, cyc, b, c, d);
    run<SEL>(nsm, out, cyc, b, c, d);
    run<SHFL>(nsm, out, cyc, b, c, d);
    run<MIX_DPX_IMAD>(nsm, out, cyc, b, c, d);
    run<MIX_CELL>(nsm, out, cyc, b, c, d);
    run<MIX_CELL16>(nsm, out, cyc, b, c, d);
    run<UMULHI>(nsm, out, cyc, b, c, d);
    run<IADD3>(nsm, out, cyc, b, c, d);
    run<VMINU2>(nsm, out, cyc, b, c, d);
    run<LDS_RAND>(nsm, out, cyc, b, c, d);
    run<MIX_ALU2_FMA2>(nsm, out, cyc, b, c, d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}

// clq_api.cu -- host side of libclq: context, device buffers, stream slots, kernel dispatch (C ABI of include/clq.h).
// There is no CPU path in this file: every alignment result is produced by the kernels in clq_kernels.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "clq_kernels.cuh"
#include "clq_convex.cuh"
#include "clq_pack.cuh"
#include "clq_convex_pack.cuh"
#include "clq_pack_adapt.cuh"
#include "clq_reads2bit.cuh"

namespace clq {
size_t pack2_plain_run(const uint8_t* bytes, size_t n_words, uint32_t* packed);  // clq_pack2_host.cpp
}

using namespace clq;

namespace {

// -DCLQ_LEAN=1: development build that instantiates only the (8,40) and (32,32) geometries (convex: (8,16), (32,32)); any other
// geometry falls through to (32,32).  Cuts the compile time of a kernel experiment from ~90 s to ~25 s.  Never shipped.
#ifndef CLQ_LEAN
#define CLQ_LEAN 0
#endif
#define CLQ_LEAN_NO8 0
#if CLQ_LEAN
#define CLQ_FULL_CASE(n, stmt)
#else
#define CLQ_FULL_CASE(n, stmt) case n: stmt;
#endif

struct Cfg { int G, C; };
// wavefront geometries: G lanes per pair x C columns per lane (stripe width W = G*C)
const Cfg kCfgs[] = {{8, 16}, {8, 24}, {8, 40}, {16, 24}, {32, 16}, {32, 32}};
constexpr int kNumCfgs = sizeof(kCfgs) / sizeof(kCfgs[0]);

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // overlapped sub-batches (clq_launch): odd sub-batches run here, in the other half of the scratch
    cudaEvent_t fork = nullptr, join = nullptr;
    cudaEvent_t ev[6] = {};
    cudaEvent_t done = nullptr;      // recorded after the last kernel of a launch (see clq_ctx::last_done)
    DevBuf read_bytes, read_off, fixed_ref, order, results, scores, cand_mask, single_ref, ref_of_read, votes;
    DevBuf cigar_pool, bits, cig_scratch, col_scratch, tb_rec, bits_off, tags, ref_groups;
    std::vector<uint32_t> h_len;     // read lengths in processing order (only when lengths vary: have_order)
    std::vector<int32_t> h_ref;      // fixed_ref in processing order (same condition, when given)
    std::vector<uint32_t> h_order;   // the processing order itself (read index per position, 0xffffffff = padding)
    DevBuf order2;                   // the order of one launch after its sub-batches were dealt (clq_launch)
    DevBuf counters;                 // [0..3] task counters (u32, padded to 8 B each), [4] cigar cursor, [5] cells, [6] retry count, [7] retries of
                                     // the launch, [8..10] fill / retry task counters and retry count of the sub-batches on stream2
    unsigned long long* h_counters = nullptr;  // pinned mirror
    uint32_t n_reads = 0;
    uint64_t n_read_bytes = 0;
    uint32_t max_len = 0, min_len = 0;
    bool have_order = false, have_fixed = false;
    bool order_by_ref = false;       // the processing order groups the reads by their fixed reference (each group padded to an even
                                     // number of positions with 0xffffffff) so that the reads of a PACK pair share theirs
    uint32_t n_pos = 0;              // processing positions incl. that padding (== n_reads otherwise)
    DevBuf retry_list;               // pack_adapt_kernel -> int32 retry pass
    DevBuf packed2, exc_pos, exc_byte;  // clq_upload_packed2: the 2-bit stream and its exception list before the expansion
    int state = 0;  // 0 empty, 1 uploaded, 2 launched, 3 downloaded
    uint32_t flags = 0;
    clq_stats_t stats = {};
    int n_dp = 0;
    bool launched = false;
};

}  // namespace

struct clq_ctx {
    int device = 0;
    int sm_count = 0;
    clq_limits_t lim = {};
    std::string err;
    DevBuf ref_bytes, ref_off, kmer_keys, kmer_owner, kmer_keys64, tag_slot;
    uint32_t tag_stride = 0;         // most tag columns ('0'..'9') of any reference, rounded up to 4
    std::vector<uint8_t> h_ref_bytes;
    std::vector<uint64_t> h_ref_off;
    uint32_t n_refs = 0, max_ref_len = 0;
    uint32_t kmer_k = 0, kmer_skip = 0, n_keys = 0;
    std::vector<Slot> slots;
    int force_cfg = -1;
    int debug_flags = 0;
    int no_pack = 0;                 // option "no_pack": never take the s16x2 PACK kernels
    int no_madd = 0;                 // option "no_madd": PACK kernels without the static row slope (M step on the ALU pipe)
    int no_adapt = 0;                // option "no_adapt": long pairs beyond the static 15-bit window stay on the int32 kernels
    int no_overlap = 0;              // option "no_overlap": sub-batches run one after the other (dealt round-robin) instead of two at a time
    int no_long8 = 0;                // option "no_long8": long reads keep the geometry their length picks instead of (8,40) with column stripes
    int adapt_guard = 0;             // option "adapt_guard" (tests): overrides the guard band of pack_adapt_kernel; a huge value forces every pair through the retry pass
    int no_group = 0;                // option "no_group": multi-reference traceback stays on the int32 kernels (no bucketing by reference)
    int force_generic = 0;           // option "force_generic": never take the FAST (PRMT/DPX) kernel variant
    bool fast_ok = false;            // the reference set has <= 6 distinct non-special bytes
    uint8_t cls[256] = {};           // byte -> class: 0 special, 1 other, 2..7 reference bytes (| row << 3, clq_kernels.cuh)
    DevBuf cls_lut;
    bool rb_ok = false;              // rust-bio mode: the reference set fits the 16-row / 8-column profile
    uint8_t cls_rb[256] = {};        // rust-bio classes: 0 = 'N', 1 other, 2..7 reference bytes; row << 3; bit 7 = unscorable in a read
    DevBuf cls_lut_rb;
    // Direction bits of one fill + walk round.  With the slot chain (serialize_slots, the default) the kernels of different slots
    // never overlap, so ONE scratch serves every slot and can be large: a round must hold several waves of tasks, and a pair of
    // 5 kb reads alone is 25 MB of bits and ~60 ms of one warp -- rounds shorter than that are bounded by their longest task.
    int64_t max_scratch_bytes = 96ll << 30;
    DevBuf sh_bits, sh_cig, sh_tbrec;        // the shared scratch (per-slot buffers are used instead when serialize_slots = 0)
    // Kernels of different slots are chained in submission order: every fill kernel is a persistent grid that fills the GPU, so
    // two launches sharing the SMs only finish together and leave the host nothing to overlap with.  With the chain slot B's
    // H2D copy runs under slot A's kernels, A finishes first, and its results are handled while B computes.
    cudaEvent_t last_done = nullptr;
    int serialize = 1;               // option "serialize_slots"
};

namespace {

int32_t fail(clq_ctx* c, int32_t code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(c, CLQ_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

int32_t ensure(clq_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return CLQ_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) return fail(c, CLQ_E_NOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    b.cap = want;
    return CLQ_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

template <int G, int C, bool TB, bool FIN, bool FAST, bool RB = false>
cudaError_t launch_one(const KParams& p, int sm_count, size_t smem, cudaStream_t st, int* grid_out, bool query_only) {
    auto kern = gotoh_kernel<G, C, TB, FIN, FAST, RB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    int grid = nb * sm_count;
    if (*grid_out > 0) grid = std::min(grid, *grid_out);  // caller-imposed cap (scratch budget)
    *grid_out = grid;
    if (query_only) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool TB, bool FIN, bool FAST, bool RB = false>
cudaError_t launch_cfg(int cfg, const KParams& p, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    switch (cfg) {
        CLQ_FULL_CASE(0, return (launch_one<8, 16, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(1, return (launch_one<8, 24, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q)))
        case 2: return launch_one<8, 40, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q);
        CLQ_FULL_CASE(3, return (launch_one<16, 24, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(4, return (launch_one<32, 16, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q)))
        default: return launch_one<32, 32, TB, FIN, FAST, RB>(p, sm, smem, st, grid, q);
    }
}

// variant = (traceback?, final-gap multiplier?, fast PRMT/DPX path?, rust-bio semantics?)
cudaError_t launch_any(int cfg, bool tb, bool fin, bool fast, bool rb, const KParams& p, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    if (rb) return launch_cfg<true, false, true, true>(cfg, p, sm, smem, st, grid, q);
    if (fast) return tb ? launch_cfg<true, false, true>(cfg, p, sm, smem, st, grid, q) : launch_cfg<false, false, true>(cfg, p, sm, smem, st, grid, q);
    if (tb) return fin ? launch_cfg<true, true, false>(cfg, p, sm, smem, st, grid, q) : launch_cfg<true, false, false>(cfg, p, sm, smem, st, grid, q);
    return fin ? launch_cfg<false, true, false>(cfg, p, sm, smem, st, grid, q) : launch_cfg<false, false, false>(cfg, p, sm, smem, st, grid, q);
}

template <int G, int C, bool TB, bool RB = false, bool MADD = false>
cudaError_t launch_pack_one(const KParams& p, const PackParams& pp, int sm_count, size_t smem, cudaStream_t st, int* grid_out, bool query_only) {
    auto kern = pack_kernel<G, C, TB, RB, MADD>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    int grid = nb * sm_count;
    if (*grid_out > 0) grid = std::min(grid, *grid_out);
    *grid_out = grid;
    if (query_only) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(p, pp);
    return cudaGetLastError();
}

template <bool TB, bool RB = false, bool MADD = false>
cudaError_t launch_pack(int cfg, const KParams& p, const PackParams& pp, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    switch (cfg) {
        CLQ_FULL_CASE(0, return (launch_pack_one<8, 16, TB, RB, MADD>(p, pp, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(1, return (launch_pack_one<8, 24, TB, RB, MADD>(p, pp, sm, smem, st, grid, q)))
        case 2: return launch_pack_one<8, 40, TB, RB, MADD>(p, pp, sm, smem, st, grid, q);
        CLQ_FULL_CASE(3, return (launch_pack_one<16, 24, TB, RB, false>(p, pp, sm, smem, st, grid, q)))  // MADD: G <= 8 only (clq_launch)
        CLQ_FULL_CASE(4, return (launch_pack_one<32, 16, TB, RB, false>(p, pp, sm, smem, st, grid, q)))
        default: return launch_pack_one<32, 32, TB, RB, false>(p, pp, sm, smem, st, grid, q);
    }
}

template <int G, int C>
cudaError_t launch_adapt_one(const KParams& p, const AdaptParams& ap, int sm_count, size_t smem, cudaStream_t st, int* grid_out, bool query_only) {
    auto kern = pack_adapt_kernel<G, C>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    int grid = nb * sm_count;
    if (*grid_out > 0) grid = std::min(grid, *grid_out);
    *grid_out = grid;
    if (query_only) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(p, ap);
    return cudaGetLastError();
}

// adaptive-bias PACK: (8,40) with column stripes by default, the long-read geometries by force_cfg
cudaError_t launch_adapt(int cfg, const KParams& p, const AdaptParams& ap, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    switch (cfg) {
        case 2: return launch_adapt_one<8, 40>(p, ap, sm, smem, st, grid, q);
        CLQ_FULL_CASE(3, return (launch_adapt_one<16, 24>(p, ap, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(4, return (launch_adapt_one<32, 16>(p, ap, sm, smem, st, grid, q)))
        default: return launch_adapt_one<32, 32>(p, ap, sm, smem, st, grid, q);
    }
}

// two-piece affine ("convex") geometries: C must be a multiple of 16 (8 direction bits per cell, 128-bit row stores)
const Cfg kCvxCfgs[] = {{8, 16}, {16, 16}, {16, 32}, {32, 32}, {32, 16}};  // the last one only by force_cfg (experiments)
constexpr int kNumCvxCfgs = sizeof(kCvxCfgs) / sizeof(kCvxCfgs[0]);

template <int G, int C, bool TB>
cudaError_t launch_cvx_one(const KParams& p, const ConvexParams& cp, int sm_count, size_t smem, cudaStream_t st, int* grid_out, bool query_only) {
    auto kern = convex_kernel<G, C, TB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    int grid = nb * sm_count;
    if (*grid_out > 0) grid = std::min(grid, *grid_out);
    *grid_out = grid;
    if (query_only) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(p, cp);
    return cudaGetLastError();
}

template <bool TB>
cudaError_t launch_cvx(int cfg, const KParams& p, const ConvexParams& cp, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    switch (cfg) {
        case 0: return launch_cvx_one<8, 16, TB>(p, cp, sm, smem, st, grid, q);
        CLQ_FULL_CASE(1, return (launch_cvx_one<16, 16, TB>(p, cp, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(2, return (launch_cvx_one<16, 32, TB>(p, cp, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(4, return (launch_cvx_one<32, 16, TB>(p, cp, sm, smem, st, grid, q)))
        default: return launch_cvx_one<32, 32, TB>(p, cp, sm, smem, st, grid, q);
    }
}

template <int G, int C, bool TB>
cudaError_t launch_cvx_pack_one(const KParams& p, const ConvexParams& cp, const PackParams& pp, int sm_count, size_t smem, cudaStream_t st, int* grid_out, bool query_only) {
    auto kern = convex_pack_kernel<G, C, TB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    int grid = nb * sm_count;
    if (*grid_out > 0) grid = std::min(grid, *grid_out);
    *grid_out = grid;
    if (query_only) return cudaSuccess;
    kern<<<grid, kThreads, smem, st>>>(p, cp, pp);
    return cudaGetLastError();
}

template <bool TB>
cudaError_t launch_cvx_pack(int cfg, const KParams& p, const ConvexParams& cp, const PackParams& pp, int sm, size_t smem, cudaStream_t st, int* grid, bool q) {
    switch (cfg) {
        case 0: return launch_cvx_pack_one<8, 16, TB>(p, cp, pp, sm, smem, st, grid, q);
        CLQ_FULL_CASE(1, return (launch_cvx_pack_one<16, 16, TB>(p, cp, pp, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(2, return (launch_cvx_pack_one<16, 32, TB>(p, cp, pp, sm, smem, st, grid, q)))
        CLQ_FULL_CASE(4, return (launch_cvx_pack_one<32, 16, TB>(p, cp, pp, sm, smem, st, grid, q)))
        default: return launch_cvx_pack_one<32, 32, TB>(p, cp, pp, sm, smem, st, grid, q);
    }
}

template <int G, int C>
cudaError_t launch_cvx_walk_one(const KParams& p, uint32_t cnt, cudaStream_t st) {
    convex_walk_kernel<G, C><<<(cnt + 127) / 128, 128, 0, st>>>(p.tb_rec, cnt, p.bits, p.bits_stride, p.bits_off, p.task_base, p.cig_scratch, p.cig_stride, p.cigar_pool,
                                                                p.cigar_cap, p.cigar_cursor, p.results, p.ref_bytes, p.ref_off, p.read_bytes, p.read_off);
    return cudaGetLastError();
}

cudaError_t launch_cvx_walk(int cfg, const KParams& p, uint32_t cnt, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_cvx_walk_one<8, 16>(p, cnt, st);
        CLQ_FULL_CASE(1, return (launch_cvx_walk_one<16, 16>(p, cnt, st)))
        CLQ_FULL_CASE(2, return (launch_cvx_walk_one<16, 32>(p, cnt, st)))
        CLQ_FULL_CASE(4, return (launch_cvx_walk_one<32, 16>(p, cnt, st)))
        default: return launch_cvx_walk_one<32, 32>(p, cnt, st);
    }
}

template <int G, int C>
cudaError_t launch_walk_one(const KParams& p, uint32_t cnt, cudaStream_t st) {
    // one thread per pair.  (The warp-per-pair mode of walk_kernel -- a window of 32 diagonal cells fetched at once -- was measured
    // on the long-read geometries: C5 unchanged, C3 15 % slower: with 10^4..10^5 pairs per sub-batch the 32x instruction count
    // costs more than the latency it hides.  It stays in the kernel for sub-batches of very few, very long pairs.)
    const bool warp_mode = G >= 16 && cnt <= 1024;
    if (warp_mode) {
        walk_kernel<G, C, (G >= 16)><<<(uint32_t)(((uint64_t)cnt * 32 + 127) / 128), 128, 0, st>>>(p.tb_rec, cnt, p.bits, p.bits_stride, p.bits_off, p.task_base, p.cig_scratch, p.cig_stride, p.cigar_pool,
                                                         p.cigar_cap, p.cigar_cursor, p.results, p.ref_bytes, p.ref_off, p.read_bytes, p.read_off,
                                                         p.tag_slot, p.tags, p.tag_stride, p.rustbio, p.band_mode, p.band_k);
        return cudaGetLastError();
    }
    walk_kernel<G, C, false><<<(cnt + 127) / 128, 128, 0, st>>>(p.tb_rec, cnt, p.bits, p.bits_stride, p.bits_off, p.task_base, p.cig_scratch, p.cig_stride, p.cigar_pool,
                                                         p.cigar_cap, p.cigar_cursor, p.results, p.ref_bytes, p.ref_off, p.read_bytes, p.read_off,
                                                         p.tag_slot, p.tags, p.tag_stride, p.rustbio, p.band_mode, p.band_k);
    return cudaGetLastError();
}

cudaError_t launch_walk(int cfg, const KParams& p, uint32_t cnt, cudaStream_t st) {
    switch (cfg) {
        CLQ_FULL_CASE(0, return (launch_walk_one<8, 16>(p, cnt, st)))
        CLQ_FULL_CASE(1, return (launch_walk_one<8, 24>(p, cnt, st)))
        case 2: return launch_walk_one<8, 40>(p, cnt, st);
        CLQ_FULL_CASE(3, return (launch_walk_one<16, 24>(p, cnt, st)))
        CLQ_FULL_CASE(4, return (launch_walk_one<32, 16>(p, cnt, st)))
        default: return launch_walk_one<32, 32>(p, cnt, st);
    }
}

int pick_cfg(const clq_ctx* c, uint32_t max_len) {
    if (c->force_cfg >= 0 && c->force_cfg < kNumCfgs) return c->force_cfg;
    for (int i = 0; i < kNumCfgs; i++) {
        if (CLQ_LEAN && i != 2 && i != kNumCfgs - 1) continue;
        if ((uint32_t)(kCfgs[i].G * kCfgs[i].C) >= max_len) return i;
    }
    return kNumCfgs - 1;
}

bool is_integral(double v) { return v == std::floor(v) && std::fabs(v) < 2.0e8; }

Slot* get_slot(clq_ctx* c, int32_t slot) {
    if (!c || slot < 0 || slot >= (int32_t)c->slots.size()) return nullptr;
    return &c->slots[slot];
}

}  // namespace

extern "C" {

int32_t clq_version(void) { return CLQ_VERSION; }

const char* clq_strerror(int32_t code) {
    switch (code) {
        case CLQ_OK: return "ok";
        case CLQ_READ_TOO_LONG: return "read too long";
        case CLQ_SCORING_NOT_REPRESENTABLE: return "scoring not representable as scaled integers";
        case CLQ_TRACEBACK_DIVERGED: return "traceback diverged (reference would not terminate)";
        case CLQ_CIGAR_POOL_FULL: return "cigar pool full";
        case CLQ_NO_CANDIDATE: return "no candidate reference";
        case CLQ_E_INVALID: return "invalid argument";
        case CLQ_E_CUDA: return "CUDA error";
        case CLQ_E_NOMEM: return "out of device memory";
        case CLQ_E_LIMIT: return "batch exceeds context limits";
        case CLQ_E_STATE: return "call out of order for this slot";
        case CLQ_E_UNSUPPORTED: return "unsupported mode";
        default: return "unknown";
    }
}

/* debug builds (-DCLQ_PACK_CANARY=1): tasks whose stored s16 values were tracked / tasks with a value outside [64, 32767] that
 * no retry pass covered; {0, 0} in production builds.  Not part of include/clq.h. */
int32_t clq_debug_canary(unsigned long long* checked, unsigned long long* violations) {
    unsigned long long h[2] = {0, 0};
    if (cudaMemcpyFromSymbol(h, g_pack_canary, sizeof(h)) != cudaSuccess) return CLQ_E_CUDA;
    if (checked) *checked = h[0];
    if (violations) *violations = h[1];
    return CLQ_PACK_CANARY ? CLQ_OK : CLQ_E_UNSUPPORTED;
}

int32_t clq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int32_t clq_affine_from_f64(double match_score, double mismatch_score, double special_character_score, double gap_open,
                            double gap_extend, double final_gap_multiplier, clq_affine_t* out) {
    if (!out) return CLQ_E_INVALID;
    if (!(gap_open < 0.0)) return CLQ_SCORING_NOT_REPRESENTABLE;
    for (int scale = 1; scale <= 64; scale *= 2) {
        const double s = (double)scale;
        const double v[10] = {match_score * s,
                              mismatch_score * s,
                              special_character_score * s,
                              (gap_open + gap_extend) * s,
                              gap_extend * s,
                              (gap_open + gap_extend * final_gap_multiplier) * s,
                              (gap_extend * final_gap_multiplier) * s,
                              (gap_open * final_gap_multiplier) * s,
                              (gap_extend * final_gap_multiplier) * s,
                              -100000.0 * s};
        bool ok = true;
        for (double x : v) ok = ok && is_integral(x);
        if (!ok) continue;
        out->scale = scale;
        out->match = (int32_t)v[0]; out->mismatch = (int32_t)v[1]; out->special = (int32_t)v[2];
        out->oe_in = (int32_t)v[3]; out->e_in = (int32_t)v[4];
        out->oe_fin = (int32_t)v[5]; out->e_fin = (int32_t)v[6];
        out->b0 = (int32_t)v[7]; out->b1 = (int32_t)v[8];
        out->max_neg = (int32_t)v[9];
        return CLQ_OK;
    }
    return CLQ_SCORING_NOT_REPRESENTABLE;
}

int32_t clq_rustbio_scoring(int32_t match_score, int32_t mismatch_score, int32_t gap_open, int32_t gap_extend, clq_affine_t* out) {
    if (!out) return CLQ_E_INVALID;
    if (!(gap_open < 0) || gap_extend > 0) return CLQ_SCORING_NOT_REPRESENTABLE;
    out->scale = 1;
    out->match = match_score; out->mismatch = mismatch_score; out->special = match_score;
    out->oe_in = out->oe_fin = gap_open + gap_extend;
    out->e_in = out->e_fin = gap_extend;
    out->b0 = gap_open; out->b1 = gap_extend;
    out->max_neg = -858993459;  // rust-bio's MIN_SCORE
    return CLQ_OK;
}

int32_t clq_host_alloc(size_t bytes, void** out) {
    if (!out) return CLQ_E_INVALID;
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 64), cudaHostAllocDefault);
    return e == cudaSuccess ? CLQ_OK : CLQ_E_CUDA;
}

int32_t clq_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? CLQ_OK : CLQ_E_CUDA; }

int32_t clq_ctx_create(int32_t device, const clq_limits_t* limits, clq_ctx** out) {
    if (!out || !limits) return CLQ_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return CLQ_E_CUDA;
    clq_ctx* c = new clq_ctx();
    c->device = device;
    c->lim = *limits;
    if (c->lim.n_slots < 1) c->lim.n_slots = 1;
    if (c->lim.n_slots > 4) c->lim.n_slots = 4;
    if (c->lim.cigar_pool_ops > 0xffffffffull) c->lim.cigar_pool_ops = 0xffffffffull;  // clq_result_t.cigar_off is 32 bits
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return CLQ_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return CLQ_E_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    c->slots.resize(c->lim.n_slots);
    for (auto& s : c->slots) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        for (auto& e : s.ev)
            if (cudaEventCreate(&e) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (cudaHostAlloc((void**)&s.h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess) { clq_ctx_destroy(c); return CLQ_E_CUDA; }
        if (ensure(c, s.counters, 16 * sizeof(unsigned long long)) != CLQ_OK) { clq_ctx_destroy(c); return CLQ_E_NOMEM; }
    }
    *out = c;
    return CLQ_OK;
}

void clq_ctx_destroy(clq_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& s : c->slots) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.stream2) { cudaStreamSynchronize(s.stream2); cudaStreamDestroy(s.stream2); }
        if (s.fork) cudaEventDestroy(s.fork);
        if (s.join) cudaEventDestroy(s.join);
        for (auto& e : s.ev) if (e) cudaEventDestroy(e);
        if (s.done) cudaEventDestroy(s.done);
        for (DevBuf* b : {&s.read_bytes, &s.read_off, &s.fixed_ref, &s.order, &s.results, &s.scores, &s.cand_mask, &s.single_ref,
                          &s.ref_of_read, &s.votes, &s.cigar_pool, &s.bits, &s.cig_scratch, &s.col_scratch, &s.tb_rec, &s.bits_off, &s.tags, &s.ref_groups, &s.retry_list, &s.packed2, &s.exc_pos, &s.exc_byte, &s.order2, &s.counters})
            release(*b);
        if (s.h_counters) cudaFreeHost(s.h_counters);
    }
    release(c->ref_bytes); release(c->ref_off); release(c->kmer_keys); release(c->kmer_owner); release(c->kmer_keys64); release(c->sh_bits); release(c->sh_cig); release(c->sh_tbrec); release(c->cls_lut); release(c->cls_lut_rb); release(c->tag_slot);
    delete c;
}

const char* clq_ctx_last_error(const clq_ctx* c) { return c ? c->err.c_str() : "null context"; }

int32_t clq_set_option(clq_ctx* c, const char* key, int64_t value) {
    if (!c || !key) return CLQ_E_INVALID;
    if (!strcmp(key, "force_cfg")) { c->force_cfg = (int)value; return CLQ_OK; }
    if (!strcmp(key, "force_generic")) { c->force_generic = (int)value; return CLQ_OK; }
    if (!strcmp(key, "debug_flags")) { c->debug_flags = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_pack")) { c->no_pack = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_madd")) { c->no_madd = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_adapt")) { c->no_adapt = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_overlap")) { c->no_overlap = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_long8")) { c->no_long8 = (int)value; return CLQ_OK; }
    if (!strcmp(key, "adapt_guard")) { c->adapt_guard = (int)value; return CLQ_OK; }
    if (!strcmp(key, "no_group")) { c->no_group = (int)value; return CLQ_OK; }
    if (!strcmp(key, "serialize_slots")) { c->serialize = (int)value; return CLQ_OK; }
    if (!strcmp(key, "max_scratch_bytes")) {
        if (value < (1 << 20)) return fail(c, CLQ_E_INVALID, "max_scratch_bytes must be at least 1 MiB");
        c->max_scratch_bytes = value;
        return CLQ_OK;
    }
    return fail(c, CLQ_E_INVALID, std::string("unknown option ") + key);
}

int32_t clq_refs_set(clq_ctx* c, uint32_t n_refs, const uint8_t* bytes, const uint64_t* off) {
    if (!c || (n_refs && (!bytes || !off))) return CLQ_E_INVALID;
    if (n_refs > c->lim.max_refs) return fail(c, CLQ_E_LIMIT, "too many references");
    const uint64_t total = n_refs ? off[n_refs] : 0;
    if (total > c->lim.max_ref_bytes) return fail(c, CLQ_E_LIMIT, "reference bytes exceed limit");
    CU(c, cudaSetDevice(c->device));
    if (n_refs) {
        c->h_ref_bytes.assign(bytes, bytes + total);
        c->h_ref_off.assign(off, off + n_refs + 1);
    } else {  // empty set (bytes / off may be NULL)
        c->h_ref_bytes.clear();
        c->h_ref_off.assign(1, 0);
    }
    c->n_refs = n_refs;
    c->max_ref_len = 0;
    for (uint32_t r = 0; r < n_refs; r++) {
        if (off[r + 1] < off[r]) return fail(c, CLQ_E_INVALID, "reference offsets must be non-decreasing");
        c->max_ref_len = std::max<uint32_t>(c->max_ref_len, (uint32_t)(off[r + 1] - off[r]));
    }
    int32_t rc;
    if ((rc = ensure(c, c->ref_bytes, total + 16)) != CLQ_OK) return rc;
    if ((rc = ensure(c, c->ref_off, (n_refs + 1) * sizeof(uint64_t))) != CLQ_OK) return rc;
    for (auto& s : c->slots) CU(c, cudaStreamSynchronize(s.stream));
    if (total) CU(c, cudaMemcpy(c->ref_bytes.p, bytes, total, cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->ref_off.p, c->h_ref_off.data(), (n_refs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    c->n_keys = 0;
    c->kmer_k = 0;
    // byte classes for the FAST kernels: every special byte (N or < 58, alignment/scoring_functions.rs:100-102) is class 0,
    // each distinct non-special reference byte gets its own class 2..7, every other byte can only mismatch: class 1.
    int next = 2;
    c->fast_ok = true;
    for (int b = 0; b < 256; b++) c->cls[b] = (b == 'N' || b < 58) ? 0 : 1;
    for (uint64_t i = 0; i < total; i++) {
        const uint8_t b = bytes[i];
        if (c->cls[b] == 1) {
            if (next > 7) { c->fast_ok = false; break; }
            c->cls[b] = (uint8_t)next++;
        }
    }
    for (int b = 0; b < 256; b++) c->cls[b] = (uint8_t)(c->cls[b] | (c->cls[b] << 3));  // row = class in clique's own scoring
    if ((rc = ensure(c, c->cls_lut, 256)) != CLQ_OK) return rc;
    CU(c, cudaMemcpy(c->cls_lut.p, c->cls, 256, cudaMemcpyHostToDevice));
    // rust-bio mode (alignment_functions.rs:55: a == b || a == b'N', a = read byte): column 0 = read 'N' (wildcard), column 1 =
    // a byte no reference holds, columns 2..7 = the first six distinct reference bytes (letters before tag symbols); further
    // reference bytes get rows 8..15 of their own but no column: a read that holds one is reported CLQ_SCORING_NOT_REPRESENTABLE.
    {
        for (int b = 0; b < 256; b++) c->cls_rb[b] = 1 | (1 << 3);
        c->cls_rb['N'] = 0;
        bool seen[256] = {};
        std::vector<uint8_t> distinct;
        for (int pass = 0; pass < 2; pass++)
            for (uint64_t i = 0; i < total; i++) {
                const uint8_t b = bytes[i];
                if (b == 'N' || seen[b] || ((b >= 'A') != (pass == 0))) continue;
                seen[b] = true;
                distinct.push_back(b);
            }
        c->rb_ok = distinct.size() <= 6 + 8;
        for (size_t k = 0; k < distinct.size() && c->rb_ok; k++)
            c->cls_rb[distinct[k]] = k < 6 ? (uint8_t)((2 + k) | ((2 + k) << 3)) : (uint8_t)(1 | ((8 + (k - 6)) << 3) | 0x80);
        if ((rc = ensure(c, c->cls_lut_rb, 256)) != CLQ_OK) return rc;
        CU(c, cudaMemcpy(c->cls_lut_rb.p, c->cls_rb, 256, cudaMemcpyHostToDevice));
    }
    // tag columns: reference bytes '0'..'9' (SPECIAL_CHARACTERS, extractor.rs:19-34); slot = rank within its reference
    std::vector<uint16_t> slot(total + 1, 0xffffu);
    uint32_t most = 0;
    for (uint32_t r = 0; r < n_refs; r++) {
        uint32_t k = 0;
        for (uint64_t i = off[r]; i < off[r + 1]; i++)
            if (bytes[i] >= '0' && bytes[i] <= '9' && k < 0xffffu) slot[i] = (uint16_t)k++;
        most = std::max(most, k);
    }
    c->tag_stride = (most + 3) / 4 * 4;
    if ((rc = ensure(c, c->tag_slot, (total + 1) * sizeof(uint16_t))) != CLQ_OK) return rc;
    CU(c, cudaMemcpy(c->tag_slot.p, slot.data(), (total + 1) * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return CLQ_OK;
}

// ReferenceManager::unique_kmers (reference/fasta_reference.rs:159-202): upper-case, windows(k).step_by(skip),
// consecutive run-length dedup, k-mer unique <=> its run counts over all references sum to exactly 1.
int32_t clq_kmer_index_set(clq_ctx* c, uint32_t k, uint32_t skip) {
    if (!c || k == 0 || skip == 0) return CLQ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    struct Ent { const uint8_t* p; uint32_t ref, count; };
    std::vector<uint8_t> up(c->h_ref_bytes.size() + 1);
    for (size_t i = 0; i < c->h_ref_bytes.size(); i++) {
        uint8_t ch = c->h_ref_bytes[i];
        up[i] = (ch >= 'a' && ch <= 'z') ? ch - 32 : ch;
    }
    std::vector<Ent> ents;
    for (uint32_t r = 0; r < c->n_refs; r++) {
        const uint8_t* s = up.data() + c->h_ref_off[r];
        const size_t len = (size_t)(c->h_ref_off[r + 1] - c->h_ref_off[r]);
        const uint8_t* run = nullptr;
        uint32_t cnt = 0;
        for (size_t pos = 0; pos + k <= len; pos += skip) {
            const uint8_t* w = s + pos;
            if (run && memcmp(run, w, k) == 0) cnt++;
            else {
                if (run) ents.push_back({run, r, cnt});
                run = w;
                cnt = 1;
            }
        }
        if (run) ents.push_back({run, r, cnt});
    }
    std::sort(ents.begin(), ents.end(), [k](const Ent& a, const Ent& b) {
        int cmp = memcmp(a.p, b.p, k);
        return cmp ? cmp < 0 : a.ref < b.ref;
    });
    std::vector<uint8_t> keys;
    std::vector<uint32_t> owner;
    for (size_t i = 0; i < ents.size();) {
        size_t j = i;
        uint64_t tot = 0;
        while (j < ents.size() && memcmp(ents[j].p, ents[i].p, k) == 0) { tot += ents[j].count; j++; }
        if (tot == 1) {
            keys.insert(keys.end(), ents[i].p, ents[i].p + k);
            owner.push_back(ents[i].ref);
        }
        i = j;
    }
    int32_t rc;
    for (auto& s : c->slots) CU(c, cudaStreamSynchronize(s.stream));  // the slot streams are non-blocking: an in-flight CLQ_SEARCH_QUICK launch reads the old table
    if ((rc = ensure(c, c->kmer_keys, keys.size() + 16)) != CLQ_OK) return rc;
    if ((rc = ensure(c, c->kmer_owner, (owner.size() + 1) * sizeof(uint32_t))) != CLQ_OK) return rc;
    if (!keys.empty()) {
        CU(c, cudaMemcpy(c->kmer_keys.p, keys.data(), keys.size(), cudaMemcpyHostToDevice));
        CU(c, cudaMemcpy(c->kmer_owner.p, owner.data(), owner.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    if (k <= 8) {  // packed big-endian copy for kmer_vote_packed_kernel (same order: memcmp order == numeric order)
        std::vector<uint64_t> k64(owner.size() + 1, 0);
        for (size_t i = 0; i < owner.size(); i++) {
            uint64_t v = 0;
            for (uint32_t b = 0; b < k; b++) v = (v << 8) | keys[i * k + b];
            k64[i] = v;
        }
        if ((rc = ensure(c, c->kmer_keys64, k64.size() * sizeof(uint64_t))) != CLQ_OK) return rc;
        CU(c, cudaMemcpy(c->kmer_keys64.p, k64.data(), k64.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    c->kmer_k = k;
    c->kmer_skip = skip;
    c->n_keys = (uint32_t)owner.size();
    return CLQ_OK;
}

}  // extern "C"

namespace {

// the 2-bit form of a batch (clq_upload_packed2); null = raw ASCII bytes
struct Packed2Src {
    const uint32_t* words;
    const uint64_t* exc_pos;
    const uint8_t* exc_byte;
    uint64_t n_exc;
};

int32_t upload_common(clq_ctx* c, int32_t slot, uint32_t n_reads, const uint8_t* read_bytes, const uint64_t* read_off,
                      const int32_t* fixed_ref, const Packed2Src* pk) {
    Slot* s = get_slot(c, slot);
    if (!s || (n_reads && (!read_off))) return CLQ_E_INVALID;
    if (n_reads > c->lim.max_reads) return fail(c, CLQ_E_LIMIT, "too many reads for one batch");
    const uint64_t total = n_reads ? read_off[n_reads] - read_off[0] : 0;
    if (total > c->lim.max_read_bytes) return fail(c, CLQ_E_LIMIT, "read bytes exceed limit");
    if (total && !(pk ? (const void*)pk->words : (const void*)read_bytes)) return CLQ_E_INVALID;
    if (pk && pk->n_exc && (!pk->exc_pos || !pk->exc_byte)) return CLQ_E_INVALID;
    if (n_reads && read_off[0] != 0) return fail(c, CLQ_E_INVALID, "read_off[0] must be 0");
    CU(c, cudaSetDevice(c->device));
    int32_t rc;
    if ((rc = ensure(c, s->read_bytes, total + 16)) != CLQ_OK) return rc;
    if ((rc = ensure(c, s->read_off, ((size_t)n_reads + 1) * sizeof(uint64_t))) != CLQ_OK) return rc;
    if ((rc = ensure(c, s->results, ((size_t)n_reads + 1) * sizeof(clq_result_t))) != CLQ_OK) return rc;
    if ((rc = ensure(c, s->ref_of_read, ((size_t)n_reads + 1) * sizeof(int32_t))) != CLQ_OK) return rc;
    // host pass over the offsets: length range + (when lengths vary) a longest-first processing order
    uint32_t mx = 0, mn = 0xffffffffu;
    for (uint32_t i = 0; i < n_reads; i++) {
        if (read_off[i + 1] < read_off[i]) return fail(c, CLQ_E_INVALID, "read offsets must be non-decreasing");
        const uint64_t l = read_off[i + 1] - read_off[i];
        if (l >= c->lim.max_read_len) continue;  // dropped on the device with CLQ_READ_TOO_LONG
        mx = std::max<uint32_t>(mx, (uint32_t)l);
        mn = std::min<uint32_t>(mn, (uint32_t)l);
    }
    if (mn == 0xffffffffu) mn = 0;
    s->max_len = mx;
    s->min_len = mn;
    s->have_order = false;
    s->order_by_ref = false;
    s->n_pos = n_reads;
    static thread_local std::vector<uint32_t> order, bucket, padded;
    if (n_reads && mx > mn + mn / 8 + 16) {
        const uint32_t nb = mx / 16 + 2;
        // multi-reference batches with a known reference per read: group by reference first (longest reference first), then
        // longest read first inside a group, so that the two reads of a PACK task share their reference
        const bool by_ref = fixed_ref != nullptr && c->n_refs > 1;
        std::vector<uint32_t> ref_rank;
        uint32_t n_groups = 1;
        if (by_ref) {
            std::vector<uint32_t> idx(c->n_refs);
            for (uint32_t r = 0; r < c->n_refs; r++) idx[r] = r;
            std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
                return c->h_ref_off[a + 1] - c->h_ref_off[a] > c->h_ref_off[b + 1] - c->h_ref_off[b];
            });
            ref_rank.assign(c->n_refs + 1, 0);
            for (uint32_t k = 0; k < c->n_refs; k++) ref_rank[idx[k]] = k;
            ref_rank[c->n_refs] = c->n_refs;  // reads without a usable reference come last
            n_groups = c->n_refs + 1;
        }
        auto group_of = [&](uint32_t i) -> uint32_t {
            if (!by_ref) return 0;
            const int32_t r = fixed_ref[i];
            return ref_rank[(r >= 0 && (uint32_t)r < c->n_refs) ? (uint32_t)r : c->n_refs];
        };
        bucket.assign((size_t)nb * n_groups + 1, 0);
        auto key = [&](uint32_t i) -> uint32_t {
            const uint64_t l = read_off[i + 1] - read_off[i];
            const uint32_t b = l >= c->lim.max_read_len ? 0 : (uint32_t)l / 16 + 1;
            return group_of(i) * nb + (nb - 1 - std::min(b, nb - 1));  // longest first
        };
        for (uint32_t i = 0; i < n_reads; i++) bucket[key(i) + 1]++;
        for (size_t b = 0; b < (size_t)nb * n_groups; b++) bucket[b + 1] += bucket[b];
        order.resize(n_reads);
        for (uint32_t i = 0; i < n_reads; i++) order[bucket[key(i)]++] = i;
        const std::vector<uint32_t>* use = &order;
        if (by_ref) {  // pad every group to an even number of positions
            padded.clear();
            padded.reserve((size_t)n_reads + n_groups);
            uint32_t prev_g = 0xffffffffu;
            for (uint32_t i = 0; i < n_reads; i++) {
                const uint32_t g = group_of(order[i]);
                if (g != prev_g && (padded.size() & 1)) padded.push_back(0xffffffffu);
                prev_g = g;
                padded.push_back(order[i]);
            }
            use = &padded;
            s->order_by_ref = true;
        }
        const size_t npos = use->size();
        s->n_pos = (uint32_t)npos;
        if ((rc = ensure(c, s->order, (npos + 2) * sizeof(uint32_t))) != CLQ_OK) return rc;
        CU(c, cudaMemcpyAsync(s->order.p, use->data(), npos * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
        CU(c, cudaStreamSynchronize(s->stream));  // the host vectors are reused
        s->have_order = true;
        s->h_order.assign(use->begin(), use->end());
        s->h_len.resize(npos);
        s->h_ref.clear();
        if (fixed_ref) s->h_ref.resize(npos);
        for (size_t i = 0; i < npos; i++) {
            const uint32_t r = (*use)[i];
            s->h_len[i] = r == 0xffffffffu ? 0u : (uint32_t)(read_off[r + 1] - read_off[r]);
            if (fixed_ref) s->h_ref[i] = r == 0xffffffffu ? -1 : fixed_ref[r];
        }
    }
    uint64_t h2d = 0;
    if (total && !pk) {
        CU(c, cudaMemcpyAsync(s->read_bytes.p, read_bytes + read_off[0], total, cudaMemcpyHostToDevice, s->stream));
        h2d += total;
    }
    if (total && pk) {
        // 2-bit stream + exception list -> the same ASCII read buffer (clq_reads2bit.cuh); everything behind it is shared
        const uint64_t n_words = (total + 15) / 16;
        if ((rc = ensure(c, s->packed2, ((n_words + 3) / 4) * 16)) != CLQ_OK) return rc;
        CU(c, cudaMemcpyAsync(s->packed2.p, pk->words, n_words * 4, cudaMemcpyHostToDevice, s->stream));
        h2d += n_words * 4;
        const uint64_t n_vec = (n_words + 3) / 4;
        const unsigned grid = (unsigned)std::min<uint64_t>((n_vec + 255) / 256, (uint64_t)c->sm_count * 8);
        clq::unpack2_kernel<<<grid, 256, 0, s->stream>>>((const uint4*)s->packed2.p, (uint4*)s->read_bytes.p, n_words);
        if (pk->n_exc) {
            if ((rc = ensure(c, s->exc_pos, pk->n_exc * sizeof(uint64_t))) != CLQ_OK) return rc;
            if ((rc = ensure(c, s->exc_byte, pk->n_exc)) != CLQ_OK) return rc;
            CU(c, cudaMemcpyAsync(s->exc_pos.p, pk->exc_pos, pk->n_exc * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
            CU(c, cudaMemcpyAsync(s->exc_byte.p, pk->exc_byte, pk->n_exc, cudaMemcpyHostToDevice, s->stream));
            h2d += pk->n_exc * 9;
            const unsigned g2 = (unsigned)std::min<uint64_t>((pk->n_exc + 255) / 256, (uint64_t)c->sm_count * 8);
            clq::patch2_kernel<<<g2, 256, 0, s->stream>>>((uint8_t*)s->read_bytes.p, (const uint64_t*)s->exc_pos.p, (const uint8_t*)s->exc_byte.p, pk->n_exc, total);
        }
        CU(c, cudaGetLastError());
    }
    if (n_reads) {
        CU(c, cudaMemcpyAsync(s->read_off.p, read_off, ((size_t)n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
        h2d += ((size_t)n_reads + 1) * sizeof(uint64_t);
    }
    s->have_fixed = false;
    if (fixed_ref && n_reads) {
        if ((rc = ensure(c, s->fixed_ref, (size_t)n_reads * sizeof(int32_t))) != CLQ_OK) return rc;
        CU(c, cudaMemcpyAsync(s->fixed_ref.p, fixed_ref, (size_t)n_reads * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
        h2d += (size_t)n_reads * sizeof(int32_t);
        s->have_fixed = true;
    }
    if (s->have_order) h2d += (size_t)s->n_pos * sizeof(uint32_t);
    s->n_reads = n_reads;
    s->n_read_bytes = total;
    s->stats = clq_stats_t{};
    s->launched = false;
    s->stats.h2d_bytes = h2d;
    s->state = 1;
    return CLQ_OK;
}

}  // namespace

extern "C" {

int32_t clq_upload(clq_ctx* c, int32_t slot, uint32_t n_reads, const uint8_t* read_bytes, const uint64_t* read_off,
                   const int32_t* fixed_ref) {
    return upload_common(c, slot, n_reads, read_bytes, read_off, fixed_ref, nullptr);
}

int32_t clq_upload_packed2(clq_ctx* c, int32_t slot, uint32_t n_reads, const uint32_t* packed, const uint64_t* read_off,
                           const uint64_t* exc_pos, const uint8_t* exc_byte, uint64_t n_exc, const int32_t* fixed_ref) {
    const Packed2Src pk{packed, exc_pos, exc_byte, n_exc};
    return upload_common(c, slot, n_reads, nullptr, read_off, fixed_ref, &pk);
}

// Host side of the 2-bit form (no CUDA call): A C G T -> 0 1 2 3, base i in bits 2 (i % 16) of word i / 16; every other byte
// is stored as code 0 and listed (position, byte) in ascending position order.
int32_t clq_pack2(const uint8_t* bytes, uint64_t n_bytes, uint32_t* packed, uint64_t* exc_pos, uint8_t* exc_byte, uint64_t exc_cap,
                  uint64_t* n_exc) {
    if ((n_bytes && (!bytes || !packed)) || !n_exc) return CLQ_E_INVALID;
    static const struct Lut {
        uint8_t v[256];
        Lut() {
            for (int i = 0; i < 256; i++) v[i] = 0x80;
            v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3;
        }
    } lut;
    uint64_t ne = 0;
    bool overflow = false;
    const uint64_t n_words = (n_bytes + 15) / 16;
    const uint64_t n_full = n_bytes / 16;
    uint64_t w = 0;
    while (w < n_words) {
        if (w < n_full) w += clq::pack2_plain_run(bytes + 16 * w, (size_t)(n_full - w), packed + w);  // runs of plain ACGT words (clq_pack2_host.cpp)
        if (w >= n_words) break;
        // a word holding other bytes, or the partial last word
        const uint64_t b0 = w * 16;
        const unsigned m = (unsigned)std::min<uint64_t>(16, n_bytes - b0);
        uint32_t word = 0, flags = 0;
        for (unsigned k = 0; k < m; k++) {
            const uint32_t code = lut.v[bytes[b0 + k]];
            word |= (code & 3u) << (2 * k);
            flags |= code;
        }
        if (flags & 0x80u) {
            for (unsigned k = 0; k < m; k++)
                if (lut.v[bytes[b0 + k]] & 0x80u) {
                    if (ne < exc_cap && exc_pos && exc_byte) { exc_pos[ne] = b0 + k; exc_byte[ne] = bytes[b0 + k]; }
                    else overflow = true;
                    ne++;
                }
        }
        packed[w] = word;
        w++;
    }
    *n_exc = ne;  // on CLQ_E_LIMIT: the capacity the list needs
    return overflow ? CLQ_E_LIMIT : CLQ_OK;
}

int32_t clq_launch(clq_ctx* c, int32_t slot, const void* scoring, uint32_t flags, double match_threshold) {
    Slot* s = get_slot(c, slot);
    if (!s || !scoring) return CLQ_E_INVALID;
    if (s->state < 1) return fail(c, CLQ_E_STATE, "clq_launch before clq_upload");
    const bool convex = (flags & CLQ_CONVEX) != 0;
    const uint32_t band = flags & CLQ_BAND_MASK;
    const uint32_t search = flags & CLQ_SEARCH_MASK;
    if (band > CLQ_BAND_K) return CLQ_E_INVALID;
    const uint32_t band_k = band == CLQ_BAND_K ? (flags >> CLQ_BAND_K_SHIFT) : 0u;
    if (search == CLQ_SEARCH_FIXED && !s->have_fixed && s->n_reads) return fail(c, CLQ_E_INVALID, "CLQ_SEARCH_FIXED needs fixed_ref");
    if (search == CLQ_SEARCH_QUICK && c->kmer_k == 0) return fail(c, CLQ_E_STATE, "CLQ_SEARCH_QUICK needs clq_kmer_index_set");
    if (search > CLQ_SEARCH_QUICK) return CLQ_E_INVALID;
    if (search != CLQ_SEARCH_FIXED && s->order_by_ref)
        return fail(c, CLQ_E_STATE, "this batch was uploaded with fixed_ref (its processing order groups the reads by reference): launch it with CLQ_SEARCH_FIXED or upload it without fixed_ref");
    clq_affine_t sc = {};
    ConvexParams cp = {};
    auto fits8 = [](int v) { return v >= -128 && v <= 127; };
    bool fin = false, fast = false;
    const bool rb = (flags & CLQ_RUSTBIO) != 0;
    if (rb) {
        if (convex || search != CLQ_SEARCH_FIXED || (flags & CLQ_SCORE_ONLY))
            return fail(c, CLQ_E_UNSUPPORTED, "CLQ_RUSTBIO is the single-reference branch: CLQ_SEARCH_FIXED with traceback, affine scoring");
        if (!c->rb_ok) return fail(c, CLQ_E_UNSUPPORTED, "CLQ_RUSTBIO: the reference set has more than 14 distinct bytes");
    }
    if (convex) {
        cp.cv = *(const clq_convex_t*)scoring;
        const clq_convex_t& cv = cp.cv;
        if (!(cv.o1 < 0 && cv.o2 < 0 && cv.e1 <= 0 && cv.e2 <= 0)) return fail(c, CLQ_SCORING_NOT_REPRESENTABLE, "convex: gap opens must be negative");
        if (!c->fast_ok || !fits8(cv.match) || !fits8(cv.mismatch) || !fits8(cv.special))
            return fail(c, CLQ_E_UNSUPPORTED, "convex mode needs <= 6 distinct non-special reference bytes and int8 substitution scores");
        if (band != CLQ_BAND_MAXLEN && band != CLQ_BAND_READLEN) return CLQ_E_INVALID;
        sc.scale = 1; sc.match = cv.match; sc.mismatch = cv.mismatch; sc.special = cv.special;  // profile table below
        fast = true;
    } else {
        sc = *(const clq_affine_t*)scoring;
        if (!(sc.oe_in - sc.e_in < 0) || sc.scale < 1) return fail(c, CLQ_SCORING_NOT_REPRESENTABLE, "gap_open must be negative");
        fin = sc.oe_fin != sc.oe_in || sc.e_fin != sc.e_in;
        fast = c->fast_ok && !c->force_generic && !fin && fits8(sc.match) && fits8(sc.mismatch) && fits8(sc.special);
        if (band == CLQ_BAND_K) fast = false;  // explicit bandwidth: the generic kernels check every cell against the row's window
        if (rb) {
            if (fin || !fits8(sc.match) || !fits8(sc.mismatch)) return fail(c, CLQ_E_UNSUPPORTED, "CLQ_RUSTBIO needs int8 substitution scores and no final-gap multiplier");
            fast = true;
        }
    }
    if (band == CLQ_BAND_K && convex) return fail(c, CLQ_E_UNSUPPORTED, "explicit bandwidth with two-piece affine gaps");
    CU(c, cudaSetDevice(c->device));
    const bool score_only = (flags & CLQ_SCORE_ONLY) != 0;
    const bool want_tags = (flags & CLQ_EXTRACT_TAGS) != 0 && c->tag_stride > 0;
    if ((flags & CLQ_EXTRACT_TAGS) && (convex || score_only)) return fail(c, CLQ_E_UNSUPPORTED, "CLQ_EXTRACT_TAGS needs the affine traceback");
    const uint32_t n = s->n_reads;
    int cfg = pick_cfg(c, s->max_len);
    // 15-bit window of the two-piece affine PACK kernel for stripe width Wx (clq_convex_pack.cuh)
    auto cvx_window = [&](int Wx, int32_t* bias_out) -> bool {
        const clq_convex_t& cv = cp.cv;
        const int64_t smin = std::min<int64_t>(std::min(cv.match, cv.mismatch), std::min(cv.special, 0));
        const int64_t smax = std::max<int64_t>(std::max(cv.match, cv.special), 0);
        const int64_t xa = cv.o1 + cv.e1, xb = cv.o2 + cv.e2;
        const int64_t emin = std::min(cv.e1, cv.e2), omin = std::min(cv.o1, cv.o2);
        const int64_t low = 2 * omin + (int64_t)(c->max_ref_len + s->max_len + Wx) * emin + std::min(std::min(smin, emin), std::min(xa, xb)) - 16;
        const int64_t high = smax * std::min<int64_t>(c->max_ref_len, s->max_len) + std::max(-xa, -xb) + 16;
        if (bias_out) *bias_out = (int32_t)(64 - low);
        return high - low + 128 <= 32767;
    };
    if (convex) {
        cfg = 3;  // (32,32), multi-stripe beyond 1024 columns
        for (int i = 0; i < 4; i++)
            if ((uint32_t)(kCvxCfgs[i].G * kCvxCfgs[i].C) >= s->max_len) { cfg = i; break; }
        // the PACK kernel is fastest on the narrow (8,16) geometry whatever the read length (168 registers, 3 CTAs/SM, four
        // pairs per warp: C3 990 GCUPS against 837 on (32,32)); the stripe-boundary column costs less than the occupancy
        if (search == CLQ_SEARCH_FIXED && !c->no_pack && n && cvx_window(128, nullptr)) cfg = 0;
        if (c->force_cfg >= 0 && c->force_cfg < kNumCvxCfgs) cfg = c->force_cfg;
        if (CLQ_LEAN && cfg != 0) cfg = 3;
    } else if (CLQ_LEAN && cfg != 2) cfg = kNumCfgs - 1;
    const uint32_t L1max = c->max_ref_len, L2max = s->max_len;
    // ---- which kernel family, on which geometry ----
    // s16x2 PACK kernels: two reads per lane group.  Needs the FAST preconditions plus a proof that every cell value of
    // this batch fits a 15-bit window: B >= 2*g(0) + (L1+L2)*e (the all-gap corner path), everything else is within a gap
    // open / one substitution of B, and nothing exceeds max(match, special) * min(L1, L2).  Pairs beyond the static window
    // (long reads) take the adaptive-bias kernel + int32 retry pass (clq_pack_adapt.cuh): traceback stage only, pair mode.
    // Pair mode needs one reference for both reads of a task: a single-reference panel, or a fixed assignment that the upload
    // grouped by reference (every group padded to whole pairs).
    const bool pairs_ok = c->n_refs == 1 || (s->order_by_ref && search == CLQ_SEARCH_FIXED);
    struct Plan { bool pack = false, madd = false, adapt = false; PackParams pkp = {}; AdaptParams adp = {}; };
    auto plan_for = [&](int cfgx) -> Plan {
        Plan pl;
        if (convex) return pl;
        const int Gx = kCfgs[cfgx].G, Cx = kCfgs[cfgx].C, Wx = Gx * Cx;
        if (fast && !c->no_pack && n) {
            const int64_t smin = std::min<int64_t>(std::min(sc.match, sc.mismatch), std::min(sc.special, 0));
            const int64_t smax = std::max<int64_t>(std::max(sc.match, sc.special), 0);
            const int64_t low = 2ll * sc.b0 + (int64_t)(L1max + L2max + Wx) * sc.b1 + 2ll * sc.oe_in - 1 + smin - 16;
            const int64_t high = smax * std::min<int64_t>(L1max, L2max) - sc.oe_in + 16;
            if (sc.b1 <= 0 && high - low + 128 <= 32767) {
                pl.pack = true;
                pl.pkp.bias = (int32_t)(64 - low);
                // static row slope (MADD kernels, clq_pack.cuh): row x stored with x * slope more, slope = -min(substitution score, 0), so
                // that the profile bytes are >= 0 and M = diag + m is a plain add.  Needs L1 * slope more head-room (Eh' of row 0 sits
                // one slope lower) and bytes that still fit int8.
                const int64_t slope = -smin;
                // (short-read geometries only: on the long-read ones, 168 registers at 3 CTAs/SM, the two extra constants cost more
                // than the ALU-pipe relief gains -- C3 1485 -> 1452 GCUPS with it)
                if (!rb && !c->no_madd && Gx <= 8 && slope > 0 && smax + slope <= 127 && high + slope * (int64_t)L1max - (low - slope) + 128 <= 32767) {
                    pl.madd = true;
                    pl.pkp.slope = (int32_t)slope;
                    pl.pkp.bias = (int32_t)(64 - (low - slope));
                }
            }
        }
        if (fast && !rb && !pl.pack && !c->no_pack && !c->no_adapt && !score_only && n && (Gx >= 16 || cfgx == 2) && sc.b1 <= 0 &&
            search == CLQ_SEARCH_FIXED && pairs_ok) {
            const int smin = std::min(std::min(sc.match, sc.mismatch), sc.special), smax = std::max(std::max(sc.match, sc.special), 0);
            // adjacent cells of a row differ by at most smax - 2 * x1 (clq_pack_adapt.cuh); a lane holds C columns, E / F / M sit within a few opens of B
            const int guard = c->adapt_guard > 0 ? c->adapt_guard : (Cx + 4) * (smax - 2 * sc.oe_in) + 64;
            if (fits8(smax + kAdaptSigmaMax) && fits8(smin + kAdaptSigmaMin) && sc.oe_in >= -512 && (c->adapt_guard > 0 || 2 * guard + 64 + 16384 <= 32767)) {
                pl.adapt = true;
                pl.adp.guard = guard;
            }
        }
        return pl;
    };
    Plan plan = convex ? Plan() : plan_for(cfg);
    // Long reads: the s16x2 kernels run fastest on the SHORT-read geometry (8,40) with as many column stripes as the read needs
    // (C3: 1870 GCUPS on (8,40) with four stripes against 1511 on (32,32); all 8 lanes of a group busy, MADD, 2 CTAs/SM with 255
    // registers).  Taken when every stage of the launch stays on the s16x2 family; the int32 kernels prefer the wide geometries.
    if (!convex && c->force_cfg < 0 && !CLQ_LEAN_NO8 && !c->no_long8 && cfg > 2 && fast && n) {
        const Plan p2 = plan_for(2);
        const bool tb_on_pack = score_only || pairs_ok || (p2.pack && !s->have_order && c->n_refs > 1 && !c->no_group);
        // (the adaptive-bias kernel stays on the geometry the length picks: both reach the same cells/s on 5 kb pairs, but a pair
        // takes ~4x longer on one 8-lane group than on a 32-lane one, and the bits of the 4736 pairs in flight on (8,40) do not
        // fit the scratch -- C5 60k reads: 232 ms on (8,40), 212 ms on (32,32), both with overlapped sub-batches)
        if (p2.pack && tb_on_pack) { cfg = 2; plan = p2; }
    }
    const int G = convex ? kCvxCfgs[cfg].G : kCfgs[cfg].G, C = convex ? kCvxCfgs[cfg].C : kCfgs[cfg].C, W = G * C, GPW = 32 / G;
    const int bits_per_cell = convex ? 8 : 4;
    const uint32_t ns_max = std::max<uint32_t>(1, (L2max + W - 1) / W);
    const uint32_t ref_sm_stride = (L1max + 15) / 16 * 16 + 16;
    // + per-warp transposition buffers of the direction bits (C/8 KiB per warp, twice for the PACK kernels); the
    // convex kernels keep the plain row layout
    const size_t tt_bytes = (!convex && G <= kTransposeMaxG) ? (size_t)(kThreads / 32) * (C / 8) * 1024 * 2 : 0;
    const size_t smem = (size_t)(kThreads / 32) * GPW * ref_sm_stride + (fast ? kLutBytes + kTabBytes : 0) + tt_bytes;
    if (smem > 200 * 1024) return fail(c, CLQ_E_LIMIT, "references too long for this geometry's shared-memory staging");
    const uint32_t ref_stride_adapt = ((L1max + 1) / 2 + 15) / 16 * 16 + 16;  // pack_adapt_kernel stages two reference classes per byte
    const size_t smem_adapt = (size_t)(kThreads / 32) * GPW * ref_stride_adapt + kLutBytes + (size_t)kAdaptTabs * kTabBytes + tt_bytes;  // pack_adapt_kernel: one profile table per slope
    PackParams pkp = plan.pkp;
    bool pack = plan.pack;
    const bool madd = plan.madd;
    if (convex && !c->no_pack && n) pack = cvx_window(W, &pkp.bias);  // two-piece affine: same proof with the gap states of both pieces
    AdaptParams adp = plan.adp;
    adp.ref_stride = ref_stride_adapt;
    const bool adapt = plan.adapt;
    const bool pack_pairs = (pack && pairs_ok) || adapt;
    // multi-reference batches: the traceback stage buckets the reads by reference on the device (ref_scatter_kernel) so that
    // the two reads of a PACK task share theirs; only with the natural read order (uniform lengths, uniform scratch slots)
    const bool group_pairs = pack && !pack_pairs && !score_only && !s->have_order && n > 0 && c->n_refs > 1 && !c->no_group;
    const uint64_t n_pos = group_pairs ? (((uint64_t)n + c->n_refs + 2) & ~1ull) : (s->have_order ? s->n_pos : n);  // processing positions incl. group padding
    uint32_t madd_tab[32] = {};  // the MADD kernels take the profile table with the row slope already added (every byte >= 0)
    cudaStream_t dp_stream = s->stream;  // the stream launch_dp queues on (odd overlapped sub-batches: stream2)
    auto launch_dp = [&](bool tb, const KParams& kp, int* grid, bool query) -> cudaError_t {
        if (convex) {
            if (pack && !kp.all_pairs && (pack_pairs || (tb && group_pairs)))  // the convex PACK kernel is pair-mode only
                return tb ? launch_cvx_pack<true>(cfg, kp, cp, pkp, c->sm_count, smem, dp_stream, grid, query) : launch_cvx_pack<false>(cfg, kp, cp, pkp, c->sm_count, smem, dp_stream, grid, query);
            return tb ? launch_cvx<true>(cfg, kp, cp, c->sm_count, smem, dp_stream, grid, query) : launch_cvx<false>(cfg, kp, cp, c->sm_count, smem, dp_stream, grid, query);
        }
        if (adapt && tb && !kp.all_pairs) return launch_adapt(cfg, kp, adp, c->sm_count, smem_adapt, dp_stream, grid, query);
        if (pack && (kp.all_pairs || pack_pairs || (tb && group_pairs))) {
            if (rb) return launch_pack<true, true>(cfg, kp, pkp, c->sm_count, smem, dp_stream, grid, query);
            if (madd) {
                KParams km = kp;
                memcpy(km.tab, madd_tab, sizeof(madd_tab));
                return tb ? launch_pack<true, false, true>(cfg, km, pkp, c->sm_count, smem, dp_stream, grid, query) : launch_pack<false, false, true>(cfg, km, pkp, c->sm_count, smem, dp_stream, grid, query);
            }
            return tb ? launch_pack<true>(cfg, kp, pkp, c->sm_count, smem, dp_stream, grid, query) : launch_pack<false>(cfg, kp, pkp, c->sm_count, smem, dp_stream, grid, query);
        }
        return launch_any(cfg, tb, fin, fast, rb && tb, kp, c->sm_count, smem, dp_stream, grid, query);
    };

    KParams p = {};
    p.ref_bytes = (const uint8_t*)c->ref_bytes.p;
    p.ref_off = (const uint64_t*)c->ref_off.p;
    p.n_refs = c->n_refs;
    p.read_bytes = (const uint8_t*)s->read_bytes.p;
    p.read_off = (const uint64_t*)s->read_off.p;
    p.n_reads = n;
    p.order = s->have_order ? (const uint32_t*)s->order.p : nullptr;
    p.sc = sc;
    p.band_mode = rb ? 0xffu : band;  // rust-bio's global alignment is unbanded
    p.band_k = band_k;
    p.rustbio = rb ? 1u : 0u;
    p.max_read_len = c->lim.max_read_len;
    p.ref_sm_stride = ref_sm_stride;
    p.results = (clq_result_t*)s->results.p;
    unsigned long long* ctr = (unsigned long long*)s->counters.p;
    p.cigar_cursor = ctr + 4;
    p.cells = ctr + 5;
    p.cls_lut = (const uint8_t*)(rb ? c->cls_lut_rb.p : c->cls_lut.p);
    p.debug_flags = (uint32_t)c->debug_flags;
    if (fast) {  // profile table: row = reference class, column = read class
        int8_t tab[16][8];
        for (int r = 0; r < 16; r++)
            for (int q = 0; q < 8; q++) {
                if (rb) tab[r][q] = (int8_t)((q == 0 || (r == q && r >= 2 && r < 8)) ? sc.match : sc.mismatch);  // read 'N' matches anything
                else tab[r][q] = (int8_t)((r == 0 || q == 0) ? sc.special : ((r == q && r >= 2) ? sc.match : sc.mismatch));
            }
        memcpy(p.tab, tab, 128);
    }
    if (madd) {
        int8_t tab[16][8];
        memcpy(tab, p.tab, 128);
        for (int r = 0; r < 16; r++)
            for (int q = 0; q < 8; q++) tab[r][q] = (int8_t)(tab[r][q] + pkp.slope);
        memcpy(madd_tab, tab, 128);
    }

    // grid + scratch sizing (per resident group)
    int grid_tb = 0, grid_sc = 0;
    cudaError_t ce;
    KParams pq = p;
    pq.all_pairs = (search != CLQ_SEARCH_FIXED) ? 1 : 0;   // which kernel family the score stage will use
    if ((ce = launch_dp(true, p, &grid_tb, true)) != cudaSuccess)
        return fail(c, CLQ_E_CUDA, std::string("occupancy(tb): ") + cudaGetErrorString(ce));
    if ((ce = launch_dp(false, pq, &grid_sc, true)) != cudaSuccess)
        return fail(c, CLQ_E_CUDA, std::string("occupancy(score): ") + cudaGetErrorString(ce));
    int grid_retry = 0;
    if (adapt && (ce = launch_any(cfg, true, false, true, false, p, c->sm_count, smem, s->stream, &grid_retry, true)) != cudaSuccess)
        return fail(c, CLQ_E_CUDA, std::string("occupancy(retry): ") + cudaGetErrorString(ce));
    // direction bits per pair: G <= 8 geometries store blocks of 8 steps x G lanes x WPL words x 8 (time-transposed,
    // BitsLayout in clq_kernels.cuh), the others and the convex kernels one row per step
    const bool transposed = !convex && G <= kTransposeMaxG;
    const uint64_t bits_stride = transposed ? (uint64_t)ns_max * ((L1max + G + 7) / 8) * G * (C / 8) * 8
                                            : ((uint64_t)ns_max * (L1max + G) * G * (C * bits_per_cell / 32) + 7) / 8 * 8;
    const uint32_t cig_stride = L1max + L2max + 8;
    const uint32_t col_stride = (L1max + 8 + 3) & ~3u;  // rows of the stripe-boundary column (F, E, M, B interleaved: 16-byte rows)
    // traceback scratch is per task of a sub-batch: direction bits + CIGAR scratch + walker record.  When the read lengths
    // vary (processing order = longest first) every position gets a slot of its own size and sub-batches are cut by bytes,
    // so a batch of mixed 300 bp - 5 kb reads is not split into tiny sub-batches sized for its longest pair.
    auto bits_words = [&](uint64_t l1, uint64_t l2) -> uint64_t {
        const uint64_t ns = std::max<uint64_t>(1, (l2 + W - 1) / W), T = l1 + G - 1;
        const uint64_t wds = transposed ? ns * ((T + 7) / 8) * G * (C / 8) * 8 : ns * T * G * (C * bits_per_cell / 32);
        return (wds + 7) / 8 * 8;
    };
    int32_t rc;
    const uint64_t per_task = bits_stride * 4 + (uint64_t)cig_stride * 4 + sizeof(TbRec);
    uint64_t sub = std::max<uint64_t>(2, std::min<uint64_t>(n_pos + 1, (uint64_t)c->max_scratch_bytes / per_task)) & ~1ull;  // even: PACK tasks are read pairs
    std::vector<uint64_t> cuts;          // sub-batch boundaries in processing positions
    const uint64_t np = s->have_order ? s->n_pos : n;  // processing positions of a host-ordered batch (incl. reference-group padding)
    const uint32_t* order_tb = nullptr;
    bool overlap = false;  // sub-batches two at a time: even ones on the slot's stream, odd ones on stream2, each stream its half of the scratch
    const bool var_slots = !score_only && n && s->have_order && s->h_len.size() == np;
    uint64_t max_sub_words = sub * bits_stride, max_sub_tasks = sub;
    if (var_slots) {
        static thread_local std::vector<uint64_t> off;
        off.resize((size_t)np + 1);
        off[0] = 0;
        const bool fixed_known = search == CLQ_SEARCH_FIXED && s->h_ref.size() == np;
        for (uint64_t i = 0; i < np; i++) {
            uint64_t l1 = L1max;
            if (fixed_known && s->h_ref[i] >= 0 && (uint32_t)s->h_ref[i] < c->n_refs)
                l1 = c->h_ref_off[s->h_ref[i] + 1] - c->h_ref_off[s->h_ref[i]];
            const uint64_t l2 = s->h_len[i] >= c->lim.max_read_len ? 0 : s->h_len[i];
            off[i + 1] = off[i] + (l2 ? bits_words(l1, l2) : 0);  // empty reads and padding positions store nothing
        }
        const uint64_t budget_words = (uint64_t)c->max_scratch_bytes / 4;
        // Sub-batches.  The positions are sorted longest first; cutting that list into contiguous pieces would make the first
        // sub-batches all-long (a 40 GB piece of 5 kb pairs is ~1600 tasks: not even one wave of the persistent grid, and every
        // task ends at the same time).  Instead the position PAIRS are dealt round-robin over K sub-batches: every sub-batch gets
        // the same mix, still longest first inside, and its short pairs fill the SMs the long ones leave.
        const uint64_t npairs = (np + 1) / 2;
        auto cost = [&](uint64_t i) -> uint64_t { return i < np ? (off[i + 1] - off[i]) + cig_stride + 4 : 0; };
        uint64_t total_cost = 0;
        for (uint64_t i = 0; i < np; i++) total_cost += cost(i);
        uint64_t K = std::max<uint64_t>(1, (total_cost + budget_words - 1) / budget_words);
        static thread_local std::vector<uint64_t> sub_cost;
        for (; K < npairs; K++) {
            sub_cost.assign(K, 0);
            for (uint64_t pr = 0; pr < npairs; pr++) sub_cost[pr % K] += cost(2 * pr) + cost(2 * pr + 1);
            if (*std::max_element(sub_cost.begin(), sub_cost.end()) <= budget_words) break;
        }
        K = std::min<uint64_t>(K, std::max<uint64_t>(npairs, 1));
        cuts.push_back(0);
        if (K > 1 && !c->no_overlap) {
            // Overlapped sub-batches: the positions stay in their order (longest reference, then longest read first) and are cut
            // into contiguous pieces of HALF the scratch budget; piece k runs on stream k % 2 in scratch half k % 2, so the
            // persistent CTAs a piece frees while its last pairs finish are taken by the next piece at once.  The whole launch
            // then behaves like one longest-first task list: the 5 kb pairs (one pair: ~100 ms on one lane group) all start in
            // the first pieces, and the launch ends on short pairs.  (A dealt sequence of sub-batches pays that critical path once
            // per sub-batch: C5 271 ms dealt, 2 sub-batches.)
            overlap = true;
            const uint64_t half = budget_words / 2;
            uint64_t acc = 0;
            for (uint64_t pr = 0; pr < npairs; pr++) {
                const uint64_t cst = cost(2 * pr) + cost(2 * pr + 1);
                if (acc && acc + cst > half) { cuts.push_back(2 * pr); acc = 0; }
                acc += cst;
            }
            cuts.push_back(np);
        } else if (K > 1 && s->h_order.size() == np) {
            static thread_local std::vector<uint32_t> order2, len2;
            static thread_local std::vector<uint64_t> off2;
            order2.clear(); len2.clear();
            off2.assign(1, 0);
            for (uint64_t k = 0; k < K; k++) {
                for (uint64_t pr = k; pr < npairs; pr += K)
                    for (uint64_t i = 2 * pr; i < 2 * pr + 2; i++) {
                        order2.push_back(i < np ? s->h_order[i] : 0xffffffffu);  // an odd tail is padded: every sub-batch holds whole pairs
                        off2.push_back(off2.back() + (i < np ? off[i + 1] - off[i] : 0));
                    }
                cuts.push_back(order2.size());
            }
            off.assign(off2.begin(), off2.end());
            if ((rc = ensure(c, s->order2, (order2.size() + 2) * sizeof(uint32_t))) != CLQ_OK) return rc;
            CU(c, cudaMemcpyAsync(s->order2.p, order2.data(), order2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
            order_tb = (const uint32_t*)s->order2.p;  // the traceback stage walks the dealt order; a score stage keeps the upload's
        } else {
            cuts.push_back(np);
        }
        const uint64_t npos2 = cuts.back();
        max_sub_words = 0; max_sub_tasks = 0;
        for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
            max_sub_words = std::max(max_sub_words, off[cuts[ci + 1]] - off[cuts[ci]]);
            max_sub_tasks = std::max(max_sub_tasks, cuts[ci + 1] - cuts[ci]);
        }
        max_sub_tasks = (max_sub_tasks + 1) & ~1ull;
        if ((rc = ensure(c, s->bits_off, ((size_t)npos2 + 1) * sizeof(uint64_t))) != CLQ_OK) return rc;
        CU(c, cudaMemcpyAsync(s->bits_off.p, off.data(), ((size_t)npos2 + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
        CU(c, cudaStreamSynchronize(s->stream));  // `off` / `order2` are reused host vectors
    } else {
        for (uint64_t b = 0; b < n_pos; b += sub) cuts.push_back(b);
        cuts.push_back(n_pos);
    }
    const uint64_t groups = (uint64_t)std::max(std::max(grid_tb, grid_sc), grid_retry) * (kThreads / 32) * GPW;
    DevBuf& sbits = c->serialize ? c->sh_bits : s->bits;
    DevBuf& scig = c->serialize ? c->sh_cig : s->cig_scratch;
    DevBuf& stbrec = c->serialize ? c->sh_tbrec : s->tb_rec;
    const size_t nh = overlap ? 2 : 1;
    const size_t half_bits = (size_t)max_sub_words * 4, half_cig = ((size_t)max_sub_tasks * cig_stride * 4 + 15) & ~(size_t)15, half_tbrec = (size_t)max_sub_tasks * sizeof(TbRec);
    const size_t half_col = (size_t)groups * col_stride * (adapt ? 20 : 16), half_retry = (((size_t)max_sub_tasks + 2) * sizeof(uint32_t) + 15) & ~(size_t)15;
    if (!score_only && n) {
        if ((rc = ensure(c, sbits, nh * half_bits + 64)) != CLQ_OK) return rc;
        if ((rc = ensure(c, scig, nh * half_cig)) != CLQ_OK) return rc;
        if ((rc = ensure(c, stbrec, nh * half_tbrec)) != CLQ_OK) return rc;
        if ((rc = ensure(c, s->cigar_pool, (size_t)c->lim.cigar_pool_ops * 4 + 16)) != CLQ_OK) return rc;
    }
    if ((rc = ensure(c, s->col_scratch, nh * half_col)) != CLQ_OK) return rc;  // pack_adapt_kernel keeps a fifth array (the bias of every row)
    if (adapt && (rc = ensure(c, s->retry_list, nh * half_retry)) != CLQ_OK) return rc;
    p.bits = (uint32_t*)sbits.p;
    p.bits_stride = bits_stride;
    p.bits_off = var_slots ? (const uint64_t*)s->bits_off.p : nullptr;
    p.cig_scratch = (uint32_t*)scig.p;
    p.cig_stride = cig_stride;
    p.col_scratch = (int32_t*)s->col_scratch.p;
    p.col_stride = col_stride;
    p.cigar_pool = (uint32_t*)s->cigar_pool.p;
    p.cigar_cap = c->lim.cigar_pool_ops;
    p.tb_rec = (TbRec*)stbrec.p;
    if (want_tags && n) {
        if ((rc = ensure(c, s->tags, (size_t)n * c->tag_stride)) != CLQ_OK) return rc;
        p.tag_slot = (const uint16_t*)c->tag_slot.p;
        p.tags = (uint8_t*)s->tags.p;
        p.tag_stride = c->tag_stride;
    }

    s->flags = flags;
    s->stats.variant = (rb ? 16u : 0u) | (fast ? 1u : 0u) | ((pack && (pack_pairs || group_pairs || (search != CLQ_SEARCH_FIXED && !convex))) ? 2u : 0u) | (convex ? 4u : 0u) | (fin ? 8u : 0u) | (madd ? 32u : 0u) | (adapt ? (2u | 64u) : 0u) | ((uint32_t)cfg << 8);
    s->stats.sub_batches = 0;
    s->stats.launches = 0;
    s->stats.dp_launches = 0;
    s->n_dp = 0;
    if (c->serialize && c->last_done && c->last_done != s->done) CU(c, cudaStreamWaitEvent(s->stream, c->last_done, 0));
    CU(c, cudaMemsetAsync(s->counters.p, 0, 16 * sizeof(unsigned long long), s->stream));
    CU(c, cudaEventRecord(s->ev[0], s->stream));

    const int32_t* ref_of_read = nullptr;
    if (search == CLQ_SEARCH_FIXED) {
        ref_of_read = (const int32_t*)s->fixed_ref.p;
    } else if (n) {
        const uint32_t nrefs = c->n_refs;
        const uint32_t mask_words = (nrefs + 31) / 32;
        const uint32_t* cand = nullptr;
        const int32_t* single = nullptr;
        if (search == CLQ_SEARCH_QUICK && nrefs) {
            if ((rc = ensure(c, s->votes, (size_t)n * nrefs * 4)) != CLQ_OK) return rc;
            if ((rc = ensure(c, s->cand_mask, (size_t)n * mask_words * 4)) != CLQ_OK) return rc;
            if ((rc = ensure(c, s->single_ref, (size_t)n * 4)) != CLQ_OK) return rc;
            const size_t vote_smem = (size_t)c->n_keys * 12 + (size_t)nrefs * 128 * 2;
            if (c->kmer_k <= 8 && vote_smem <= 160 * 1024 && !c->force_generic) {
                if (vote_smem > 48 * 1024) CU(c, cudaFuncSetAttribute(kmer_vote_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vote_smem));
                kmer_vote_packed_kernel<<<(n + 127) / 128, 128, vote_smem, s->stream>>>(
                    p.read_bytes, p.read_off, n, (const uint64_t*)c->kmer_keys64.p, (const uint32_t*)c->kmer_owner.p, c->n_keys, c->kmer_k,
                    c->kmer_skip, nrefs, match_threshold, (uint32_t*)s->cand_mask.p, mask_words, (int32_t*)s->single_ref.p);
            } else
            kmer_vote_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(
                p.read_bytes, p.read_off, n, (const uint8_t*)c->kmer_keys.p, (const uint32_t*)c->kmer_owner.p, c->n_keys, c->kmer_k,
                c->kmer_skip, nrefs, match_threshold, (uint32_t*)s->votes.p, (uint32_t*)s->cand_mask.p, mask_words, (int32_t*)s->single_ref.p);
            CU(c, cudaGetLastError());
            s->stats.launches++;
            cand = (const uint32_t*)s->cand_mask.p;
            single = (const int32_t*)s->single_ref.p;
        }
        if (nrefs) {
            if ((rc = ensure(c, s->scores, (size_t)n * nrefs * 4)) != CLQ_OK) return rc;
            KParams q = p;
            q.all_pairs = 1;
            q.n_tasks = (pack && !convex) ? ((n + 1) / 2) * nrefs : n * nrefs;  // the convex all-pairs stage stays on the int32 kernel
            if ((uint64_t)n * nrefs > 0xfffffff0ull) return fail(c, CLQ_E_LIMIT, "reads x references exceeds 2^32 tasks per batch");
            q.cand_mask = cand;
            q.mask_words = mask_words;
            q.scores = (int32_t*)s->scores.p;
            q.task_counter = (unsigned int*)(ctr + 0);
            CU(c, cudaEventRecord(s->ev[1], s->stream));
            int g = grid_sc;
            if ((ce = launch_dp(false, q, &g, false)) != cudaSuccess)
                return fail(c, CLQ_E_CUDA, std::string("score kernel: ") + cudaGetErrorString(ce));
            CU(c, cudaEventRecord(s->ev[2], s->stream));
            s->stats.launches++;
            s->stats.dp_launches++;
            s->n_dp |= 1;
        }
        select_best_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>((const int32_t*)s->scores.p, n, nrefs, cand, mask_words,
                                                                   (int32_t*)s->ref_of_read.p, single);
        CU(c, cudaGetLastError());
        s->stats.launches++;
        ref_of_read = (const int32_t*)s->ref_of_read.p;
    }
    if (n) {
        KParams q = p;
        q.all_pairs = 0;
        q.ref_of_read = ref_of_read;
        q.task_counter = (unsigned int*)(ctr + 1);
        if (order_tb) q.order = order_tb;
        if (group_pairs) {
            const uint32_t nb = c->n_refs + 1;
            if ((rc = ensure(c, s->ref_groups, (size_t)2 * nb * sizeof(uint32_t))) != CLQ_OK) return rc;
            if ((rc = ensure(c, s->order, n_pos * sizeof(uint32_t))) != CLQ_OK) return rc;
            uint32_t* hist = (uint32_t*)s->ref_groups.p;
            uint32_t* cursor = hist + nb;
            CU(c, cudaMemsetAsync(hist, 0, (size_t)2 * nb * sizeof(uint32_t), s->stream));
            CU(c, cudaMemsetAsync(s->order.p, 0xff, n_pos * sizeof(uint32_t), s->stream));
            ref_hist_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(ref_of_read, n, c->n_refs, hist);
            ref_scan_kernel<<<1, 32, 0, s->stream>>>(hist, nb, cursor);
            ref_scatter_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(ref_of_read, n, c->n_refs, cursor, (uint32_t*)s->order.p);
            CU(c, cudaGetLastError());
            s->stats.launches += 3;
            q.order = (const uint32_t*)s->order.p;
        }
        const bool tb_pairs = pack_pairs || group_pairs;
        CU(c, cudaEventRecord(s->ev[3], s->stream));
        if (score_only) {
            // positions, not reads: a host order that groups by reference holds padding positions
            // (the adaptive kernel is traceback-only: a score-only launch of long pairs stays on the int32 kernel)
            const bool sc_pairs = pack_pairs && !adapt;
            q.n_tasks = (uint32_t)(sc_pairs ? (np + 1) / 2 : np);
            q.task_base = 0;
            q.task_end = (uint32_t)np;
            int g = grid_sc;
            if ((ce = launch_dp(false, q, &g, false)) != cudaSuccess)
                return fail(c, CLQ_E_CUDA, std::string("score kernel: ") + cudaGetErrorString(ce));
            s->stats.launches++;
            s->stats.dp_launches++;
        } else {
            // fill (direction bits into the sub-batch's slots) then walk (one thread per pair), sub-batch by sub-batch
            if (overlap) {
                CU(c, cudaEventRecord(s->fork, s->stream));
                CU(c, cudaStreamWaitEvent(s->stream2, s->fork, 0));
            }
            for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
                const uint64_t base = cuts[ci];
                const uint32_t cnt = (uint32_t)(cuts[ci + 1] - base);
                if (!cnt) continue;
                const size_t h = overlap ? (ci & 1) : 0;   // scratch half + stream of this sub-batch
                cudaStream_t st = h ? s->stream2 : s->stream;
                dp_stream = st;
                unsigned long long* c_fill = ctr + (h ? 8 : 1);
                unsigned long long* c_retry = ctr + (h ? 9 : 2);
                unsigned long long* c_count = ctr + (h ? 10 : 6);
                q.n_tasks = tb_pairs ? (cnt + 1) / 2 : cnt;
                q.task_base = (uint32_t)base;
                q.task_end = (uint32_t)(base + cnt);
                q.task_counter = (unsigned int*)c_fill;
                q.bits = (uint32_t*)((char*)sbits.p + h * half_bits);
                q.cig_scratch = (uint32_t*)((char*)scig.p + h * half_cig);
                q.tb_rec = (TbRec*)((char*)stbrec.p + h * half_tbrec);
                q.col_scratch = (int32_t*)((char*)s->col_scratch.p + h * half_col);
                const bool again = ci >= nh;  // this half's counters were used by an earlier sub-batch (same stream: ordered)
                if (again) CU(c, cudaMemsetAsync(c_fill, 0, sizeof(unsigned long long), st));
                if (adapt) {  // retry pass bookkeeping: its task counter and the number of listed positions
                    if (again) {
                        CU(c, cudaMemsetAsync(c_retry, 0, sizeof(unsigned long long), st));
                        CU(c, cudaMemsetAsync(c_count, 0, sizeof(unsigned long long), st));
                    }
                    adp.retry_list = (uint32_t*)((char*)s->retry_list.p + h * half_retry);
                    adp.retry_count = (unsigned int*)c_count;
                    adp.retry_total = (unsigned int*)(ctr + 7);
                }
                int g = grid_tb;
                if ((ce = launch_dp(true, q, &g, false)) != cudaSuccess)
                    return fail(c, CLQ_E_CUDA, std::string("fill kernel: ") + cudaGetErrorString(ce));
                s->stats.launches++;
                s->stats.dp_launches++;
                s->stats.sub_batches++;
                if (adapt && !(c->debug_flags & 2)) {
                    // pairs that left the guarded window are redone by the int32 kernel, which writes their records, walker records and
                    // direction bits into the same slots (same geometry); the number of tasks is read on the device
                    KParams q2 = q;
                    q2.retry_list = adp.retry_list;
                    q2.n_tasks_dev = (const unsigned int*)c_count;
                    q2.n_tasks = 0;
                    q2.task_counter = (unsigned int*)c_retry;
                    int g2 = grid_retry;
                    if ((ce = launch_any(cfg, true, false, true, false, q2, c->sm_count, smem, st, &g2, false)) != cudaSuccess)
                        return fail(c, CLQ_E_CUDA, std::string("retry kernel: ") + cudaGetErrorString(ce));
                    s->stats.launches++;
                    s->stats.dp_launches++;
                }
                if (!(c->debug_flags & 1)) {
                    if ((ce = convex ? launch_cvx_walk(cfg, q, tb_pairs ? 2 * q.n_tasks : cnt, st) : launch_walk(cfg, q, tb_pairs ? 2 * q.n_tasks : cnt, st)) != cudaSuccess)
                        return fail(c, CLQ_E_CUDA, std::string("walk kernel: ") + cudaGetErrorString(ce));
                    s->stats.launches++;
                }
            }
            dp_stream = s->stream;
            if (overlap) {
                CU(c, cudaEventRecord(s->join, s->stream2));
                CU(c, cudaStreamWaitEvent(s->stream, s->join, 0));
            }
        }
        CU(c, cudaEventRecord(s->ev[4], s->stream));
        s->n_dp |= 2;
    }
    CU(c, cudaEventRecord(s->ev[5], s->stream));
    CU(c, cudaEventRecord(s->done, s->stream));
    c->last_done = s->done;
    s->launched = true;
    s->state = 2;
    return CLQ_OK;
}

int32_t clq_download(clq_ctx* c, int32_t slot) {
    Slot* s = get_slot(c, slot);
    if (!s) return CLQ_E_INVALID;
    if (s->state < 2) return fail(c, CLQ_E_STATE, "clq_download before clq_launch");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(s->h_counters, s->counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    s->state = 3;
    return CLQ_OK;
}

int32_t clq_sync(clq_ctx* c, int32_t slot) {
    Slot* s = get_slot(c, slot);
    if (!s) return CLQ_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(s->stream));
    return CLQ_OK;
}

int32_t clq_slot_stats(clq_ctx* c, int32_t slot, clq_stats_t* out) {
    Slot* s = get_slot(c, slot);
    if (!s || !out) return CLQ_E_INVALID;
    if (!s->launched) return fail(c, CLQ_E_STATE, "no launch on this slot");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    CU(c, cudaEventElapsedTime(&ms, s->ev[0], s->ev[5]));
    s->stats.kernel_ms = ms;
    float dp = 0.f;
    if (s->n_dp & 1) { CU(c, cudaEventElapsedTime(&ms, s->ev[1], s->ev[2])); dp += ms; }
    if (s->n_dp & 2) { CU(c, cudaEventElapsedTime(&ms, s->ev[3], s->ev[4])); dp += ms; }
    s->stats.dp_ms = dp;
    unsigned long long cells = 0;
    CU(c, cudaMemcpy(&cells, (unsigned long long*)s->counters.p + 5, sizeof(cells), cudaMemcpyDeviceToHost));
    s->stats.cells = cells;
    *out = s->stats;
    return CLQ_OK;
}

int32_t clq_wait(clq_ctx* c, int32_t slot, clq_result_t* results, uint32_t* cigar_pool, uint64_t cigar_cap, uint64_t* cigar_used) {
    Slot* s = get_slot(c, slot);
    if (!s) return CLQ_E_INVALID;
    if (s->state < 2) return fail(c, CLQ_E_STATE, "clq_wait before clq_launch");
    int32_t rc;
    if (s->state < 3 && (rc = clq_download(c, slot)) != CLQ_OK) return rc;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(s->stream));
    const uint64_t cursor = s->h_counters[4];
    const uint64_t used = std::min<uint64_t>(cursor, c->lim.cigar_pool_ops);
    s->stats.cells = s->h_counters[5];
    s->stats.pack_retries = (uint32_t)(s->h_counters[7] & 0xffffffffull);
    uint64_t d2h = 8 * sizeof(unsigned long long);
    if (results && s->n_reads) {
        CU(c, cudaMemcpyAsync(results, s->results.p, (size_t)s->n_reads * sizeof(clq_result_t), cudaMemcpyDeviceToHost, s->stream));
        d2h += (size_t)s->n_reads * sizeof(clq_result_t);
    }
    const bool has_pool = !(s->flags & CLQ_SCORE_ONLY);
    if (cigar_used) *cigar_used = has_pool ? used : 0;
    if (has_pool && used) {
        if (!cigar_pool || used > cigar_cap) {
            CU(c, cudaStreamSynchronize(s->stream));
            return fail(c, CLQ_E_LIMIT, "caller's CIGAR buffer is smaller than the ops produced");
        }
        CU(c, cudaMemcpyAsync(cigar_pool, s->cigar_pool.p, used * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
        d2h += used * sizeof(uint32_t);
    }
    CU(c, cudaStreamSynchronize(s->stream));
    s->stats.d2h_bytes = d2h;
    s->state = 1;  // inputs stay resident: the slot can be launched again
    return CLQ_OK;
}

int32_t clq_tags_download(clq_ctx* c, int32_t slot, uint8_t* tags, uint64_t cap, uint32_t* tag_stride) {
    Slot* s = get_slot(c, slot);
    if (!s || !tag_stride) return CLQ_E_INVALID;
    if (!s->launched) return fail(c, CLQ_E_STATE, "clq_tags_download before clq_launch");
    *tag_stride = 0;
    if (!(s->flags & CLQ_EXTRACT_TAGS)) return fail(c, CLQ_E_STATE, "the last launch on this slot did not ask for CLQ_EXTRACT_TAGS");
    *tag_stride = c->tag_stride;
    const uint64_t bytes = (uint64_t)s->n_reads * c->tag_stride;
    if (!bytes || !tags) return CLQ_OK;  // tags == NULL: only report the stride
    if (cap < bytes) return fail(c, CLQ_E_LIMIT, "caller's tag buffer is smaller than n_reads * tag_stride");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpyAsync(tags, s->tags.p, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU(c, cudaStreamSynchronize(s->stream));
    s->stats.d2h_bytes += bytes;
    return CLQ_OK;
}

int32_t clq_submit(clq_ctx* c, int32_t slot, uint32_t n_reads, const uint8_t* read_bytes, const uint64_t* read_off,
                   const int32_t* fixed_ref, const void* scoring, uint32_t flags, double match_threshold) {
    int32_t rc;
    if ((rc = clq_upload(c, slot, n_reads, read_bytes, read_off, fixed_ref)) != CLQ_OK) return rc;
    if ((rc = clq_launch(c, slot, scoring, flags, match_threshold)) != CLQ_OK) return rc;
    return clq_download(c, slot);
}

int32_t clq_submit_packed2(clq_ctx* c, int32_t slot, uint32_t n_reads, const uint32_t* packed, const uint64_t* read_off,
                           const uint64_t* exc_pos, const uint8_t* exc_byte, uint64_t n_exc, const int32_t* fixed_ref,
                           const void* scoring, uint32_t flags, double match_threshold) {
    int32_t rc;
    if ((rc = clq_upload_packed2(c, slot, n_reads, packed, read_off, exc_pos, exc_byte, n_exc, fixed_ref)) != CLQ_OK) return rc;
    if ((rc = clq_launch(c, slot, scoring, flags, match_threshold)) != CLQ_OK) return rc;
    return clq_download(c, slot);
}

}  // extern "C"

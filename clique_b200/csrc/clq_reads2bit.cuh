// clq_reads2bit.cuh -- 2-bit packed read ingestion (BASELINE.json north_star (1): "read batching, 2-bit packing").
//
// The host layer may ship a batch as a 2-bit stream (A C G T = 0 1 2 3, base i of the concatenated batch in bits
// 2 (i % 16) .. of 32-bit word i / 16) plus a sorted exception list (position, byte) for every byte that is not an upper-case
// A / C / G / T: the reference's match_mismatch compares raw bytes, case included (alignment/scoring_functions.rs:100-102), so
// 'N', IUPAC codes and soft-masked lower-case bases must reach the DP kernels unchanged.  The stream is expanded on the device
// into the same ASCII read buffer the ASCII upload fills; every kernel behind it is shared with that path, so results are
// identical by construction (tests/test_reads2bit.py checks the round trip on the CPU and the batch results on the GPU).
//
// unpack2_kernel is a pure streaming kernel: one 128-bit load of packed words per lane (64 bases), redistributed inside the
// warp so that each of the four store instructions writes 512 contiguous bytes (one 128-bit store per lane).  Traffic: 0.25 B
// read + 1 B written per base; at C2's 300 MB per step it is ~60 us of a 31 ms step.
#pragma once

#include <cstdint>

namespace clq {

// 8 bits = 4 bases -> 4 ASCII bytes: the codes become the nibbles of a PRMT selector into the constant "ACGT"
__device__ __forceinline__ uint32_t expand4(uint32_t b) {
    const uint32_t sel = (b & 3u) | ((b & 0xcu) << 2) | ((b & 0x30u) << 4) | ((b & 0xc0u) << 6);
    return __byte_perm(0x54474341u /* 'A' 'C' 'G' 'T' */, 0u, sel);
}

__device__ __forceinline__ uint4 expand16(uint32_t w) {
    return make_uint4(expand4(w & 0xffu), expand4((w >> 8) & 0xffu), expand4((w >> 16) & 0xffu), expand4(w >> 24));
}

// packed: n_words 32-bit words (the buffer is padded to a multiple of 4 words); out: 16 bytes per word
__global__ void __launch_bounds__(256) unpack2_kernel(const uint4* __restrict__ packed, uint4* __restrict__ out, uint64_t n_words) {
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_vec = (n_words + 3) >> 2;
    for (uint64_t wv = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; wv < n_vec; wv += warps * 32) {
        const uint64_t iv = wv + lane;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (iv < n_vec) v = __ldg(packed + iv);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // store k covers words 32 k .. 32 k + 31 of the warp's 128: lane l takes word 32 k + l = component l % 4 of lane 8 k + l / 4
            const int src = 8 * k + (int)(lane >> 2);
            const uint32_t x = __shfl_sync(0xffffffffu, v.x, src), y = __shfl_sync(0xffffffffu, v.y, src);
            const uint32_t z = __shfl_sync(0xffffffffu, v.z, src), w = __shfl_sync(0xffffffffu, v.w, src);
            const unsigned q = lane & 3u;
            const uint32_t word = q == 0 ? x : q == 1 ? y : q == 2 ? z : w;
            const uint64_t iw = wv * 4 + 32u * k + lane;
            if (iw < n_words) out[iw] = expand16(word);
        }
    }
}

// the bytes the 2-bit alphabet cannot carry
__global__ void __launch_bounds__(256) patch2_kernel(uint8_t* __restrict__ out, const uint64_t* __restrict__ pos, const uint8_t* __restrict__ byte,
                                                     uint64_t n_exc, uint64_t n_bytes) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_exc; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t at = pos[i];
        if (at < n_bytes) out[at] = byte[i];
    }
}

}  // namespace clq

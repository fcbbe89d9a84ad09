// clq_pack2_host.cpp -- the plain-run kernel of clq_pack2 (include/clq.h): packs consecutive 16-base words of upper-case A C G T into
// 2-bit words until it meets a word holding any other byte.  Host code only (compiled by the host compiler, no CUDA): an AVX2
// body picked at run time (32 bases per iteration) and a portable 64-bit SWAR body.
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define CLQ_HAVE_AVX2_BODY 1
#else
#define CLQ_HAVE_AVX2_BODY 0
#endif

namespace clq {

namespace {

// 8 bases in a 64-bit register: code = ((b >> 1) ^ (b >> 2)) & 3 maps A C G T to 0 1 2 3; the chunk is plain iff rebuilding the
// letters from the codes ('A' + 2 lo + 6 hi + 11 (lo & hi), no carry between bytes) gives the chunk back
inline bool pack8(const uint8_t* q, uint32_t& out) {
    const uint64_t K01 = 0x0101010101010101ull;
    uint64_t x;
    std::memcpy(&x, q, 8);
    uint64_t t = ((x >> 1) ^ (x >> 2)) & (3 * K01);
    const uint64_t lo = t & K01, hi = (t >> 1) & K01;
    const uint64_t e = 0x41 * K01 + 2 * lo + 6 * hi + 11 * (lo & hi);
    t = (t | (t >> 6)) & 0x000f000f000f000full;
    t = (t | (t >> 12)) & 0x000000ff000000ffull;
    t = (t | (t >> 24)) & 0xffffull;
    out = (uint32_t)t;
    return e == x;
}

size_t plain_run_swar(const uint8_t* bytes, size_t n_words, uint32_t* packed) {
    for (size_t w = 0; w < n_words; w++) {
        uint32_t l, h;
        const bool okl = pack8(bytes + 16 * w, l), okh = pack8(bytes + 16 * w + 8, h);
        if (!(okl && okh)) return w;
        packed[w] = l | (h << 16);
    }
    return n_words;
}

#if CLQ_HAVE_AVX2_BODY
__attribute__((target("avx2"))) size_t plain_run_avx2(const uint8_t* bytes, size_t n_words, uint32_t* packed) {
    const __m256i m3 = _mm256_set1_epi8(3);
    const __m256i letters = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w14 = _mm256_set1_epi16(0x0401);       // bytes (1, 4): c0 + 4 c1 per 16-bit lane
    const __m256i w116 = _mm256_set1_epi32(0x00100001);  // words (1, 16): (c0 + 4 c1) + 16 (c2 + 4 c3) per 32-bit lane
    size_t w = 0;
    for (; w + 2 <= n_words; w += 2) {
        const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bytes + 16 * w));
        const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), m3);
        const __m256i back = _mm256_shuffle_epi8(letters, t);
        if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(back, x)) != -1) break;  // a byte outside ACGT in these two words
        const __m256i p16 = _mm256_maddubs_epi16(t, w14);
        const __m256i p32 = _mm256_madd_epi16(p16, w116);                    // one byte (4 bases) per 32-bit lane
        const __m256i b16 = _mm256_packus_epi32(p32, p32);
        const __m256i b8 = _mm256_packus_epi16(b16, b16);                    // per 128-bit lane: its 4 bytes, repeated
        packed[w] = (uint32_t)_mm256_extract_epi32(b8, 0);
        packed[w + 1] = (uint32_t)_mm256_extract_epi32(b8, 4);
    }
    return w + plain_run_swar(bytes + 16 * w, (n_words - w) < 2 ? (n_words - w) : 2, packed + w);  // the odd last word / the word before the break
}
#endif

}  // namespace

// Packs words 0 .. r-1 and returns r: the first word that holds a byte outside ACGT (r == n_words: none).
size_t pack2_plain_run(const uint8_t* bytes, size_t n_words, uint32_t* packed) {
#if CLQ_HAVE_AVX2_BODY
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return plain_run_avx2(bytes, n_words, packed);
#endif
    return plain_run_swar(bytes, n_words, packed);
}

}  // namespace clq

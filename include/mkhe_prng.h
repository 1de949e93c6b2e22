/* mkhe_prng.h -- "mkhe-ctr-1": the counter-based generator and the three samplers behind the device-side key generation and
 * encryption (mkhe_keygen_*, mkhe_encrypt, mkhe_crs_sample in mkhe.h).
 *
 * The reference draws from lattigo's Blake2b-keyed PRNG with a fresh random key per object (mkrlwe/params.go:28, keygen.go:26,
 * encryptor.go:32): its streams cannot be reproduced by anybody, so what a replacement has to preserve is the DISTRIBUTIONS and the
 * key equations, not the bits.  This generator is stateless -- word t of element idx of stream `stream` under `seed` is a pure
 * function -- so a GPU thread per coefficient, the CPU oracle (oracle/mkhe_oracle.c includes this very file) and a Go caller that
 * wants to reproduce a key all get the same values in any order.
 *
 *   uniform mod q    lattigo ring.UniformSampler: mask to bitlen(q), reject x >= q (words t = 0, 1, ... of the element)
 *   ternary          lattigo ring.TernarySampler: P(0) = p, P(+1) = P(-1) = (1 - p) / 2; p as a 53-bit threshold (keygen.go:58-60: p = 1/2)
 *   gaussian         lattigo ring.GaussianSampler(sigma = 3.2 = rlwe.DefaultSigma, bound = int(6 sigma) = 19; keygen.go:35): the rounded
 *                    normal restricted to |x| <= 19, here by inversion of the exact cumulative table of |x| (tools/gen_cdt.py;
 *                    integer only, hence identical on CPU and GPU)
 * Plain C, no dependencies; usable from CUDA device code (MKHE_PRNG_FN). */
#ifndef MKHE_PRNG_H
#define MKHE_PRNG_H
#include <stdint.h>

#ifdef __CUDACC__
#define MKHE_PRNG_FN static __host__ __device__ __forceinline__
#else
#define MKHE_PRNG_FN static inline
#endif

#define MKHE_PRNG_GOLD 0x9E3779B97F4A7C15ull
#define MKHE_GAUSS_BOUND 19

/* splitmix64 finaliser */
MKHE_PRNG_FN uint64_t mkhe_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* key of element idx of a stream (hoisted out of the per-word function: a sampler draws several words per element) */
MKHE_PRNG_FN uint64_t mkhe_ctr_key(uint64_t seed, uint64_t stream, uint64_t idx) {
    const uint64_t k = mkhe_mix64(seed + MKHE_PRNG_GOLD * (stream + 1));
    return mkhe_mix64(k + MKHE_PRNG_GOLD * (idx + 1));
}
MKHE_PRNG_FN uint64_t mkhe_ctr_word(uint64_t key, uint64_t t) { return mkhe_mix64(key + MKHE_PRNG_GOLD * (t + 1)); }

/* uniform in [0, q), q < 2^63: masked rejection (success probability > 1/2 per word; 64 words never run out in practice) */
MKHE_PRNG_FN uint64_t mkhe_sample_uniform(uint64_t seed, uint64_t stream, uint64_t idx, uint64_t q) {
    const uint64_t key = mkhe_ctr_key(seed, stream, idx);
    uint64_t mask = q;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t x = 0;
    for (uint64_t t = 0; t < 64; t++) {
        x = mkhe_ctr_word(key, t) & mask;
        if (x < q) return x;
    }
    return x % q;
}
/* ternary: -1, 0, +1; thr53 = floor(P(0) * 2^53) */
MKHE_PRNG_FN int mkhe_sample_ternary(uint64_t seed, uint64_t stream, uint64_t idx, uint64_t thr53) {
    const uint64_t key = mkhe_ctr_key(seed, stream, idx);
    if ((mkhe_ctr_word(key, 0) >> 11) < thr53) return 0;
    return (mkhe_ctr_word(key, 1) & 1) ? 1 : -1;
}
/* cumulative distribution of |x|, x = round(N(0, 3.2^2)) given |x| <= 19, scaled to 2^64 (tools/gen_cdt.py) */
MKHE_PRNG_FN int mkhe_sample_gaussian(uint64_t seed, uint64_t stream, uint64_t idx) {
    const uint64_t cdt[MKHE_GAUSS_BOUND + 1] = {
        0x1fc936d04f96104aull, 0x5c5a387a5ace6ef8ull, 0x90ba6b482d08ad5bull, 0xb9d6e65f9ea011c1ull, 0xd7212f219d8a79a4ull,
        0xea123150a10ceeffull, 0xf5307037160ba631ull, 0xfb1cdace999bbc51ull, 0xfdfa2ad2ef32b011ull, 0xff3c09d58f775249ull,
        0xffbc4515c6317697ull, 0xffeaa370f8d8fe80ull, 0xfff9db54855945feull, 0xfffe63de5d682fa9ull, 0xffff9da4e5787e12ull,
        0xffffeaa4776e9ad3ull, 0xfffffbcaa4d6ac10ull, 0xffffff4214991c7cull, 0xffffffe4e419af7full, 0xffffffffffffffffull};
    const uint64_t key = mkhe_ctr_key(seed, stream, idx);
    const uint64_t u = mkhe_ctr_word(key, 0);
    int m = 0;
    for (int j = 0; j < MKHE_GAUSS_BOUND; j++) m += u >= cdt[j];
    if (m == 0) return 0;
    return (mkhe_ctr_word(key, 1) & 1) ? m : -m;
}
#endif

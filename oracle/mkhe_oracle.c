/*
 * mkhe_oracle.c -- CPU ORACLE (test infrastructure only; see mkhe_oracle.h header comment).
 *
 * Every function cites the reference file:line it restates.  Bare file names are
 * /root/reference/mkrlwe/<file>; "lattigo" = github.com/ldsec/lattigo/v2 v2.3.0 (not vendored,
 * restated from its published algorithm; the reference call sites are given instead).
 */
#include "mkhe_oracle.h"
#include "../include/mkhe_prng.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

#define PRAGMA(x) _Pragma(#x)
#ifdef _OPENMP
#define PAR_FOR PRAGMA(omp parallel for schedule(static))
#else
#define PAR_FOR
#endif

int ork_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "mkhe_oracle: out of memory (%zu bytes)\n", n); abort(); }
    return p;
}
static void *xcalloc(size_t n, size_t s) {
    void *p = calloc(n ? n : 1, s);
    if (!p) { fprintf(stderr, "mkhe_oracle: out of memory\n"); abort(); }
    return p;
}

/* ======================================================================================
 * lattigo ring/modular_reduction.go (used at basis_extension.go:44-48,110-148,220,489,551)
 * ====================================================================================== */

static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) >> 64); }

/* MRed: x*y*2^-64 mod q, canonical */
static inline uint64_t mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv) {
    u128 m = (u128)x * y;
    uint64_t mhi = (uint64_t)(m >> 64), mlo = (uint64_t)m;
    uint64_t hhi = mulhi64(mlo * qinv, q);
    uint64_t r = mhi - hhi + q;
    if (r >= q) r -= q;
    return r;
}
/* MRedConstant: same without the final subtraction, result in [0,2q) */
static inline uint64_t mred_const(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv) {
    u128 m = (u128)x * y;
    uint64_t mhi = (uint64_t)(m >> 64), mlo = (uint64_t)m;
    uint64_t hhi = mulhi64(mlo * qinv, q);
    return mhi - hhi + q;
}
/* MForm: a*2^64 mod q using BRedParams u = [hi,lo] of floor(2^128/q) */
static inline uint64_t mform(uint64_t a, uint64_t q, const uint64_t *u) {
    uint64_t mhi = mulhi64(a, u[1]);
    uint64_t r = -(a * u[0] + mhi) * q;
    if (r >= q) r -= q;
    return r;
}
/* BRedAdd: a mod q for any 64-bit a */
static inline uint64_t bred_add(uint64_t a, uint64_t q, const uint64_t *u) {
    uint64_t mhi = mulhi64(a, u[0]);
    uint64_t r = a - mhi * q;
    if (r >= q) r -= q;
    return r;
}
static inline uint64_t cred(uint64_t a, uint64_t q) { return a >= q ? a - q : a; }
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

uint64_t ork_modexp(uint64_t x, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    x %= q;
    while (e) {
        if (e & 1) r = mulmod(r, x, q);
        x = mulmod(x, x, q);
        e >>= 1;
    }
    return r;
}
uint64_t ork_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv) { return mred(x, y, q, qinv); }
uint64_t ork_mform(uint64_t a, uint64_t q, const uint64_t *bred) { return mform(a, q, bred); }
uint64_t ork_bred_add(uint64_t a, uint64_t q, const uint64_t *bred) { return bred_add(a, q, bred); }

/* MRedParams: q^-1 mod 2^64 (Newton iteration) */
static uint64_t mred_params(uint64_t q) {
    uint64_t x = 1;
    for (int i = 0; i < 63; i++) { x *= q; q *= q; }
    return x; /* lattigo: qInv = q^(2^63-1) = q^-1 mod 2^64 */
}
static void bred_params(uint64_t q, uint64_t *u) {
    /* floor(2^128/q) = hi*2^64 + lo */
    u128 r = ((u128)1 << 127);
    /* compute floor(2^128/q) via two-step long division */
    uint64_t hi = (uint64_t)((((u128)1) << 64) / q);            /* floor(2^64/q) */
    u128 rem = (((u128)1) << 64) - (u128)hi * q;                /* 2^64 mod q    */
    uint64_t lo = (uint64_t)((rem << 64) / q);
    (void)r;
    u[0] = hi; u[1] = lo;
}

/* ======================================================================================
 * lattigo ring/ring.go : primitiveRoot, genNTTParams (tables are public fields of ring.Ring,
 * read e.g. at basis_extension.go:263-264)
 * ====================================================================================== */
static int is_prime64(uint64_t n) {
    if (n < 2) return 0;
    static const uint64_t small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (int i = 0; i < 12; i++) { if (n % small[i] == 0) return n == small[i]; }
    uint64_t d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (int i = 0; i < 12; i++) {
        uint64_t a = small[i];
        uint64_t x = ork_modexp(a, d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int r = 1; r < s; r++) { x = mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}
static uint64_t gcd64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
static uint64_t pollard_rho(uint64_t n) {
    if (!(n & 1)) return 2;
    for (uint64_t c = 1;; c++) {
        uint64_t x = 2, y = 2, d = 1;
        while (d == 1) {
            x = (mulmod(x, x, n) + c) % n;
            y = (mulmod(y, y, n) + c) % n;
            y = (mulmod(y, y, n) + c) % n;
            d = gcd64(x > y ? x - y : y - x, n);
        }
        if (d != n) return d;
    }
}
static void factor_rec(uint64_t n, uint64_t *f, int *nf) {
    if (n == 1) return;
    if (is_prime64(n)) {
        for (int i = 0; i < *nf; i++) if (f[i] == n) return;
        f[(*nf)++] = n;
        return;
    }
    uint64_t d = pollard_rho(n);
    factor_rec(d, f, nf);
    factor_rec(n / d, f, nf);
}
/* primitiveRoot: g=2; loop { g++; test g^((q-1)/f) != 1 for every prime factor f of q-1 } */
uint64_t ork_primitive_root(uint64_t q) {
    uint64_t f[64]; int nf = 0;
    factor_rec(q - 1, f, &nf);
    uint64_t g = 2;
    for (;;) {
        g++;
        int ok = 1;
        for (int i = 0; i < nf; i++) {
            if (ork_modexp(g, (q - 1) / f[i], q) == 1) { ok = 0; break; }
        }
        if (ok) return g;
    }
}
static inline uint64_t bitrev(uint64_t x, int bits) {
    uint64_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | ((x >> i) & 1); }
    return r;
}

ork_ring *ork_ring_new(int logN, const uint64_t *moduli, int nmod) {
    ork_ring *r = (ork_ring *)xcalloc(1, sizeof(*r));
    int N = 1 << logN;
    r->logN = logN; r->N = N; r->nmod = nmod;
    r->q = (uint64_t *)xmalloc(sizeof(uint64_t) * nmod);
    r->qinv = (uint64_t *)xmalloc(sizeof(uint64_t) * nmod);
    r->bred = (uint64_t *)xmalloc(sizeof(uint64_t) * 2 * nmod);
    r->ninv = (uint64_t *)xmalloc(sizeof(uint64_t) * nmod);
    r->psi = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)nmod * N);
    r->psiinv = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)nmod * N);
    r->rescale = (uint64_t *)xcalloc((size_t)(nmod > 1 ? nmod - 1 : 1) * nmod, sizeof(uint64_t));
    for (int i = 0; i < nmod; i++) {
        uint64_t qi = moduli[i];
        r->q[i] = qi;
        r->qinv[i] = mred_params(qi);
        bred_params(qi, &r->bred[2 * i]);
        const uint64_t *u = &r->bred[2 * i];
        r->ninv[i] = mform(ork_modexp((uint64_t)N, qi - 2, qi), qi, u);
        uint64_t g = ork_primitive_root(qi);
        uint64_t twoN = (uint64_t)N << 1;
        uint64_t power = (qi - 1) / twoN;
        uint64_t powerInv = (qi - 1) - power;
        uint64_t psiMont = mform(ork_modexp(g, power, qi), qi, u);
        uint64_t psiInvMont = mform(ork_modexp(g, powerInv, qi), qi, u);
        uint64_t *psi = r->psi + (size_t)i * N, *psiinv = r->psiinv + (size_t)i * N;
        psi[0] = mform(1, qi, u);
        psiinv[0] = mform(1, qi, u);
        for (uint64_t j = 1; j < (uint64_t)N; j++) {
            uint64_t prev = bitrev(j - 1, logN), next = bitrev(j, logN);
            psi[next] = mred(psi[prev], psiMont, qi, r->qinv[i]);
            psiinv[next] = mred(psiinv[prev], psiInvMont, qi, r->qinv[i]);
        }
    }
    /* RescaleParams[l-1][i] = MForm(q_l^-1 mod q_i) for i < l (lattigo ring.go genRescaleParams) */
    for (int l = 1; l < nmod; l++)
        for (int i = 0; i < l; i++) {
            uint64_t qi = r->q[i];
            r->rescale[(size_t)(l - 1) * nmod + i] = mform(ork_modexp(r->q[l] % qi, qi - 2, qi), qi, &r->bred[2 * i]);
        }
    return r;
}
void ork_ring_free(ork_ring *r) {
    if (!r) return;
    free(r->q); free(r->qinv); free(r->bred); free(r->ninv); free(r->psi); free(r->psiinv); free(r->rescale);
    free(r);
}
void ork_ring_set_tables(ork_ring *r, int i, const uint64_t *psi, const uint64_t *psiinv, uint64_t ninv) {
    memcpy(r->psi + (size_t)i * r->N, psi, sizeof(uint64_t) * r->N);
    memcpy(r->psiinv + (size_t)i * r->N, psiinv, sizeof(uint64_t) * r->N);
    r->ninv[i] = ninv;
}

/* ======================================================================================
 * lattigo ring/ntt.go : NTTLazy + final BRedAdd ; InvNTTLazy + final MRed / MRedConstant
 * (call sites keyswitch.go:29-30,58,114-115,183-206; keyswitch_hoisted.go:36-37,120-143)
 * ====================================================================================== */
static int bitlen64(uint64_t x) { int n = 0; while (x) { n++; x >>= 1; } return n; }

static void ntt_lazy(const uint64_t *in, uint64_t *out, int N, const uint64_t *psi, uint64_t Q, uint64_t qinv) {
    uint64_t fourQ = 4 * Q, twoQ = 2 * Q;
    int t = N >> 1;
    uint64_t F = psi[1];
    for (int j = 0; j < t; j++) {
        uint64_t U = in[j];
        uint64_t V = mred_const(in[j + t], F, Q, qinv);
        out[j] = U + V;
        out[j + t] = U + twoQ - V;
    }
    for (int m = 2; m < N; m <<= 1) {
        int reduce = (bitlen64((uint64_t)m) & 1) == 1;
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int j1 = (i * t) << 1;
            F = psi[m + i];
            for (int j = j1; j < j1 + t; j++) {
                uint64_t U = out[j];
                if (reduce && U >= fourQ) U -= fourQ;
                uint64_t V = mred_const(out[j + t], F, Q, qinv);
                out[j] = U + V;
                out[j + t] = U + twoQ - V;
            }
        }
    }
}
static void ntt_one(const ork_ring *r, int i, const uint64_t *in, uint64_t *out) {
    int N = r->N;
    uint64_t Q = r->q[i];
    ntt_lazy(in, out, N, r->psi + (size_t)i * N, Q, r->qinv[i]);
    const uint64_t *u = &r->bred[2 * i];
    for (int j = 0; j < N; j++) out[j] = bred_add(out[j], Q, u);
}
static void intt_core(const ork_ring *r, int i, const uint64_t *in, uint64_t *out, int lazy) {
    int N = r->N;
    uint64_t Q = r->q[i], qinv = r->qinv[i], twoQ = Q << 1, fourQ = Q << 2;
    const uint64_t *psiinv = r->psiinv + (size_t)i * N;
    int h = N >> 1, t = 1;
    for (int k = 0; k < h; k++) {
        uint64_t U = in[2 * k], V = in[2 * k + 1];
        uint64_t X = U + V;
        if (X >= twoQ) X -= twoQ;
        out[2 * k] = X;
        out[2 * k + 1] = mred_const(U + fourQ - V, psiinv[h + k], Q, qinv);
    }
    t <<= 1;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        for (int g = 0, j1 = 0; g < h; g++, j1 += 2 * t) {
            uint64_t F = psiinv[h + g];
            for (int j = j1; j < j1 + t; j++) {
                uint64_t U = out[j], V = out[j + t];
                uint64_t X = U + V;
                if (X >= twoQ) X -= twoQ;
                out[j] = X;
                out[j + t] = mred_const(U + fourQ - V, F, Q, qinv);
            }
        }
        t <<= 1;
    }
    uint64_t ninv = r->ninv[i];
    if (lazy) for (int j = 0; j < N; j++) out[j] = mred_const(out[j], ninv, Q, qinv);
    else      for (int j = 0; j < N; j++) out[j] = mred(out[j], ninv, Q, qinv);
}
void ork_ntt_single(const ork_ring *r, int mod, const uint64_t *in, uint64_t *out) { ntt_one(r, mod, in, out); }
void ork_ntt_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out) {
    PAR_FOR
    for (int i = 0; i <= level; i++) ntt_one(r, i, in + (size_t)i * r->N, out + (size_t)i * r->N);
}
void ork_intt_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out) {
    PAR_FOR
    for (int i = 0; i <= level; i++) intt_core(r, i, in + (size_t)i * r->N, out + (size_t)i * r->N, 0);
}
void ork_intt_lazy_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out) {
    PAR_FOR
    for (int i = 0; i <= level; i++) intt_core(r, i, in + (size_t)i * r->N, out + (size_t)i * r->N, 1);
}

/* ======================================================================================
 * lattigo ring/ring_operations.go (call sites keyswitch_hoisted.go:28-30,84-138,153)
 * ====================================================================================== */
#define LIMB_LOOP(body)                                                        \
    PAR_FOR                                                                    \
    for (int i = 0; i <= level; i++) {                                         \
        const uint64_t Q = r->q[i], qinv = r->qinv[i];                         \
        const uint64_t *u = &r->bred[2 * i];                                   \
        const size_t off = (size_t)i * r->N;                                   \
        (void)qinv; (void)u; (void)Q;                                          \
        for (int j = 0; j < r->N; j++) { body; }                               \
    }

void ork_mform_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out) {
    LIMB_LOOP(out[off + j] = mform(in[off + j], Q, u))
}
void ork_invmform_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out) {
    LIMB_LOOP(out[off + j] = mred(in[off + j], 1, Q, qinv))
}
void ork_mul_mont_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    LIMB_LOOP(out[off + j] = mred(a[off + j], b[off + j], Q, qinv))
}
void ork_mul_mont_add_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    LIMB_LOOP(out[off + j] = cred(out[off + j] + mred(a[off + j], b[off + j], Q, qinv), Q))
}
void ork_mul_mont_sub_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    LIMB_LOOP(out[off + j] = cred(out[off + j] + (Q - mred(a[off + j], b[off + j], Q, qinv)), Q))
}
void ork_add_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    LIMB_LOOP(out[off + j] = cred(a[off + j] + b[off + j], Q))
}
void ork_sub_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    LIMB_LOOP(out[off + j] = cred(a[off + j] + Q - b[off + j], Q))
}
void ork_neg_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t *out) {
    LIMB_LOOP(out[off + j] = Q - a[off + j])   /* lattigo Neg: q - x (0 -> q), not reduced */
}
void ork_reduce_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t *out) {
    LIMB_LOOP(out[off + j] = bred_add(a[off + j], Q, u))
}
/* MulScalar: MRed(x, MForm(BRedAdd(scalar))) */
void ork_mul_scalar_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t scalar, uint64_t *out) {
    PAR_FOR
    for (int i = 0; i <= level; i++) {
        uint64_t Q = r->q[i], qinv = r->qinv[i];
        const uint64_t *u = &r->bred[2 * i];
        uint64_t s = mform(bred_add(scalar, Q, u), Q, u);
        size_t off = (size_t)i * r->N;
        for (int j = 0; j < r->N; j++) out[off + j] = mred(a[off + j], s, Q, qinv);
    }
}
void ork_mul_residues_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *residues, uint64_t *out) {
    PAR_FOR
    for (int i = 0; i <= level; i++) {
        uint64_t Q = r->q[i], qinv = r->qinv[i];
        const uint64_t *u = &r->bred[2 * i];
        uint64_t s = mform(residues[i] % Q, Q, u);
        size_t off = (size_t)i * r->N;
        for (int j = 0; j < r->N; j++) out[off + j] = mred(a[off + j], s, Q, qinv);
    }
}

/* ring.Permute == the inline loop at keyswitch.go:268-296 / keyswitch_hoisted.go:217-245.
 * Negated positions store q - c WITHOUT reduction (c == 0 becomes q). `in` and `out` must differ. */
void ork_permute(const ork_ring *r, int level, const uint64_t *in, uint64_t galEl, uint64_t *out) {
    uint64_t mask = (uint64_t)r->N - 1;
    int logN = r->logN;
    for (uint64_t i = 0; i < (uint64_t)r->N; i++) {
        uint64_t indexRaw = i * galEl;
        uint64_t index = indexRaw & mask;
        uint64_t tmp = (indexRaw >> logN) & 1;
        for (int j = 0; j <= level; j++) {
            uint64_t qi = r->q[j];
            uint64_t c = in[(size_t)j * r->N + i];
            out[(size_t)j * r->N + index] = c * (tmp ^ 1) | (qi - c) * tmp;
        }
    }
}
/* ring.PermuteNTTIndex + PermuteNTTWithIndexLvl (keygen.go:212-214,243-245) */
void ork_permute_ntt(const ork_ring *r, int level, const uint64_t *in, uint64_t galEl, uint64_t *out) {
    uint64_t N = (uint64_t)r->N, mask = (N << 1) - 1;
    int logN = r->logN;
    for (uint64_t i = 0; i < N; i++) {
        uint64_t tmp1 = 2 * bitrev(i, logN) + 1;
        uint64_t tmp2 = (((galEl * tmp1) & mask) - 1) >> 1;
        uint64_t idx = bitrev(tmp2, logN);
        for (int j = 0; j <= level; j++) out[(size_t)j * N + i] = in[(size_t)j * N + idx];
    }
}
/* rlwe.Parameters.GaloisElementForColumnRotationBy: 5^(k mod 2N) mod 2N (keyswitch_hoisted.go:216) */
uint64_t ork_galois_element_for_rotation(int logN, int k) {
    uint64_t twoN = (uint64_t)1 << (logN + 1);
    uint64_t kRed = (uint64_t)((int64_t)k) & (twoN - 1);
    uint64_t r = 1, b = 5;
    while (kRed) { if (kRed & 1) r = (r * b) & (twoN - 1); b = (b * b) & (twoN - 1); kRed >>= 1; }
    return r;
}

/* lattigo ring/scaling.go DivRoundByLastModulus[Many]Lvl (call site mkckks/evaluator.go:388).
 * NOTE: adds (q_l-1)/2 to the input's last limb IN PLACE. */
static void div_round_by_last_modulus(const ork_ring *r, int level, uint64_t *p0, uint64_t *p1) {
    int N = r->N;
    uint64_t pj = r->q[level];
    uint64_t pHalf = (pj - 1) >> 1;
    uint64_t *last = p0 + (size_t)level * N;
    for (int j = 0; j < N; j++) last[j] = cred(last[j] + pHalf, pj);
    PAR_FOR
    for (int i = 0; i < level; i++) {
        uint64_t qi = r->q[i], qinv = r->qinv[i], twoqi = qi << 1;
        const uint64_t *u = &r->bred[2 * i];
        uint64_t rescaleParams = qi - r->rescale[(size_t)(level - 1) * r->nmod + i];
        uint64_t pHalfNegQi = qi - bred_add(pHalf, qi, u);
        const uint64_t *src = p0 + (size_t)i * N;
        uint64_t *dst = p1 + (size_t)i * N;
        for (int j = 0; j < N; j++) {
            /* (x_last' - h - x_i) * (-(q_l^-1)) = (x_i + h - x_last') * q_l^-1 mod q_i */
            uint64_t t = bred_add(last[j], qi, u);
            dst[j] = mred(t + pHalfNegQi + twoqi - src[j], rescaleParams, qi, qinv);
        }
    }
}
void ork_div_round_by_last_modulus_many(const ork_ring *r, int level, int nb, uint64_t *p0, uint64_t *p1) {
    int N = r->N;
    if (nb == 0) {
        if (p0 != p1) memcpy(p1, p0, sizeof(uint64_t) * (size_t)(level + 1) * N);
        return;
    }
    if (nb == 1) { div_round_by_last_modulus(r, level, p0, p1); return; }
    uint64_t *pool = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)(level + 1) * N);
    div_round_by_last_modulus(r, level, p0, pool);
    for (int i = 1; i < nb; i++) div_round_by_last_modulus(r, level - i, pool, pool);
    memcpy(p1, pool, sizeof(uint64_t) * (size_t)(level + 1 - nb) * N);
    free(pool);
}

/* ======================================================================================
 * samplers -- own PRNG (xoshiro256**), documented draw order; NOT lattigo's Blake2b XOF
 * (the reference cannot be seeded anyway: params.go:28, keygen.go:26, encryptor.go:32)
 * ====================================================================================== */
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void ork_prng_seed(ork_prng *p, uint64_t seed) {
    for (int i = 0; i < 4; i++) {
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        p->s[i] = z ^ (z >> 31);
    }
}
uint64_t ork_prng_next(ork_prng *p) {
    uint64_t *s = p->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
/* uniform in [0,q): masked rejection sampling, limb by limb, coefficient by coefficient */
void ork_sample_uniform(ork_prng *p, const ork_ring *r, int level, uint64_t *out) {
    for (int i = 0; i <= level; i++) {
        uint64_t q = r->q[i];
        uint64_t mask = ((uint64_t)1 << bitlen64(q)) - 1;
        for (int j = 0; j < r->N; j++) {
            uint64_t x;
            do { x = ork_prng_next(p) & mask; } while (x >= q);
            out[(size_t)i * r->N + j] = x;
        }
    }
}
static double prng_unit(ork_prng *p) { return (double)(ork_prng_next(p) >> 11) * (1.0 / 9007199254740992.0); }
/* ternary: P(0)=pzero, P(+1)=P(-1)=(1-pzero)/2  (keygen.go:58-60: p = 1/2) */
void ork_sample_ternary(ork_prng *p, int N, double pzero, int64_t *out) {
    for (int j = 0; j < N; j++) {
        double x = prng_unit(p);
        if (x < pzero) out[j] = 0;
        else out[j] = (ork_prng_next(p) & 1) ? 1 : -1;
    }
}
/* rounded Gaussian, sigma (3.2), truncated at |x| <= bound (19) (keygen.go:35) */
void ork_sample_gaussian(ork_prng *p, int N, double sigma, int bound, int64_t *out) {
    for (int j = 0; j < N; j++) {
        for (;;) {
            double u1 = prng_unit(p), u2 = prng_unit(p);
            if (u1 <= 0.0) continue;
            double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
            long v = lround(z * sigma);
            if (v > bound || v < -bound) continue;
            out[j] = v;
            break;
        }
    }
}
/* ---- the counter-based samplers of include/mkhe_prng.h ("mkhe-ctr-1"): what the device-side key generation and encryption
 * draw from; the oracle's KeyGenerator / Encryptor run unchanged on top of them (oracle.CtrPRNG), so a key made on the GPU can be
 * compared with the oracle's bit for bit.  Limb i of a uniform poly is stream `stream + i`. */
void ork_ctr_uniform(uint64_t seed, uint64_t stream, const ork_ring *r, int level, uint64_t *out) {
    for (int i = 0; i <= level; i++)
        for (int j = 0; j < r->N; j++) out[(size_t)i * r->N + j] = mkhe_sample_uniform(seed, stream + (uint64_t)i, (uint64_t)j, r->q[i]);
}
void ork_ctr_ternary(uint64_t seed, uint64_t stream, int N, uint64_t thr53, int64_t *out) {
    for (int j = 0; j < N; j++) out[j] = mkhe_sample_ternary(seed, stream, (uint64_t)j, thr53);
}
void ork_ctr_gaussian(uint64_t seed, uint64_t stream, int N, int64_t *out) {
    for (int j = 0; j < N; j++) out[j] = mkhe_sample_gaussian(seed, stream, (uint64_t)j);
}
/* write a small signed vector into every limb (sampler ReadLvl / ExtendBasisSmallNormAndCenter) */
void ork_lift_small(const ork_ring *r, int level, const int64_t *small, uint64_t *out) {
    for (int i = 0; i <= level; i++) {
        uint64_t q = r->q[i];
        for (int j = 0; j < r->N; j++) {
            int64_t v = small[j];
            out[(size_t)i * r->N + j] = v >= 0 ? (uint64_t)v : q - (uint64_t)(-v);
        }
    }
}

/* ======================================================================================
 * mkrlwe/basis_extension.go
 * ====================================================================================== */

/* basisextenderparameters, basis_extension.go:83-153 */
static void modup_params_gen(ork_modup_params *mp, const uint64_t *Q, int nq, const uint64_t *P, int np) {
    mp->nq = nq; mp->np = np;
    mp->qoverqiinvqi = (uint64_t *)xmalloc(sizeof(uint64_t) * nq);
    mp->qoverqimodp = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)np * nq);
    mp->vtimesqmodp = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)np * (nq + 1));
    for (int i = 0; i < nq; i++) {
        uint64_t qi = Q[i], bp[2];
        bred_params(qi, bp);
        /* (Q/qi) mod qi, then inverse, Montgomery form (:110-119) */
        uint64_t star = 1;
        for (int j = 0; j < nq; j++) if (j != i) star = mulmod(star, Q[j] % qi, qi);
        mp->qoverqiinvqi[i] = mform(ork_modexp(star, qi - 2, qi), qi, bp);
        for (int j = 0; j < np; j++) {
            uint64_t pj = P[j], bpp[2];
            bred_params(pj, bpp);
            uint64_t s = 1;
            for (int u = 0; u < nq; u++) if (u != i) s = mulmod(s, Q[u] % pj, pj);
            mp->qoverqimodp[(size_t)j * nq + i] = mform(s, pj, bpp);      /* (:121-131) */
        }
    }
    for (int j = 0; j < np; j++) {
        uint64_t pj = P[j];
        uint64_t QmodP = 1;
        for (int i = 0; i < nq; i++) QmodP = mulmod(QmodP, Q[i] % pj, pj);
        uint64_t v = pj - QmodP;
        uint64_t *vt = mp->vtimesqmodp + (size_t)j * (nq + 1);
        vt[0] = 0;
        for (int i = 1; i < nq + 1; i++) vt[i] = cred(vt[i - 1] + v, pj);    /* (:134-150) */
    }
}
static void modup_params_free(ork_modup_params *mp) {
    free(mp->qoverqiinvqi); free(mp->qoverqimodp); free(mp->vtimesqmodp);
}
/* genModDownParams, basis_extension.go:34-54: params[j][i] = prod_{u<=j} P_u^-1 mod q_i (Montgomery) */
static uint64_t *moddown_params_gen(const ork_ring *ringQ, const ork_ring *ringP) {
    int nq = ringQ->nmod, np = ringP->nmod;
    uint64_t *params = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)np * nq);
    for (int j = 0; j < np; j++)
        for (int i = 0; i < nq; i++) {
            uint64_t qi = ringQ->q[i];
            uint64_t v = ork_modexp(ringP->q[j] % qi, qi - 2, qi);
            v = mform(v, qi, &ringQ->bred[2 * i]);
            if (j > 0) v = mred(v, params[(size_t)(j - 1) * nq + i], qi, ringQ->qinv[i]);
            params[(size_t)j * nq + i] = v;
        }
    return params;
}
/* NewFastBasisExtender, basis_extension.go:57-81 */
ork_basis_extender *ork_be_new(const ork_ring *ringQ, const ork_ring *ringP) {
    ork_basis_extender *be = (ork_basis_extender *)xcalloc(1, sizeof(*be));
    be->ringQ = ringQ; be->ringP = ringP;
    be->paramsQtoP = (ork_modup_params *)xcalloc(ringQ->nmod, sizeof(ork_modup_params));
    for (int i = 0; i < ringQ->nmod; i++) modup_params_gen(&be->paramsQtoP[i], ringQ->q, i + 1, ringP->q, ringP->nmod);
    be->paramsPtoQ = (ork_modup_params *)xcalloc(ringP->nmod, sizeof(ork_modup_params));
    for (int i = 0; i < ringP->nmod; i++) modup_params_gen(&be->paramsPtoQ[i], ringP->q, i + 1, ringQ->q, ringQ->nmod);
    be->modDownPtoQ = moddown_params_gen(ringQ, ringP);
    be->modDownQtoP = moddown_params_gen(ringP, ringQ);
    return be;
}
void ork_be_free(ork_basis_extender *be) {
    if (!be) return;
    for (int i = 0; i < be->ringQ->nmod; i++) modup_params_free(&be->paramsQtoP[i]);
    for (int i = 0; i < be->ringP->nmod; i++) modup_params_free(&be->paramsPtoQ[i]);
    free(be->paramsQtoP); free(be->paramsPtoQ); free(be->modDownPtoQ); free(be->modDownQtoP);
    free(be);
}

/* modUpExact = reconstructRNS + multSum, basis_extension.go:337-357,537-646 (one coefficient at a time;
 * the reference's 8-wide unrolling does not change any value).  Output limbs are LAZY: [0, ~3p). */
static void modup_exact(const uint64_t *p1, int n1, uint64_t *p2, int n2, int N,
                        const ork_ring *ringQ, const ork_ring *ringP, const ork_modup_params *mp) {
    PAR_FOR
    for (int x = 0; x < N; x++) {
        uint64_t y[64];
        double vi = 0.0;
        for (int i = 0; i < n1; i++) {                                       /* reconstructRNS :543-569 */
            uint64_t qi = ringQ->q[i];
            y[i] = mred(p1[(size_t)i * N + x], mp->qoverqiinvqi[i], qi, ringQ->qinv[i]);
            vi += (double)y[i] / (double)qi;
        }
        uint64_t v = (uint64_t)vi;                                           /* :571 */
        for (int j = 0; j < n2; j++) {                                       /* multSum :582-646 */
            uint64_t pj = ringP->q[j], qInv = ringP->qinv[j];
            const uint64_t *qoq = mp->qoverqimodp + (size_t)j * mp->nq;
            const uint64_t *vt = mp->vtimesqmodp + (size_t)j * (mp->nq + 1);
            uint64_t rlo = 0, rhi = 0;
            for (int i = 0; i < n1; i++) {
                u128 m = (u128)y[i] * qoq[i];
                uint64_t mhi = (uint64_t)(m >> 64), mlo = (uint64_t)m;
                uint64_t s = rlo + mlo;
                uint64_t c = s < rlo;
                rlo = s;
                rhi += mhi + c;
            }
            uint64_t hhi = mulhi64(rlo * qInv, pj);
            p2[(size_t)j * N + x] = rhi - hhi + pj + vt[v];
        }
    }
}
/* ModUpQtoP :177-179 */
void ork_be_modup_q_to_p(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *polQ, uint64_t *polP) {
    modup_exact(polQ, levelQ + 1, polP, levelP + 1, be->ringQ->N, be->ringQ, be->ringP, &be->paramsQtoP[levelQ]);
}
/* ModUpPtoQ :184-186 */
void ork_be_modup_p_to_q(const ork_basis_extender *be, int levelP, int levelQ, const uint64_t *polP, uint64_t *polQ) {
    modup_exact(polP, levelP + 1, polQ, levelQ + 1, be->ringQ->N, be->ringP, be->ringQ, &be->paramsPtoQ[levelP]);
}
/* ModDownQPtoQ :192-232 (the lattigo twin behind ks.Baseconverter, keyswitch_hoisted.go:39, is the same algorithm) */
void ork_be_moddown_qp_to_q(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, const uint64_t *p1P, uint64_t *p2Q) {
    const ork_ring *ringQ = be->ringQ;
    int N = ringQ->N;
    uint64_t *pool = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)(levelQ + 1) * N);
    ork_be_modup_p_to_q(be, levelP, levelQ, p1P, pool);
    PAR_FOR
    for (int i = 0; i <= levelQ; i++) {
        uint64_t qi = ringQ->q[i], twoqi = qi << 1;
        uint64_t params = qi - be->modDownPtoQ[(size_t)levelP * ringQ->nmod + i];
        uint64_t qinv = ringQ->qinv[i];
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            p2Q[o] = mred(pool[o] + twoqi - p1Q[o], params, qi, qinv);
        }
    }
    free(pool);
}
/* ModDownQPtoQNTT :239-290: inputs in the NTT domain; p1P is overwritten by its lazy inverse transform like in the reference */
void ork_be_moddown_qp_to_q_ntt(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, uint64_t *p1P, uint64_t *p2Q) {
    const ork_ring *ringQ = be->ringQ;
    int N = ringQ->N;
    uint64_t *pool = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)(levelQ + 1) * N);
    ork_intt_lazy_lvl(be->ringP, levelP, p1P, p1P);                          /* :247 */
    ork_be_modup_p_to_q(be, levelP, levelQ, p1P, pool);                      /* :251 */
    PAR_FOR
    for (int i = 0; i <= levelQ; i++) {
        uint64_t qi = ringQ->q[i], twoqi = qi << 1;
        uint64_t params = qi - be->modDownPtoQ[(size_t)levelP * ringQ->nmod + i];
        uint64_t qinv = ringQ->qinv[i];
        uint64_t *p3 = pool + (size_t)i * N;
        ntt_lazy(p3, p3, N, ringQ->psi + (size_t)i * N, qi, qinv);           /* :268, in place like ring.NTTLazy(p3tmp, p3tmp, ..) */
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            p2Q[o] = mred(pool[o] + twoqi - p1Q[o], params, qi, qinv);       /* :277-284 */
        }
    }
    free(pool);
}
/* ModDownQPtoP :294-334.  NOTE the reference indexes modDownParams[levelP][i] (not [levelQ][i]); kept. */
void ork_be_moddown_qp_to_p(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, const uint64_t *p1P, uint64_t *p2P) {
    const ork_ring *ringP = be->ringP;
    int N = ringP->N;
    uint64_t *pool = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)(levelP + 1) * N);
    ork_be_modup_q_to_p(be, levelQ, levelP, p1Q, pool);
    PAR_FOR
    for (int i = 0; i <= levelP; i++) {
        uint64_t qi = ringP->q[i], twoqi = qi << 1;
        uint64_t params = qi - be->modDownQtoP[(size_t)levelP * ringP->nmod + i];
        uint64_t qinv = ringP->qinv[i];
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)i * N + j;
            p2P[o] = mred(pool[o] + twoqi - p1P[o], params, qi, qinv);
        }
    }
    free(pool);
}

/* ======================================================================================
 * mkrlwe/keyswitch.go + keyswitch_hoisted.go
 * ====================================================================================== */
#define SWK_DIGIT(ks, p, i) ((p) + (size_t)(i) * ((ks)->nQ + (ks)->nP) * (ks)->N)
#define SWK_Q(ks, p, i) SWK_DIGIT(ks, p, i)
#define SWK_P(ks, p, i) (SWK_DIGIT(ks, p, i) + (size_t)(ks)->nQ * (ks)->N)

/* NewKeySwitcher keyswitch.go:33-47 */
ork_keyswitcher *ork_ks_new(const ork_ring *ringQ, const ork_ring *ringP, int gamma) {
    ork_keyswitcher *ks = (ork_keyswitcher *)xcalloc(1, sizeof(*ks));
    ks->ringQ = ringQ; ks->ringP = ringP;
    ks->nQ = ringQ->nmod; ks->nP = ringP->nmod; ks->gamma = gamma; ks->N = ringQ->N;
    ks->alpha = ks->nP / gamma;                                    /* params.go:63-65 */
    if (ks->alpha < 1) { fprintf(stderr, "mkhe_oracle: alpha = #P/gamma must be >= 1\n"); abort(); }
    ks->be = ork_be_new(ringQ, ringP);
    size_t swk = (size_t)ks->nQ * (ks->nQ + ks->nP) * ks->N;       /* beta_max = nQ digits */
    ks->swkPool1 = (uint64_t *)xcalloc(swk, 8);
    ks->swkPool2 = (uint64_t *)xcalloc(swk, 8);
    ks->swkPool3 = (uint64_t *)xcalloc(swk, 8);
    for (int i = 0; i < 3; i++) ks->polyQPool[i] = (uint64_t *)xcalloc((size_t)ks->nQ * ks->N, 8);
    ks->poolQP0 = (uint64_t *)xcalloc((size_t)(ks->nQ + ks->nP) * ks->N, 8);
    ks->poolQP1 = (uint64_t *)xcalloc((size_t)(ks->nQ + ks->nP) * ks->N, 8);
    return ks;
}
void ork_ks_free(ork_keyswitcher *ks) {
    if (!ks) return;
    ork_be_free(ks->be);
    free(ks->swkPool1); free(ks->swkPool2); free(ks->swkPool3);
    for (int i = 0; i < 3; i++) free(ks->polyQPool[i]);
    free(ks->poolQP0); free(ks->poolQP1);
    free(ks);
}
/* Beta, params.go:67-71 */
int ork_ks_beta(const ork_keyswitcher *ks, int levelQ) {
    return (int)ceil((double)(levelQ + 1) / (double)ks->alpha);
}

/* DecomposeSingleNTT keyswitch.go:21-31 with DecomposeAndSplit's alpha=1 copy branch
 * (basis_extension.go:443-451): the digit limb is copied UNREDUCED into every Q limb <= levelQ and
 * every P limb, then NTT'd (the lazy NTT + final BRedAdd absorb the unreduced input). */
void ork_ks_decompose_single_ntt(ork_keyswitcher *ks, int levelQ, const uint64_t *digitLimb, uint64_t *outQP) {
    int N = ks->N, levelP = ks->nP - 1;
    uint64_t *outQ = outQP, *outP = outQP + (size_t)ks->nQ * N;
    for (int j = 0; j <= levelQ; j++) memcpy(outQ + (size_t)j * N, digitLimb, sizeof(uint64_t) * N);
    for (int j = 0; j <= levelP; j++) memcpy(outP + (size_t)j * N, digitLimb, sizeof(uint64_t) * N);
    ork_ntt_lvl(ks->ringQ, levelQ, outQ, outQ);
    ork_ntt_lvl(ks->ringP, levelP, outP, outP);
}
/* DecomposeAndSplit, general branch (basis_extension.go:428-535) for digit `beta` when alpha > 1: the digit's limbs
 * [alpha*beta, alpha*beta + decompLvl + 2) are lifted exactly (reconstructRNS + multSum with the fp64 overflow estimate)
 * to every Q limb <= levelQ and every P limb.  Like the reference, the loop over "greater" limbs starts at alpha*beta
 * (:525), so the digit's own limbs are overwritten by their (congruent, lazy) multSum values as well. */
static void decompose_and_split_general(ork_keyswitcher *ks, int levelQ, int beta, int decompLvl, const uint64_t *a, uint64_t *outQP) {
    const ork_ring *ringQ = ks->ringQ, *ringP = ks->ringP;
    int N = ks->N, nQ = ks->nQ, nP = ks->nP, alpha = ks->alpha, levelP = nP - 1;
    int lvlQStart = beta * alpha, nsrc = decompLvl + 2;
    /* modUpParams[gamma*alpha-2][beta][decompLvl] = basisextenderparameters(Q[beta*alpha .. +nsrc), Q u P)  (:392-416) */
    uint64_t *Pi = (uint64_t *)xmalloc(sizeof(uint64_t) * (size_t)(nQ + nP));
    for (int k = 0; k < nQ; k++) Pi[k] = ringQ->q[k];
    for (int k = 0; k < nP; k++) Pi[nQ + k] = ringP->q[k];
    ork_modup_params mp;
    modup_params_gen(&mp, ringQ->q + lvlQStart, nsrc, Pi, nQ + nP);
    uint64_t *outQ = outQP, *outP = outQP + (size_t)nQ * N;
    PAR_FOR
    for (int x = 0; x < N; x++) {
        uint64_t y[64];
        double vi = 0.0;
        for (int i = 0, j = lvlQStart; i < nsrc; i++, j++) {                 /* :471-503 */
            uint64_t qi = ringQ->q[j];
            uint64_t px = a[(size_t)j * N + x];
            outQ[(size_t)j * N + x] = px;
            y[i] = mred(px, mp.qoverqiinvqi[i], qi, ringQ->qinv[j]);
            vi += (double)y[i] / (double)qi;
        }
        uint64_t v = (uint64_t)vi;                                           /* :506-513 */
        for (int u = 0; u < nQ + nP; u++) {
            int isP = u >= nQ, j = isP ? u - nQ : u;
            if (!isP && (j > levelQ || (j >= lvlQStart && j < alpha * beta))) continue;   /* :516-527 (second range is empty) */
            if (isP && j > levelP) continue;
            uint64_t pj = isP ? ringP->q[j] : ringQ->q[j], qInv = isP ? ringP->qinv[j] : ringQ->qinv[j];
            const uint64_t *qoq = mp.qoverqimodp + (size_t)u * mp.nq;
            const uint64_t *vt = mp.vtimesqmodp + (size_t)u * (mp.nq + 1);
            uint64_t rlo = 0, rhi = 0;
            for (int i = 0; i < nsrc; i++) {                                 /* multSum :582-646 */
                u128 m = (u128)y[i] * qoq[i];
                uint64_t mhi = (uint64_t)(m >> 64), mlo = (uint64_t)m;
                uint64_t s2 = rlo + mlo;
                uint64_t c = s2 < rlo;
                rlo = s2;
                rhi += mhi + c;
            }
            uint64_t hhi = mulhi64(rlo * qInv, pj);
            (isP ? outP : outQ)[(size_t)j * N + x] = rhi - hhi + pj + vt[v];
        }
    }
    modup_params_free(&mp);
    free(Pi);
}
/* DecomposeSingleNTT keyswitch.go:21-31 for any alpha: `a` is the whole coefficient-domain poly */
static void decompose_single_ntt_any(ork_keyswitcher *ks, int levelQ, int beta, const uint64_t *a, uint64_t *outQP) {
    int alpha = ks->alpha, levelP = ks->nP - 1;
    int decompLvl = levelQ > alpha * (beta + 1) - 1 ? alpha - 2 : (levelQ % alpha) - 1;    /* :435-440 */
    if (decompLvl == -1) { ork_ks_decompose_single_ntt(ks, levelQ, a + (size_t)beta * alpha * ks->N, outQP); return; }
    decompose_and_split_general(ks, levelQ, beta, decompLvl, a, outQP);
    ork_ntt_lvl(ks->ringQ, levelQ, outQP, outQP);
    ork_ntt_lvl(ks->ringP, levelP, outQP + (size_t)ks->nQ * ks->N, outQP + (size_t)ks->nQ * ks->N);
}
/* Decompose keyswitch.go:49-73 (a.IsNTT == false branch: the only one the CKKS/BFV flows take) */
void ork_ks_decompose(ork_keyswitcher *ks, int levelQ, const uint64_t *a, uint64_t *ad) {
    int beta = ork_ks_beta(ks, levelQ);
    for (int i = 0; i < beta; i++)
        decompose_single_ntt_any(ks, levelQ, i, a, SWK_DIGIT(ks, ad, i));
}
static void qp_mul_mont(ork_keyswitcher *ks, int levelQ, int levelP, const uint64_t *a, const uint64_t *b, uint64_t *c, int add) {
    size_t po = (size_t)ks->nQ * ks->N;
    if (add) {
        ork_mul_mont_add_lvl(ks->ringQ, levelQ, a, b, c);
        ork_mul_mont_add_lvl(ks->ringP, levelP, a + po, b + po, c + po);
    } else {
        ork_mul_mont_lvl(ks->ringQ, levelQ, a, b, c);
        ork_mul_mont_lvl(ks->ringP, levelP, a + po, b + po, c + po);
    }
}
static void qp_mform(ork_keyswitcher *ks, int levelQ, int levelP, uint64_t *a) {
    size_t po = (size_t)ks->nQ * ks->N;
    ork_mform_lvl(ks->ringQ, levelQ, a, a);
    ork_mform_lvl(ks->ringP, levelP, a + po, a + po);
}
static void ks_finish_external_product(ork_keyswitcher *ks, int levelQ, uint64_t *c1QP, uint64_t *c) {
    int levelP = ks->nP - 1;
    uint64_t *c1Q = c1QP, *c1P = c1QP + (size_t)ks->nQ * ks->N;
    ork_intt_lazy_lvl(ks->ringQ, levelQ, c1Q, c1Q);
    ork_intt_lazy_lvl(ks->ringP, levelP, c1P, c1P);
    ork_be_moddown_qp_to_q(ks->be, levelQ, levelP, c1Q, c1P, c);
}
/* ExternalProductHoisted keyswitch_hoisted.go:10-40 */
void ork_ks_external_product_hoisted(ork_keyswitcher *ks, int levelQ, const uint64_t *aHoisted, const uint64_t *bg, uint64_t *c) {
    int levelP = ks->nP - 1, beta = ork_ks_beta(ks, levelQ);
    uint64_t *c1QP = ks->poolQP1;
    for (int i = 0; i < beta; i++)
        qp_mul_mont(ks, levelQ, levelP, SWK_DIGIT(ks, bg, i), SWK_DIGIT(ks, aHoisted, i), c1QP, i != 0);
    ks_finish_external_product(ks, levelQ, c1QP, c);
}
/* ExternalProduct keyswitch.go:79-118 */
void ork_ks_external_product(ork_keyswitcher *ks, int levelQ, const uint64_t *a, const uint64_t *bg, uint64_t *c) {
    int levelP = ks->nP - 1, beta = ork_ks_beta(ks, levelQ);
    uint64_t *c0QP = ks->poolQP0, *c1QP = ks->poolQP1;
    for (int i = 0; i < beta; i++) {
        decompose_single_ntt_any(ks, levelQ, i, a, c0QP);
        qp_mul_mont(ks, levelQ, levelP, SWK_DIGIT(ks, bg, i), c0QP, c1QP, i != 0);
    }
    ks_finish_external_product(ks, levelQ, c1QP, c);
}

static int find_id(int n, const int *ids, int id) {
    for (int i = 0; i < n; i++) if (ids[i] == id) return i;
    return -1;
}

/* MulAndRelinHoisted keyswitch_hoisted.go:44-179.  h0 / h1 == NULL selects the reference's
 * `op*Hoisted == nil` branches (:80-85,100-105,148-149,165-166). */
void ork_ks_mul_and_relin_hoisted(ork_keyswitcher *ks, int level,
        int n0, const int *ids0, uint64_t *const *op0, uint64_t *const *h0,
        int n1, const int *ids1, uint64_t *const *op1, uint64_t *const *h1,
        uint64_t *const *rlk_b, uint64_t *const *rlk_d, uint64_t *const *rlk_v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out) {
    const ork_ring *ringQ = ks->ringQ;
    int levelP = ks->nP - 1, beta = ork_ks_beta(ks, level), N = ks->N;
    size_t digit = (size_t)(ks->nQ + ks->nP) * N;
    uint64_t *x = ks->swkPool1, *y = ks->swkPool2;
    memset(x, 0, sizeof(uint64_t) * digit * beta);                                   /* :70-76 */
    memset(y, 0, sizeof(uint64_t) * digit * beta);

    for (int t = 0; t < n0; t++) {                                                    /* :79-92 */
        const uint64_t *hh;
        if (!h0) { ork_ks_decompose(ks, level, op0[1 + t], ks->swkPool3); hh = ks->swkPool3; }
        else hh = h0[t];
        const uint64_t *d = rlk_d[ids0[t]];
        for (int i = 0; i < beta; i++)
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, d, i), SWK_DIGIT(ks, hh, i), SWK_DIGIT(ks, x, i), 1);
    }
    for (int i = 0; i < beta; i++) qp_mform(ks, level, levelP, SWK_DIGIT(ks, x, i));   /* :94-96 */

    for (int t = 0; t < n1; t++) {                                                    /* :99-113 */
        const uint64_t *hh;
        if (!h1) { ork_ks_decompose(ks, level, op1[1 + t], ks->swkPool3); hh = ks->swkPool3; }
        else hh = h1[t];
        const uint64_t *b = rlk_b[ids1[t]];
        for (int i = 0; i < beta; i++)
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, b, i), SWK_DIGIT(ks, hh, i), SWK_DIGIT(ks, y, i), 1);
    }
    for (int i = 0; i < beta; i++) qp_mform(ks, level, levelP, SWK_DIGIT(ks, y, i));   /* :115-117 */

    uint64_t *pool0 = ks->polyQPool[0], *pool1 = ks->polyQPool[1], *pool2 = ks->polyQPool[2];
    ork_ntt_lvl(ringQ, level, op0[0], pool0);                                         /* :120-124 */
    ork_ntt_lvl(ringQ, level, op1[0], pool1);
    ork_mform_lvl(ringQ, level, pool0, pool0);
    ork_mul_mont_lvl(ringQ, level, pool0, pool1, out[0]);

    ork_mform_lvl(ringQ, level, pool1, pool1);                                        /* :127-140 */
    for (int t = 0; t < n0; t++) {
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids0[t])];
        ork_ntt_lvl(ringQ, level, op0[1 + t], pool2);
        ork_mul_mont_lvl(ringQ, level, pool1, pool2, o);
    }
    for (int t = 0; t < n1; t++) {
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids1[t])];
        ork_ntt_lvl(ringQ, level, op1[1 + t], pool2);
        if (find_id(n0, ids0, ids1[t]) >= 0) ork_mul_mont_add_lvl(ringQ, level, pool0, pool2, o);
        else ork_mul_mont_lvl(ringQ, level, pool0, pool2, o);
    }
    for (int t = 0; t <= nOut; t++) ork_intt_lvl(ringQ, level, out[t], out[t]);        /* :142-144 */

    for (int t = 0; t < n1; t++) {                                                    /* :147-154 */
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids1[t])];
        if (!h1) ork_ks_external_product(ks, level, op1[1 + t], x, pool0);
        else ork_ks_external_product_hoisted(ks, level, h1[t], x, pool0);
        ork_add_lvl(ringQ, level, o, pool0, o);
    }
    for (int t = 0; t < n0; t++) {                                                    /* :161-178 */
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids0[t])];
        const uint64_t *v = rlk_v[ids0[t]];
        if (!h0) ork_ks_external_product(ks, level, op0[1 + t], y, pool0);
        else ork_ks_external_product_hoisted(ks, level, h0[t], y, pool0);
        ork_ks_decompose(ks, level, pool0, ks->swkPool3);
        ork_ks_external_product_hoisted(ks, level, ks->swkPool3, v, pool1);
        ork_add_lvl(ringQ, level, out[0], pool1, out[0]);
        ork_ks_external_product_hoisted(ks, level, ks->swkPool3, u, pool2);
        ork_add_lvl(ringQ, level, o, pool2, o);
    }
}
/* MulAndRelin keyswitch.go:122-230: same arithmetic, decomposing on the fly. */
void ork_ks_mul_and_relin(ork_keyswitcher *ks, int level,
        int n0, const int *ids0, uint64_t *const *op0,
        int n1, const int *ids1, uint64_t *const *op1,
        uint64_t *const *rlk_b, uint64_t *const *rlk_d, uint64_t *const *rlk_v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out) {
    ork_ks_mul_and_relin_hoisted(ks, level, n0, ids0, op0, NULL, n1, ids1, op1, NULL,
                                 rlk_b, rlk_d, rlk_v, u, nOut, idsOut, out);
}

/* permutation epilogue shared by Rotate / RotateHoisted (keyswitch.go:266-296, keyswitch_hoisted.go:215-245) */
static void ks_permute_all(ork_keyswitcher *ks, int level, int n, uint64_t galEl, uint64_t *const *out) {
    for (int t = 0; t <= n; t++) {
        ork_permute(ks->ringQ, level, out[t], galEl, ks->polyQPool[0]);
        memcpy(out[t], ks->polyQPool[0], sizeof(uint64_t) * (size_t)(level + 1) * ks->N);
    }
}
/* RotateHoisted keyswitch_hoisted.go:183-247 */
void ork_ks_rotate_hoisted(ork_keyswitcher *ks, int level, int rotidx,
        int n, const int *ids, uint64_t *const *ctIn, uint64_t *const *hoisted,
        uint64_t *const *rk, const uint64_t *a, uint64_t *const *out) {
    (void)ids;
    int N = ks->N;
    while (rotidx < 0) rotidx += N / 2;                                              /* :196-198 */
    memcpy(out[0], ctIn[0], sizeof(uint64_t) * (size_t)(level + 1) * N);              /* :205 */
    for (int t = 0; t < n; t++) {                                                    /* :207-213 */
        ork_ks_external_product_hoisted(ks, level, hoisted[t], rk[t], ks->polyQPool[0]);
        ork_add_lvl(ks->ringQ, level, out[0], ks->polyQPool[0], out[0]);
        ork_ks_external_product_hoisted(ks, level, hoisted[t], a, out[1 + t]);
    }
    ks_permute_all(ks, level, n, ork_galois_element_for_rotation(ks->ringQ->logN, rotidx), out);
}
/* Rotate keyswitch.go:234-298 */
void ork_ks_rotate(ork_keyswitcher *ks, int level, int rotidx,
        int n, const int *ids, uint64_t *const *ctIn,
        uint64_t *const *rk, const uint64_t *a, uint64_t *const *out) {
    (void)ids;
    int N = ks->N;
    while (rotidx < 0) rotidx += N / 2;
    memcpy(out[0], ctIn[0], sizeof(uint64_t) * (size_t)(level + 1) * N);
    for (int t = 0; t < n; t++) {
        ork_ks_external_product(ks, level, ctIn[1 + t], rk[t], ks->polyQPool[0]);
        ork_add_lvl(ks->ringQ, level, out[0], ks->polyQPool[0], out[0]);
        ork_ks_external_product(ks, level, ctIn[1 + t], a, out[1 + t]);
    }
    ks_permute_all(ks, level, n, ork_galois_element_for_rotation(ks->ringQ->logN, rotidx), out);
}
/* Conjugate keyswitch.go:302-332 (galEl = 2N-1) */
void ork_ks_conjugate(ork_keyswitcher *ks, int level,
        int n, const int *ids, uint64_t *const *ctIn,
        uint64_t *const *ck, const uint64_t *a, uint64_t *const *out) {
    (void)ids;
    uint64_t galEl = ((uint64_t)ks->N << 1) - 1;
    for (int t = 0; t <= n; t++) ork_permute(ks->ringQ, level, ctIn[t], galEl, out[t]);      /* :315-317 */
    for (int t = 0; t < n; t++) {                                                          /* :320-324 */
        ork_ks_external_product(ks, level, out[1 + t], ck[t], ks->polyQPool[0]);
        ork_add_lvl(ks->ringQ, level, out[0], ks->polyQPool[0], out[0]);
    }
    for (int t = 0; t < n; t++) {                                                          /* :327-331 */
        ork_ks_external_product(ks, level, out[1 + t], a, ks->polyQPool[0]);
        memcpy(out[1 + t], ks->polyQPool[0], sizeof(uint64_t) * (size_t)(level + 1) * ks->N);
    }
}

/* ======================================================================================
 * mkbfv/basis_extension.go, mkbfv/keyswitch.go, mkbfv/keyswitch_hoisted.go
 * ====================================================================================== */
/* NewFastBasisExtender mkbfv/basis_extension.go:20-46 + NewKeySwitcher mkbfv/keyswitch.go:33-60.
 * qmulModQ[i] = QMul mod q_i (the caller does the big-integer reduction of AddScalarBigint, :42). */
ork_bfv *ork_bfv_new(const ork_ring *ringQ, const ork_ring *ringQMul, const ork_ring *ringR, const ork_ring *ringP,
                     int gamma, uint64_t T, const uint64_t *qmulModQ) {
    ork_bfv *b = (ork_bfv *)xcalloc(1, sizeof(*b));
    b->ringQ = ringQ; b->ringQMul = ringQMul; b->ringR = ringR; b->ringP = ringP; b->T = T;
    b->convQQMul = ork_be_new(ringQ, ringQMul);
    b->ks = ork_ks_new(ringQ, ringP, gamma);
    int nQ = ringQ->nmod, N = ringQ->N;
    b->mFormQMul = (uint64_t *)xmalloc(sizeof(uint64_t) * nQ);
    for (int i = 0; i < nQ; i++) b->mFormQMul[i] = mform(qmulModQ[i], ringQ->q[i], &ringQ->bred[2 * i]);
    size_t swk = (size_t)nQ * (nQ + ringP->nmod) * N;
    for (int i = 0; i < 6; i++) b->swkPool[i] = (uint64_t *)xcalloc(swk, 8);
    for (int i = 0; i < 4; i++) b->polyR[i] = (uint64_t *)xcalloc((size_t)2 * nQ * N, 8);
    for (int i = 0; i < 2; i++) b->polyQ[i] = (uint64_t *)xcalloc((size_t)nQ * N, 8);
    b->poolQ = (uint64_t *)xcalloc((size_t)nQ * N, 8);
    b->poolQMul = (uint64_t *)xcalloc((size_t)nQ * N, 8);
    b->poolR = (uint64_t *)xcalloc((size_t)2 * nQ * N, 8);
    return b;
}
void ork_bfv_free(ork_bfv *b) {
    if (!b) return;
    ork_be_free(b->convQQMul); ork_ks_free(b->ks); free(b->mFormQMul);
    for (int i = 0; i < 6; i++) free(b->swkPool[i]);
    for (int i = 0; i < 4; i++) free(b->polyR[i]);
    for (int i = 0; i < 2; i++) free(b->polyQ[i]);
    free(b->poolQ); free(b->poolQMul); free(b->poolR);
    free(b);
}
/* ModUpQtoR mkbfv/basis_extension.go:49-63: R = [Q limbs verbatim | lazy multSum lift to QMul] */
void ork_bfv_modup_q_to_r(ork_bfv *b, const uint64_t *polyQ, uint64_t *polyR) {
    int nQ = b->ringQ->nmod, N = b->ringQ->N, levelQ = nQ - 1;
    ork_be_modup_q_to_p(b->convQQMul, levelQ, levelQ, polyQ, b->poolQMul);
    memcpy(polyR, polyQ, sizeof(uint64_t) * (size_t)nQ * N);
    memcpy(polyR + (size_t)nQ * N, b->poolQMul, sizeof(uint64_t) * (size_t)nQ * N);
}
/* Quantize mkbfv/basis_extension.go:66-80 */
void ork_bfv_quantize(ork_bfv *b, const uint64_t *polyR, uint64_t *polyQ) {
    int nQ = b->ringQ->nmod, N = b->ringQ->N, levelQ = nQ - 1;
    ork_mul_scalar_lvl(b->ringR, 2 * nQ - 1, polyR, b->T, b->poolR);
    ork_intt_lvl(b->ringR, 2 * nQ - 1, b->poolR, b->poolR);
    memcpy(b->poolQ, b->poolR, sizeof(uint64_t) * (size_t)nQ * N);
    memcpy(b->poolQMul, b->poolR + (size_t)nQ * N, sizeof(uint64_t) * (size_t)nQ * N);
    ork_be_moddown_qp_to_q(b->convQQMul, levelQ, levelQ, b->poolQ, b->poolQMul, polyQ);
}
/* Rescale mkbfv/basis_extension.go:83-97: R = [lazy lift of the QMul part | canonical floor(QMul*c/Q) mod QMul] */
void ork_bfv_rescale(ork_bfv *b, const uint64_t *polyQ, uint64_t *polyR) {
    const ork_ring *r = b->ringQ;
    int nQ = r->nmod, N = r->N, levelQ = nQ - 1;
    for (int i = 0; i < nQ; i++)                                   /* MulCoeffsMontgomery(polyQ, mFormQMul) :88 */
        for (int j = 0; j < N; j++)
            b->poolQ[(size_t)i * N + j] = mred(polyQ[(size_t)i * N + j], b->mFormQMul[i], r->q[i], r->qinv[i]);
    memset(b->poolQMul, 0, sizeof(uint64_t) * (size_t)nQ * N);     /* MulScalar(.,0,.) :89 */
    ork_be_moddown_qp_to_p(b->convQQMul, levelQ, levelQ, b->poolQ, b->poolQMul, b->poolQMul);
    ork_be_modup_p_to_q(b->convQQMul, levelQ, levelQ, b->poolQMul, b->poolQ);
    memcpy(polyR, b->poolQ, sizeof(uint64_t) * (size_t)nQ * N);
    memcpy(polyR + (size_t)nQ * N, b->poolQMul, sizeof(uint64_t) * (size_t)nQ * N);
}
/* DecomposeBFV mkbfv/keyswitch.go:57-81: 2*beta digits, one per R limb (alpha = 1), each broadcast
 * UNREDUCED (incl. the lazy multSum limbs) to the Q and P limbs and NTT'd. */
void ork_bfv_decompose(ork_bfv *b, int levelQ, const uint64_t *aR, uint64_t *ad1, uint64_t *ad2) {
    ork_keyswitcher *ks = b->ks;
    int beta = ork_ks_beta(ks, levelQ), N = ks->N;
    for (int i = 0; i < beta; i++) {
        ork_ks_decompose_single_ntt(ks, levelQ, aR + (size_t)i * N, SWK_DIGIT(ks, ad1, i));
        ork_ks_decompose_single_ntt(ks, levelQ, aR + (size_t)(i + beta) * N, SWK_DIGIT(ks, ad2, i));
    }
}
/* ExternalProductBFVHoisted mkbfv/keyswitch_hoisted.go:7-35 */
void ork_bfv_external_product_hoisted(ork_bfv *b, int levelQ, const uint64_t *ah1, const uint64_t *ah2,
        const uint64_t *bg1, const uint64_t *bg2, uint64_t *c) {
    ork_keyswitcher *ks = b->ks;
    int levelP = ks->nP - 1, beta = ork_ks_beta(ks, levelQ);
    uint64_t *c1QP = ks->poolQP1;
    for (int i = 0; i < beta; i++) {
        qp_mul_mont(ks, levelQ, levelP, SWK_DIGIT(ks, bg1, i), SWK_DIGIT(ks, ah1, i), c1QP, i != 0);
        qp_mul_mont(ks, levelQ, levelP, SWK_DIGIT(ks, bg2, i), SWK_DIGIT(ks, ah2, i), c1QP, 1);
    }
    ks_finish_external_product(ks, levelQ, c1QP, c);
}
/* MulAndRelinBFVHoisted mkbfv/keyswitch_hoisted.go:39-207 (nil-hoisted branches via h*a == NULL) */
void ork_bfv_mul_and_relin_hoisted(ork_bfv *bf, int level,
        int n0, const int *ids0, uint64_t *const *op0, uint64_t *const *h0a, uint64_t *const *h0b,
        int n1, const int *ids1, uint64_t *const *op1, uint64_t *const *h1a, uint64_t *const *h1b,
        uint64_t *const *b1, uint64_t *const *b2, uint64_t *const *d1, uint64_t *const *d2,
        uint64_t *const *v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out) {
    ork_keyswitcher *ks = bf->ks;
    const ork_ring *ringQ = bf->ringQ, *ringR = bf->ringR;
    int levelP = ks->nP - 1, beta = ork_ks_beta(ks, level), N = ks->N, lvlR = ringR->nmod - 1;
    size_t digit = (size_t)(ks->nQ + ks->nP) * N;
    uint64_t *x1 = bf->swkPool[2], *x2 = bf->swkPool[3], *y1 = bf->swkPool[4], *y2 = bf->swkPool[5];
    memset(x1, 0, 8 * digit * beta); memset(x2, 0, 8 * digit * beta);                 /* :77-89 */
    memset(y1, 0, 8 * digit * beta); memset(y2, 0, 8 * digit * beta);

    for (int t = 0; t < n0; t++) {                                                    /* :92-110 */
        const uint64_t *ha, *hb;
        if (!h0a) { ork_bfv_decompose(bf, level, op0[1 + t], bf->swkPool[0], bf->swkPool[1]); ha = bf->swkPool[0]; hb = bf->swkPool[1]; }
        else { ha = h0a[t]; hb = h0b[t]; }
        for (int i = 0; i < beta; i++) {
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, d1[ids0[t]], i), SWK_DIGIT(ks, ha, i), SWK_DIGIT(ks, x1, i), 1);
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, d2[ids0[t]], i), SWK_DIGIT(ks, hb, i), SWK_DIGIT(ks, x2, i), 1);
        }
    }
    for (int i = 0; i < beta; i++) { qp_mform(ks, level, levelP, SWK_DIGIT(ks, x1, i)); qp_mform(ks, level, levelP, SWK_DIGIT(ks, x2, i)); }
    for (int t = 0; t < n1; t++) {                                                    /* :118-136 */
        const uint64_t *ha, *hb;
        if (!h1a) { ork_bfv_decompose(bf, level, op1[1 + t], bf->swkPool[0], bf->swkPool[1]); ha = bf->swkPool[0]; hb = bf->swkPool[1]; }
        else { ha = h1a[t]; hb = h1b[t]; }
        for (int i = 0; i < beta; i++) {
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, b1[ids1[t]], i), SWK_DIGIT(ks, ha, i), SWK_DIGIT(ks, y1, i), 1);
            qp_mul_mont(ks, level, levelP, SWK_DIGIT(ks, b2[ids1[t]], i), SWK_DIGIT(ks, hb, i), SWK_DIGIT(ks, y2, i), 1);
        }
    }
    for (int i = 0; i < beta; i++) { qp_mform(ks, level, levelP, SWK_DIGIT(ks, y1, i)); qp_mform(ks, level, levelP, SWK_DIGIT(ks, y2, i)); }

    uint64_t *R1 = bf->polyR[0], *R2 = bf->polyR[1], *R3 = bf->polyR[2], *R4 = bf->polyR[3];
    ork_ntt_lvl(ringR, lvlR, op0[0], R1);                                             /* :144-150 */
    ork_ntt_lvl(ringR, lvlR, op1[0], R2);
    ork_mform_lvl(ringR, lvlR, R1, R1);
    ork_mul_mont_lvl(ringR, lvlR, R1, R2, R3);
    ork_bfv_quantize(bf, R3, out[0]);
    ork_mform_lvl(ringR, lvlR, R2, R2);                                               /* :153 */
    for (int t = 0; t < n0; t++) {                                                    /* :155-161 */
        if (find_id(n1, ids1, ids0[t]) < 0) {
            ork_ntt_lvl(ringR, lvlR, op0[1 + t], R3);
            ork_mul_mont_lvl(ringR, lvlR, R2, R3, R3);
            ork_bfv_quantize(bf, R3, out[1 + find_id(nOut, idsOut, ids0[t])]);
        }
    }
    for (int t = 0; t < n1; t++) {                                                    /* :163-169 */
        if (find_id(n0, ids0, ids1[t]) < 0) {
            ork_ntt_lvl(ringR, lvlR, op1[1 + t], R3);
            ork_mul_mont_lvl(ringR, lvlR, R1, R3, R3);
            ork_bfv_quantize(bf, R3, out[1 + find_id(nOut, idsOut, ids1[t])]);
        }
    }
    for (int t = 0; t < n1; t++) {                                                    /* :171-181 */
        int s = find_id(n0, ids0, ids1[t]);
        if (s >= 0) {
            ork_ntt_lvl(ringR, lvlR, op1[1 + t], R3);
            ork_mul_mont_lvl(ringR, lvlR, R1, R3, R3);
            ork_ntt_lvl(ringR, lvlR, op0[1 + s], R4);
            ork_mul_mont_add_lvl(ringR, lvlR, R2, R4, R3);
            ork_bfv_quantize(bf, R3, out[1 + find_id(nOut, idsOut, ids1[t])]);
        }
    }
    uint64_t *Q1 = bf->polyQ[0], *Q2 = bf->polyQ[1];
    for (int t = 0; t < n1; t++) {                                                    /* :184-191 */
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids1[t])];
        if (!h1a) {
            ork_bfv_decompose(bf, level, op1[1 + t], bf->swkPool[0], bf->swkPool[1]);
            ork_bfv_external_product_hoisted(bf, level, bf->swkPool[0], bf->swkPool[1], x1, x2, Q1);
        } else ork_bfv_external_product_hoisted(bf, level, h1a[t], h1b[t], x1, x2, Q1);
        ork_add_lvl(ringQ, level, o, Q1, o);
    }
    for (int t = 0; t < n0; t++) {                                                    /* :198-216 */
        uint64_t *o = out[1 + find_id(nOut, idsOut, ids0[t])];
        if (!h0a) {
            ork_bfv_decompose(bf, level, op0[1 + t], bf->swkPool[0], bf->swkPool[1]);
            ork_bfv_external_product_hoisted(bf, level, bf->swkPool[0], bf->swkPool[1], y1, y2, Q1);
        } else ork_bfv_external_product_hoisted(bf, level, h0a[t], h0b[t], y1, y2, Q1);
        ork_ks_decompose(ks, level, Q1, ks->swkPool3);
        ork_ks_external_product_hoisted(ks, level, ks->swkPool3, v[ids0[t]], Q2);
        ork_add_lvl(ringQ, level, out[0], Q2, out[0]);
        ork_ks_external_product_hoisted(ks, level, ks->swkPool3, u, Q2);
        ork_add_lvl(ringQ, level, o, Q2, o);
    }
}

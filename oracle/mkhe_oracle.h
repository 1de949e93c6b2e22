/*
 * mkhe_oracle.h -- CPU ORACLE for the MKHE-KKLSS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's MulRelin / key-switch / hoisted-Rotate /
 * BFV basis-extension arithmetic (SNUCP/MKHE-KKLSS, Go) and of the lattigo v2.3.0 `ring`
 * primitives it bottoms out in.  It exists to CHECK the CUDA path; it is never linked,
 * imported or executed by the product (`mkhe_kklss_b200/`).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may use it.
 *
 * PARITY UNPINNED: the reference holds no golden vectors / known-answer tests for this path
 *  and neither Go nor lattigo v2.3.0 is available in this environment, so
 * the oracle cannot be checked against reference outputs.  It is pinned instead by (i) line-by-line
 * restatement of the in-tree files cited on every function, (ii) self-consistency tests
 * (schoolbook negacyclic products, big-integer CRT) and (iii) the reference's own semantic test
 * thresholds (decrypt precision / exact BFV equality), see tests/test_oracle_*.py.
 *
 * lattigo v2.3.0 internals (github.com/ldsec/lattigo/v2 v2.3.0, go.mod:7 -- NOT vendored) are
 * restated from its published algorithm; call sites in the reference are cited per function.
 *
 * Data conventions (identical to the C-ABI in include/mkhe.h):
 *   poly        : uint64_t[nlimbs][N]                       limb-major, contiguous
 *   polyQP/swk  : uint64_t[beta][nQ+nP][N]                  digit-major; Q limbs first, then P limbs
 */
#ifndef MKHE_ORACLE_H
#define MKHE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- ring (lattigo ring.Ring) ---------------- */
typedef struct ork_ring {
    int logN, N, nmod;
    uint64_t *q;        /* Modulus        */
    uint64_t *qinv;     /* MredParams: q^-1 mod 2^64 */
    uint64_t *bred;     /* BredParams: [2*i]=hi, [2*i+1]=lo of floor(2^128/q) */
    uint64_t *ninv;     /* NttNInv (Montgomery form) */
    uint64_t *psi;      /* NttPsi    [nmod][N] Montgomery form, bit-reversed order */
    uint64_t *psiinv;   /* NttPsiInv [nmod][N] */
    uint64_t *rescale;  /* RescaleParams [nmod-1][nmod]: [l-1][i] = MForm(q_l^-1 mod q_i), i<l */
} ork_ring;

ork_ring *ork_ring_new(int logN, const uint64_t *moduli, int nmod);
void ork_ring_free(ork_ring *r);
/* replace generated tables of modulus i by externally supplied lattigo tables */
void ork_ring_set_tables(ork_ring *r, int i, const uint64_t *psi, const uint64_t *psiinv, uint64_t ninv);
uint64_t ork_primitive_root(uint64_t q);

/* scalar primitives (exported for tests) */
uint64_t ork_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv);
uint64_t ork_mform(uint64_t a, uint64_t q, const uint64_t *bred);
uint64_t ork_bred_add(uint64_t a, uint64_t q, const uint64_t *bred);
uint64_t ork_modexp(uint64_t x, uint64_t e, uint64_t q);

/* per-limb vector ops; `level` = last limb index processed */
void ork_ntt_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out);
void ork_intt_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out);
void ork_intt_lazy_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out);
void ork_ntt_single(const ork_ring *r, int mod, const uint64_t *in, uint64_t *out);
void ork_mform_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out);
void ork_invmform_lvl(const ork_ring *r, int level, const uint64_t *in, uint64_t *out);
void ork_mul_mont_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ork_mul_mont_add_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ork_mul_mont_sub_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ork_add_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ork_sub_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *b, uint64_t *out);
void ork_neg_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t *out);
void ork_reduce_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t *out);
void ork_mul_scalar_lvl(const ork_ring *r, int level, const uint64_t *a, uint64_t scalar, uint64_t *out);
/* multiply limb i by residues[i] (plain, < q_i): MulScalarBigint restated on residues */
void ork_mul_residues_lvl(const ork_ring *r, int level, const uint64_t *a, const uint64_t *residues, uint64_t *out);
void ork_permute(const ork_ring *r, int level, const uint64_t *in, uint64_t galEl, uint64_t *out);
void ork_permute_ntt(const ork_ring *r, int level, const uint64_t *in, uint64_t galEl, uint64_t *out);
void ork_div_round_by_last_modulus_many(const ork_ring *r, int level, int nb, uint64_t *p0, uint64_t *p1);
uint64_t ork_galois_element_for_rotation(int logN, int k);

/* ---------------- samplers (own PRNG; lattigo's Blake2b PRNG is not reproduced) ---------------- */
typedef struct ork_prng { uint64_t s[4]; } ork_prng;
void ork_prng_seed(ork_prng *p, uint64_t seed);
uint64_t ork_prng_next(ork_prng *p);
void ork_sample_uniform(ork_prng *p, const ork_ring *r, int level, uint64_t *out);
/* small-norm integer vectors (signed), then lifted to any ring */
void ork_sample_ternary(ork_prng *p, int N, double pzero, int64_t *out);
void ork_sample_gaussian(ork_prng *p, int N, double sigma, int bound, int64_t *out);
/* counter-based samplers (include/mkhe_prng.h) */
void ork_ctr_uniform(uint64_t seed, uint64_t stream, const ork_ring *r, int level, uint64_t *out);
void ork_ctr_ternary(uint64_t seed, uint64_t stream, int N, uint64_t thr53, int64_t *out);
void ork_ctr_gaussian(uint64_t seed, uint64_t stream, int N, int64_t *out);
void ork_lift_small(const ork_ring *r, int level, const int64_t *small, uint64_t *out);

/* ---------------- FastBasisExtender (mkrlwe/basis_extension.go) ---------------- */
typedef struct ork_modup_params {
    int nq, np;
    uint64_t *qoverqiinvqi; /* [nq]            */
    uint64_t *qoverqimodp;  /* [np][nq]        */
    uint64_t *vtimesqmodp;  /* [np][nq+1]      */
} ork_modup_params;

typedef struct ork_basis_extender {
    const ork_ring *ringQ, *ringP;
    ork_modup_params *paramsQtoP;   /* [nQ] */
    ork_modup_params *paramsPtoQ;   /* [nP] */
    uint64_t *modDownPtoQ;          /* [nP][nQ] */
    uint64_t *modDownQtoP;          /* [nQ][nP] */
} ork_basis_extender;

ork_basis_extender *ork_be_new(const ork_ring *ringQ, const ork_ring *ringP);
void ork_be_free(ork_basis_extender *be);
void ork_be_modup_q_to_p(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *polQ, uint64_t *polP);
void ork_be_modup_p_to_q(const ork_basis_extender *be, int levelP, int levelQ, const uint64_t *polP, uint64_t *polQ);
void ork_be_moddown_qp_to_q(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, const uint64_t *p1P, uint64_t *p2Q);
void ork_be_moddown_qp_to_q_ntt(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, uint64_t *p1P, uint64_t *p2Q);
void ork_be_moddown_qp_to_p(const ork_basis_extender *be, int levelQ, int levelP, const uint64_t *p1Q, const uint64_t *p1P, uint64_t *p2P);

/* ---------------- KeySwitcher (mkrlwe/keyswitch.go, keyswitch_hoisted.go) ---------------- */
typedef struct ork_keyswitcher {
    const ork_ring *ringQ, *ringP;
    ork_basis_extender *be;       /* rlwe.KeySwitcher.Baseconverter */
    int nQ, nP, gamma, alpha;
    int N;
    /* pools (mirrors swkPool1..3, polyQPool[3], Pool[0..1]) */
    uint64_t *swkPool1, *swkPool2, *swkPool3;
    uint64_t *polyQPool[3];
    uint64_t *poolQP0, *poolQP1;
} ork_keyswitcher;

ork_keyswitcher *ork_ks_new(const ork_ring *ringQ, const ork_ring *ringP, int gamma);
void ork_ks_free(ork_keyswitcher *ks);
int  ork_ks_beta(const ork_keyswitcher *ks, int levelQ);
/* a: coefficient-domain poly (>= levelQ+1 limbs); ad: swk-shaped output */
void ork_ks_decompose(ork_keyswitcher *ks, int levelQ, const uint64_t *a, uint64_t *ad);
/* digit source is limb `digit` of a poly with arbitrary limb count (used by BFV's kswRP) */
void ork_ks_decompose_single_ntt(ork_keyswitcher *ks, int levelQ, const uint64_t *digitLimb, uint64_t *outQP);
void ork_ks_external_product(ork_keyswitcher *ks, int levelQ, const uint64_t *a, const uint64_t *bg, uint64_t *c);
void ork_ks_external_product_hoisted(ork_keyswitcher *ks, int levelQ, const uint64_t *aHoisted, const uint64_t *bg, uint64_t *c);

/*
 * Ciphertexts cross this API as (ids[], polys[]) : polys[0] is component "0", polys[1+t] belongs
 * to party ids[t].  hoisted[t] may be NULL as a whole array pointer (the reference's `nil`).
 * rlk_b / rlk_d / rlk_v are indexed by party id (0..maxId).  out uses idsOut (the union).
 */
void ork_ks_mul_and_relin_hoisted(ork_keyswitcher *ks, int level,
        int n0, const int *ids0, uint64_t *const *op0, uint64_t *const *h0,
        int n1, const int *ids1, uint64_t *const *op1, uint64_t *const *h1,
        uint64_t *const *rlk_b, uint64_t *const *rlk_d, uint64_t *const *rlk_v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out);
void ork_ks_mul_and_relin(ork_keyswitcher *ks, int level,
        int n0, const int *ids0, uint64_t *const *op0,
        int n1, const int *ids1, uint64_t *const *op1,
        uint64_t *const *rlk_b, uint64_t *const *rlk_d, uint64_t *const *rlk_v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out);
/* rk[t] = rotation key of party ids[t] for this rotidx; a = CRS[rotidx] */
void ork_ks_rotate_hoisted(ork_keyswitcher *ks, int level, int rotidx,
        int n, const int *ids, uint64_t *const *ctIn, uint64_t *const *hoisted,
        uint64_t *const *rk, const uint64_t *a, uint64_t *const *out);
void ork_ks_rotate(ork_keyswitcher *ks, int level, int rotidx,
        int n, const int *ids, uint64_t *const *ctIn,
        uint64_t *const *rk, const uint64_t *a, uint64_t *const *out);
void ork_ks_conjugate(ork_keyswitcher *ks, int level,
        int n, const int *ids, uint64_t *const *ctIn,
        uint64_t *const *ck, const uint64_t *a, uint64_t *const *out);

/* ---------------- MK-BFV (mkbfv/basis_extension.go, keyswitch.go, keyswitch_hoisted.go) ---------------- */
typedef struct ork_bfv {
    const ork_ring *ringQ, *ringQMul, *ringR, *ringP;
    ork_basis_extender *convQQMul;
    ork_keyswitcher *ks;            /* over (Q,P) */
    uint64_t T;
    uint64_t *mFormQMul;            /* [nQ] : MForm(QMul mod q_i) */
    uint64_t *swkPool[6];           /* swkPool1..6 */
    uint64_t *polyR[4];
    uint64_t *polyQ[2];
    uint64_t *poolQ, *poolQMul, *poolR;
} ork_bfv;

ork_bfv *ork_bfv_new(const ork_ring *ringQ, const ork_ring *ringQMul, const ork_ring *ringR, const ork_ring *ringP,
                     int gamma, uint64_t T, const uint64_t *qmulModQ);
void ork_bfv_free(ork_bfv *b);
void ork_bfv_modup_q_to_r(ork_bfv *b, const uint64_t *polyQ, uint64_t *polyR);
void ork_bfv_rescale(ork_bfv *b, const uint64_t *polyQ, uint64_t *polyR);
void ork_bfv_quantize(ork_bfv *b, const uint64_t *polyR, uint64_t *polyQ);
void ork_bfv_decompose(ork_bfv *b, int levelQ, const uint64_t *aR, uint64_t *ad1, uint64_t *ad2);
void ork_bfv_external_product_hoisted(ork_bfv *b, int levelQ, const uint64_t *ah1, const uint64_t *ah2,
        const uint64_t *bg1, const uint64_t *bg2, uint64_t *c);
/* op0/op1 are R-basis ciphertexts (2*nQ limbs per poly); h*a/h*b the two hoisted halves per party (may be NULL) */
void ork_bfv_mul_and_relin_hoisted(ork_bfv *b, int level,
        int n0, const int *ids0, uint64_t *const *op0, uint64_t *const *h0a, uint64_t *const *h0b,
        int n1, const int *ids1, uint64_t *const *op1, uint64_t *const *h1a, uint64_t *const *h1b,
        uint64_t *const *b1, uint64_t *const *b2, uint64_t *const *d1, uint64_t *const *d2,
        uint64_t *const *v, const uint64_t *u,
        int nOut, const int *idsOut, uint64_t *const *out);

int ork_set_threads(int n);   /* OpenMP threads used by the limb loops; returns the value in effect */

#ifdef __cplusplus
}
#endif
#endif

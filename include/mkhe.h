/*
 * mkhe.h -- C ABI of libmkhe_b200.so: the B200 (sm_100a) backend for the MKHE-KKLSS hot path.
 *
 * The reference (SNUCP/MKHE-KKLSS, pure Go) has no FFI seam; this header IS the seam a maintainer
 * adds under the method bodies of mkrlwe.KeySwitcher / mkbfv.KeySwitcher / mkbfv.FastBasisExtender /
 * mkckks.Evaluator.Rescale.  Each entry point cites the reference method whose
 * body it replaces (paths relative to the reference root).  The cgo binding is in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative mkhe_status otherwise; mkhe_last_error() gives
 *     the message.  CUDA failures are sticky on the context.
 *   - host pointers are borrowed for the duration of the call only (cgo rule): uploads/downloads
 *     complete before the function returns (the *_async pair on library-owned pinned memory is the
 *     documented exception).  Ops only ENQUEUE work on the context's stream; results become visible
 *     to the host through *_download / mkhe_sync.
 *   - a context is not thread safe (neither are the reference's KeySwitcher / Evaluator: shared pools,
 *     mkrlwe/keyswitch.go:12-15); one context = one CUDA device = one stream.
 *   - limb  = N little-endian uint64.   poly = [nlimbs][N] (limb-major).
 *     swk   = [beta_max][nQ+nP][N]: digit-major; Q limbs first, then P limbs (mkrlwe/keys.go:23-25,
 *     rlwe.PolyQP{Q,P}); always allocated at max level like the reference (keys.go:245-256).
 *   - ciphertext limbs: coefficient domain, values as the reference stores them (canonical [0,q) except
 *     the documented lazy forms).  key limbs: NTT domain in lattigo's bit-reversed
 *     psi ordering, Montgomery form (value * 2^64 mod q) -- exactly the bytes of the Go slices.
 *   - a ciphertext crosses the ABI as an array of poly handles: element 0 is component "0",
 *     element 1+t belongs to the t-th party of the accompanying id list (mkrlwe/elements.go:17-33).
 */
#ifndef MKHE_H
#define MKHE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mkhe_ctx mkhe_ctx;
typedef uint64_t mkhe_poly;   /* device-resident ring.Poly            */
typedef uint64_t mkhe_swk;    /* device-resident mkrlwe.SwitchingKey / hoisted form */

typedef enum mkhe_status {
    MKHE_OK = 0,
    MKHE_ERR_INVALID = -1,      /* bad argument (level mismatch, null handle, ...) */
    MKHE_ERR_CUDA = -2,         /* CUDA runtime failure (sticky) */
    MKHE_ERR_NOMEM = -3,
    MKHE_ERR_UNSUPPORTED = -4,  /* e.g. alpha = #P/gamma != 1, logN out of range */
    MKHE_ERR_NCCL = -5
} mkhe_status;

#define MKHE_MAX_PARTIES 64

/* ---- context: replaces mkrlwe.NewParameters + NewKeySwitcher scratch (params.go:16-61, keyswitch.go:33-47) */
int mkhe_ctx_create(int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP, int gamma,
                    int device, mkhe_ctx **out);
/* MK-BFV: adds ring QMul (R = Q u QMul) and t (mkbfv/params.go:36-83, mkbfv/basis_extension.go:20-46) */
int mkhe_ctx_set_bfv(mkhe_ctx *ctx, const uint64_t *QMul, int nQMul, uint64_t T);
/* optional: lattigo's own ring.Ring tables (public fields NttPsi/NttPsiInv/NttNInv, Montgomery form) for
 * modulus index m (0..nQ-1 = Q, nQ..nQ+nP-1 = P, then QMul).  Must be called before any op. */
int mkhe_ctx_set_ntt_tables(mkhe_ctx *ctx, int m, const uint64_t *nttPsi, const uint64_t *nttPsiInv,
                            uint64_t nttNInv);
/* A lane: a second evaluator over the same parameters, keys and ciphertexts -- what constructing another mkrlwe.KeySwitcher
 * over one Parameters value is in the reference (private pools, shared read-only tables: mkrlwe/keyswitch.go:33-47, and
 * FastBasisExtender.ShallowCopy, mkrlwe/basis_extension.go:155-175).  The fork shares every handle of its root (polys, switching
 * keys, hoisted forms) and owns a CUDA stream, two copy streams and its scratch pools, so independent ops issued on different
 * lanes overlap on the device (the bandwidth-bound multiply-accumulates of one beside the integer-bound transforms of the
 * other).  Uses of ONE object on different lanes are ordered automatically with events (a reader waits for the last writer, a
 * writer for every earlier user).  A root and its forks are driven from one host thread at a time; at most 3 forks per root;
 * mkhe_ctx_set_bfv / mkhe_ctx_set_ntt_tables come first.  Destroying the root destroys its forks. */
int mkhe_ctx_fork(mkhe_ctx *ctx, mkhe_ctx **out);
/* ctx's stream waits for everything enqueued so far on `other` (a lane of the same root): joins lanes without blocking the host */
int mkhe_ctx_wait(mkhe_ctx *ctx, mkhe_ctx *other);
void mkhe_ctx_destroy(mkhe_ctx *ctx);
const char *mkhe_last_error(const mkhe_ctx *ctx);
int mkhe_sync(mkhe_ctx *ctx);
int mkhe_device_count(void);

/* ---- ring.Poly storage (Ciphertext.Value[id], mkrlwe/elements.go:17-62) */
int mkhe_poly_alloc(mkhe_ctx *ctx, int nlimbs_capacity, mkhe_poly *out);
int mkhe_poly_free(mkhe_ctx *ctx, mkhe_poly p);
int mkhe_poly_set_nlimbs(mkhe_ctx *ctx, mkhe_poly p, int nlimbs);   /* DropLevel / Rescale trim: a view, no realloc (mkckks/evaluator.go:106-112,389) */
int mkhe_poly_get_nlimbs(mkhe_ctx *ctx, mkhe_poly p, int *nlimbs);
int mkhe_poly_upload_limb(mkhe_ctx *ctx, mkhe_poly p, int limb, const uint64_t *src);     /* Coeffs[limb] */
int mkhe_poly_download_limb(mkhe_ctx *ctx, mkhe_poly p, int limb, uint64_t *dst);
int mkhe_poly_upload(mkhe_ctx *ctx, mkhe_poly p, const uint64_t *src, int nlimbs);        /* contiguous [nlimbs][N] */
int mkhe_poly_download(mkhe_ctx *ctx, mkhe_poly p, uint64_t *dst, int nlimbs);
/* one limb, asynchronous, on the copy streams (same ordering rules as mkhe_poly_upload_async / _download_async) */
int mkhe_poly_upload_limb_async(mkhe_ctx *ctx, mkhe_poly h, int limb, const uint64_t *src);
int mkhe_poly_download_limb_async(mkhe_ctx *ctx, mkhe_poly h, int limb, uint64_t *dst);
/* the limbs of [0, nlimbs) this rank of a team owns (all of them without a team), between a host buffer in whole-poly layout and the
 * poly: one strided DMA.  With mkhe_team_allgather this is the multi-GPU upload path: 1 / nranks of a ciphertext per rank over PCIe. */
int mkhe_poly_upload_owned_async(mkhe_ctx *ctx, mkhe_poly h, const uint64_t *src, int nlimbs);
int mkhe_poly_download_owned_async(mkhe_ctx *ctx, mkhe_poly h, uint64_t *dst, int nlimbs);
int mkhe_poly_copy(mkhe_ctx *ctx, mkhe_poly dst, mkhe_poly src);                          /* ring.Poly.Copy */
int mkhe_poly_copy_lvl(mkhe_ctx *ctx, int level, mkhe_poly dst, mkhe_poly src);           /* ring.CopyValuesLvl (mkckks/evaluator.go:297-301): limbs 0..level, views unchanged */
/* Asynchronous transfers for pipelines that keep the device busy while ciphertexts stream over PCIe (the Go shim's lazy
 * syncToDevice / syncToHost, SURVEY 8b "residency model").  The host buffer must be page-locked memory obtained from
 * mkhe_host_alloc (C-owned, so the cgo pointer rule does not apply) and must stay valid and untouched until mkhe_sync.
 * Ordering is automatic: the copy starts after every op enqueued so far, and any later op that uses the poly waits for it;
 * ops that do not touch the poly run concurrently with the copy. */
int mkhe_host_alloc(mkhe_ctx *ctx, size_t bytes, void **out);
int mkhe_host_free(mkhe_ctx *ctx, void *p);
int mkhe_poly_upload_async(mkhe_ctx *ctx, mkhe_poly p, const uint64_t *src, int nlimbs);
int mkhe_poly_download_async(mkhe_ctx *ctx, mkhe_poly p, uint64_t *dst, int nlimbs);

/* ---- SwitchingKey / HoistedCiphertext entry storage (mkrlwe/keys.go:23-62, elements.go:5-15) */
int mkhe_swk_alloc(mkhe_ctx *ctx, mkhe_swk *out);
int mkhe_swk_free(mkhe_ctx *ctx, mkhe_swk k);
int mkhe_swk_upload_limb(mkhe_ctx *ctx, mkhe_swk k, int digit, int is_p, int limb, const uint64_t *src);   /* Value[digit].{Q,P}.Coeffs[limb] */
int mkhe_swk_download_limb(mkhe_ctx *ctx, mkhe_swk k, int digit, int is_p, int limb, uint64_t *dst);
int mkhe_swk_upload(mkhe_ctx *ctx, mkhe_swk k, const uint64_t *src);      /* contiguous [beta_max][nQ+nP][N] */
int mkhe_swk_download(mkhe_ctx *ctx, mkhe_swk k, uint64_t *dst);

/* ---- lattigo ring primitives exposed for parity tests (ring.NTTLvl / InvNTTLvl) */
int mkhe_ntt(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out);
int mkhe_intt(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out);

/* ---- mkrlwe.KeySwitcher */
/* Decompose(levelQ, a, ad)                                   mkrlwe/keyswitch.go:49-73 (= HoistedForm body, mkckks/evaluator.go:543-553) */
int mkhe_decompose(mkhe_ctx *ctx, int levelQ, mkhe_poly a, mkhe_swk ad);
/* FastBasisExtender.ModDownQPtoQNTT(levelQ, levelP = max, p1Q, p1P, p2Q)      mkrlwe/basis_extension.go:239-290
 * p1 = one rlwe.PolyQP (nQ + nP limbs: Q limbs, then P limbs) in the NTT domain; p2Q = round(p1 / P), NTT domain */
int mkhe_moddown_qp_to_q_ntt(mkhe_ctx *ctx, int levelQ, mkhe_poly p1, mkhe_poly p2Q);
/* ExternalProduct(levelQ, a, bg, c)                          mkrlwe/keyswitch.go:79-118 */
int mkhe_external_product(mkhe_ctx *ctx, int levelQ, mkhe_poly a, mkhe_swk bg, mkhe_poly c);
/* ExternalProductHoisted(levelQ, aHoisted, bg, c)            mkrlwe/keyswitch_hoisted.go:10-40 */
int mkhe_external_product_hoisted(mkhe_ctx *ctx, int levelQ, mkhe_swk a_hoisted, mkhe_swk bg, mkhe_poly c);
/* MulAndRelinHoisted(op0, op1, op0Hoisted, op1Hoisted, rlkSet, ctOut)   mkrlwe/keyswitch_hoisted.go:44-179
 *   op0[0..n0], op1[0..n1], out[0..nOut]: ciphertext component handles (element 0 = "0").
 *   h0 / h1: hoisted forms per party, or NULL for the reference's nil branches (:80-85,100-105,148-149,165-166).
 *   rlk_d[n0], rlk_v[n0] belong to ids0[t]; rlk_b[n1] to ids1[t]  (rlkSet.Value[id].Value[1|2|0]).
 *   u = params.CRS[-1].  idsOut must be the union of ids0 and ids1; level = ctOut.Level().
 *   MulAndRelin (mkrlwe/keyswitch.go:122-230) is the same call with h0 = h1 = NULL. */
int mkhe_mul_relin_hoisted(mkhe_ctx *ctx, int level,
                           int n0, const int *ids0, const mkhe_poly *op0, const mkhe_swk *h0,
                           int n1, const int *ids1, const mkhe_poly *op1, const mkhe_swk *h1,
                           const mkhe_swk *rlk_b, const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u,
                           int nOut, const int *idsOut, const mkhe_poly *out);
/* RotateHoisted(ctIn, rotidx, ctInHoisted, rkSet, ctOut)     mkrlwe/keyswitch_hoisted.go:183-247
 *   rk[t] = rkSet.GetRotationKey(id_t, rotidx).Value ; a = params.CRS[rotidx] */
int mkhe_rotate_hoisted(mkhe_ctx *ctx, int level, int rotidx, int n,
                        const mkhe_poly *ct_in, const mkhe_swk *hoisted, const mkhe_swk *rk, mkhe_swk a,
                        const mkhe_poly *ct_out);
/* Rotate(ctIn, rotidx, rkSet, ctOut)                         mkrlwe/keyswitch.go:234-298 */
int mkhe_rotate(mkhe_ctx *ctx, int level, int rotidx, int n,
                const mkhe_poly *ct_in, const mkhe_swk *rk, mkhe_swk a, const mkhe_poly *ct_out);
/* Conjugate(ctIn, ckSet, ctOut)                              mkrlwe/keyswitch.go:302-332 ; a = CRS[-2] */
int mkhe_conjugate(mkhe_ctx *ctx, int level, int n,
                   const mkhe_poly *ct_in, const mkhe_swk *ck, mkhe_swk a, const mkhe_poly *ct_out);

/* ---- mkckks.Evaluator */
/* ringQ.DivRoundByLastModulusManyLvl(level, nbRescales, in, pool, out) + Coeffs trim   mkckks/evaluator.go:385-390.
 * Like lattigo it adds (q_l-1)/2 to the INPUT's last limb in place. out view = level+1-nb limbs. */
int mkhe_rescale(mkhe_ctx *ctx, int level, int nb_rescales, mkhe_poly in, mkhe_poly out);
/* ringQ.AddLvl / SubLvl(level, a, b, out)                    mkckks/evaluator.go:200-351 call sites */
int mkhe_poly_add(mkhe_ctx *ctx, int level, mkhe_poly a, mkhe_poly b, mkhe_poly out);
int mkhe_poly_sub(mkhe_ctx *ctx, int level, mkhe_poly a, mkhe_poly b, mkhe_poly out);
/* ringQ.NegLvl(level, a, out): q - x, unreduced (0 is stored as q)       mkckks/evaluator.go:340 (Sub's negation) */
int mkhe_poly_neg(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out);
/* MultByConst(ct0, constant, ctOut)                          mkckks/evaluator.go:117-198
 *   c_real, c_imag, scale: what getConstAndScale (:39-93, host bookkeeping) returns for the constant.  The library derives the
 *   per-limb multipliers exactly like the reference (scaleUpExact, mkckks/utils.go:59-86; MRed by NttPsi[i][1] for the imaginary
 *   part) and multiplies coefficients [0, N/2) and [N/2, N) of every limb <= level by the first / second one.  in[t] -> out[t]
 *   for the n components of ct0 (out may alias in).  ctOut.Scale = ct0.Scale * scale stays on the host. */
int mkhe_ckks_mult_by_const(mkhe_ctx *ctx, int level, int n, const mkhe_poly *in, const mkhe_poly *out,
                            double c_real, double c_imag, double scale);
/* MulPtxtNew's device work before its Rescale                mkckks/evaluator.go:465-478
 *   pt = ckks.Plaintext.Value (coefficient domain): NTT + MForm once, then per component NTT, MulCoeffsMontgomery, InvNTT. */
int mkhe_ckks_mul_ptxt(mkhe_ctx *ctx, int level, mkhe_poly pt, int n, const mkhe_poly *in, const mkhe_poly *out);
/* MulRelinNew's device work in one call: hoist both operands (once if same_operand), MulAndRelinHoisted,
 * Rescale by nb_rescales                                     mkckks/evaluator.go:416-443,558-581
 * (this is what mkckks_benchmark_test.go:78-82 times).  Hoisting uses context-owned pools, mirroring
 * rlkSet.HoistPool (mkrlwe/keys.go:53-57). out views are trimmed to level+1-nb_rescales. */
int mkhe_ckks_mul_relin(mkhe_ctx *ctx, int level, int nb_rescales, int same_operand,
                        int n0, const int *ids0, const mkhe_poly *op0,
                        int n1, const int *ids1, const mkhe_poly *op1,
                        const mkhe_swk *rlk_b, const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u,
                        int nOut, const int *idsOut, const mkhe_poly *out);

/* ---- mkrlwe.Decryptor (the step after the path)
 * Decrypt(ciphertext, skSet, plaintext)                      mkrlwe/decryptor.go:48-66 (PartialDecrypt :26-43 per party, ReduceLvl)
 *   ct[0] = component "0", ct[1+t] = component of party t; sk[t] = that party's SecretKey.Value.Q (NTT domain, Montgomery form)
 *   uploaded as a poly; pt receives the coefficient-domain plaintext, canonical, level+1 limbs.  The ciphertext is not modified
 *   (the reference works on a CopyNew).  Secret keys only ever reach the device if the caller puts them there. */
int mkhe_decrypt(mkhe_ctx *ctx, int level, int n, const mkhe_poly *ct, const mkhe_poly *sk, mkhe_poly pt);

/* ---- mkbfv (polys in basis R have 2*nQ limbs: Q limbs then QMul limbs) */
/* FastBasisExtender.ModUpQtoR(polyQ, polyR)                  mkbfv/basis_extension.go:49-63 */
int mkhe_bfv_modup_q_to_r(mkhe_ctx *ctx, mkhe_poly polyQ, mkhe_poly polyR);
/* FastBasisExtender.Rescale(polyQ, polyR)                    mkbfv/basis_extension.go:83-97 */
int mkhe_bfv_rescale_q_to_r(mkhe_ctx *ctx, mkhe_poly polyQ, mkhe_poly polyR);
/* FastBasisExtender.Quantize(polyR, polyQ, t)                mkbfv/basis_extension.go:66-80 (polyR in NTT domain) */
int mkhe_bfv_quantize(mkhe_ctx *ctx, mkhe_poly polyR, mkhe_poly polyQ);
/* KeySwitcher.DecomposeBFV(levelQ, aR, ad1, ad2)             mkbfv/keyswitch.go:57-81 */
int mkhe_bfv_decompose(mkhe_ctx *ctx, int levelQ, mkhe_poly aR, mkhe_swk ad1, mkhe_swk ad2);
/* KeySwitcher.MulAndRelinBFVHoisted(...)                     mkbfv/keyswitch_hoisted.go:39-207
 *   op0/op1 are R-basis ciphertexts; h*a / h*b the two hoisted halves per party (NULL = nil branches);
 *   d1,d2,v aligned with ids0 ; b1,b2 aligned with ids1. */
int mkhe_bfv_mul_relin_hoisted(mkhe_ctx *ctx, int level,
                               int n0, const int *ids0, const mkhe_poly *op0, const mkhe_swk *h0a, const mkhe_swk *h0b,
                               int n1, const int *ids1, const mkhe_poly *op1, const mkhe_swk *h1a, const mkhe_swk *h1b,
                               const mkhe_swk *b1, const mkhe_swk *b2, const mkhe_swk *d1, const mkhe_swk *d2,
                               const mkhe_swk *v, mkhe_swk u,
                               int nOut, const int *idsOut, const mkhe_poly *out);
/* Evaluator.MulRelinNew's device work: ModUpQtoR / Rescale every component, DecomposeBFV, MulAndRelinBFVHoisted
 *                                                            mkbfv/evaluator.go:84-150 */
int mkhe_bfv_mul_relin(mkhe_ctx *ctx,
                       int n0, const int *ids0, const mkhe_poly *ct0,
                       int n1, const int *ids1, const mkhe_poly *ct1,
                       const mkhe_swk *b1, const mkhe_swk *b2, const mkhe_swk *d1, const mkhe_swk *d2,
                       const mkhe_swk *v, mkhe_swk u,
                       int nOut, const int *idsOut, const mkhe_poly *out);

/* ---- multi-GPU: party-sharded MulRelin; one context per rank, NCCL over NVLink.
 * unique_id = the 128 bytes of an ncclUniqueId created by rank 0 (mkhe_comm_unique_id) and broadcast by the host. */
int mkhe_comm_unique_id(uint8_t out[128]);
int mkhe_comm_init(mkhe_ctx *ctx, int nranks, int rank, const uint8_t unique_id[128]);
int mkhe_comm_destroy(mkhe_ctx *ctx);
/* Fused exchange over peer memory (NVLink / NVSwitch) for the sharded MulRelin: after mkhe_comm_init every rank calls
 * mkhe_p2p_export (allocates its two exchange buffers, returns their CUDA IPC handles: 2 x 64 bytes), the host hands the
 * concatenation of all ranks' 128 bytes to mkhe_p2p_import on every rank.  From then on the partial x, y of
 * mkhe_ckks_mul_relin_sharded are written by the multiply-accumulate kernel straight into their owner's memory, summed there and
 * the sums written straight into every rank's x||y -- no 112 MiB all-reduce; NCCL only provides two tiny barriers per op and the
 * all-reduce of the c_0 contributions.  One node, one process per GPU, at most 8 ranks. */
int mkhe_p2p_export(mkhe_ctx *ctx, uint8_t out[128]);
int mkhe_p2p_import(mkhe_ctx *ctx, int nranks, int rank, const uint8_t *all_handles);
/* ---- key generation and encryption on the device.  Randomness = the counter-based samplers of
 * mkhe_prng.h ("mkhe-ctr-1"): every sampled polynomial is one stream of (seed, stream); a call documents how many streams it
 * takes from `stream` on, in the order mkrlwe.KeyGenerator draws them.  Secret and public keys are polys of nQ + nP limbs
 * (rlwe.PolyQP flattened: Q limbs, then P limbs), NTT domain; secret keys in Montgomery form like the reference's.
 * Nothing crosses PCIe: at k = 32, logN = 15 that is 5.4 GiB of keys made where they are used. */
int mkhe_sample_crs(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_swk out);           /* mkrlwe/params.go:47-58,77-99 (AddCRS): uniform then MForm, beta_max * (nQ + nP) streams (digit-major, Q limbs then P limbs) */
int mkhe_keygen_secret(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, double pzero, mkhe_poly sk);   /* GenSecretKeyWithDistrib keygen.go:44-76 (GenSecretKey: pzero = 0.5): 1 stream */
int mkhe_keygen_public(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk crs_a,
                       mkhe_poly pk0, mkhe_poly pk1);                                          /* GenPublicKey keygen.go:88-109: 1 stream */
int mkhe_keygen_switching_key(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk out);   /* GenSwitchingKey keygen.go:269-327 (g s + e): beta streams */
int mkhe_keygen_relin(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_poly r, mkhe_swk crs_a, mkhe_swk crs_u,
                      mkhe_swk b, mkhe_swk d, mkhe_swk v);                                     /* GenRelinearizationKey keygen.go:137-187: 3 beta streams (b | d | v) */
int mkhe_keygen_rotation(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, int rotidx, mkhe_poly sk, mkhe_swk crs_rot,
                         mkhe_swk rk);                                                         /* GenRotationKey keygen.go:190-229: beta streams */
int mkhe_keygen_conjugation(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk crs_cj,
                            mkhe_swk ck);                                                      /* GenConjugationKey keygen.go:240-266: beta streams */
/* mkbfv.KeyGenerator.GenRelinearizationKey with GenBFVSwitchingKey (mkbfv/keygen.go:24-162): (b1, d1, v) | (b2, d2); crs_a1 = CRS[0],
 * crs_a2 = CRS[-3], crs_u = CRS[-1].  5 beta streams: b1[i] = stream + 2i, b2[i] = stream + 2i + 1, then d1 | d2 | v. */
int mkhe_keygen_bfv_relin(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_poly r, mkhe_swk crs_a1, mkhe_swk crs_a2,
                          mkhe_swk crs_u, mkhe_swk b1, mkhe_swk b2, mkhe_swk d1, mkhe_swk d2, mkhe_swk v);
/* Encryptor.Encrypt, coefficient-domain ciphertext (mkrlwe/encryptor.go:55-118); pt = 0: an encryption of zero; 3 streams (w, e0, e1) */
int mkhe_encrypt(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, int level, mkhe_poly pt, mkhe_poly pk0, mkhe_poly pk1,
                 mkhe_poly c0, mkhe_poly c1);

/* ---- multi-GPU: limb-sharded MulRelinNew.  A team = the contexts ("ranks") that execute ONE op together: rank r
 * owns the limb slots s with s mod nranks == r of every key, hoisted form and accumulator (mkrlwe/keyswitch_hoisted.go:44-179 is
 * limb-local apart from ModDown's P limbs, the re-decomposed p_id and the final Rescale).  The three exchanges are peer stores from
 * the epilogues of the kernels that produce the data (NVLink / NVSwitch), ordered by an in-kernel flag barrier; no collective
 * library is involved.  Every rank passes the same (replicated) operands and key handles and ends with the whole result,
 * bit-identical to mkhe_ckks_mul_relin.
 *   ranks in different processes: mkhe_team_export on every rank (allocates the rank's team memory for ciphertexts of up to
 *     max_parties parties and returns its 64-byte CUDA IPC handle); the host hands the concatenation of all handles (rank order)
 *     to mkhe_team_import on every rank.
 *   ranks inside one process (several GPUs, or several contexts on one GPU as in the tests): mkhe_team_join_local with the
 *     contexts of all ranks, called once per rank.
 * Every lane (mkhe_ctx_fork) that issues limb-sharded ops joins a team of its own.  mkhe_team_status reports a barrier that gave
 * up waiting for a rank (5 s): the op's results are then undefined. */
int mkhe_team_export(mkhe_ctx *ctx, int max_parties, uint8_t out[64]);
int mkhe_team_import(mkhe_ctx *ctx, int nranks, int rank, const uint8_t *all_handles);
int mkhe_team_join_local(mkhe_ctx *ctx, int max_parties, int nranks, int rank, mkhe_ctx *const *members);
int mkhe_team_status(mkhe_ctx *ctx, int *timed_out);
/* every rank holds valid data in the limbs it owns (limb mod nranks == rank, mkhe_team_owns_limb) of each listed poly -- e.g.
 * uploaded there, 1 / nranks of the ciphertext per rank over PCIe; afterwards (in stream order) limbs 0..level of every poly are
 * valid on every rank: the rank's limbs go to every rank's staging area by peer stores, barrier, local copy. */
int mkhe_team_allgather(mkhe_ctx *ctx, int level, int npolys, const mkhe_poly *polys);
int mkhe_team_owns_limb(mkhe_ctx *ctx, int limb);
int mkhe_team_flags(mkhe_ctx *ctx, uint64_t out[17]);   /* diagnostics: flag / status words of this rank + barriers issued */
int mkhe_ckks_mul_relin_limbs(mkhe_ctx *ctx, int level, int nb_rescales,
                              int n0, const int *ids0, const mkhe_poly *op0,
                              int n1, const int *ids1, const mkhe_poly *op1,
                              const mkhe_swk *rlk_b, const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u,
                              int nOut, const int *idsOut, const mkhe_poly *out);
/* party-sharded MulRelinNew: every rank passes the same (replicated) operand ciphertexts and id lists, but only holds
 * the relinearization keys of the parties in own_ids (other entries of rlk_* may be 0).  Partial x, y and the c_0
 * contributions are summed with ncclAllReduce(uint64, sum) and reduced mod q -- bit-identical to the single-GPU result
 * because every accumulation of the reference is an exact modular add.  On return the rank holds component "0" and
 * the components of its own parties (the others are left incomplete).   mkckks/evaluator.go:416-443 */
int mkhe_ckks_mul_relin_sharded(mkhe_ctx *ctx, int level, int nb_rescales,
                                int n0, const int *ids0, const mkhe_poly *op0,
                                int n1, const int *ids1, const mkhe_poly *op1,
                                int nown, const int *own_ids,
                                const mkhe_swk *rlk_b, const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u,
                                int nOut, const int *idsOut, const mkhe_poly *out);

/* ---- measurement helpers (CUDA events on the context's stream) */
int mkhe_timer_start(mkhe_ctx *ctx);
int mkhe_timer_stop(mkhe_ctx *ctx, float *elapsed_ms);     /* synchronises */
/* per-kernel timing: between begin and end every launch is bracketed by CUDA events on the context's stream;
 * end() writes one line "<kernel> <launches> <total_ms>" per kernel into buf.  Not used on the timed path. */
int mkhe_profile_begin(mkhe_ctx *ctx);
int mkhe_profile_end(mkhe_ctx *ctx, char *buf, size_t cap);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t mkhe_launch_count(const mkhe_ctx *ctx);
/* CUDA graphs (opt-in per lane; MKHE_GRAPHS=1 in the environment turns it on for every new context): an op issued again on the same
 * lane with the very same operands (device addresses, ids, level) replays the launch sequence captured at its third occurrence --
 * one graph launch instead of up to 22 kernel launches.  mkhe_launch_count keeps counting the kernels inside replayed graphs.
 * Off by default: measured on B200 the ops are GPU bound even at logN = 14, k = 2, and two lanes interleave better kernel by
 * kernel than graph by graph (profiles/r02p_cuda_graphs.txt). */
int mkhe_ctx_set_graphs(mkhe_ctx *ctx, int on);
int mkhe_graph_stats(const mkhe_ctx *ctx, uint64_t *captures, uint64_t *replays);
/* register-resident Shoup-butterfly throughput microbenchmark: butterflies per second on this device
 * (the integer-pipe roofline denominator) */
int mkhe_bench_butterfly_peak(mkhe_ctx *ctx, double *butterflies_per_s);
const char *mkhe_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MKHE_H */

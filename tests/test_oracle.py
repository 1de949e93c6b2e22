"""CPU tests of the ORACLE itself (oracle/ is test infrastructure).  The reference ships no golden vectors
(SURVEY.md section 4), so the restatement is pinned by (i) self-consistency against independent big-integer
arithmetic, (ii) the reference's own semantic thresholds, (iii) committed fixtures (tests/golden) that
freeze its behaviour."""
import hashlib
import json
import os
from functools import reduce

import numpy as np
import pytest

from oracle import oracle as O
from mkhe_kklss_b200 import params as PR

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rand_poly(rng, moduli, N):
    return np.array([[int(x) % q for x in rng.integers(0, 2**63, N)] for q in moduli], dtype=np.uint64)


@pytest.mark.parametrize("lit", [PR.CKKS_PN15QP880, PR.CKKS_PN14QP439, PR.CNN_PN14QP433, PR.BFV_PN15QP880, PR.BFV_PN14QP439],
                         ids=lambda l: l.name)
def test_primes_and_roots(lit):
    """every prime of the five in-tree sets is prime and = 1 mod 2N; psi is a primitive 2N-th root"""
    for q in list(lit.Q) + list(lit.P) + list(lit.QMul):
        assert (q - 1) % (2 << lit.logN) == 0 and q < 2**60
        g = int(O.lib().ork_primitive_root(q))
        psi = pow(g, (q - 1) // (2 * lit.N), q)
        assert pow(psi, lit.N, q) == q - 1


def test_ntt_is_negacyclic_product():
    """NTT-based product == schoolbook product mod (X^N + 1, q)  (lattigo NTT/InvNTT/MForm/MRed restatement)"""
    N, logN = 64, 6
    r = O.Ring(logN, PR.CKKS_PN15QP880.Q[:3])
    rng = np.random.default_rng(1)
    a, b = _rand_poly(rng, r.moduli, N), _rand_poly(rng, r.moduli, N)
    c = r.intt(r.mul_mont(r.ntt(a), r.mform(r.ntt(b))))
    for i, q in enumerate(r.moduli):
        ref = [0] * N
        for x in range(N):
            for y in range(N):
                v = int(a[i, x]) * int(b[i, y])
                if x + y >= N:
                    ref[x + y - N] -= v
                else:
                    ref[x + y] += v
        assert [v % q for v in ref] == [int(v) for v in c[i]]
    assert np.array_equal(r.intt(r.ntt(a)), a)
    lazy = r.intt(r.ntt(a), lazy=True)
    assert all(np.array_equal(lazy[i] % np.uint64(q), a[i]) for i, q in enumerate(r.moduli))


def test_ntt_accepts_unreduced_digits():
    """digit broadcast feeds values up to 2^60 into 54-bit limbs (basis_extension.go:443-451)"""
    r = O.Ring(8, PR.CKKS_PN15QP880.Q[:3])
    rng = np.random.default_rng(2)
    big = rng.integers(0, 2**60, (1, r.N), dtype=np.uint64)
    x = np.repeat(big, 3, axis=0)
    red = np.array([x[i] % np.uint64(q) for i, q in enumerate(r.moduli)], dtype=np.uint64)
    assert np.array_equal(r.ntt(x), r.ntt(red))


def test_modup_exact_matches_bigint_crt():
    """modUpExact (reconstructRNS + multSum) == exact CRT lift; outputs are lazy but congruent"""
    lit = PR.BFV_PN14QP439
    rq, rp = O.Ring(6, lit.Q), O.Ring(6, lit.QMul)
    be = O.BasisExtender(rq, rp)
    rng = np.random.default_rng(3)
    a = _rand_poly(rng, rq.moduli, rq.N)
    lifted = be.modup_q_to_p(rq.nmod - 1, rp.nmod - 1, a)
    Qb = reduce(lambda x, y: x * y, lit.Q, 1)
    for x in range(rq.N):
        v = 0
        for i, q in enumerate(lit.Q):
            Qi = Qb // q
            v += Qi * ((int(a[i, x]) * pow(Qi, -1, q)) % q)
        v %= Qb
        for j, p in enumerate(lit.QMul):
            assert int(lifted[j, x]) % p == v % p
            assert int(lifted[j, x]) < 3 * p
    # ModDownQPtoQ == floor division by P after the exact lift
    aP = _rand_poly(rng, rp.moduli, rq.N)
    down = be.moddown_qp_to_q(rq.nmod - 1, rp.nmod - 1, a, aP)
    Pb = reduce(lambda x, y: x * y, lit.QMul, 1)
    for x in range(0, rq.N, 7):
        vq = O.crt_centered(lit.Q, a[:, x:x + 1])[0] % Qb
        vp = O.crt_centered(lit.QMul, aP[:, x:x + 1])[0] % Pb
        for i, q in enumerate(lit.Q):
            assert int(down[i, x]) == ((vq - vp) * pow(Pb, -1, q)) % q


def test_div_round_by_last_modulus():
    lit = PR.CKKS_PN14QP439
    r = O.Ring(6, lit.Q)
    rng = np.random.default_rng(4)
    a = _rand_poly(rng, r.moduli, r.N)
    out = np.zeros_like(a)
    level = r.nmod - 1
    r.div_round_by_last_modulus_many(level, 1, a.copy(), out)
    ql = lit.Q[level]
    Qb = reduce(lambda x, y: x * y, lit.Q, 1)
    for x in range(r.N):
        v = O.crt_centered(lit.Q, a[:, x:x + 1])[0] % Qb
        ref = (v + (ql - 1) // 2) // ql
        for i in range(level):
            assert int(out[i, x]) == ref % lit.Q[i]


def _ckks_world(logN=11, k=2, lit=None):
    lit = (lit or PR.CKKS_PN14QP439).at_logn(logN)
    p = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, crs_rots=[1, 2])
    kg = O.KeyGenerator(p)
    sks, pks, rlks, rks, cks = {}, {}, {}, {}, {}
    for i in range(k):
        sk, r = kg.gen_secret_key(i), kg.gen_secret_key(i)
        sks[i], pks[i], rlks[i] = sk, kg.gen_public_key(sk), kg.gen_relin_key(sk, r)
        rks[i] = {1: kg.gen_rotation_key(1, sk), 2: kg.gen_rotation_key(2, sk)}
        cks[i] = kg.gen_conjugation_key(sk)
    return lit, p, sks, pks, rlks, rks, cks


@pytest.mark.parametrize("lit", [PR.CKKS_PN14QP439, PR.PN16QP1761_Q7, PR.PN16QP1761_Q7_ALPHA4], ids=lambda l: l.name)
def test_ckks_semantic_thresholds(lit):
    """the reference's own assertions on the oracle's outputs: Enc/Dec (mkckks_test.go:221), MulRelin square of a
    sum of k fresh ciphertexts (:357-358, +12), Rotate / Conjugate (+11).  PN16QP1761[:7] has alpha = 2: the general
    DecomposeAndSplit branch (exact lift of two-limb digits, a single-limb last digit) and the alpha-limb gadget of
    GenSwitchingKey must decrypt correctly together.  With gamma = 1 (alpha = 4: digits of four limbs, a partial digit of
    three) the relinearisation noise ~ digit^2 / P no longer fits -- which is why the reference hard-wires gamma = 2 -- so
    only the single key switches (Rotate, Conjugate) are held to their thresholds there."""
    lit, p, sks, pks, rlks, rks, cks = _ckks_world(lit=lit)
    assert p.alpha() == len(lit.P) // lit.gamma
    n = p.N // 2
    logslots = np.log2(n)
    logscale = np.log2(lit.scale)
    rng = np.random.default_rng(7)
    enc, dec, ev = O.Encryptor(p), O.Decryptor(p), O.CKKSEvaluator(p, lit.scale)
    msgs = [(rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)) / 2 for _ in range(2)]
    cts = [enc.encrypt(O.ckks_encode(p, msgs[i], lit.scale), pks[i], i, lit.scale) for i in range(2)]
    err = np.abs(O.ckks_decode(p, dec.decrypt(cts[0], sks), lit.scale) - msgs[0]).max()
    assert np.log2(err) <= -logscale + logslots + 8
    ct = O.Ciphertext({"0": p.ringQ.add(cts[0].value["0"], cts[1].value["0"]), 0: cts[0].value[0], 1: cts[1].value[1]}, lit.scale)
    msum = msgs[0] + msgs[1]
    res = ev.mul_relin_new(ct, ct, rlks)
    assert res.level() == ct.level() - 1
    if lit.gamma == 2:
        err = np.abs(O.ckks_decode(p, dec.decrypt(res, sks), res.scale) - msum * msum).max()
        assert np.log2(err) <= -logscale + logslots + 12
        res = ev.mul_relin_new(cts[0], cts[1], rlks)
        err = np.abs(O.ckks_decode(p, dec.decrypt(res, sks), res.scale) - msgs[0] * msgs[1]).max()
        assert np.log2(err) <= -logscale + logslots + 12
    # hoisted MulRelin with one nil hoisted operand (mkckks_test.go:506-550) equals the fully hoisted one
    h = ev.hoisted_form(ct)
    a = ev.mul_relin_hoisted_new(ct, ct, h, None, rlks)
    b = ev.mul_relin_hoisted_new(ct, ct, h, h, rlks)
    assert all(np.array_equal(a.value[k], b.value[k]) for k in a.value)
    rot = ev.rotate_hoisted_new(ct, 2, h, rks)
    err = np.abs(O.ckks_decode(p, dec.decrypt(rot, sks), rot.scale) - np.roll(msum, -2)).max()
    assert np.log2(err) <= -logscale + logslots + 11
    rot2 = ev.rotate_new(ct, 2, rks)
    assert all(np.array_equal(rot.value[k], rot2.value[k]) for k in rot.value)
    rot3 = ev.rotate_new(ct, 3, rks)
    err = np.abs(O.ckks_decode(p, dec.decrypt(rot3, sks), rot3.scale) - np.roll(msum, -3)).max()
    assert np.log2(err) <= -logscale + logslots + 11
    with pytest.raises(RuntimeError, match="Hoisted rotation only works"):
        ev.rotate_hoisted_new(ct, 3, h, rks)
    cj = ev.conjugate_new(ct, cks)
    err = np.abs(O.ckks_decode(p, dec.decrypt(cj, sks), cj.scale) - np.conj(msum)).max()
    assert np.log2(err) <= -logscale + logslots + 11


@pytest.mark.parametrize("level", ["max", 0])
def test_external_product_and_decompose_noise(level):
    """mkrlwe_test.go:456-505 (ExternalProductMaxLevel) and :507-610 (DecomposeMaxLevel / MinLevel) restated: with sg =
    GenSwitchingKey(sk), ExternalProduct(c, sg) and Decompose(c) . sg followed by ModDown both equal c * s up to noise whose
    inner sum (log2OfInnerSum, :92-150) stays below 2^(10 + logN)"""
    lit = PR.CKKS_PN14QP439.at_logn(12)
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000021, crs_rots=[])
    kg = O.KeyGenerator(p, seed=5)
    sk = kg.gen_secret_key(0)
    sg = kg.gen_switching_key(sk)
    L = p.max_level() if level == "max" else level
    c = O.PRNG(77).uniform(p.ringQ, L)
    ks = O.KeySwitcher(p)
    rq = p.ringQ
    cs = rq.intt(rq.mul_mont(rq.ntt(c, L), np.ascontiguousarray(sk.Q[:L + 1]), L), L)          # c * s

    def inner_sum_log2(poly):
        return sum(abs(v) for v in O.crt_centered(list(lit.Q[:L + 1]), poly)).bit_length()

    ext = ks.external_product(L, c, sg)
    assert inner_sum_log2(rq.sub(ext, cs, L)) <= 10 + lit.logN
    # Decompose, the digit-by-digit product accumulated over QP, ModDown: the same value as ExternalProduct, bit for bit
    ext2 = ks.external_product_hoisted(L, ks.decompose(L, c), sg)
    assert np.array_equal(ext, ext2)


def test_rotation_stores_negated_zero_as_q():
    """the permutation writes q - c unreduced, so c = 0 becomes q_j"""
    lit, p, sks, pks, rlks, rks, cks = _ckks_world(logN=10)
    r = p.ringQ
    z = r.new_poly()
    out = r.permute(z, p.galois_element_for_rotation(1))
    for j, q in enumerate(p.Q):
        vals = set(int(v) for v in out[j])
        assert vals == {0, q}


def test_bfv_exact_plaintext_equality():
    """mkbfv_test.go:412: exact equality after decrypt for MulRelin(ct, ct'); also the lazy multSum limbs are common"""
    lit = PR.BFV_PN14QP439.at_logn(10)
    p = O.BFVParams(lit.logN, lit.Q, lit.QMul, lit.P, lit.T)
    kg = O.BFVKeyGenerator(p)
    sks, pks, rlks = {}, {}, {}
    for i in range(2):
        sk, r = kg.gen_secret_key(i), kg.gen_secret_key(i)
        sks[i], pks[i], rlks[i] = sk, kg.gen_public_key(sk), kg.gen_bfv_relin_key(sk, r)
    rng = np.random.default_rng(3)
    N, T = p.N, p.T
    ms = [rng.integers(0, T, N) for _ in range(2)]
    enc, dec, ev = O.Encryptor(p), O.Decryptor(p), O.BFVEvaluator(p)
    cts = [enc.encrypt(O.bfv_encode(p, ms[i]), pks[i], i) for i in range(2)]
    assert np.array_equal(O.bfv_decode(p, dec.decrypt(cts[0], sks)), ms[0])

    def negacyclic(a, b):
        full = np.convolve(np.array([int(x) for x in a], dtype=object), np.array([int(x) for x in b], dtype=object))
        c = np.zeros(N, dtype=object)
        for i, v in enumerate(full):
            if i < N:
                c[i] += v
            else:
                c[i - N] -= v
        return np.array([int(v) % T for v in c], dtype=np.int64)

    res = ev.mul_relin_new(cts[0], cts[1], rlks)
    assert np.array_equal(O.bfv_decode(p, dec.decrypt(res, sks)), negacyclic(ms[0], ms[1]))
    res = ev.mul_relin_new(cts[0], cts[0], rlks)
    assert np.array_equal(O.bfv_decode(p, dec.decrypt(res, sks)), negacyclic(ms[0], ms[0]))
    R = ev.modup_q_to_r(cts[0].value["0"])
    frac = np.mean([(R[p.nQ + i] >= np.uint64(q)).mean() for i, q in enumerate(p.QMul)])
    assert 0.2 < frac < 0.9, "SURVEY App. A.3.2: roughly half of the lifted coefficients are >= p_j"


def test_golden_fixtures():
    """tests/golden/*.json freeze the oracle's outputs on seeded inputs (generated by tools/gen_golden.py)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "gen_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with open(os.path.join(GOLDEN, "oracle_digests.json")) as f:
        golden = json.load(f)
    got = gen.compute_digests()
    assert got == golden


def test_counter_based_samplers_have_the_reference_distributions():
    """include/mkhe_prng.h ("mkhe-ctr-1"), what the device-side key generation and encryption draw from: the distributions of
    lattigo's samplers as the reference configures them (keygen.go:35,58-60; params.go:32-33) -- uniform below q, ternary with
    P(0) = 1/2, rounded Gaussian sigma = 3.2 cut at 19 -- and statelessness (element j of a stream does not depend on the rest)."""
    lit = PR.CKKS_PN14QP439.at_logn(14)
    ring = O.Ring(lit.logN, lit.Q[:3])
    prng = O.CtrPRNG(0x1234, 7)
    N = ring.N
    u = prng.uniform(ring)
    for i, q in enumerate(lit.Q[:3]):
        assert u[i].max() < q
        assert abs(float(u[i].astype(np.float64).mean()) / q - 0.5) < 0.02
    assert prng.stream == 7 + 3
    assert np.array_equal(O.CtrPRNG(0x1234, 8).uniform(O.Ring(lit.logN, lit.Q[:3]))[0][:8] < np.uint64(lit.Q[0]), np.ones(8, bool))
    t = np.concatenate([prng.ternary(N) for _ in range(8)])
    assert set(np.unique(t)) == {-1, 0, 1}
    assert abs((t == 0).mean() - 0.5) < 0.01 and abs((t == 1).mean() - 0.25) < 0.01
    t3 = O.CtrPRNG(1, 0).ternary(8 * N, 1.0 / 3)
    assert abs((t3 == 0).mean() - 1.0 / 3) < 0.01
    g = np.concatenate([prng.gaussian(N) for _ in range(16)])
    assert np.abs(g).max() <= 19
    assert abs(g.mean()) < 0.05 and abs(g.std() - 3.2) < 0.05
    # counter-based: the same (seed, stream) gives the same polynomial, another stream an unrelated one
    a, b, c = O.CtrPRNG(5, 100).gaussian(N), O.CtrPRNG(5, 100).gaussian(N), O.CtrPRNG(5, 101).gaussian(N)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    # known answers: the committed constants ARE the specification (tools/gen_cdt.py regenerates the table)
    assert [int(x) for x in O.CtrPRNG(0, 0).gaussian(16)] == [-1, -1, 0, 4, 4, 1, -1, -4, 4, 0, 1, 4, 1, 2, -1, 0]
    assert [int(x) for x in O.CtrPRNG(0, 0).ternary(16)] == [0, 0, 0, 1, 1, 0, 0, -1, 1, 0, 0, 1, 0, 0, 0, 0]
    assert [int(x) for x in O.CtrPRNG(0, 0).uniform(O.Ring(12, [0x10000000006e0001]))[0][:4]] == [
        0x38275bc38fcbe91, 0xf101fe21496ea20, 0xb91752fd22fb56a, 0xda3e176f37bc9fb]


def test_cnn_inference_semantics():
    """BASELINE config 5, semantically: the oracle's keys, encryptions, evaluator and decryptor run the reference's CNN op sequence
    (cnn/cnn.go) on the reference's slot packing (cnn/cnn_test.go:345-545) and must reproduce the PLAIN forward pass of the same
    random network on the same random image -- rotation directions, packing, scale management and all (tests/cnn_semantic.py)."""
    import cnn_semantic as S
    got, want = S.run_oracle(PR.CNN_PN14QP433)
    assert np.abs(got - want).max() < 1e-4, (got, want)
    assert got.argmax() == want.argmax()

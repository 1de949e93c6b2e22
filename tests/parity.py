"""Shared parity harness: drives the C ABI (through the reference-shaped host mirror) and the CPU oracle
on the same seeded inputs and compares every output limb bit for bit.

Used by tests/test_gpu_parity.py (the real CUDA library, `-m gpu`) and by tests/test_emu_kernels.py
(the same kernel sources compiled for the CPU emulator, development aid).
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from mkhe_kklss_b200 import mkbfv, mkckks, mkrlwe
from mkhe_kklss_b200.params import ParamLiteral


def assert_same(a, b, what):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        first = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} words differ; first at {first}: "
                             f"device={int(a[first])} oracle={int(b[first])}; per-leading-index counts "
                             f"{np.unique(bad[:, 0], return_counts=True)}")


def uniform_poly(prng, ring, level):
    return prng.uniform(ring, level)


def uniform_swk(prng, p: O.MKParams):
    swk = p.new_swk()
    for i in range(swk.shape[0]):
        swk[i, :p.nQ] = prng.uniform(p.ringQ)
        swk[i, p.nQ:] = prng.uniform(p.ringP)
    return swk


class CKKSWorld:
    """oracle objects + their device twins for one parameter set with uniform-random key material"""

    def __init__(self, lit: ParamLiteral, nparties, seed=0xB2000001, lib=None, rots=(1, 2), real_keys=False):
        self.lit = lit
        self.op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, seed=seed, crs_rots=list(rots))
        self.prng = O.PRNG(seed ^ 0x5EED)
        self.dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma)
        self.ctx = self.dp.ctx
        for idx, arr in self.op.CRS.items():
            self.dp.SetCRS(idx, arr)
        self.oev = O.CKKSEvaluator(self.op, lit.scale)
        self.dev = mkckks.Evaluator(self.dp)
        self.ids = list(range(nparties))
        self.o_rlk, self.d_rlk = {}, mkrlwe.RelinearizationKeySet()
        self.o_rk, self.d_rk = {}, mkrlwe.RotationKeySet()
        self.o_ck, self.d_ck = {}, mkrlwe.ConjugationKeySet()
        self.sks, self.pks = {}, {}
        kg = O.KeyGenerator(self.op) if real_keys else None
        for i in self.ids:
            if real_keys:
                sk, r = kg.gen_secret_key(i), kg.gen_secret_key(i)
                self.sks[i], self.pks[i] = sk, kg.gen_public_key(sk)
                rl = kg.gen_relin_key(sk, r)
                rks = {rot: kg.gen_rotation_key(rot, sk) for rot in rots}
                ck = kg.gen_conjugation_key(sk)
            else:
                rl = O.RelinKey(i, uniform_swk(self.prng, self.op), uniform_swk(self.prng, self.op), uniform_swk(self.prng, self.op))
                rks = {rot: uniform_swk(self.prng, self.op) for rot in rots}
                ck = uniform_swk(self.prng, self.op)
            self.o_rlk[i] = rl
            self.d_rlk.AddRelinearizationKey(mkrlwe.RelinearizationKey(self.ctx, i, rl.b, rl.d, rl.v))
            self.o_rk[i] = rks
            for rot, a in rks.items():
                self.d_rk.AddRotationKey(i, rot, mkrlwe.SwitchingKey(self.ctx, a))
            self.o_ck[i] = ck
            self.d_ck.AddConjugationKey(i, mkrlwe.SwitchingKey(self.ctx, ck))

    def random_ct(self, ids, level, scale=None):
        """uniform-random ciphertext (throughput-style input) on both sides"""
        scale = self.lit.scale if scale is None else scale
        val = {"0": uniform_poly(self.prng, self.op.ringQ, level)}
        for i in ids:
            val[i] = uniform_poly(self.prng, self.op.ringQ, level)
        oct_ = O.Ciphertext({k: v.copy() for k, v in val.items()}, scale)
        dct = mkckks.Ciphertext.from_numpy(self.ctx, val, scale)
        return oct_, dct

    def compare_ct(self, dct, oct_, what):
        dv = dct.numpy()
        assert set(dv) == set(oct_.value), f"{what}: component sets differ {sorted(map(str, dv))} vs {sorted(map(str, oct_.value))}"
        for k in oct_.value:
            assert_same(dv[k], oct_.value[k], f"{what}[{k}]")

    def close(self):
        self.ctx.close()


# ---- individual checks (each returns nothing, raises AssertionError on mismatch) -------------------
def check_ntt(w: CKKSWorld):
    level = w.op.max_level()
    a = uniform_poly(w.prng, w.op.ringQ, level)
    pin = mkrlwe.Poly.from_numpy(w.ctx, a)
    pout = mkrlwe.Poly(w.ctx, level + 1)
    w.ctx.ntt(level, pin.h, pout.h)
    assert_same(pout.numpy(), w.op.ringQ.ntt(a), "NTTLvl")
    w.ctx.intt(level, pout.h, pin.h)
    assert_same(pin.numpy(), a, "InvNTTLvl(NTTLvl(a))")
    b = uniform_poly(w.prng, w.op.ringQ, level)
    pb = mkrlwe.Poly.from_numpy(w.ctx, b)
    w.ctx.intt(level, pb.h, pout.h)
    assert_same(pout.numpy(), w.op.ringQ.intt(b), "InvNTTLvl")


def check_decompose(w: CKKSWorld, level=None):
    level = w.op.max_level() if level is None else level
    a = uniform_poly(w.prng, w.op.ringQ, level)
    # non-canonical limbs as produced by the rotation quirk (App. A.3.1): a few coefficients equal to q
    a[0, :3] = np.uint64(w.op.Q[0])
    pa = mkrlwe.Poly.from_numpy(w.ctx, a)
    sk = mkrlwe.SwitchingKey(w.ctx)
    w.dev.ksw.Decompose(level, pa, sk)
    ref = w.oev.ksw.decompose(level, a)
    got = sk.numpy()
    beta = level + 1
    assert_same(got[:beta, :level + 1], ref[:beta, :level + 1], f"Decompose(level={level}) Q limbs")
    assert_same(got[:beta, w.op.nQ:], ref[:beta, w.op.nQ:], f"Decompose(level={level}) P limbs")


def check_moddown_ntt(w: CKKSWorld, level=None):
    """FastBasisExtender.ModDownQPtoQNTT (basis_extension.go:239-290)"""
    level = w.op.max_level() if level is None else level
    p = w.op
    a = np.concatenate([uniform_poly(w.prng, p.ringQ, p.max_level()), uniform_poly(w.prng, p.ringP, p.nP - 1)])
    pin = mkrlwe.Poly.from_numpy(w.ctx, a)
    pout = mkrlwe.Poly(w.ctx, level + 1)
    mkrlwe.FastBasisExtender(w.dp).ModDownQPtoQNTT(level, p.nP - 1, pin, pout)
    be = O.BasisExtender(p.ringQ, p.ringP)
    ref = be.moddown_qp_to_q_ntt(level, p.nP - 1, a[:level + 1], a[p.nQ:])
    assert_same(pout.numpy(), ref, f"ModDownQPtoQNTT(level={level})")
    # ... and it agrees with the coefficient-domain ModDownQPtoQ of the same value
    coef = be.moddown_qp_to_q(level, p.nP - 1, p.ringQ.intt(np.ascontiguousarray(a[:level + 1]), level), p.ringP.intt(np.ascontiguousarray(a[p.nQ:])))
    assert_same(p.ringQ.intt(ref, level), coef, "ModDownQPtoQNTT against ModDownQPtoQ")


def check_external_product(w: CKKSWorld, level=None):
    level = w.op.max_level() if level is None else level
    a = uniform_poly(w.prng, w.op.ringQ, level)
    bg = uniform_swk(w.prng, w.op)
    pa = mkrlwe.Poly.from_numpy(w.ctx, a)
    dbg = mkrlwe.SwitchingKey(w.ctx, bg)
    h = mkrlwe.SwitchingKey(w.ctx)
    w.dev.ksw.Decompose(level, pa, h)
    pc = mkrlwe.Poly(w.ctx, level + 1)
    w.dev.ksw.ExternalProductHoisted(level, h, dbg, pc)
    ref = w.oev.ksw.external_product_hoisted(level, w.oev.ksw.decompose(level, a), bg)
    assert_same(pc.numpy(), ref, f"ExternalProductHoisted(level={level})")
    pc2 = mkrlwe.Poly(w.ctx, level + 1)
    w.dev.ksw.ExternalProduct(level, pa, dbg, pc2)
    assert_same(pc2.numpy(), w.oev.ksw.external_product(level, a, bg), f"ExternalProduct(level={level})")


def check_mul_relin_hoisted(w: CKKSWorld, ids0, ids1, level=None, nil0=False, nil1=False, same=False):
    level = w.op.max_level() if level is None else level
    o0, d0 = w.random_ct(ids0, level)
    if same:
        o1, d1 = o0, d0
    else:
        o1, d1 = w.random_ct(ids1, level)
    oh0 = None if nil0 else w.oev.hoisted_form(o0)
    oh1 = None if nil1 else (oh0 if same and not nil0 else w.oev.hoisted_form(o1))
    dh0 = None if nil0 else w.dev.HoistedForm(d0)
    dh1 = None if nil1 else (dh0 if same and not nil0 else w.dev.HoistedForm(d1))
    oout = w.oev._new_binary(o0, o1)
    w.oev.ksw.mul_and_relin_hoisted(o0, o1, oh0, oh1, w.o_rlk, oout)
    dout = w.dev.newCiphertextBinary(d0, d1)
    w.dev.ksw.MulAndRelinHoisted(d0, d1, dh0, dh1, w.d_rlk, dout)
    w.compare_ct(dout, oout, f"MulAndRelinHoisted(ids0={ids0}, ids1={ids1}, level={level}, nil=({nil0},{nil1}), same={same})")


def check_mul_relin_new(w: CKKSWorld, ids0, ids1, level=None, same=False):
    level = w.op.max_level() if level is None else level
    o0, d0 = w.random_ct(ids0, level)
    if same:
        o1, d1 = o0, d0
    else:
        o1, d1 = w.random_ct(ids1, level)
    oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    dout = w.dev.MulRelinNew(d0, d1, w.d_rlk)
    assert dout.Level() == oout.level(), (dout.Level(), oout.level())
    assert dout.Scale == oout.scale, (dout.Scale, oout.scale)
    w.compare_ct(dout, oout, f"MulRelinNew(ids0={ids0}, ids1={ids1}, level={level}, same={same})")
    return dout, oout


def check_rescale(w: CKKSWorld, level=None, nb=1):
    level = w.op.max_level() if level is None else level
    a = uniform_poly(w.prng, w.op.ringQ, level)
    pin = mkrlwe.Poly.from_numpy(w.ctx, a)
    pout = mkrlwe.Poly(w.ctx, level + 1)
    w.ctx.rescale(level, nb, pin.h, pout.h)
    ain = a.copy()
    ref = np.zeros_like(a)
    w.op.ringQ.div_round_by_last_modulus_many(level, nb, ain, ref)
    assert pout.nlimbs() == level + 1 - nb
    assert_same(pout.numpy(), ref[:level + 1 - nb], f"DivRoundByLastModulusManyLvl(level={level}, nb={nb})")
    if nb == 1:
        # lattigo mutates the input's last limb
        assert_same(w.ctx.poly_download(pin.h, level + 1), ain, "Rescale input side effect")


def check_rotate(w: CKKSWorld, ids, rot, level=None, zero_component=False, hoisted=True):
    level = w.op.max_level() if level is None else level
    oct_, dct = w.random_ct(ids, level)
    if not hoisted:
        # rot without a CRS entry: RotateNew chains powers of two, RotateHoistedNew panics (mkckks/evaluator.go:615)
        oo2 = w.oev.rotate_new(oct_, rot, w.o_rk)
        do2 = w.dev.RotateNew(dct, rot, w.d_rk)
        w.compare_ct(do2, oo2, f"RotateNew(rot={rot}, chained)")
        try:
            w.dev.RotateHoistedNew(dct, rot, w.dev.HoistedForm(dct), w.d_rk)
        except RuntimeError as e:
            assert "Hoisted rotation only works for precomputed rotation keys" in str(e)
        else:
            raise AssertionError("RotateHoistedNew must panic for a rotation without CRS")
        return
    if zero_component:
        # quirk test: a zero component goes through the permutation as q_j
        oct_.value["0"][:] = 0
        w.ctx.poly_upload(dct.Value["0"].h, oct_.value["0"])
    oh = w.oev.hoisted_form(oct_)
    dh = w.dev.HoistedForm(dct)
    oo = w.oev.rotate_hoisted_new(oct_, rot, oh, w.o_rk)
    do = w.dev.RotateHoistedNew(dct, rot, dh, w.d_rk)
    w.compare_ct(do, oo, f"RotateHoistedNew(rot={rot}, level={level})")
    oo2 = w.oev.rotate_new(oct_, rot, w.o_rk)
    do2 = w.dev.RotateNew(dct, rot, w.d_rk)
    w.compare_ct(do2, oo2, f"RotateNew(rot={rot}, level={level})")


def check_lanes(w: CKKSWorld, rounds=3, level=None):
    """two evaluators on two lanes of one context (mkhe_ctx_fork): independent ops interleaved, then chains whose operands
    cross lanes -- read-after-write (a product made on one lane is rotated and squared on the other), write-after-read (an
    operand still being read on one lane is overwritten from the other) -- everything compared with the oracle"""
    level = w.op.max_level() if level is None else level
    ev = [w.dev, w.dev.ShallowCopy()]
    ids = w.ids
    rot = sorted(r for r in w.op.CRS if r > 0)[0]
    pend = []
    for r in range(rounds):
        for lane in (0, 1):
            o0, d0 = w.random_ct(ids, level)
            o1, d1 = w.random_ct(ids, level)
            dout = ev[lane].MulRelinNew(d0, d1, w.d_rlk)                 # independent ops, alternating lanes
            other = ev[1 - lane]
            drot = other.RotateNew(dout, rot, w.d_rk)                   # consumer on the OTHER lane (RAW)
            dsq = other.MulRelinNew(dout, dout, w.d_rlk) if dout.Level() > 0 else None
            # WAR: overwrite one operand from the other lane while the first product may still be reading it
            o2, _ = w.random_ct(ids, level)
            for kk, poly in d0.Value.items():
                other.ctx.poly_upload(poly.h, o2.value[kk])
            drot2 = other.RotateNew(d0, rot, w.d_rk)
            pend.append((o0, o1, o2, dout, drot, dsq, drot2))
    # asynchronous transfers across lanes: lane 0 uploads from pinned memory without waiting, lane 1 consumes at once (it has to
    # wait for the COPY, not for the call), then lane 0 overwrites the operand while lane 1's download of the result is in flight
    o3, _ = w.random_ct(ids, level)
    fresh = mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale)
    pins = {}
    for kk, poly in fresh.Value.items():
        pins[kk] = ev[0].ctx.host_alloc(o3.value[kk].shape)
        pins[kk][...] = o3.value[kk]
        ev[0].ctx.poly_upload_async(poly.h, pins[kk])
    drot3 = ev[1].RotateNew(fresh, rot, w.d_rk)
    back = {kk: ev[1].ctx.host_alloc(o3.value[kk].shape) for kk in drot3.Value}
    for kk, poly in drot3.Value.items():
        ev[1].ctx.poly_download_async(poly.h, back[kk])
    o4, _ = w.random_ct(ids, level)
    for kk, poly in fresh.Value.items():
        ev[0].ctx.poly_upload(poly.h, o4.value[kk])                      # write-after-read across lanes
    drot4 = ev[0].RotateNew(fresh, rot, w.d_rk)
    for e in ev:
        e.ctx.sync()
    want3 = w.oev.rotate_new(o3, rot, w.o_rk)
    for kk in want3.value:
        assert_same(back[kk], want3.value[kk], f"lanes: asynchronous upload on one lane, use + asynchronous download on the other [{kk}]")
    w.compare_ct(drot4, w.oev.rotate_new(o4, rot, w.o_rk), "lanes: overwrite after the other lane's read")
    for n, (o0, o1, o2, dout, drot, dsq, drot2) in enumerate(pend):
        oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
        w.compare_ct(dout, oout, f"lanes[{n}] MulRelinNew")
        w.compare_ct(drot, w.oev.rotate_new(oout, rot, w.o_rk), f"lanes[{n}] RotateNew of the other lane's product")
        if dsq is not None:
            w.compare_ct(dsq, w.oev.mul_relin_new(oout, oout, w.o_rlk), f"lanes[{n}] square of the other lane's product")
        w.compare_ct(drot2, w.oev.rotate_new(o2, rot, w.o_rk), f"lanes[{n}] RotateNew after a cross-lane overwrite")
    ev[1].ctx.close()


def check_mixed_levels(w: CKKSWorld):
    """alpha > 1 with operands at DIFFERENT levels: the reference hoists each operand at its own level
    (mkckks/evaluator.go:431-437), so the last digit in use is lifted from a different limb group than a decomposition at
    the output level would take"""
    L = w.op.max_level()
    for l0, l1 in ((L, L - 1), (L - 2, L), (L, L - 3)):
        o0, d0 = w.random_ct(w.ids, l0)
        o1, d1 = w.random_ct(w.ids, l1)
        oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
        dout = w.dev.MulRelinNew(d0, d1, w.d_rlk)
        assert dout.Level() == oout.level() and dout.Scale == oout.scale
        w.compare_ct(dout, oout, f"MulRelinNew at levels ({l0}, {l1}), alpha = {w.dp.Alpha()}")


def check_abi_negative(w: CKKSWorld):
    """a switching key holds beta_max = ceil(nQ / alpha) digits: the per-limb upload / download must reject digit >= beta_max
    (alpha > 1 sets have beta_max < nQ) and null host pointers instead of touching memory past the allocation; a transform
    over R = Q u QMul needs BFV parameters"""
    import ctypes as C
    ctx, dll = w.ctx, w.ctx.dll
    alpha, nQ = w.dp.Alpha(), len(w.lit.Q)
    beta_max = (nQ + alpha - 1) // alpha
    assert alpha > 1 and beta_max < nQ
    sk = mkrlwe.SwitchingKey(ctx)
    buf = np.zeros(w.lit.N, dtype=np.uint64)
    ptr = buf.ctypes.data_as(C.POINTER(C.c_uint64))
    h = C.c_uint64(sk.h)
    assert dll.mkhe_swk_upload_limb(ctx.ptr, h, C.c_int(beta_max - 1), C.c_int(0), C.c_int(0), ptr) == 0
    for digit in (beta_max, nQ - 1, nQ, -1):
        assert dll.mkhe_swk_upload_limb(ctx.ptr, h, C.c_int(digit), C.c_int(0), C.c_int(0), ptr) != 0, digit
        assert dll.mkhe_swk_download_limb(ctx.ptr, h, C.c_int(digit), C.c_int(0), C.c_int(0), ptr) != 0, digit
    assert dll.mkhe_swk_upload_limb(ctx.ptr, h, C.c_int(0), C.c_int(0), C.c_int(0), None) != 0
    assert dll.mkhe_swk_download_limb(ctx.ptr, h, C.c_int(0), C.c_int(0), C.c_int(0), None) != 0
    pq = mkrlwe.Poly(ctx, nQ + 1)
    assert dll.mkhe_ntt(ctx.ptr, C.c_int(nQ), C.c_uint64(pq.h), C.c_uint64(pq.h)) != 0


def check_fork_after_queued_work(w: CKKSWorld, rounds=4):
    """a lane forked AFTER work was queued on the root is ordered behind that work: the product of an asynchronous MulRelinNew
    on the root is consumed at once by the new lane (no synchronisation in between)"""
    L = w.op.max_level()
    rot = sorted(r for r in w.op.CRS if r > 0)[0]
    pend = []
    for r in range(rounds):
        o0, d0 = w.random_ct(w.ids, L)
        o1, d1 = w.random_ct(w.ids, L)
        dout = w.dev.MulRelinNew(d0, d1, w.d_rlk)        # queued on the root, still running
        ev2 = w.dev.ShallowCopy() if r < 3 else pend[0][4]   # fork now (at most 4 lanes per root)
        drot = ev2.RotateNew(dout, rot, w.d_rk)          # first use of the pending output on the new lane
        pend.append((o0, o1, dout, drot, ev2))
    for o0, o1, dout, drot, ev2 in pend:
        oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
        w.compare_ct(dout, oout, "MulRelinNew before the fork")
        w.compare_ct(drot, w.oev.rotate_new(oout, rot, w.o_rk), "RotateNew on a lane forked after the product was queued")
    for *_, ev2 in pend[:3]:
        ev2.ctx.close()


def check_limb_sharded(lit, nparties, nranks, lib=None, level=None, rounds=2, ids0=None, ids1=None, seed=0xB2000061):
    """limb-sharded MulRelinNew (mkhe_ckks_mul_relin_limbs) on `nranks` ranks living in ONE process on ONE device (every rank is a
    context of its own; mkhe_team_join_local): every rank's result is the oracle's whole result, bit for bit.  The ops of all
    ranks are enqueued without any host synchronisation, op after op: the in-kernel barriers are the only ordering."""
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, seed=seed, crs_rots=[])
    prng = O.PRNG(seed ^ 0x7EA3)
    oev = O.CKKSEvaluator(op, lit.scale)
    ids = list(range(nparties))
    ids0 = ids if ids0 is None else ids0
    ids1 = ids if ids1 is None else ids1
    o_rlk = {i: O.RelinKey(i, uniform_swk(prng, op), uniform_swk(prng, op), uniform_swk(prng, op)) for i in ids}
    level = op.max_level() if level is None else level
    ranks = []
    for r in range(nranks):
        dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma)
        dp.SetCRS(-1, op.CRS[-1])
        rl = mkrlwe.RelinearizationKeySet()
        for i in ids:
            rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, o_rlk[i].b, o_rlk[i].d, o_rlk[i].v))
        ranks.append((dp, rl))
    ctxs = [dp.ctx for dp, _ in ranks]
    for r, c in enumerate(ctxs):
        c.team_join_local(nparties, r, ctxs)
    # Everything that allocates happens BEFORE the first sharded op: a device memory allocation is an implicit synchronisation point
    # between the streams of one process (CUDA programming guide, "Implicit Synchronization"), and with all ranks in one process a
    # spinning barrier of one rank would then sit in front of the other ranks' kernels.  (Ranks in processes of their own, one per
    # GPU, do not share a device.)  The warm-up op sizes every scratch pool of the context.
    want, ops = [], []
    evs = [mkckks.Evaluator(dp) for dp, _ in ranks]
    for n in range(rounds):
        val0 = {"0": uniform_poly(prng, op.ringQ, level), **{i: uniform_poly(prng, op.ringQ, level) for i in ids0}}
        val1 = {"0": uniform_poly(prng, op.ringQ, level), **{i: uniform_poly(prng, op.ringQ, level) for i in ids1}}
        o0, o1 = O.Ciphertext({k: v.copy() for k, v in val0.items()}, lit.scale), O.Ciphertext({k: v.copy() for k, v in val1.items()}, lit.scale)
        want.append(oev.mul_relin_new(o0, o1, o_rlk))
        row = []
        for (dp, rl), ev in zip(ranks, evs):
            d0 = mkckks.Ciphertext.from_numpy(dp.ctx, val0, lit.scale)
            d1 = mkckks.Ciphertext.from_numpy(dp.ctx, val1, lit.scale)
            row.append((d0, d1, ev.newCiphertextBinary(d0, d1)))
        ops.append(row)
    for (dp, rl), ev, (d0, d1, _) in zip(ranks, evs, ops[0]):
        ev.MulRelinNew(d0, d1, rl).free()                          # warm-up on one rank alone (not sharded)
        dp.ctx.sync()
    outs = []
    for row in ops:
        res = []
        for (dp, rl), ev, (d0, d1, dout) in zip(ranks, evs, row):   # enqueue on every rank, op after op, no synchronisation
            res.append(ev.MulRelinLimbSharded(d0, d1, rl, dout))
        outs.append(res)
    for c in ctxs:
        c.sync()
    for c in ctxs:
        assert not c.team_timed_out(), f"a team barrier timed out; flags per rank: {[x.team_flags() for x in ctxs]}"
    for n, row in enumerate(outs):
        for r, dout in enumerate(row):
            assert dout.Level() == want[n].level() and dout.Scale == want[n].scale
            dv = dout.numpy()
            assert set(dv) == set(want[n].value)
            for k in want[n].value:
                assert_same(dv[k], want[n].value[k], f"limb-sharded MulRelinNew, {nranks} ranks, op {n}, rank {r} [{k}]")
    for c in ctxs:
        c.close()


def check_team_allgather(lit, nranks, npolys=3, lib=None, level=None, rounds=3):
    """mkhe_team_allgather: every rank uploads only the limbs it owns (limb mod nranks == rank), afterwards every rank holds
    every limb -- ranks = contexts on one device, several gathers back to back (the two staging areas alternate)"""
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, seed=0xB2000071, crs_rots=[])
    prng = O.PRNG(0xA11)
    level = op.max_level() if level is None else level
    dps = [mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma) for _ in range(nranks)]
    ctxs = [dp.ctx for dp in dps]
    for r, c in enumerate(ctxs):
        c.team_join_local(max(npolys, 2), r, ctxs)
    data = [[uniform_poly(prng, op.ringQ, level) for _ in range(npolys)] for _ in range(rounds)]
    polys = [[[mkrlwe.Poly(c, level + 1) for _ in range(npolys)] for c in ctxs] for _ in range(rounds)]
    for n in range(rounds):
        for r, c in enumerate(ctxs):
            for i in range(npolys):
                junk = np.full((level + 1, lit.N), 0xDEAD0000 + r, dtype=np.uint64)
                c.poly_upload(polys[n][r][i].h, junk)
                for j in range(level + 1):
                    assert c.team_owns_limb(j) == (j % nranks == r)
                    if c.team_owns_limb(j):
                        c.poly_upload_limb(polys[n][r][i].h, j, data[n][i][j])
    for n in range(rounds):
        for r, c in enumerate(ctxs):                    # enqueued rank after rank, no synchronisation
            c.team_allgather(level, [p.h for p in polys[n][r]])
    for c in ctxs:
        c.sync()
    for n in range(rounds):
        for r, c in enumerate(ctxs):
            for i in range(npolys):
                assert_same(c.poly_download(polys[n][r][i].h, level + 1), data[n][i], f"allgather round {n} rank {r} poly {i}")
    for c in ctxs:
        c.close()


def check_elementwise(w: CKKSWorld):
    """the evaluator ops either side of the key switches: AddNew / SubNew over different id sets, levels
    and scales (scale alignment through MultByConst), MultByConst with integer / fractional / negative / complex constants,
    MulPtxtNew, DropLevelNew -- bit for bit against the oracle"""
    L = w.op.max_level()
    ids = w.ids
    cases = [(ids, ids, L, L, 1.0, 1.0), (ids[:1], ids[1:], L, L, 1.0, 1.0), (ids[:1], ids, L, L - 1, 1.0, 1.0),
             (ids, ids[1:], L - 1, L, 1.0, 1.0), (ids, ids, L, L, 1.0, 7.3), (ids[:1], ids, L, L, 5.0, 1.0),
             (ids, ids, L, L, 1.0, 1.9)]
    for n, (i0, i1, l0, l1, f0, f1) in enumerate(cases):
        o0, d0 = w.random_ct(i0, l0, w.lit.scale * f0)
        o1, d1 = w.random_ct(i1, l1, w.lit.scale * f1)
        if n == 0:      # non-canonical inputs as rotations produce them (q stored for a negated zero)
            o0.value["0"][:, :7] = np.array(w.op.Q[:l0 + 1], dtype=np.uint64)[:, None]
            w.ctx.poly_upload(d0.Value["0"].h, o0.value["0"])
        for sub in (False, True):
            oo = (w.oev.sub_new if sub else w.oev.add_new)(o0, o1)
            do = (w.dev.SubNew if sub else w.dev.AddNew)(d0, d1)
            assert do.Level() == oo.level() and do.Scale == oo.scale, (n, sub, do.Level(), oo.level(), do.Scale, oo.scale)
            w.compare_ct(do, oo, f"{'SubNew' if sub else 'AddNew'} case {n}")
    for const in (3, -5, 2.5, -0.37, 1e-3, complex(0.5, -1.25), complex(0, 2), float(2 ** 40) + 0.5, 0):
        o0, d0 = w.random_ct(ids, L)
        oo = w.oev.new_ciphertext(ids, L, 0.0)
        do = mkckks.Ciphertext.new(w.dp, ids, L, 0.0)
        w.oev.mult_by_const(o0, const, oo)
        w.dev.MultByConst(d0, const, do)
        assert do.Scale == oo.scale, (const, do.Scale, oo.scale)
        w.compare_ct(do, oo, f"MultByConst({const})")
    w.oev.mult_by_const(o0, -7, o0)                      # in place
    w.dev.MultByConst(d0, -7, d0)
    w.compare_ct(d0, o0, "MultByConst in place")
    for lvl in (L, 1):
        o0, d0 = w.random_ct(ids, lvl)
        pt = uniform_poly(w.prng, w.op.ringQ, lvl)
        dpt = mkrlwe.Poly.from_numpy(w.ctx, pt)
        oo = w.oev.mul_ptxt_new(o0, pt, w.lit.scale)
        do = w.dev.MulPtxtNew(d0, dpt, w.lit.scale)
        assert do.Level() == oo.level() and do.Scale == oo.scale
        w.compare_ct(do, oo, f"MulPtxtNew(level={lvl})")
    o0, d0 = w.random_ct(ids, L)
    dd = w.dev.DropLevelNew(d0, 2)
    w.oev.drop_level(o0, 2)
    w.compare_ct(dd, o0, "DropLevelNew")


def check_cnn_flow(lit, lib=None):
    """the reference's encrypted-CNN op sequence (tests/cnn_flow.py) on the oracle and on the device, compared bit for bit"""
    import cnn_flow as F
    inp = F.make_inputs(lit)
    oout, omid = F.run_oracle(lit, inp)
    dout, dmid, fac, ctx = F.run_device(lit, inp, lib)
    for name in omid:
        for k in omid[name].value:
            assert_same(dmid[name].numpy()[k], omid[name].value[k], f"cnn {name}[{k}]")
    assert dout.Level() == oout.level() == 0 and dout.Scale == oout.scale, (dout.Level(), oout.level(), dout.Scale, oout.scale)
    dv = dout.numpy()
    assert set(dv) == set(oout.value)
    for k in oout.value:
        assert_same(dv[k], oout.value[k], f"cnn fc2Out[{k}]")
    assert fac.counts == {"HoistedForm": 27, "MulRelinHoistedNew": 14, "MulRelinNew": 1, "RotateHoistedNew": 11, "RotateNew": 19,
                          "AddNew": 31, "MulPtxtNew": 1}, fac.counts
    ctx.close()


def check_conjugate(w: CKKSWorld, ids, level=None):
    level = w.op.max_level() if level is None else level
    oct_, dct = w.random_ct(ids, level)
    oo = w.oev.conjugate_new(oct_, w.o_ck)
    do = w.dev.ConjugateNew(dct, w.d_ck)
    w.compare_ct(do, oo, f"ConjugateNew(level={level})")


def check_ckks_semantics(w: CKKSWorld):
    """T3: decrypt the DEVICE result and apply the reference's own precision threshold
    (mkckks_test.go:357-358: log2(err) <= -logScale + logSlots + 12)."""
    assert w.sks, "needs real keys"
    p, lit = w.op, w.lit
    n = p.N // 2
    rng = np.random.default_rng(11)
    enc, dec = O.Encryptor(p), O.Decryptor(p)
    k = len(w.ids)
    msgs, cts = [], []
    for i in w.ids:
        m = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)) / k
        msgs.append(m)
        cts.append(enc.encrypt(O.ckks_encode(p, m, lit.scale), w.pks[i], i, lit.scale))
    acc = cts[0]
    for c in cts[1:]:
        ids = sorted(set(acc.ids()) | set(c.ids()))
        nv = {"0": p.ringQ.add(acc.value["0"], c.value["0"])}
        for i in ids:
            nv[i] = acc.value[i].copy() if i in acc.value else c.value[i].copy()
        acc = O.Ciphertext(nv, lit.scale)
    msum = sum(msgs)
    dct = mkckks.Ciphertext.from_numpy(w.ctx, acc.value, lit.scale)
    dres = w.dev.MulRelinNew(dct, dct, w.d_rlk)
    res = O.Ciphertext(dres.numpy(), dres.Scale)
    got = O.ckks_decode(p, dec.decrypt(res, w.sks), res.scale)
    err = np.abs(got - msum * msum).max()
    bound = 2.0 ** (-np.log2(lit.scale) + np.log2(n) + 12)
    assert err <= bound, f"MulRelin precision {np.log2(err):.1f} > {np.log2(bound):.1f}"
    ores = w.oev.mul_relin_new(acc, acc, w.o_rlk)
    for kk in ores.value:
        assert_same(res.value[kk], ores.value[kk], f"semantic MulRelinNew[{kk}]")
    # hoisted rotation by 2 (mkckks_test.go:552-598: +11)
    dh = w.dev.HoistedForm(dct)
    drot = w.dev.RotateHoistedNew(dct, 2, dh, w.d_rk)
    rr = O.Ciphertext(drot.numpy(), drot.Scale)
    got = O.ckks_decode(p, dec.decrypt(rr, w.sks), rr.scale)
    err = np.abs(got - np.roll(msum, -2)).max()
    bound = 2.0 ** (-np.log2(lit.scale) + np.log2(n) + 11)
    assert err <= bound, f"Rotate precision {np.log2(err):.1f} > {np.log2(bound):.1f}"
    # Decrypt on the device (mkrlwe/decryptor.go:48-66): the plaintext poly is the oracle's, bit for bit
    dsk = {i: mkrlwe.Poly.from_numpy(w.ctx, w.sks[i].Q) for i in w.ids}
    ddec = mkrlwe.Decryptor(w.dp)
    for dc in (dres, drot, dct):
        got_pt = ddec.Decrypt(dc, dsk).numpy()
        want_pt = dec.decrypt(O.Ciphertext(dc.numpy(), dc.Scale), w.sks)
        assert_same(got_pt, want_pt, "Decrypt")
    try:
        ddec.Decrypt(dct, {w.ids[0]: dsk[w.ids[0]]} if len(w.ids) > 1 else {})
    except RuntimeError as e:
        assert "missing secretkey" in str(e)
    else:
        raise AssertionError("Decrypt must panic when a secret key is missing")
    # element-wise ops on device outputs, the reference's thresholds (mkckks_test.go:228-318: Add / Sub +11; the product
    # thresholds +12 for the plaintext product and the constant multiples)
    bound = 2.0 ** (-np.log2(lit.scale) + np.log2(n) + 12)

    def dec_dev(dc):
        r = O.Ciphertext(dc.numpy(), dc.Scale)
        return O.ckks_decode(p, dec.decrypt(r, w.sks), r.scale)

    d0 = mkckks.Ciphertext.from_numpy(w.ctx, cts[0].value, lit.scale)
    d1 = mkckks.Ciphertext.from_numpy(w.ctx, cts[-1].value, lit.scale)
    assert np.abs(dec_dev(w.dev.AddNew(d0, d1)) - (msgs[0] + msgs[-1])).max() <= bound, "AddNew precision"
    assert np.abs(dec_dev(w.dev.SubNew(d0, d1)) - (msgs[0] - msgs[-1])).max() <= bound, "SubNew precision"
    # real constants only: for a complex one the reference applies lattigo's NTT-domain formula (first / second half of the
    # SLOTS) to a coefficient-domain ciphertext (evaluator.go:123-127 vs the InvNTT'd ciphertexts of this library), which is
    # reproduced bit for bit (check_elementwise) but is not a multiplication by that constant
    for const in (3, -2.5):
        dm = mkckks.Ciphertext.new(w.dp, d0.IDSet(), d0.Level(), 0.0)
        w.dev.MultByConst(d0, const, dm)
        if dm.Scale > lit.scale * 2:                       # a scaled constant: bring the scale back like a caller would
            assert w.dev.Rescale(dm, lit.scale, dm) is None
        assert np.abs(dec_dev(dm) - const * msgs[0]).max() <= bound * 8, f"MultByConst({const}) precision"
    # scale alignment: the product (scale ~ Delta after Rescale... here: not rescaled, Delta^2) plus a fresh ciphertext
    big = mkckks.Ciphertext.new(w.dp, d0.IDSet(), d0.Level(), 0.0)
    w.dev.MultByConst(d0, 1000, big)                       # same scale, message * 1000
    big.Scale = lit.scale * 1000                           # ... reinterpreted: message at scale 1000 * Delta
    aligned = w.dev.AddNew(big, d1)                        # d1 is multiplied by floor(1000) first
    assert aligned.Scale == lit.scale * 1000
    assert np.abs(dec_dev(aligned) - (msgs[0] + msgs[-1])).max() <= bound, "AddNew with scale alignment"
    mp = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    dpt = mkrlwe.Poly.from_numpy(w.ctx, O.ckks_encode(p, mp, lit.scale))
    assert np.abs(dec_dev(w.dev.MulPtxtNew(d0, dpt, lit.scale)) - mp * msgs[0]).max() <= bound, "MulPtxtNew precision"


# ---- key generation and encryption on the device ----------------------------------
KG_SEED = 0x6B65795F67656E21


def _ctr_crs(op: O.MKParams, dp, idxs):
    """the same CRS on both sides from the counter-based streams (Parameters.AddCRS); compared bit for bit"""
    for idx in idxs:
        op.CRS.pop(idx, None)
        op.add_crs(idx, prng=O.CtrPRNG(KG_SEED, (idx & 0xFFFFFFFF) * dp.CRS_STREAM_STRIDE))
        assert_same(dp.AddCRS(idx, KG_SEED).numpy(), op.CRS[idx], f"AddCRS({idx})")


def check_keygen(lit: ParamLiteral, lib=None, nparties=2, rots=(1, 2), semantics=True):
    """every key mkrlwe.KeyGenerator makes + Encryptor.Encrypt, device against the oracle's KeyGenerator / Encryptor drawing from
    the same streams; then the whole flow on device-made material only: encrypt -> MulRelinNew -> RotateHoistedNew -> Decrypt
    with the reference's precision thresholds (mkckks_test.go:357-358, 552-598)."""
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, crs_rots=list(rots))
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma)
    try:
        _ctr_crs(op, dp, [0, -1, -2] + list(rots))
        okg = O.KeyGenerator(op, prng=O.CtrPRNG(KG_SEED, 1 << 40))
        dkg = mkrlwe.KeyGenerator(dp, KG_SEED, 1 << 40)
        d_rlk, d_rk, d_ck = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet(), mkrlwe.ConjugationKeySet()
        o_sk, d_sk, o_pk, d_pk, o_rlk, o_rk = {}, {}, {}, {}, {}, {}
        for i in range(nparties):
            osk, orr = okg.gen_secret_key(i), okg.gen_secret_key(i)
            dsk, drr = dkg.GenSecretKey(i), dkg.GenSecretKey(i)
            assert_same(dsk.Value.numpy(), okg.sk_qp(osk), f"GenSecretKey({i})")
            assert_same(drr.Value.numpy(), okg.sk_qp(orr), f"GenSecretKey(r{i})")
            opk, dpk = okg.gen_public_key(osk), dkg.GenPublicKey(dsk)
            assert_same(dpk.Value[0].numpy(), opk[0], f"GenPublicKey({i})[0]")
            assert_same(dpk.Value[1].numpy(), opk[1], f"GenPublicKey({i})[1]")
            orl, drl = okg.gen_relin_key(osk, orr), dkg.GenRelinearizationKey(dsk, drr)
            for name, dv, ov in zip("bdv", drl.Value, (orl.b, orl.d, orl.v)):
                assert_same(dv.numpy(), ov, f"GenRelinearizationKey({i}).{name}")
            d_rlk.AddRelinearizationKey(drl)
            o_rk[i] = {}
            for rot in rots:
                ork, drk = okg.gen_rotation_key(rot, osk), dkg.GenRotationKey(rot, dsk)
                assert_same(drk.numpy(), ork, f"GenRotationKey({rot}, {i})")
                d_rk.AddRotationKey(i, rot, drk)
                o_rk[i][rot] = ork
            ock, dck = okg.gen_conjugation_key(osk), dkg.GenConjugationKey(dsk)
            assert_same(dck.numpy(), ock, f"GenConjugationKey({i})")
            d_ck.AddConjugationKey(i, dck)
            oswk, dswk = okg.gen_switching_key(osk), mkrlwe.SwitchingKey(dp.ctx)
            dkg.GenSwitchingKey(dsk, dswk)
            assert_same(dswk.numpy(), oswk, f"GenSwitchingKey({i})")
            o_sk[i], d_sk[i], o_pk[i], d_pk[i], o_rlk[i] = osk, dsk, opk, dpk, orl
        # Encrypt: fresh ciphertexts at two levels, with and without a plaintext
        oenc, denc = O.Encryptor(op, prng=O.CtrPRNG(KG_SEED, 1 << 41)), mkrlwe.Encryptor(dp, KG_SEED, 1 << 41)
        n = op.N // 2
        rng = np.random.default_rng(5)
        msgs, octs, dcts = [], [], []
        for i in range(nparties):
            level = op.max_level() if i == 0 or semantics else max(op.max_level() - 1, 1)
            m = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)) / nparties
            pt = O.ckks_encode(op, m, lit.scale, level)
            oc = oenc.encrypt(pt, o_pk[i], i, lit.scale)
            dc = mkckks.Ciphertext.new(dp, [i], level, lit.scale)
            denc.Encrypt(mkrlwe.Poly.from_numpy(dp.ctx, pt), d_pk[i], dc)
            for kk in oc.value:
                assert_same(dc.Value[kk].numpy(), oc.value[kk], f"Encrypt(party {i})[{kk}]")
            msgs.append(m); octs.append(oc); dcts.append(dc)
        zero_pt = np.zeros((op.max_level() + 1, op.N), dtype=np.uint64)
        oz = oenc.encrypt(zero_pt, o_pk[0], 0)
        dz = mkckks.Ciphertext.new(dp, [0], op.max_level(), lit.scale)
        denc.Encrypt(None, d_pk[0], dz)
        for kk in oz.value:
            assert_same(dz.Value[kk].numpy(), oz.value[kk], f"Encrypt(zero)[{kk}]")
        if not semantics:
            return
        # the flow on device-made material only
        dev = mkckks.Evaluator(dp)
        acc = dcts[0]
        for c in dcts[1:]:
            acc = dev.AddNew(acc, c)
        msum = sum(msgs)
        dres = dev.MulRelinNew(acc, acc, d_rlk)
        ddec = mkrlwe.Decryptor(dp)
        got = O.ckks_decode(op, ddec.Decrypt(dres, d_sk).numpy(), dres.Scale)
        err, bound = np.abs(got - msum * msum).max(), 2.0 ** (-np.log2(lit.scale) + np.log2(n) + 12)
        assert err <= bound, f"MulRelin precision on device-made keys {np.log2(err):.1f} > {np.log2(bound):.1f}"
        drot = dev.RotateHoistedNew(acc, rots[-1], dev.HoistedForm(acc), d_rk)
        got = O.ckks_decode(op, ddec.Decrypt(drot, d_sk).numpy(), drot.Scale)
        err, bound = np.abs(got - np.roll(msum, -rots[-1])).max(), 2.0 ** (-np.log2(lit.scale) + np.log2(n) + 11)
        assert err <= bound, f"Rotate precision on device-made keys {np.log2(err):.1f} > {np.log2(bound):.1f}"
        dcj = dev.ConjugateNew(acc, d_ck)
        got = O.ckks_decode(op, ddec.Decrypt(dcj, d_sk).numpy(), dcj.Scale)
        err = np.abs(got - np.conj(msum)).max()
        assert err <= bound, f"Conjugate precision on device-made keys {np.log2(err):.1f} > {np.log2(bound):.1f}"
    finally:
        dp.ctx.close()


def _negacyclic_mod(a, b, T):
    N = len(a)
    full = np.convolve(np.array([int(x) for x in a], dtype=object), np.array([int(x) for x in b], dtype=object))
    ref = np.zeros(N, dtype=object)
    ref[:N] += full[:N]
    ref[:len(full) - N] -= full[N:]
    return np.array([int(v) % T for v in ref], dtype=np.int64)


def check_bfv_keygen(lit: ParamLiteral, lib=None, nparties=2):
    """mkbfv.KeyGenerator.GenRelinearizationKey (mkbfv/keygen.go:24-162) on the device against the oracle, then an exact BFV product
    on device-made keys and encryptions (mkbfv_test.go: plaintexts multiply exactly modulo T)"""
    op = O.BFVParams(lit.logN, lit.Q, lit.QMul, lit.P, lit.T)
    dp = mkbfv.Parameters(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, lib=lib)
    try:
        _ctr_crs(op, dp, [0, -1, -3])
        okg = O.BFVKeyGenerator(op, prng=O.CtrPRNG(KG_SEED, 1 << 40))
        dkg = mkbfv.KeyGenerator(dp, KG_SEED, 1 << 40)
        d_rlk, d_sk, d_pk = mkbfv.RelinearizationKeySet(), {}, {}
        for i in range(nparties):
            osk, orr = okg.gen_secret_key(i), okg.gen_secret_key(i)
            dsk, drr = dkg.GenSecretKey(i), dkg.GenSecretKey(i)
            opk, dpk = okg.gen_public_key(osk), dkg.GenPublicKey(dsk)
            assert_same(dpk.Value[0].numpy(), opk[0], f"GenPublicKey({i})[0]")
            orl, drl = okg.gen_bfv_relin_key(osk, orr), dkg.GenRelinearizationKey(dsk, drr)
            for name in ("b1", "b2", "d1", "d2", "v"):
                assert_same(getattr(drl, name).numpy(), getattr(orl, name), f"mkbfv GenRelinearizationKey({i}).{name}")
            d_rlk.AddRelinearizationKey(drl)
            d_sk[i], d_pk[i] = dsk, dpk
        denc = mkrlwe.Encryptor(dp, KG_SEED, 1 << 41)
        rng = np.random.default_rng(9)
        ms = [rng.integers(0, lit.T, op.N) for _ in range(2)]
        cts = []
        for j in range(2):
            i = j % nparties
            c = mkrlwe.Ciphertext.new(dp.ctx, [i], op.max_level())
            denc.Encrypt(mkrlwe.Poly.from_numpy(dp.ctx, O.bfv_encode(op, ms[j])), d_pk[i], c)
            cts.append(c)
        dev = mkbfv.Evaluator(dp)
        res = dev.MulRelinNew(cts[0], cts[1], d_rlk)
        got = O.bfv_decode(op, mkrlwe.Decryptor(dp).Decrypt(res, d_sk).numpy())
        assert np.array_equal(got, _negacyclic_mod(ms[0], ms[1], lit.T)), "BFV product on device-made keys"
    finally:
        dp.ctx.close()


# ---- BFV --------------------------------------------------------------------------------------------
class BFVWorld:
    def __init__(self, lit: ParamLiteral, nparties, seed=0xB2000003, lib=None, real_keys=False):
        self.lit = lit
        self.op = O.BFVParams(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, seed=seed)
        self.prng = O.PRNG(seed ^ 0xBF5EED)
        self.dp = mkbfv.Parameters(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, lib=lib)
        self.ctx = self.dp.ctx
        for idx, arr in self.op.CRS.items():
            self.dp.SetCRS(idx, arr)
        self.oev = O.BFVEvaluator(self.op)
        self.dev = mkbfv.Evaluator(self.dp)
        self.ids = list(range(nparties))
        self.o_rlk, self.d_rlk = {}, mkbfv.RelinearizationKeySet()
        self.sks, self.pks = {}, {}
        kg = O.BFVKeyGenerator(self.op) if real_keys else None
        for i in self.ids:
            if real_keys:
                sk, r = kg.gen_secret_key(i), kg.gen_secret_key(i)
                self.sks[i], self.pks[i] = sk, kg.gen_public_key(sk)
                rl = kg.gen_bfv_relin_key(sk, r)
            else:
                u = lambda: uniform_swk(self.prng, self.op)
                rl = O.BFVRelinKey(i, u(), u(), u(), u(), u())
            self.o_rlk[i] = rl
            self.d_rlk.AddRelinearizationKey(mkbfv.RelinearizationKey(self.ctx, i, rl.b1, rl.d1, rl.v, rl.b2, rl.d2))

    def random_ct(self, ids):
        level = self.op.max_level()
        val = {"0": uniform_poly(self.prng, self.op.ringQ, level)}
        for i in ids:
            val[i] = uniform_poly(self.prng, self.op.ringQ, level)
        return O.Ciphertext({k: v.copy() for k, v in val.items()}), mkrlwe.Ciphertext.from_numpy(self.ctx, val)

    def close(self):
        self.ctx.close()


def check_bfv_conv(w: BFVWorld):
    p = w.op
    a = uniform_poly(w.prng, p.ringQ, p.max_level())
    pq = mkrlwe.Poly.from_numpy(w.ctx, a)
    pr = mkrlwe.Poly(w.ctx, 2 * p.nQ)
    w.dev.conv.ModUpQtoR(pq, pr)
    refR = w.oev.modup_q_to_r(a)
    assert_same(pr.numpy(), refR, "ModUpQtoR (lazy multSum limbs must match exactly)")
    pr2 = mkrlwe.Poly(w.ctx, 2 * p.nQ)
    w.dev.conv.Rescale(pq, pr2)
    refR2 = w.oev.rescale_q_to_r(a)
    assert_same(pr2.numpy(), refR2, "Rescale Q->R")
    # Quantize takes an NTT-domain R poly
    r = np.concatenate([uniform_poly(w.prng, p.ringQ, p.max_level()), uniform_poly(w.prng, p.ringQMul, p.max_level())])
    prn = mkrlwe.Poly.from_numpy(w.ctx, r)
    pq2 = mkrlwe.Poly(w.ctx, p.nQ)
    w.dev.conv.Quantize(prn, pq2)
    assert_same(pq2.numpy(), w.oev.quantize(r), "Quantize")
    # DecomposeBFV on the lazy R poly
    h1, h2 = mkrlwe.SwitchingKey(w.ctx), mkrlwe.SwitchingKey(w.ctx)
    w.dev.ksw.DecomposeBFV(p.max_level(), pr, h1, h2)
    r1, r2 = w.oev.decompose_bfv(p.max_level(), refR)
    assert_same(h1.numpy(), r1, "DecomposeBFV ad1")
    assert_same(h2.numpy(), r2, "DecomposeBFV ad2")


def check_bfv_mul_relin(w: BFVWorld, ids0, ids1, same=False):
    o0, d0 = w.random_ct(ids0)
    if same:
        o1, d1 = o0, d0
    else:
        o1, d1 = w.random_ct(ids1)
    oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    dout = w.dev.MulRelinNew(d0, d1, w.d_rlk)
    dv = dout.numpy()
    assert set(dv) == set(oout.value)
    for k in oout.value:
        assert_same(dv[k], oout.value[k], f"BFV MulRelinNew(ids0={ids0}, ids1={ids1}, same={same})[{k}]")
    return dout


def check_bfv_linear_ops(w: BFVWorld):
    """mkbfv.Evaluator.{AddNew, SubNew, RotateNew, ConjugateNew} (mkbfv/evaluator.go:25-82,152-226): ring-Q element-wise ops and the
    mkrlwe key switches at the top level; the oracle side is the same arithmetic through its CKKS-agnostic evaluator"""
    p = w.op
    for rot in (1, 2):
        if rot not in p.CRS:
            p.add_crs(rot)
            w.dp.SetCRS(rot, p.CRS[rot])
    if -2 not in w.dp.CRS:
        w.dp.SetCRS(-2, p.CRS[-2])
    oev = O.CKKSEvaluator(p, 1.0)
    ork, drk = {}, mkrlwe.RotationKeySet()
    ock, dck = {}, mkrlwe.ConjugationKeySet()
    for i in w.ids:
        ork[i] = {r: uniform_swk(w.prng, p) for r in (1, 2)}
        for r, a in ork[i].items():
            drk.AddRotationKey(i, r, mkrlwe.SwitchingKey(w.ctx, a))
        ock[i] = uniform_swk(w.prng, p)
        dck.AddConjugationKey(i, mkrlwe.SwitchingKey(w.ctx, ock[i]))
    wrap = lambda c: O.Ciphertext(c.value, 1.0)

    def cmp(dct, oct_, what):
        dv = dct.numpy()
        assert set(dv) == set(oct_.value), what
        for k in oct_.value:
            assert_same(dv[k], oct_.value[k], f"{what}[{k}]")

    for i0, i1 in ((w.ids, w.ids), (w.ids[:1], w.ids[1:] or w.ids), (w.ids[:1], w.ids)):
        o0, d0 = w.random_ct(i0)
        o1, d1 = w.random_ct(i1)
        cmp(w.dev.AddNew(d0, d1), oev.add_new(wrap(o0), wrap(o1)), f"bfv AddNew {i0} {i1}")
        cmp(w.dev.SubNew(d0, d1), oev.sub_new(wrap(o0), wrap(o1)), f"bfv SubNew {i0} {i1}")
    o0, d0 = w.random_ct(w.ids)
    cmp(w.dev.RotateNew(d0, 2, drk), oev.rotate_new(wrap(o0), 2, ork), "bfv RotateNew(2)")
    cmp(w.dev.RotateNew(d0, 3, drk), oev.rotate_new(wrap(o0), 3, ork), "bfv RotateNew(3, chained)")
    cmp(w.dev.RotateNew(d0, 0, drk), wrap(o0), "bfv RotateNew(0)")
    cmp(w.dev.ConjugateNew(d0, dck), oev.conjugate_new(wrap(o0), ock), "bfv ConjugateNew")


def check_bfv_semantics(w: BFVWorld):
    """T3: exact plaintext equality after decrypting the DEVICE result (mkbfv_test.go:412)"""
    p = w.op
    N, T = p.N, p.T
    rng = np.random.default_rng(5)
    enc, dec = O.Encryptor(p), O.Decryptor(p)
    ms = [rng.integers(0, T, N) for _ in w.ids[:2]]
    cts = [enc.encrypt(O.bfv_encode(p, ms[j]), w.pks[i], i) for j, i in enumerate(w.ids[:2])]
    d0 = mkrlwe.Ciphertext.from_numpy(w.ctx, cts[0].value)
    d1 = mkrlwe.Ciphertext.from_numpy(w.ctx, cts[1].value)
    dres = w.dev.MulRelinNew(d0, d1, w.d_rlk)
    got = O.bfv_decode(p, dec.decrypt(O.Ciphertext(dres.numpy()), w.sks))
    full = np.convolve(np.array([int(x) for x in ms[0]], dtype=object), np.array([int(x) for x in ms[1]], dtype=object))
    ref = np.zeros(N, dtype=object)
    for i, v in enumerate(full):
        if i < N:
            ref[i] += v
        else:
            ref[i - N] -= v
    ref = np.array([int(v) % T for v in ref], dtype=np.int64)
    assert np.array_equal(got, ref), "BFV MulRelin: decrypted product differs from the plaintext product"


def run_ckks_suite(w: CKKSWorld, quick=False):
    if len(w.ids) >= 2 and w.op.max_level() >= 2:
        check_elementwise(w)
    L = w.op.max_level()
    check_ntt(w)
    check_decompose(w)
    check_decompose(w, level=max(L - 2, 1))
    check_external_product(w)
    check_external_product(w, level=1)
    check_moddown_ntt(w)
    check_moddown_ntt(w, level=1)
    check_rescale(w)
    ids = w.ids
    check_mul_relin_hoisted(w, ids, ids)
    check_mul_relin_new(w, ids, ids)
    check_mul_relin_new(w, ids, ids, same=True)
    check_rotate(w, ids, 2)
    if quick:
        return
    check_rescale(w, nb=2)
    check_rescale(w, level=1)
    check_mul_relin_hoisted(w, ids, ids, nil0=True)
    check_mul_relin_hoisted(w, ids, ids, nil1=True)
    check_mul_relin_hoisted(w, ids[:1], ids[1:])             # disjoint idsets
    check_mul_relin_hoisted(w, ids[:1], ids)                 # overlapping idsets
    check_mul_relin_hoisted(w, ids, ids, level=max(L - 2, 1))
    check_mul_relin_new(w, ids, ids[:1], level=1)
    check_rotate(w, ids, 1, level=max(L - 1, 1))
    check_rotate(w, ids, 3, hoisted=False)                   # rot 3: power-of-two chaining in RotateNew; hoisted panics
    check_rotate(w, ids, 2, zero_component=True)
    check_conjugate(w, ids)

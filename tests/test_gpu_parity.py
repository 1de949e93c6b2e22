"""-m gpu: the parity tests proper.  Every call goes through the C ABI of libmkhe_b200.so (via the reference-shaped
host mirror) and is compared bit for bit with the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import parity
from mkhe_kklss_b200 import params as PR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lit,logN,k", [
    (PR.CKKS_PN14QP439, 12, 2), (PR.CKKS_PN14QP439, 13, 3), (PR.CKKS_PN14QP439, 14, 2),
    (PR.CNN_PN14QP433, 14, 2), (PR.CKKS_PN15QP880, 12, 4), (PR.CKKS_PN15QP880, 15, 2),
], ids=lambda x: getattr(x, "name", str(x)))
def test_ckks_parity(lit, logN, k):
    w = parity.CKKSWorld(lit.at_logn(logN), k)
    parity.run_ckks_suite(w, quick=(logN >= 15))
    w.close()


def test_ckks_many_parties():
    """k = 8 parties (config 2) at a size the oracle finishes in seconds"""
    w = parity.CKKSWorld(PR.CKKS_PN15QP880.at_logn(12), 8)
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_mul_relin_hoisted(w, w.ids[:5], w.ids[3:])
    parity.check_rotate(w, w.ids, 1)
    w.close()


@pytest.mark.parametrize("logN,k,rounds", [(12, 2, 6), (14, 2, 3), (15, 4, 2)])
def test_lanes(logN, k, rounds):
    """two lanes of one context: concurrent independent ops and cross-lane read-after-write / write-after-read chains"""
    w = parity.CKKSWorld(PR.CKKS_PN15QP880.at_logn(logN), k, rots=(1,))
    parity.check_lanes(w, rounds=rounds, level=None if logN < 15 else 3)
    w.close()


def test_cnn_inference_flow():
    """BASELINE config 5: the whole encrypted-CNN op sequence of the reference (cnn/cnn.go + cnn/cnn_test.go:121-162) at its own
    parameter set PN14QP433 / logN = 14, two parties (model owner, data owner), 24 rotation keys per party: 15 MulRelin, 11
    hoisted + 19 plain rotations, 27 HoistedForm, 31 AddNew, 1 MulPtxtNew, levels 6 -> 0.  Every intermediate ciphertext of
    the device is the oracle's, bit for bit."""
    parity.check_cnn_flow(PR.CNN_PN14QP433)


def test_concurrent_lanes_stress():
    """full-size MulRelinNew and hoisted Rotate calls on two lanes with NO synchronisation in between (the bench's schedule), every
    result compared with the oracle.  Guards the hazard that only concurrency exposed: with another kernel's CTAs on the SM the
    shared-memory loads of k_mac_digits could still be in flight when the TMA refill of their stage landed (hazard 1)."""
    w = parity.CKKSWorld(PR.CKKS_PN15QP880, 2, rots=(2,))
    ids, level = w.ids, w.op.max_level()
    o0, d0 = w.random_ct(ids, level)
    o1, d1 = w.random_ct(ids, level)
    want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    want_rot = w.oev.rotate_hoisted_new(o0, 2, w.oev.hoisted_form(o0), w.o_rk)
    lanes = [w.ctx, w.ctx.fork()]
    dh = w.dev.HoistedForm(d0)
    B = 16
    from mkhe_kklss_b200 import mkckks
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    rots = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
    rk = [w.d_rk.GetRotationKey(i, 2).h for i in ids]
    nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
    bad = []
    for r in range(10):
        for i in range(B):
            a, b = lanes[i % 2], lanes[(i + 1) % 2]
            a.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))
            if r % 2:
                b.rotate_hoisted(level, 2, d0.handles(ids), [dh[t].h for t in ids], rk, w.dp.CRS[2].h, rots[i].handles(ids))
            else:
                b.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, rots[i].handles(ids))
        for ln in lanes:
            ln.sync()
        for i in range(B):
            for key in ["0"] + ids:
                if not np.array_equal(w.ctx.poly_download(outs[i].Value[key].h, level + 1 - nb), want.value[key]):
                    bad.append((r, i, "mul", key))
                ref = want_rot.value[key] if r % 2 else want.value[key]
                if not np.array_equal(w.ctx.poly_download(rots[i].Value[key].h, ref.shape[0]), ref):
                    bad.append((r, i, "other lane", key))
    assert not bad, bad[:8]
    w.close()


def test_ckks_semantics_on_device_outputs():
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2, real_keys=True)
    parity.check_ckks_semantics(w)
    w.close()


@pytest.mark.parametrize("logN,k", [(12, 2), (13, 3), (14, 2)])
def test_bfv_parity(logN, k):
    w = parity.BFVWorld(PR.BFV_PN14QP439.at_logn(logN), k)
    parity.check_bfv_conv(w)
    parity.check_bfv_mul_relin(w, w.ids, w.ids)
    parity.check_bfv_mul_relin(w, w.ids, w.ids, same=True)
    parity.check_bfv_mul_relin(w, w.ids[:1], w.ids[1:])
    parity.check_bfv_mul_relin(w, w.ids[:1], w.ids)
    parity.check_bfv_linear_ops(w)
    w.close()


def test_bfv_pn15_primes():
    w = parity.BFVWorld(PR.BFV_PN15QP880.at_logn(12), 2)
    parity.check_bfv_conv(w)
    parity.check_bfv_mul_relin(w, w.ids, w.ids)
    w.close()


def test_bfv_semantics_on_device_outputs():
    w = parity.BFVWorld(PR.BFV_PN14QP439.at_logn(12), 2, real_keys=True)
    parity.check_bfv_semantics(w)
    w.close()


def test_full_size_properties():
    """BASELINE config 2 size (PN15QP880, logN = 15, k = 4): size-independent properties instead of the oracle:
    NTT round trip, MulRelin(op0,op1) == MulRelin(op1,op0), hoisted == non-hoisted, Rotate(r) then Rotate(-r)
    structure via RotateNew == RotateHoistedNew."""
    from mkhe_kklss_b200 import mkrlwe
    w = parity.CKKSWorld(PR.CKKS_PN15QP880, 4)
    L = w.op.max_level()
    a = parity.uniform_poly(w.prng, w.op.ringQ, L)
    pin, pout = mkrlwe.Poly.from_numpy(w.ctx, a), mkrlwe.Poly(w.ctx, L + 1)
    w.ctx.ntt(L, pin.h, pout.h)
    w.ctx.intt(L, pout.h, pout.h)
    parity.assert_same(pout.numpy(), a, "INTT(NTT(a)) at logN=15")
    _, d0 = w.random_ct(w.ids, L)
    _, d1 = w.random_ct(w.ids, L)
    r01 = w.dev.MulRelinNew(d0, d1, w.d_rlk).numpy()
    h0, h1 = w.dev.HoistedForm(d0), w.dev.HoistedForm(d1)
    rh = w.dev.MulRelinHoistedNew(d0, d1, h0, h1, w.d_rlk).numpy()
    for k in r01:
        parity.assert_same(rh[k], r01[k], f"hoisted vs fused MulRelinNew [{k}]")
        assert r01[k].shape[0] == L
        for j, q in enumerate(w.op.Q[:L]):
            assert int(r01[k][j].max()) < q, "outputs must be canonical"
    rot_h = w.dev.RotateHoistedNew(d0, 2, h0, w.d_rk).numpy()
    rot = w.dev.RotateNew(d0, 2, w.d_rk).numpy()
    for k in rot:
        parity.assert_same(rot_h[k], rot[k], f"RotateHoistedNew vs RotateNew [{k}]")
    w.close()


@pytest.mark.parametrize("k", [4, 8])
def test_benchmarked_ckks_config_against_oracle(k):
    """BASELINE config 2 exactly as bench.py times it (mkckks_benchmark_test.go:55-84): PN15QP880, logN = 15, level 13, k = 4 and
    k = 8 parties, op0 != op1 with all k ids -- MulRelinNew, the square, HoistedForm + RotateHoistedNew (rot 2) and RotateNew,
    every limb of every component against the oracle (OpenMP over limbs: about a second per op)."""
    w = parity.CKKSWorld(PR.CKKS_PN15QP880, k, rots=(2,))
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_mul_relin_new(w, w.ids, w.ids, same=True)
    parity.check_rotate(w, w.ids, 2)
    w.close()


def test_benchmarked_bfv_config_against_oracle():
    """BASELINE config 3 (mkbfv_bench_test.go:39-64): mkbfv MulRelinNew at PN15QP880, logN = 15, k = 4, against the oracle;
    op0 != op1 and the square (ct, ct) the reference benchmark times"""
    w = parity.BFVWorld(PR.BFV_PN15QP880, 4)
    parity.check_bfv_mul_relin(w, w.ids, w.ids)
    parity.check_bfv_mul_relin(w, w.ids, w.ids, same=True)
    w.close()


@pytest.mark.parametrize("lit,logN,k,nranks", [(PR.CKKS_PN14QP439, 12, 2, 2), (PR.CKKS_PN15QP880, 12, 3, 4), (PR.CKKS_PN15QP880, 13, 4, 8),
                                               (PR.CNN_PN14QP433, 14, 2, 3), (PR.CKKS_PN15QP880, 15, 4, 2)],
                         ids=lambda x: getattr(x, "name", str(x)))
def test_limb_sharded_mul_relin(lit, logN, k, nranks):
    """MulRelinNew executed by a team of ranks, each computing the limb slots it owns (peer-store exchanges fused into
    k_moddown_P / k_moddown_Q / k_team_gather, in-kernel flag barriers).  The ranks of these cases are contexts on ONE device, so
    the driver's single-GPU test run covers the whole protocol; every rank's result against the oracle.  The last case is the
    benchmarked size."""
    parity.check_limb_sharded(lit.at_logn(logN), k, nranks, rounds=3 if logN < 15 else 2)


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_team_allgather(nranks):
    parity.check_team_allgather(PR.CKKS_PN15QP880.at_logn(12), nranks)


def test_limb_sharded_id_sets_and_levels():
    lit = PR.CKKS_PN15QP880.at_logn(12)
    parity.check_limb_sharded(lit, 3, 4, level=5, ids0=[0, 1], ids1=[1, 2])      # overlapping id sets below the top level
    parity.check_limb_sharded(lit, 3, 8, level=2, ids0=[0], ids1=[1, 2])         # more ranks than limbs: some ranks own nothing
    parity.check_limb_sharded(lit, 2, 2, level=1)                                # the lowest level a product can be rescaled from


def test_full_size_one_party_against_oracle():
    """one full-size (logN = 15, 14 limbs) MulRelinNew with k = 1 against the oracle (a few seconds of CPU)"""
    w = parity.CKKSWorld(PR.CKKS_PN15QP880, 1, rots=(1,))
    parity.check_mul_relin_new(w, [0], [0])
    w.close()


def test_device_matches_committed_golden_digests():
    """the committed fixtures (tests/golden/oracle_digests.json, made by tools/gen_golden.py) reproduced on the GPU:
    same seeds, device results hashed, no oracle call at run time"""
    import hashlib
    import json
    import os
    from oracle import oracle as O   # PRNG only: regenerates the seeded inputs
    from mkhe_kklss_b200 import mkckks, mkrlwe
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")))
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()
    for name, lit, k, same in (("ckks_PN14QP439_logN12_k2", PR.CKKS_PN14QP439.at_logn(12), 2, False),
                               ("ckks_PN14QP439_logN12_k2_square", PR.CKKS_PN14QP439.at_logn(12), 2, True),
                               ("ckks_PN15QP880_logN12_k3", PR.CKKS_PN15QP880.at_logn(12), 3, False),
                               ("cnn_PN14QP433_logN12_k2", PR.CNN_PN14QP433.at_logn(12), 2, False)):
        p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000001, crs_rots=[1, 2])
        prng = O.PRNG(0xB2000001 ^ 0x5EED)
        dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale)
        for idx, arr in p.CRS.items():
            dp.SetCRS(idx, arr)
        rl, rk = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet()
        sw = lambda: parity.uniform_swk(prng, p)
        for i in range(k):
            rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, sw(), sw(), sw()))
        for i in range(k):
            rk.AddRotationKey(i, 2, mkrlwe.SwitchingKey(dp.ctx, sw()))
        ev = mkckks.Evaluator(dp)
        L = p.max_level()
        mk = lambda: {**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in range(k)}}
        c0 = mkckks.Ciphertext.from_numpy(dp.ctx, mk(), lit.scale)
        c1 = c0 if same else mkckks.Ciphertext.from_numpy(dp.ctx, mk(), lit.scale)
        out = ev.MulRelinNew(c0, c1, rl).numpy()
        rot = ev.RotateHoistedNew(c0, 2, ev.HoistedForm(c0), rk).numpy()
        got = {f"mul[{kk}]": dig(v) for kk, v in out.items()}
        got.update({f"rot[{kk}]": dig(v) for kk, v in rot.items()})
        assert got == golden[name], name
        dp.ctx.close()
    # the reference's CNN op sequence at logN = 12: device result and intermediates hashed against the committed digests
    import cnn_flow as F
    lit = PR.CNN_PN14QP433.at_logn(12)
    out, mid, _, ctx = F.run_device(lit, F.make_inputs(lit))
    got = {f"fc2Out[{kk}]": dig(v) for kk, v in out.numpy().items()}
    for nm, ct in mid.items():
        got.update({f"{nm}[{kk}]": dig(v) for kk, v in ct.numpy().items()})
    assert got == golden["cnn_flow_PN14QP433_logN12"]
    ctx.close()


def test_full_size_back_to_back():
    """logN = 15 (BASELINE size): MulRelinNew calls enqueued back to back without synchronisation (as bench.py does), every
    output against ONE oracle result, then the same for hoisted Rotate.  A staging / ordering error between or inside
    kernels shows up as an intermittent mismatch: this is the test that caught a TMA refill overtaking still-queued
    shared-memory loads in pass 2 (rate ~1 %, tools/stress_b2b.py)."""
    from mkhe_kklss_b200 import mkckks
    w = parity.CKKSWorld(PR.CKKS_PN15QP880, 2, rots=(2,))
    ids, level = w.ids, w.op.max_level()
    o0, d0 = w.random_ct(ids, level)
    o1, d1 = w.random_ct(ids, level)
    want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(16)]
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
    nb, new_scale = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
    for rnd in range(12):
        for out in outs:
            w.ctx.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids,
                                 out.handles(ids))
        w.ctx.sync()
        for i, out in enumerate(outs):
            out.Scale = new_scale
            w.compare_ct(out, want, f"MulRelinNew round {rnd} op {i}")
    hd = w.dev.HoistedForm(d0)
    want_r = w.oev.rotate_hoisted_new(o0, 2, w.oev.hoisted_form(o0), w.o_rk)
    for rep in range(25):
        w.compare_ct(w.dev.RotateHoistedNew(d0, 2, hd, w.d_rk), want_r, f"RotateHoistedNew repetition {rep}")
    w.close()


def test_async_transfers_pipeline():
    """mkhe_poly_upload_async / _download_async on library-owned pinned memory: a three-deep pipeline of MulRelinNew calls
    whose operands arrive and whose results leave asynchronously gives the oracle's bits for every op"""
    w = parity.CKKSWorld(PR.CKKS_PN14QP439, 2)
    ctx, level = w.ctx, w.op.max_level()
    ops = []
    for _ in range(3):
        o0, _d0 = w.random_ct(w.ids, level)
        o1, _d1 = w.random_ct(w.ids, level)
        ops.append((o0, o1, w.oev.mul_relin_new(o0, o1, w.o_rlk)))
    from mkhe_kklss_b200 import mkckks
    keys = ["0"] + w.ids
    dev = [(mkckks.Ciphertext.new(w.dp, w.ids, level, w.lit.scale), mkckks.Ciphertext.new(w.dp, w.ids, level, w.lit.scale)) for _ in range(2)]
    outs = [mkckks.Ciphertext.new(w.dp, w.ids, level, w.lit.scale) for _ in range(2)]
    pin_in = [({k: ctx.host_alloc(o0.value[k].shape) for k in keys}, {k: ctx.host_alloc(o1.value[k].shape) for k in keys}) for o0, o1, _ in ops]
    for (pa, pb), (o0, o1, _) in zip(pin_in, ops):
        for k in keys:
            pa[k][...] = o0.value[k]
            pb[k][...] = o1.value[k]
    nl = ops[0][2].value["0"].shape[0]
    pin_out = [{k: ctx.host_alloc((nl, w.lit.N)) for k in keys} for _ in range(6)]
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in w.ids] for j in range(3))
    nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)

    def upload(n):
        a, b = dev[n % 2]
        pa, pb = pin_in[n % 3]
        for k in keys:
            ctx.poly_upload_async(a.Value[k].h, pa[k])
            ctx.poly_upload_async(b.Value[k].h, pb[k])

    upload(0)
    for n in range(6):
        if n + 1 < 6:
            upload(n + 1)
        a, b = dev[n % 2]
        out = outs[n % 2]
        ctx.ckks_mul_relin(level, nb, False, w.ids, a.handles(w.ids), w.ids, b.handles(w.ids), kb, kd, kv,
                           w.dp.CRS[-1].h, w.ids, out.handles(w.ids))
        for k in keys:
            ctx.poly_download_async(out.Value[k].h, pin_out[n][k])
    ctx.sync()
    for n in range(6):
        for k in keys:
            parity.assert_same(pin_out[n][k], ops[n % 3][2].value[k], f"pipelined MulRelinNew {n}[{k}]")
    w.close()


@pytest.mark.parametrize("lit,logN", [(PR.PN16QP1761_Q7, 13), (PR.PN16QP1761_Q7_ALPHA4, 12)], ids=lambda x: getattr(x, "name", str(x)))
def test_wide_digits(lit, logN):
    """alpha = #P/gamma > 1: DecomposeAndSplit's general branch (mkrlwe/basis_extension.go:454-534), alpha-limb gadget keys"""
    w = parity.CKKSWorld(lit.at_logn(logN), 2)
    parity.run_ckks_suite(w, quick=True)
    parity.check_decompose(w, level=5)
    parity.check_decompose(w, level=4)
    parity.check_mul_relin_new(w, w.ids, w.ids, level=5)
    parity.check_rotate(w, w.ids, 1, level=4)
    w.close()


def test_wide_digits_mixed_levels():
    w = parity.CKKSWorld(PR.PN16QP1761_Q7.at_logn(12), 2)
    parity.check_mixed_levels(w)
    w.close()


def test_abi_rejects_out_of_range_key_digits():
    w = parity.CKKSWorld(PR.PN16QP1761_Q7.at_logn(12), 1, rots=())
    parity.check_abi_negative(w)
    w.close()


def test_fork_after_queued_work():
    w = parity.CKKSWorld(PR.CKKS_PN15QP880.at_logn(14), 2, rots=(1,))
    parity.check_fork_after_queued_work(w)
    w.close()


def test_pn16qp1761_full_limb_count():
    """the reference's commented-out PN16QP1761 (34 + 4 limbs, alpha = 2, 17 digits) at a size the oracle finishes quickly"""
    w = parity.CKKSWorld(PR.PN16QP1761.at_logn(12), 2, rots=(1,))
    parity.check_decompose(w)
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_rotate(w, w.ids, 1)
    w.close()


@pytest.mark.gpu
def test_keygen_and_encrypt_on_device():
    """every key of mkrlwe.KeyGenerator / mkbfv.KeyGenerator and Encryptor.Encrypt made on the device,
    bit for bit against the oracle's generators on the same counter-based streams; then encrypt -> MulRelin -> Rotate ->
    Conjugate -> Decrypt on device-made material only, with the reference's precision thresholds."""
    parity.check_keygen(PR.CKKS_PN14QP439.at_logn(13))
    parity.check_keygen(PR.PN16QP1761_Q7.at_logn(12), rots=(1,), semantics=False)
    parity.check_bfv_keygen(PR.BFV_PN14QP439.at_logn(12))


@pytest.mark.gpu
def test_keygen_full_size():
    """PN15QP880 at logN = 15: a party's whole key set made on the device (no 168 MiB upload per party), against the oracle"""
    parity.check_keygen(PR.CKKS_PN15QP880, nparties=1, rots=(1,), semantics=False)


def test_cuda_graph_replay_matches_the_oracle():
    """the opt-in graph cache (mkhe_ctx_set_graphs): ops repeated on the same operands replay a captured launch sequence; every
    replayed result is the oracle's, and a different operand (another signature) is not served from a stale graph"""
    from mkhe_kklss_b200 import mkckks
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(13), 2, rots=(2,))
    w.ctx.set_graphs(True)
    ids, level = w.ids, w.op.max_level()
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
    nb, new_scale = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
    cts = [w.random_ct(ids, level) for _ in range(3)]
    wants = [w.oev.mul_relin_new(cts[i][0], cts[(i + 1) % 3][0], w.o_rlk) for i in range(3)]
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(3)]
    for rep in range(6):
        for i in range(3):
            for p_ in outs[i].Value.values():
                p_.set_nlimbs(level + 1)
            w.ctx.ckks_mul_relin(level, nb, False, ids, cts[i][1].handles(ids), ids, cts[(i + 1) % 3][1].handles(ids), kb, kd, kv,
                                 w.dp.CRS[-1].h, ids, outs[i].handles(ids))
        for i in range(3):
            outs[i].Scale = new_scale
            w.compare_ct(outs[i], wants[i], f"graph replay {rep}, signature {i}")
    hd = w.dev.HoistedForm(cts[0][1])
    want_r = w.oev.rotate_hoisted_new(cts[0][0], 2, w.oev.hoisted_form(cts[0][0]), w.o_rk)
    out_r = mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale)
    for rep in range(5):
        w.ctx.rotate_hoisted(level, 2, cts[0][1].handles(ids), [hd[t].h for t in ids], [w.d_rk.GetRotationKey(t, 2).h for t in ids],
                             w.dp.CRS[2].h, out_r.handles(ids))
        w.compare_ct(out_r, want_r, f"graph replay of RotateHoisted {rep}")
    captures, replays = w.ctx.graph_stats()
    assert captures >= 4 and replays >= 9, (captures, replays)
    w.close()


def test_cnn_inference_semantics_on_device_made_material():
    """BASELINE config 5 end to end on the device: CRS, both parties' key sets (relinearisation + 24 rotation keys each) and every
    ciphertext are made on the GPU, the reference's CNN op sequence runs there, the result is decrypted there -- and the ten
    logits equal the plain forward pass of the same network on the same image within CKKS precision (tests/cnn_semantic.py)."""
    import cnn_semantic as S
    got, want = S.run(PR.CNN_PN14QP433)
    assert np.abs(got - want).max() < 1e-4, (got, want)
    assert got.argmax() == want.argmax()

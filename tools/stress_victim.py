#!/usr/bin/env python
"""GPU race hunt: lane 0 runs ONE small op type over and over (the victim, every result compared with the oracle), lane 1 runs
full MulRelinNew calls as noise.  usage: tools/stress_victim.py <decompose|extprod|ntt|intt|rescale> [rounds] [noise 0/1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from mkhe_kklss_b200 import params as PR, mkckks, mkrlwe

if os.environ.get("MKHE_LIB"):
    from mkhe_kklss_b200 import _lib
    _lib._default = _lib.Library(os.path.abspath(os.environ["MKHE_LIB"]))
victim = sys.argv[1]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
noise = sys.argv[3] if len(sys.argv) > 3 else "mul"          # 0 / mul / decompose / extprod / ntt / copy
B = 16
w = parity.CKKSWorld(PR.CKKS_PN15QP880, 2, rots=(2,))
ids, level = w.ids, w.op.max_level()
o0, d0 = w.random_ct(ids, level)
o1, d1 = w.random_ct(ids, level)
lanes = [w.ctx, w.ctx.fork()]
if os.environ.get("VICTIM_ON_FORK", "1") == "1":
    lanes = lanes[::-1]            # lanes[0] = victim lane (the fork), lanes[1] = noise lane (the root)
g = w.d_rlk.GetRelinearizationKey
kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
nout = mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale)
nswk, nswk2, npoly = mkrlwe.SwitchingKey(w.ctx), mkrlwe.SwitchingKey(w.ctx), mkrlwe.Poly(w.ctx, level + 1)
w.ctx.decompose(level, d1.Value["0"].h, nswk2.h)
a = o0.value["0"]
pa = d0.Value["0"]
ks = w.oev.ksw
if victim == "decompose":
    want = ks.decompose(level, a)
    outs = [mkrlwe.SwitchingKey(w.ctx) for _ in range(B)]
    run = lambda i: lanes[0].decompose(level, pa.h, outs[i].h)
    get = lambda i: outs[i].numpy()
elif victim == "extprod":
    oh = ks.decompose(level, a)
    dh = mkrlwe.SwitchingKey(w.ctx, oh)
    key = w.o_rlk[0].b
    dkey = g(0).Value[0]
    want = ks.external_product_hoisted(level, oh, key)
    outs = [mkrlwe.Poly(w.ctx, level + 1) for _ in range(B)]
    run = lambda i: lanes[0].external_product_hoisted(level, dh.h, dkey.h, outs[i].h)
    get = lambda i: outs[i].numpy()
elif victim in ("ntt", "intt"):
    fn = w.op.ringQ.ntt if victim == "ntt" else w.op.ringQ.intt
    want = fn(a, level)
    outs = [mkrlwe.Poly(w.ctx, level + 1) for _ in range(B)]
    run = lambda i: (lanes[0].ntt if victim == "ntt" else lanes[0].intt)(level, pa.h, outs[i].h)
    get = lambda i: outs[i].numpy()
elif victim == "rotate":
    ohs = w.oev.hoisted_form(o0)
    dhs = w.dev.HoistedForm(d0)
    want_ct = w.oev.rotate_hoisted_new(o0, 2, ohs, w.o_rk)
    want = np.stack([want_ct.value[k] for k in ["0"] + ids])
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    rk = [w.d_rk.GetRotationKey(i, 2).h for i in ids]
    run = lambda i: lanes[0].rotate_hoisted(level, 2, d0.handles(ids), [dhs[t].h for t in ids], rk, w.dp.CRS[2].h, outs[i].handles(ids))
    get = lambda i: np.stack([outs[i].Value[k].numpy() for k in ["0"] + ids])
elif victim == "mulhoisted":
    ohs0, ohs1 = w.oev.hoisted_form(o0), w.oev.hoisted_form(o1)
    dhs0, dhs1 = w.dev.HoistedForm(d0), w.dev.HoistedForm(d1)
    wct = w.oev.new_ciphertext(ids, level, 1.0)
    w.oev.ksw.mul_and_relin_hoisted(o0, o1, ohs0, ohs1, w.o_rlk, wct)
    want = np.stack([wct.value[k] for k in ["0"] + ids])
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    run = lambda i: lanes[0].mul_relin_hoisted(level, ids, d0.handles(ids), [dhs0[t].h for t in ids], ids, d1.handles(ids), [dhs1[t].h for t in ids],
                                               kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))
    get = lambda i: np.stack([outs[i].Value[k].numpy() for k in ["0"] + ids])
elif victim == "mul":
    wct = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    want = np.stack([wct.value[k] for k in ["0"] + ids])
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]

    def run(i):
        for p in outs[i].Value.values():
            lanes[0].poly_set_nlimbs(p.h, level + 1)
        lanes[0].ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))
    get = lambda i: np.stack([outs[i].Value[k].numpy() for k in ["0"] + ids])
else:
    raise SystemExit("unknown victim")
bad = 0
for r in range(rounds):
    for i in range(B):
        if noise in ("1", "mul"):
            for p in nout.Value.values():
                p.set_nlimbs(level + 1)
            lanes[1].ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, nout.handles(ids))
        elif noise == "decompose":
            for _ in range(6):
                lanes[1].decompose(level, d1.Value["0"].h, nswk.h)
        elif noise == "extprod":
            for _ in range(12):
                lanes[1].external_product_hoisted(level, nswk2.h, g(1).Value[1].h, npoly.h)
        elif noise == "ntt":
            for _ in range(40):
                lanes[1].ntt(level, d1.Value["0"].h, npoly.h)
        elif noise == "copy":
            for _ in range(40):
                lanes[1].poly_copy(npoly.h, d1.Value["0"].h)
        run(i)
        run(i)          # twice into the same output: more victim work per noise op
    for ln in lanes:
        ln.sync()
    for i in range(B):
        got = get(i)
        if not np.array_equal(got, want):
            bad += 1
            d = np.argwhere(got != want)
            print(f"round {r} op {i}: MISMATCH n={len(d)} idx0 {np.unique(d[:, 0]).tolist()[:16]} idx1 {np.unique(d[:, 1]).tolist()[:16]} lastdim {d[:, -1].min()}..{d[:, -1].max()}", flush=True)
print(f"{victim}: {bad} of {rounds * B} victim results differ (noise={noise})")

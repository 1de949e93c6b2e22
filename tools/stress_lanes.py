#!/usr/bin/env python
"""GPU stress for the lanes: B MulRelinNew + RotateHoisted calls spread over two lanes, enqueued back to back (no
synchronisation in between), every output compared with the oracle afterwards.  usage: tools/stress_lanes.py [rounds] [B] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if os.environ.get("MKHE_LIB"):
    from mkhe_kklss_b200 import _lib
    _lib._default = _lib.Library(os.path.abspath(os.environ["MKHE_LIB"]))
import parity
from mkhe_kklss_b200 import params as PR, mkckks

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = sys.argv[4] if len(sys.argv) > 4 else "alt"      # alt: roles alternate; mulmul: both lanes multiply; fixed0 / fixed1: lane 0 / 1 multiplies, the other rotates
w = parity.CKKSWorld(PR.CKKS_PN15QP880, k, rots=(2,))
ids, level = w.ids, w.op.max_level()
o0, d0 = w.random_ct(ids, level)
o1, d1 = w.random_ct(ids, level)
want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
oh = w.oev.hoisted_form(o0)
want_rot = w.oev.rotate_hoisted_new(o0, 2, oh, w.o_rk)
lanes = [w.ctx, w.ctx.fork()]
dh = w.dev.HoistedForm(d0)
outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
rots = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
g = w.d_rlk.GetRelinearizationKey
kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
rk = [w.d_rk.GetRotationKey(i, 2).h for i in ids]
nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
bad = 0
for r in range(rounds):
    for i in range(B):
        li = {"alt": i % 2, "mulmul": i % 2, "fixed0": 0, "fixed1": 1, "forkonly": 1, "serial": i % 2}[mode]
        ln = lanes[li]
        if not os.environ.get("STRESS_NOSET"):
            for p in outs[i].Value.values():
                p.set_nlimbs(level + 1)
        ln.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))
        if mode == "serial":
            ln.sync()
        if mode in ("alt", "fixed0", "fixed1"):
            lanes[1 - li].rotate_hoisted(level, 2, d0.handles(ids), [dh[t].h for t in ids], rk, w.dp.CRS[2].h, rots[i].handles(ids))

    for ln in lanes:
        ln.sync()
    for i in range(B):
        msgs = []
        for key in ["0"] + ids:
            a = w.ctx.poly_download(outs[i].Value[key].h, level + 1 - nb)
            if not np.array_equal(a, want.value[key]):
                d = np.argwhere(a != want.value[key])
                pos = d[:, 1]
                msgs.append(f"mul {key!r}: limbs {np.unique(d[:, 0]).tolist()[:20]} n={len(d)} pos {pos.min()}..{pos.max()} "
                            f"tiles {np.unique(pos // 2048).tolist()[:8]} cols%128 {np.unique(pos % 128).size} cols%2048 {np.unique(pos % 2048).size}")
            if mode in ("alt", "fixed0", "fixed1"):
                a = w.ctx.poly_download(rots[i].Value[key].h, level + 1)
                if not np.array_equal(a, want_rot.value[key]):
                    msgs.append(f"rot {key!r}")
        if msgs:
            bad += 1
            print(f"round {r} op {i}: MISMATCH " + "; ".join(m[:60] for m in msgs), flush=True)
print(f"{mode}: {bad} of {rounds * B} two-lane op pairs differ")

#!/usr/bin/env python
"""GPU stress: B MulRelinNew calls enqueued back to back (no synchronisation in between, like bench.py), every output
compared with the oracle afterwards.  usage: tools/stress_b2b.py [rounds] [B] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from mkhe_kklss_b200 import params as PR, mkckks, _lib
if os.environ.get("MKHE_LIB"):                 # A/B runs: another build of the library
    _lib._default = _lib.Library(os.path.abspath(os.environ["MKHE_LIB"]))

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w = parity.CKKSWorld(PR.CKKS_PN15QP880, k, rots=(2,))
ids, level = w.ids, w.op.max_level()
o0, d0 = w.random_ct(ids, level)
o1, d1 = w.random_ct(ids, level)
want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
g = w.d_rlk.GetRelinearizationKey
kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
bad = 0
if os.environ.get("MKHE_STRESS_MODE") == "rot":        # hoisted rotations only: k_mac_intt + k_moddown_P / Q, no forward transforms
    hd = w.dev.HoistedForm(d0)
    want_r = w.oev.rotate_hoisted_new(o0, 2, w.oev.hoisted_form(o0), w.o_rk)
    hh, rk = [hd[t].h for t in ids], [w.d_rk.GetRotationKey(t, 2).h for t in ids]
    for out in outs:
        for p_ in out.Value.values():
            p_.set_nlimbs(level + 1)
    for r in range(rounds):
        for out in outs:
            w.ctx.rotate_hoisted(level, 2, d0.handles(ids), hh, rk, w.dp.CRS[2].h, out.handles(ids))
        w.ctx.sync()
        for i, out in enumerate(outs):
            msgs = []
            for key in ["0"] + ids:
                a = w.ctx.poly_download(out.Value[key].h, level + 1)
                if not np.array_equal(a, want_r.value[key]):
                    msgs.append(f"{key!r}: limbs {np.unique(np.argwhere(a != want_r.value[key])[:, 0]).tolist()}")
            if msgs:
                bad += 1
                print(f"round {r} op {i}: ROTATE MISMATCH " + "; ".join(msgs), flush=True)
    print(f"{bad} of {rounds * B} back-to-back hoisted rotations differ")
    sys.exit(0)
for r in range(rounds):
    for out in outs:
        w.ctx.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, out.handles(ids))
    w.ctx.sync()
    for i, out in enumerate(outs):
        msgs = []
        for key in ["0"] + ids:
            a = w.ctx.poly_download(out.Value[key].h, level + 1 - nb)
            b = want.value[key]
            if not np.array_equal(a, b):
                d = np.argwhere(a != b)
                msgs.append(f"{key!r}: limbs {np.unique(d[:, 0]).tolist()}")
        if msgs:
            bad += 1
            print(f"round {r} op {i}: MISMATCH " + "; ".join(msgs), flush=True)
print(f"{bad} of {rounds * B} back-to-back ops differ")

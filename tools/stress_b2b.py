#!/usr/bin/env python
"""GPU stress: B MulRelinNew calls enqueued back to back (no synchronisation in between, like bench.py), every output
compared with the oracle afterwards.  usage: tools/stress_b2b.py [rounds] [B] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from mkhe_kklss_b200 import params as PR, mkckks

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

#!/usr/bin/env python
"""GPU stress: the same MulRelinNew / hoisted Rotate at logN=15 repeated many times against one oracle result,
to expose intermittent (ordering) errors.  usage: tools/stress_race.py [reps] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from mkhe_kklss_b200 import params as PR

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = parity.CKKSWorld(PR.CKKS_PN15QP880, k)
ids = list(range(k))
level = w.op.max_level()
o0, d0 = w.random_ct(ids, level)
o1, d1 = w.random_ct(ids, level)
oout = w.oev.mul_relin_new(o0, o1, w.o_rlk)
bad = 0
for r in range(reps):
    dout = w.dev.MulRelinNew(d0, d1, w.d_rlk)
    msgs = []
    dv = dout.numpy()
    for key, a in dv.items():
        b = np.asarray(oout.value[key])
        if not np.array_equal(a, b):
            diff = np.argwhere(a != b)
            limbs = np.unique(diff[:, 0], return_counts=True)
            pos = diff[:, 1]
            msgs.append(f"comp {key}: {len(diff)} words, limbs {dict(zip(limbs[0].tolist(), limbs[1].tolist()))}, pos range {pos.min()}..{pos.max()}")
    if msgs:
        bad += 1
        print(f"rep {r}: MISMATCH", "; ".join(msgs), flush=True)
print(f"{bad} of {reps} repetitions differ (overlap env: {os.environ.get('MKHE_NO_OVERLAP')})")

#!/usr/bin/env python
"""GPU race hunt without lanes: TWO INDEPENDENT contexts (own keys, own ciphertexts, nothing shared) on one device, MulRelinNew
calls enqueued on both without synchronisation, every result compared with the oracle."""
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
B = 8
ws = [parity.CKKSWorld(PR.CKKS_PN15QP880, 2, rots=(2,)) for _ in range(2)]
st = []
for w in ws:
    ids, level = w.ids, w.op.max_level()
    o0, d0 = w.random_ct(ids, level)
    o1, d1 = w.random_ct(ids, level)
    want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
    nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
    st.append((w, ids, level, d0, d1, want, outs, kb, kd, kv, nb))
bad = [0, 0]
for r in range(rounds):
    for i in range(B):
        for (w, ids, level, d0, d1, want, outs, kb, kd, kv, nb) in st:
            w.ctx.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))
    for s in st:
        s[0].ctx.sync()
    for c, (w, ids, level, d0, d1, want, outs, kb, kd, kv, nb) in enumerate(st):
        for i in range(B):
            for key in ["0"] + ids:
                a = w.ctx.poly_download(outs[i].Value[key].h, level + 1 - nb)
                if not np.array_equal(a, want.value[key]):
                    d = np.argwhere(a != want.value[key])
                    bad[c] += 1
                    print(f"round {r} ctx {c} op {i} comp {key!r}: limbs {np.unique(d[:, 0]).tolist()} n={len(d)}", flush=True)
print(f"two independent contexts: mismatching components {bad} of {rounds * B * 3} each")

#!/usr/bin/env python
"""GPU stress: single ops of the C ABI repeated many times against their own first result (which is checked against the
oracle elsewhere); finds WHICH op is intermittently wrong.  usage: tools/stress_ops.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mkhe_kklss_b200 import params as PR
from mkhe_kklss_b200._lib import Context

lit = PR.CKKS_PN15QP880
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
A = Context(lit.logN, lit.Q, lit.P)
rng = np.random.default_rng(1)
mods = list(lit.Q) + list(lit.P)
level = len(lit.Q) - 1
rnd_poly = lambda: np.stack([rng.integers(0, q, size=lit.N, dtype=np.uint64) for q in lit.Q])
rnd_swk = lambda: np.stack([np.stack([rng.integers(0, q, size=lit.N, dtype=np.uint64) for q in mods]) for _ in lit.Q])

def where(a, b):
    d = np.argwhere(a != b)
    lead = [np.unique(d[:, i]).tolist()[:8] for i in range(d.shape[1] - 1)]
    return f"{len(d)} words; leading indices {lead}; pos {d[:, -1].min()}..{d[:, -1].max()}"

def stress(name, run, fetch):
    run(); A.sync(); ref = fetch()
    bad = 0
    t = time.time()
    for r in range(reps):
        run(); A.sync()
        got = fetch()
        if not np.array_equal(got, ref):
            bad += 1
            if bad <= 3: print(f"  {name} rep {r}: {where(got, ref)}", flush=True)
    print(f"{name}: {bad} of {reps} differ ({time.time() - t:.1f} s)", flush=True)

p_in = A.poly_alloc(level + 1); A.poly_upload(p_in, rnd_poly())
h = A.swk_alloc(); k1 = A.swk_alloc(); k2 = A.swk_alloc()
A.swk_upload(k1, rnd_swk()); A.swk_upload(k2, rnd_swk())
p_out = A.poly_alloc(level + 1); p_ntt = A.poly_alloc(level + 1)
stress("ntt (pass1 + pass2)", lambda: A.ntt(level, p_in, p_ntt), lambda: A.poly_download(p_ntt))
stress("decompose (bcast + pass2)", lambda: A.decompose(level, p_in, h), lambda: A.swk_download(h))
stress("external_product_hoisted (TMA MAC + inverse + ModDown)", lambda: A.external_product_hoisted(level, h, k1, p_out), lambda: A.poly_download(p_out))
stress("external_product (decompose + ...)", lambda: A.external_product(level, p_in, k2, p_out), lambda: A.poly_download(p_out))

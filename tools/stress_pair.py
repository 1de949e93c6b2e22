#!/usr/bin/env python
"""GPU stress: two contexts (two streams) on one device run Decompose (bcast + pass 2) and ExternalProductHoisted
(digit MAC + inverse NTT + ModDown) concurrently; every output is compared with the same op run alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mkhe_kklss_b200 import params as PR
from mkhe_kklss_b200._lib import Context

lit = PR.CKKS_PN15QP880
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
A = Context(lit.logN, lit.Q, lit.P)
B = Context(lit.logN, lit.Q, lit.P)
rng = np.random.default_rng(1)
mods = list(lit.Q) + list(lit.P)
level = len(lit.Q) - 1
def rnd_poly():
    return np.stack([rng.integers(0, q, size=lit.N, dtype=np.uint64) for q in lit.Q])
def rnd_swk():
    return np.stack([np.stack([rng.integers(0, q, size=lit.N, dtype=np.uint64) for q in mods]) for _ in lit.Q])
# context A: decompose polys -> swks
pa = [A.poly_alloc(level + 1) for _ in range(NB)]
for h in pa: A.poly_upload(h, rnd_poly())
ha = [A.swk_alloc() for _ in range(NB)]
# context B: external products
hb, kb = B.swk_alloc(), B.swk_alloc()
B.swk_upload(hb, rnd_swk()); B.swk_upload(kb, rnd_swk())
pb = [B.poly_alloc(level + 1) for _ in range(NB)]
# solo references
for i in range(NB): A.decompose(level, pa[i], ha[i])
A.sync()
refA = [A.swk_download(h) for h in ha]
for i in range(NB): B.external_product_hoisted(level, hb, kb, pb[i])
B.sync()
refB = [B.poly_download(h) for h in pb]
badA = badB = 0
for r in range(reps):
    for i in range(NB):
        A.decompose(level, pa[i], ha[i])
        B.external_product_hoisted(level, hb, kb, pb[i])
    A.sync(); B.sync()
    for i in range(NB):
        if not np.array_equal(A.swk_download(ha[i]), refA[i]): badA += 1
        if not np.array_equal(B.poly_download(pb[i]), refB[i]): badB += 1
print(f"concurrent decompose x external product: {badA} bad decompositions, {badB} bad products of {NB * reps} each")

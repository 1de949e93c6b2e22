"""development: host-side timing of the limb-sharded op with ranks in one process (where does a call block?)"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from oracle import oracle as O
from mkhe_kklss_b200 import params as PR, mkckks, mkrlwe
lit = PR.CKKS_PN14QP439.at_logn(12); k = 2; R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, seed=5, crs_rots=[])
prng = O.PRNG(77)
ids = list(range(k))
o_rlk = {i: O.RelinKey(i, parity.uniform_swk(prng, op), parity.uniform_swk(prng, op), parity.uniform_swk(prng, op)) for i in ids}
level = op.max_level()
ranks = []
for r in range(R):
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale)
    dp.SetCRS(-1, op.CRS[-1])
    rl = mkrlwe.RelinearizationKeySet()
    for i in ids:
        rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, o_rlk[i].b, o_rlk[i].d, o_rlk[i].v))
    ranks.append((dp, rl))
ctxs = [dp.ctx for dp, _ in ranks]
for r, c in enumerate(ctxs):
    c.team_join_local(k, r, ctxs)
val0 = {"0": parity.uniform_poly(prng, op.ringQ, level), **{i: parity.uniform_poly(prng, op.ringQ, level) for i in ids}}
val1 = {"0": parity.uniform_poly(prng, op.ringQ, level), **{i: parity.uniform_poly(prng, op.ringQ, level) for i in ids}}
evs = [mkckks.Evaluator(dp) for dp, _ in ranks]
ops = []
for (dp, rl), ev in zip(ranks, evs):
    d0 = mkckks.Ciphertext.from_numpy(dp.ctx, val0, lit.scale); d1 = mkckks.Ciphertext.from_numpy(dp.ctx, val1, lit.scale)
    ops.append((d0, d1, ev.newCiphertextBinary(d0, d1)))
for (dp, rl), ev, (d0, d1, _) in zip(ranks, evs, ops):
    ev.MulRelinNew(d0, d1, rl).free(); dp.ctx.sync()
for n in range(2):
    for r, ((dp, rl), ev, (d0, d1, dout)) in enumerate(zip(ranks, evs, ops)):
        l0 = dp.ctx.launch_count(); t0 = time.perf_counter()
        ev.MulRelinLimbSharded(d0, d1, rl, dout)
        print(f"op {n} rank {r}: host call {1e3 * (time.perf_counter() - t0):9.2f} ms, {dp.ctx.launch_count() - l0} launches", flush=True)
t0 = time.perf_counter()
for c in ctxs: c.sync()
print(f"sync {1e3 * (time.perf_counter() - t0):.2f} ms; timed out: {[c.team_timed_out() for c in ctxs]}; flags {[c.team_flags() for c in ctxs]}")

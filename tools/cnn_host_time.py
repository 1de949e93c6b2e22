"""development: is the cnn leg bound by the host (Python mirror issuing ~100 evaluator calls per image) or by the GPU?
prints, per image: time to ENQUEUE the inference (no synchronisation) and time until the device has finished"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mkhe_kklss_b200 import params as PR, cnn, mkckks, mkrlwe
import bench
lit = PR.CNN_PN14QP433
inp = bench.cnn_host_inputs(lit, 4)
dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale)
ctx = dp.ctx
for idx, arr in inp["crs"].items():
    dp.SetCRS(idx, arr)
rl, rk = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet()
for i in (cnn.MODEL, cnn.DATA):
    rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(ctx, i, *inp["rlk"][i]))
    for r, a in inp["rk"][i].items():
        rk.AddRotationKey(i, r, mkrlwe.SwitchingKey(ctx, a))
up = lambda c: mkckks.Ciphertext.from_numpy(ctx, c, lit.scale)
E = cnn.DeviceFacade(mkckks.Evaluator(dp), rl, rk)
kernels, fc1 = [up(c) for c in inp["kernels"]], [up(c) for c in inp["fc1"]]
fc2, b1, b2 = up(inp["fc2"]), up(inp["b1"]), up(inp["b2"])
mask = mkrlwe.Poly.from_numpy(ctx, inp["mask"])
kh, fh = cnn.hoist_model(E, kernels, fc1)
images = [up(c) for c in inp["images"]]
if len(sys.argv) > 1 and sys.argv[1] == "graphs":
    ctx.set_graphs(True)
for rep in range(6 if len(sys.argv) > 1 else 3):
    for im in images:
        ctx.sync()
        t0 = time.perf_counter()
        out = cnn.infer(E, im, kernels, kh, fc1, fh, fc2, b1, b2, mask, lit.scale)[0]
        t1 = time.perf_counter()
        ctx.sync()
        t2 = time.perf_counter()
        print(f"rep {rep}: enqueue {1e3 * (t1 - t0):6.2f} ms, device done after {1e3 * (t2 - t0):6.2f} ms  (one lane, one image in flight)")
print("graph cache (captures, replays):", ctx.graph_stats())
ctx.profile_begin()
out = cnn.infer(E, images[0], kernels, kh, fc1, fh, fc2, b1, b2, mask, lit.scale)[0]
prof = ctx.profile_end()
tot = sum(v[1] for v in prof.values())
print(f"sum of kernel times per image: {tot:.2f} ms over {sum(v[0] for v in prof.values())} launches")
for name, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"   {name:26s} {v[1]:7.3f} ms  x{v[0]}")

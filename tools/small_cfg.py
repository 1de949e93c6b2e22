"""development: where does a small configuration (PN14QP439, k = 2: BASELINE config 1) spend its time?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mkhe_kklss_b200 import params as PR  # noqa: E402

if __name__ == "__main__":
    sys.argv = sys.argv[:1]
    from bench import DeviceWorkload
    for lanes in (1, 2):
        wl = DeviceWorkload(PR.CKKS_PN14QP439, 2, 0, seed=3, batch=16, lanes=lanes)
        ms = wl.timed(wl.mul_relin_step, 20, 3)
        print(f"lanes={lanes}: {16 * 20 / (ms * 1e-3):.0f} ops/s, {ms / 320 * 1e3:.1f} us per op (device time)")
        t0 = time.perf_counter()
        for i in range(20):
            wl.mul_relin_step(i)
        host = (time.perf_counter() - t0) / 320
        wl.sync()
        print(f"   host enqueue time per op: {host * 1e6:.1f} us")
        if lanes == 1:
            wl.ctx.profile_begin()
            for i in range(8):
                wl.mul_relin_op(i, lane=0)
            prof = wl.ctx.profile_end()
            tot = sum(v[1] for v in prof.values())
            print(f"   sum of kernel times per op: {tot / 8 * 1e3:.1f} us over {sum(v[0] for v in prof.values()) / 8:.0f} launches")
            for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]:
                print(f"      {k:28s} {v[1] / 8 * 1e3:7.1f} us  x{v[0] / 8:.0f}")
        wl.ctx.close()

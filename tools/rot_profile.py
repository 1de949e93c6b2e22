"""development: per-kernel times of one hoisted Rotate (k parties) on one lane"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mkhe_kklss_b200 import params as PR
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
if len(sys.argv) > 2:                      # another build of the library (A/B runs)
    from mkhe_kklss_b200 import _lib
    _lib._default = _lib.Library(os.path.abspath(sys.argv[2]))
sys.argv = sys.argv[:1]
from bench import DeviceWorkload
wl = DeviceWorkload(PR.CKKS_PN15QP880, k, 0, seed=3, batch=8, lanes=1)
wl.prepare_rotate()
ms = wl.timed(wl.rotate_step, 10, 3)
print(f"k={k} one lane: {8 * 10 / (ms * 1e-3):.0f} rotate ops/s, {ms / 80 * 1e3:.1f} us per op")
wl.ctx.profile_begin()
wl.rotate_step(0)
prof = wl.ctx.profile_end()
tot = sum(v[1] for v in prof.values())
print(f"sum of kernel times per op: {tot / 8 * 1e3:.1f} us")
for name, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"   {name:26s} {v[1] / 8 * 1e3:7.1f} us  x{v[0] / 8:.0f}")

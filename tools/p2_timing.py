"""development tool: per-CTA start / end times of the persistent pass-2 kernel (build with -DMKHE_P2_TIMING).
usage (GPU box): python tools/p2_timing.py [extra nvcc -D flags]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from mkhe_kklss_b200 import params as PR  # noqa: E402
from mkhe_kklss_b200._lib import Context, Library  # noqa: E402

OUT = os.path.join(ROOT, "tools", "_build", "libmkhe_dbg.so")


def build(extra):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-DMKHE_P2_TIMING", "-I" + os.path.join(ROOT, "include"), os.path.join(g.CSRC, "mkhe_api.cu"), "-o", OUT] + g._nccl_flags() + extra
    subprocess.check_call(cmd)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build([a for a in sys.argv[1:] if a.startswith("-D")])
        sys.exit(0)
    lib = Library(OUT)
    from mkhe_kklss_b200 import _lib
    _lib._default = lib                      # every Context of this process uses the instrumented build
    sys.argv = sys.argv[:1]
    from bench import DeviceWorkload
    wl = DeviceWorkload(PR.CKKS_PN15QP880, 4, 0, seed=5, batch=1)
    ctx = wl.ctx
    os.makedirs(os.path.join(ROOT, "gpurun_out", "p2t"), exist_ok=True)
    for rep in range(3):
        for i in range(4):
            wl.mul_relin_op(i)                   # the last pass-2 launch of a MulRelin is Decompose(p): 4 polys x 14 digits x 16 limbs
        ctx.sync()
        buf = (C.c_uint64 * (4 * 4096))()
        n = C.c_int(0)
        ctx.check(lib.dll.mkhe_debug_p2_timing(ctx.ptr, buf, 4096, C.byref(n)))
        t = np.array(buf[:4 * n.value], dtype=np.uint64).reshape(-1, 4).astype(np.int64)
        np.save(os.path.join(ROOT, "gpurun_out", "p2t", f"t{rep}.npy"), t)
        t0 = t[:, 0].min()
        st, en, sm, tiles = t[:, 0] - t0, t[:, 1] - t0, t[:, 2], t[:, 3]
        print(f"CTAs {n.value}  start spread {st.max()} ns  end min/median/max {en.min()} / {int(np.median(en))} / {en.max()} ns  tiles/CTA {tiles.min()}..{tiles.max()}")
        dur = en - st
        order = np.argsort(en)
        print("slowest 8 CTAs (cta, sm, dur ns, tiles):", [(int(i), int(sm[i]), int(dur[i]), int(tiles[i])) for i in order[-8:]])
        print("fastest 8 CTAs:", [(int(i), int(sm[i]), int(dur[i]), int(tiles[i])) for i in order[:8]])
        per = dur / np.maximum(tiles, 1)
        print(f"ns per tile: min {per.min():.0f} median {np.median(per):.0f} max {per.max():.0f}")
        # by first slot (big moduli: slot 0 and the P limbs at the end)
        q = np.quantile(en, [0.1, 0.5, 0.9, 0.99])
        print("end-time quantiles 10/50/90/99 %:", [int(x) for x in q])

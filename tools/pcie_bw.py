"""development: host<->device copy rates through the library's asynchronous transfer path (pinned memory, copy streams)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mkhe_kklss_b200 import params as PR
from mkhe_kklss_b200._lib import Context

lit = PR.CKKS_PN15QP880
ctx = Context(lit.logN, lit.Q, lit.P)
nl = len(lit.Q)
polys = [ctx.poly_alloc(nl) for _ in range(10)]
hosts = [ctx.host_alloc((nl, lit.N)) for _ in range(10)]
for mode in ("h2d", "d2h", "both"):
    ctx.sync()
    t0 = time.perf_counter()
    reps = 40
    for r in range(reps):
        for p, h in zip(polys, hosts):
            if mode in ("h2d", "both"):
                ctx.poly_upload_async(p, h)
            if mode in ("d2h", "both"):
                ctx.poly_download_async(p, h)
    ctx.sync()
    dt = time.perf_counter() - t0
    nbytes = reps * 10 * hosts[0].nbytes * (2 if mode == "both" else 1)
    print(f"{mode}: {nbytes / dt / 1e9:.1f} GB/s  ({dt / reps * 1e3:.2f} ms per 10 polys of {hosts[0].nbytes / 2**20:.1f} MiB)")

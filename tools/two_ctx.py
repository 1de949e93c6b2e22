"""development experiment: independent MulRelin ops on TWO contexts (two streams, separate scratch) of one GPU, interleaved.
Does the HBM-bound multiply-accumulate work of one op overlap the integer-bound transforms of the other?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mkhe_kklss_b200 import params as PR  # noqa: E402

if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    sys.argv = sys.argv[:1]
    if os.environ.get('MKHE_LIB'):
        from mkhe_kklss_b200 import _lib
        _lib._default = _lib.Library(os.path.abspath(os.environ['MKHE_LIB']))
    from bench import DeviceWorkload
    wls = [DeviceWorkload(PR.CKKS_PN15QP880, k, 0, seed=5 + i, batch=1) for i in range(nctx)]
    nops = 64

    def run(active):
        for w in active:
            w.ctx.sync()
        t0 = time.perf_counter()
        for i in range(nops):
            for w in active:
                w.mul_relin_op(i)
        for w in active:
            w.ctx.sync()
        return nops * len(active) / (time.perf_counter() - t0)

    for rep in range(2):
        print(f"k={k}: one context {run(wls[:1]):.1f} ops/s; {nctx} contexts interleaved {run(wls):.1f} ops/s")

#!/usr/bin/env python
"""multi-GPU stress of the limb-sharded MulRelinNew: `rounds` x B ops enqueued back to back over the lanes of every rank (one process
per GPU, torchrun), every result of every rank compared with the single-GPU op of the same rank on the same operands.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/stress_team.py [rounds] [B] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from mkhe_kklss_b200 import params as PR, mkckks
import bench

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 12
k = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lit = PR.CKKS_PN15QP880
wl = bench.DeviceWorkload(lit, k, local, seed=0xB2000777, batch=B, lanes=3 if world >= 4 else 2, team=(world, rank, dist), npairs=4, rots=())
outs = [mkckks.Ciphertext.new(wl.params, wl.ids, wl.level, lit.scale) for _ in range(B)]
nl = wl.level + 1 - wl.nb
want = []
for p in range(len(wl.pairs)):                       # the single-GPU op of this rank on pair p
    wl.mul_relin_op(p, lane=0, sharded=False)
    wl.sync()
    want.append(wl.result(0))
bad = 0
for r in range(rounds):
    for n in range(B):
        a, b = wl.pairs[n % len(wl.pairs)]
        for p_ in outs[n].Value.values():
            p_.set_nlimbs(wl.level + 1)
        wl._issue(n % len(wl.lanes), a, b, outs[n], True)
    wl.sync()
    for n in range(B):
        got = {kk: wl.ctx.poly_download(p_.h, nl) for kk, p_ in outs[n].Value.items()}
        w = want[n % len(wl.pairs)]
        if not all(np.array_equal(got[kk], w[kk]) for kk in w):
            bad += 1
            print(f"rank {rank} round {r} op {n}: MISMATCH in {[kk for kk in w if not np.array_equal(got[kk], w[kk])]}", flush=True)
t = torch.tensor([bad, int(wl.ctx.team_timed_out())], dtype=torch.int64, device="cuda")
dist.all_reduce(t)
if rank == 0:
    print(f"{int(t[0])} of {rounds * B * world} limb-sharded results (all ranks) differ from the single-GPU op; barrier time-outs: {int(t[1])}")
dist.destroy_process_group()

#!/usr/bin/env python
"""bench.py -- MK-CKKS MulRelin ops/s (and hoisted Rotate ops/s) at logN=15 on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference's path

A "step" is one pass over a batch of `--batch` (default 16) ciphertext pairs; each pair is one
`MulRelinNew(ct0, ct1, rlkSet)` exactly as mkckks_benchmark_test.go:78-82 times it: hoist both operands,
MulAndRelinHoisted, Rescale -- on synthetic uniform-random ciphertexts and keys of the PN15QP880 set
(mkckks_test.go:51-72), op0 != op1, both with all k party ids, level 13.  value = MulRelin ops/s.

  value  : steps/s with ciphertexts and keys already resident in HBM (CUDA events on the library's stream,
           max over ranks).  N > 1 = independent ciphertext batches per GPU, keys replicated (weak scaling).
  e2e    : the same call through the host-buffer API: both operand ciphertexts are uploaded from pinned host memory, the
           op runs, the result ciphertext is downloaded, for every op, all inside the timed region (asynchronous
           transfers: the copies of neighbouring ops overlap the compute; host clock between two mkhe_sync).
Only the cpu_baseline leg and `--impl reference` touch oracle/ (as the thing timed on the CPU, never as
part of the GPU path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mkhe_kklss_b200 import params as PR  # noqa: E402

METRIC = "MK-CKKS MulRelin ops/s at logN=15 (PN15QP880, level 13)"
UNIT = "ops/s"


# ---------------------------------------------------------------------------------------------------
def uniform_limbs(rng, moduli, shape_prefix, N):
    """i.i.d. uniform limbs in [0, q_j) (throughput runs)"""
    out = np.empty(tuple(shape_prefix) + (len(moduli), N), dtype=np.uint64)
    for j, q in enumerate(moduli):
        out[..., j, :] = rng.integers(0, q, size=tuple(shape_prefix) + (N,), dtype=np.uint64)
    return out


def iter_host_keys(lit, k, seed, rots=()):
    """uniform key material in a fixed order, one switching key at a time (k = 32 is 5.4 GiB: nothing is kept on the host)"""
    rng = np.random.default_rng(seed)
    mods = list(lit.Q) + list(lit.P)
    beta = len(lit.Q)
    mk = lambda: uniform_limbs(rng, mods, (beta,), lit.N)
    yield ("u", None, None, mk())
    for i in range(k):
        for j in range(3):
            yield ("rlk", i, j, mk())
    for r in rots:
        yield ("a", r, None, mk())
    for i in range(k):
        for r in rots:
            yield ("rk", i, r, mk())


def host_keys(lit, k, seed, rots=()):
    keys = {"rlk": [[None] * 3 for _ in range(k)], "a": {}, "rk": [{} for _ in range(k)]}
    for kind, i, j, arr in iter_host_keys(lit, k, seed, rots):
        if kind == "u":
            keys["u"] = arr
        elif kind == "rlk":
            keys["rlk"][i][j] = arr
        elif kind == "a":
            keys["a"][i] = arr
        else:
            keys["rk"][i][j] = arr
    return keys


def host_ct(lit, k, level, rng):
    return {"0" if i < 0 else i: uniform_limbs(rng, lit.Q[:level + 1], (), lit.N) for i in range(-1, k)}


def algorithmic_model(k, ell, nP, N):
    """compulsory bytes / butterflies of one MulRelinNew and the per-kernel algorithmic bytes of
    the current (materialising) schedule, per step"""
    D, beta, limb = ell + nP, ell, 8 * N
    b_mul = (3 * k * beta * D + beta * D + 2 * (k + 1) * ell + (k + 1) * (ell - 1)) * limb
    logN = N.bit_length() - 1
    bfly_limb = (N // 2) * logN
    fwd, inv = 3 * k * beta * D + (2 * k + 2) * ell, (k + 1) * ell + 4 * k * D
    kern = {
        "k_bcast_ntt_pass1": (3 * k * ell + 3 * k * beta * D) * limb,
        # the tensor step transforms only the two "0" components: NTT(op_id) is read from the diagonal of the hoisted forms
        "k_ntt_pass2": 2 * (3 * k * beta * D + 2 * ell) * limb,
        "k_ntt_pass1": 2 * 2 * ell * limb,
        "k_mac_parties": (4 * k * beta * D + 2 * beta * D) * limb,
        # digit multiply-accumulate fused with the inverse pass A: keys + hoisted forms + x, y, u once, 4k QP accumulators out
        "k_mac_intt": (4 * k * beta * D + 3 * beta * D + 4 * k * D) * limb,
        "k_intt_passA": 2 * (k + 1) * ell * limb,                      # tensor outputs only
        "k_intt_passB": 2 * (k + 1) * ell * limb,
        "k_moddown_P": 4 * k * 3 * nP * limb,
        "k_moddown_Q": (4 * k * (ell + 2 * nP) + (2 * k + 1) * ell + (3 * k + 1) * ell) * limb,
        "k_tensor": (3 * k + 3) * ell * limb,
        "k_rescale": (k + 1) * (2 * ell - 1) * limb,
    }
    s1 = logN - 11                 # column stages (pass 1 / pass B); the tile passes do the other 11
    bfly = {
        "k_bcast_ntt_pass1": 3 * k * beta * D * (N // 2) * s1,
        "k_ntt_pass2": (3 * k * beta * D + 2 * ell) * (N // 2) * 11,
        "k_ntt_pass1": 2 * ell * (N // 2) * s1,
        "k_mac_intt": 4 * k * D * (N // 2) * 11,
        "k_intt_passA": (k + 1) * ell * (N // 2) * 11,
        "k_intt_passB": (k + 1) * ell * (N // 2) * s1,
        "k_moddown_P": 4 * k * nP * (N // 2) * s1,
        "k_moddown_Q": 4 * k * ell * (N // 2) * s1,
    }
    return {"compulsory_bytes": b_mul, "butterflies": (fwd + inv) * bfly_limb, "kernel_bytes": kern, "kernel_butterflies": bfly}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def nvlink_bytes(device):
    """cumulative NVLink data counters of one GPU: (tx bytes, rx bytes) summed over its links (`nvidia-smi nvlink -gt d`), or None"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(device)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = rx = 0
    found = False
    for ln in out.splitlines():
        f = ln.replace(":", " ").split()
        if "KiB" in f and ("Tx" in ln or "Rx" in ln):
            try:
                v = int(f[f.index("KiB") - 1]) * 1024
            except ValueError:
                continue
            found = True
            if "Tx" in ln:
                tx += v
            else:
                rx += v
    return (tx, rx) if found else None


# ---------------------------------------------------------------------------------------------------
class DeviceWorkload:
    """keys + ciphertext pool of one GPU"""

    def __init__(self, lit, k, device, seed, rots=(2,), npairs=4, batch=16, lanes=2, team=None):
        """team = (world, rank, dist): the ops are limb-sharded over the ranks (mkhe_ckks_mul_relin_limbs); every rank builds the
        SAME keys and ciphertexts (same seed) and every lane joins a team of its own"""
        from mkhe_kklss_b200 import mkckks, mkrlwe
        self.lit, self.k, self.rots, self.batch, self.team = lit, k, rots, batch, team
        self._own = {}
        self.level = len(lit.Q) - 1
        self.params = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, device=device)
        self.ctx = self.params.ctx
        self.ev = mkckks.Evaluator(self.params)
        self.rlk = mkrlwe.RelinearizationKeySet()
        self.rk = mkrlwe.RotationKeySet()
        part = {}
        for kind, i, j, arr in iter_host_keys(lit, k, seed, rots):      # uploaded one by one
            if kind == "u":
                self.params.SetCRS(-1, arr)
            elif kind == "a":
                self.params.SetCRS(i, arr)
            elif kind == "rk":
                self.rk.AddRotationKey(i, j, mkrlwe.SwitchingKey(self.ctx, arr))
            else:
                part[j] = mkrlwe.SwitchingKey(self.ctx, arr)
                if j == 2:
                    rl = mkrlwe.RelinearizationKey.__new__(mkrlwe.RelinearizationKey)
                    rl.ID, rl.Value = i, [part[0], part[1], part[2]]
                    self.rlk.AddRelinearizationKey(rl)
                    part = {}
        rng = np.random.default_rng(seed + 99)
        self.host_pairs = [(host_ct(lit, k, self.level, rng), host_ct(lit, k, self.level, rng)) for _ in range(npairs)]
        self.pairs = [(mkckks.Ciphertext.from_numpy(self.ctx, a, lit.scale), mkckks.Ciphertext.from_numpy(self.ctx, b, lit.scale))
                      for a, b in self.host_pairs]
        self.ids = list(range(k))
        # lanes: forks of the context (mkhe_ctx_fork) -- same keys and ciphertext handles, own stream and scratch pools; op n
        # runs on lane n % lanes, so the bandwidth-bound stages of one op overlap the integer-bound ones of its neighbour
        self.lanes = [self.ctx] + [self.ctx.fork() for _ in range(lanes - 1)]
        if team:
            world, rank, dist = team
            for ln in self.lanes:                       # CUDA IPC handles of every rank's team memory, lane by lane
                hs = [None] * world
                dist.all_gather_object(hs, ln.team_export(k))
                ln.team_import(world, rank, hs)
        self.outs = [mkckks.Ciphertext.new(self.params, self.ids, self.level, lit.scale) for _ in self.lanes]
        self.out = self.outs[0]
        self.nb, self.new_scale = self.ev._nb_rescales(lit.scale * lit.scale, self.level, lit.scale)
        g = self.rlk.GetRelinearizationKey
        self.kb = [g(i).Value[0].h for i in self.ids]
        self.kd = [g(i).Value[1].h for i in self.ids]
        self.kv = [g(i).Value[2].h for i in self.ids]
        self.ctx.sync()

    def _issue(self, ln, a, b, out, sharded=None):
        """one MulRelinNew on lane ln: limb-sharded over the team when there is one"""
        if (self.team is not None) if sharded is None else sharded:
            self.lanes[ln].ckks_mul_relin_limbs(self.level, self.nb, self.ids, a.handles(self.ids), self.ids, b.handles(self.ids),
                                                self.kb, self.kd, self.kv, self.params.CRS[-1].h, self.ids, out.handles(self.ids))
        else:
            self.lanes[ln].ckks_mul_relin(self.level, self.nb, False, self.ids, a.handles(self.ids), self.ids, b.handles(self.ids),
                                          self.kb, self.kd, self.kv, self.params.CRS[-1].h, self.ids, out.handles(self.ids))

    def mul_relin_op(self, i, lane=None, sharded=None):
        a, b = self.pairs[i % len(self.pairs)]
        ln = i % len(self.lanes) if lane is None else lane
        self._issue(ln, a, b, self.outs[ln], sharded)

    def result(self, lane=0):
        nl = self.level + 1 - self.nb
        return {kk: self.ctx.poly_download(p_.h, nl) for kk, p_ in self.outs[lane].Value.items()}

    def mul_relin_step(self, i):
        for j in range(self.batch):
            self.mul_relin_op(i * self.batch + j)

    def api_step(self, i):
        """the call a user of the reference API makes: Evaluator.MulRelinNew through the host mirror -- a fresh result ciphertext is
        allocated by every call and the previous one dropped (mkckks_benchmark_test.go:78-82 times exactly that, allocation included)"""
        import copy
        from mkhe_kklss_b200 import mkckks
        if not hasattr(self, "api_evs"):
            self.api_evs, self.api_keep = [], {}
            for ln in self.lanes:
                p2 = copy.copy(self.params)
                p2.ctx = ln
                self.api_evs.append(mkckks.Evaluator(p2))
        for j in range(self.batch):
            n = i * self.batch + j
            a, b = self.pairs[n % len(self.pairs)]
            ln = n % len(self.lanes)
            self.api_keep[ln] = self.api_evs[ln].MulRelinNew(a, b, self.rlk)      # the lane's previous result is released here

    def sync(self):
        for ln in self.lanes:
            ln.sync()

    def launch_count(self):
        return sum(ln.launch_count() for ln in self.lanes)

    def prepare_rotate(self):
        self.hoisted = [self.ev.HoistedForm(a) for a, _ in self.pairs]
        for o in self.outs:
            for p in o.Value.values():
                p.set_nlimbs(self.level + 1)
        self.sync()

    def rotate_step(self, i, rot=2):
        for b in range(self.batch):
            n = i * self.batch + b
            j, ln = n % len(self.pairs), n % len(self.lanes)
            a = self.pairs[j][0]
            self.lanes[ln].rotate_hoisted(self.level, rot, a.handles(self.ids), [self.hoisted[j][t].h for t in self.ids],
                                          [self.rk.GetRotationKey(t, rot).h for t in self.ids], self.params.CRS[rot].h,
                                          self.outs[ln].handles(self.ids))

    def timed(self, fn, steps, warmup, barrier=None):
        """W untimed steps, then exactly K steps between CUDA events on the root lane's stream: the other lanes start after
        the first event (mkhe_ctx_wait) and the root lane waits for them before the second one"""
        for i in range(warmup):
            fn(i)
        self.sync()
        if barrier:
            barrier()
        l0 = self.launch_count()
        self.ctx.timer_start()
        for ln in self.lanes[1:]:
            ln.wait(self.ctx)
        for i in range(steps):
            fn(i)
        for ln in self.lanes[1:]:
            self.ctx.wait(ln)
        ms = self.ctx.timer_stop()
        self.last_launches = self.launch_count() - l0
        if barrier:
            barrier()
        return ms

    # -- end to end through host buffers ------------------------------------------------------------
    def prepare_e2e(self):
        """page-locked host copies of the operand ciphertexts (library-owned, mkhe_host_alloc), one result ciphertext per lane
        on the device and its host landing buffer"""
        self.pin = []
        for a, b in self.host_pairs:
            pa, pb = {}, {}
            for src, dst in ((a, pa), (b, pb)):
                for kk, arr in src.items():
                    dst[kk] = self.ctx.host_alloc(arr.shape)
                    dst[kk][...] = arr
            self.pin.append((pa, pb))
        from mkhe_kklss_b200 import mkckks
        nl = self.level + 1 - self.nb
        # two result ciphertexts (and host landing buffers) per lane: op n+lanes writes one while result n streams back from the other
        self.e2e_outs = [[self.outs[ln], mkckks.Ciphertext.new(self.params, self.ids, self.level, self.lit.scale)] for ln in range(len(self.lanes))]
        self.res_host = [[{kk: self.ctx.host_alloc((nl, self.lit.N)) for kk in ["0"] + self.ids} for _ in range(2)] for _ in self.lanes]
        self.e2e_n = 0
        self.e2e_primed = False

    def _upload_pair(self, n):
        """asynchronous upload of host pair n into device slot n (both modulo the pool size) on op n's lane"""
        lane = self.lanes[n % len(self.lanes)]
        a, b = self.pairs[n % len(self.pairs)]
        pa, pb = self.pin[n % len(self.pin)]
        nbytes = 0
        for ct, hp in ((a, pa), (b, pb)):
            for kk, poly in ct.Value.items():
                if self.team:           # this rank's share of the ciphertext: the limbs it owns (the all-gather over NVLink follows)
                    lane.poly_upload_owned_async(poly.h, hp[kk])
                    nbytes += len(self.own_limbs(self.level + 1)) * hp[kk][0].nbytes
                else:
                    lane.poly_upload_async(poly.h, hp[kk])
                    nbytes += hp[kk].nbytes
        return nbytes

    def own_limbs(self, nlimbs):
        if nlimbs not in self._own:
            self._own[nlimbs] = [j for j in range(nlimbs) if self.ctx.team_owns_limb(j)]
        return self._own[nlimbs]

    def e2e_step(self, i):
        """`batch` times through the C ABI with HOST buffers: upload both operand ciphertexts, MulRelinNew, download the
        result ciphertext -- all three on the op's lane (op n runs on lane n % lanes).  The transfers are asynchronous
        (mkhe_poly_upload_async / _download_async on pinned memory, every lane has its own copy streams): the operands of a
        lane's NEXT op are uploaded while its current op runs, results stream back behind the compute, and the copies of one
        lane overlap the compute of the other.  Every op waits for its own operands; a device slot is not overwritten before
        its last reader has finished (the library orders uses of one object); the region ends with mkhe_sync on every lane."""
        h2d = d2h = 0
        L = len(self.lanes)
        if not self.e2e_primed:
            for j in range(L):
                self._upload_pair(self.e2e_n + j)          # pipeline prologue (first call only, inside the warm-up)
            self.e2e_primed = True
        for _ in range(self.batch):
            n = self.e2e_n
            ln = n % L
            lane = self.lanes[ln]
            h2d += self._upload_pair(n + L)                # the lane's next operands: copied beside op n
            a, b = self.pairs[n % len(self.pairs)]
            par = (n // L) % 2
            out = self.e2e_outs[ln][par]
            if self.team:               # every rank uploaded 1 / N of both operands: make them whole on every rank
                lane.team_allgather(self.level, a.handles(self.ids) + b.handles(self.ids))
            self._issue(ln, a, b, out)
            for kk, poly in out.Value.items():
                dst = self.res_host[ln][par][kk]
                if self.team:           # ... and reads back its share of the result
                    lane.poly_download_owned_async(poly.h, dst)
                    d2h += len(self.own_limbs(dst.shape[0])) * dst[0].nbytes
                else:
                    lane.poly_download_async(poly.h, dst)
                    d2h += dst.nbytes
            self.e2e_n += 1
        return h2d, d2h

    def timed_wall(self, fn, steps, warmup, barrier=None):
        """like timed(), for regions that include host<->device copies on the copy streams: host clock between two full
        synchronisations (mkhe_sync waits for the context's stream and both copy streams)"""
        for i in range(warmup):
            fn(i)
        self.sync()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        self.sync()
        ms = (time.perf_counter() - t0) * 1e3
        if barrier:
            barrier()
        return ms


def sharded_mul_relin(lit, k, rank, world, local_rank, steps, warmup, batch, dist, barrier, allmax, nlanes=2, p2p=True):
    """BASELINE config 4: ONE MulRelinNew whose per-party key switches are sharded over the ranks (mkhe_ckks_mul_relin_sharded:
    rank g holds the relinearisation keys of its parties only, the partial x, y and the c_0 contributions are summed with
    ncclAllReduce over NVLink and reduced mod q).  Operand ciphertexts are replicated (they are small); every rank issues the
    same ops.  Returns ops/s of the whole group (strong scaling: total work fixed) -- CUDA events, max over ranks."""
    from mkhe_kklss_b200 import mkckks, mkrlwe, sharding
    level = len(lit.Q) - 1
    params = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, device=local_rank)
    ctx = params.ctx
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])

    def enable_p2p(c):
        """fused exchange over peer memory (mkhe_p2p_export / _import) instead of the 112 MiB all-reduce of x || y"""
        hs = [None] * world
        dist.all_gather_object(hs, c.p2p_export())
        c.p2p_import(world, rank, hs)

    if p2p:
        enable_p2p(ctx)
    mods, beta = list(lit.Q) + list(lit.P), len(lit.Q)
    mk = lambda rng: uniform_limbs(rng, mods, (beta,), lit.N)
    params.SetCRS(-1, mk(np.random.default_rng(0xB2000040)))            # the CRS is common to all ranks
    ids = list(range(k))
    own = sharding.owned_parties(ids, world, rank)
    rlk = mkrlwe.RelinearizationKeySet()
    for i in own:
        rng = np.random.default_rng(0xB2000041 + i)
        rlk.AddRelinearizationKey(mkrlwe.RelinearizationKey(ctx, i, mk(rng), mk(rng), mk(rng)))
    rng = np.random.default_rng(0xB2000042)                             # replicated operands: same seed on every rank
    pairs = [(mkckks.Ciphertext.from_numpy(ctx, host_ct(lit, k, level, rng), lit.scale),
              mkckks.Ciphertext.from_numpy(ctx, host_ct(lit, k, level, rng), lit.scale)) for _ in range(2)]
    ev = mkckks.Evaluator(params)
    nb, _ = ev._nb_rescales(lit.scale * lit.scale, level, lit.scale)
    key = lambda i, j: rlk.Value[i].Value[j].h if i in rlk.Value else 0
    kb, kd, kv = [key(i, 0) for i in ids], [key(i, 1) for i in ids], [key(i, 2) for i in ids]
    # two lanes, each with its own NCCL communicator: the all-reduce of one op runs beside the transforms of the next one
    # (every rank issues op n on lane n % 2, so the collectives of each communicator are issued in the same order everywhere)
    lanes = [ctx]
    for _ in range(nlanes - 1):
        ln = ctx.fork()
        uid = [ln.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ln.comm_init(world, rank, uid[0])
        if p2p:
            enable_p2p(ln)
        lanes.append(ln)
    outs = [mkckks.Ciphertext.new(params, ids, level, lit.scale) for _ in lanes]

    def step(i):
        for j in range(batch):
            n = i * batch + j
            a, b = pairs[n % len(pairs)]
            ln = n % len(lanes)
            lanes[ln].ckks_mul_relin_sharded(level, nb, ids, a.handles(ids), ids, b.handles(ids), own, kb, kd, kv,
                                             params.CRS[-1].h, ids, outs[ln].handles(ids))

    for i in range(warmup):
        step(i)
    for ln in lanes:
        ln.sync()
    barrier()
    ctx.timer_start()
    for ln in lanes[1:]:
        ln.wait(ctx)
    for i in range(steps):
        step(i)
    for ln in lanes[1:]:
        ctx.wait(ln)
    ms = allmax(ctx.timer_stop())
    barrier()
    ctx.close()
    return batch * steps / (ms * 1e-3)


def bfv_mul_relin_ops(lit, k, device, steps, warmup, batch, nlanes):
    """BASELINE config 3: mkbfv MulRelinNew (ModUpQtoR / Rescale to R, DecomposeBFV, MulAndRelinBFVHoisted, Quantize), uniform
    keys and ciphertexts, ops spread over the lanes"""
    from mkhe_kklss_b200 import mkbfv, mkrlwe
    params = mkbfv.Parameters(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, device=device)
    ctx = params.ctx
    rng = np.random.default_rng(0xB2000300)
    mods, beta = list(lit.Q) + list(lit.P), len(lit.Q)
    mk = lambda: uniform_limbs(rng, mods, (beta,), lit.N)
    params.SetCRS(-1, mk())
    rlk = mkbfv.RelinearizationKeySet()
    for i in range(k):
        rlk.AddRelinearizationKey(mkbfv.RelinearizationKey(ctx, i, mk(), mk(), mk(), mk(), mk()))
    level = len(lit.Q) - 1
    ids = list(range(k))
    pairs = [(mkrlwe.Ciphertext.from_numpy(ctx, host_ct(lit, k, level, rng)), mkrlwe.Ciphertext.from_numpy(ctx, host_ct(lit, k, level, rng)))
             for _ in range(2)]
    lanes = [ctx] + [ctx.fork() for _ in range(nlanes - 1)]
    outs = [mkrlwe.Ciphertext.new(ctx, ids, level) for _ in lanes]
    g = rlk.GetRelinearizationKey
    kb1, kb2 = [g(i).b1.h for i in ids], [g(i).b2.h for i in ids]
    kd1, kd2, kv = [g(i).d1.h for i in ids], [g(i).d2.h for i in ids], [g(i).v.h for i in ids]

    def step(i):
        for j in range(batch):
            n = i * batch + j
            a, b = pairs[n % 2]
            ln = n % len(lanes)
            lanes[ln].bfv_mul_relin(ids, a.handles(ids), ids, b.handles(ids), kb1, kb2, kd1, kd2, kv, params.CRS[-1].h, ids,
                                    outs[ln].handles(ids))

    for i in range(warmup):
        step(i)
    for ln in lanes:
        ln.sync()
    ctx.timer_start()
    for ln in lanes[1:]:
        ln.wait(ctx)
    for i in range(steps):
        step(i)
    for ln in lanes[1:]:
        ctx.wait(ln)
    ms = ctx.timer_stop()
    ctx.close()
    return batch * steps / (ms * 1e-3)


# ---------------------------------------------------------------------------------------------------
# BASELINE config 5: the reference's encrypted-CNN inference (cnn/cnn.go, timed by cnn/cnn_bench_test.go) over a batch of images
def cnn_host_inputs(lit, nimages, seed=0xB2000005):
    """synthetic inputs of the flow as host arrays: CRS entries, uniform relinearisation / rotation keys of the model owner and the
    data owner (24 rotation keys each), model ciphertexts, the mask plaintext and `nimages` image ciphertexts at the top level"""
    from mkhe_kklss_b200 import cnn
    rng = np.random.default_rng(seed)
    mods, beta, L = list(lit.Q) + list(lit.P), len(lit.Q), len(lit.Q) - 1
    sw = lambda: uniform_limbs(rng, mods, (beta,), lit.N)
    rots = [r for r in cnn.cnn_rotations(lit.logN) if r < lit.N // 2]
    ct = lambda ids: {("0" if i < 0 else i): uniform_limbs(rng, lit.Q[:L + 1], (), lit.N) for i in [-1] + list(ids)}
    M, Dt = cnn.MODEL, cnn.DATA
    return {"rots": rots, "crs": {idx: sw() for idx in [-1] + rots},
            "rlk": {i: (sw(), sw(), sw()) for i in (M, Dt)}, "rk": {i: {r: sw() for r in rots} for i in (M, Dt)},
            "kernels": [ct([M]) for _ in range(4)], "fc1": [ct([M]) for _ in range(8)], "fc2": ct([M]), "b1": ct([M]), "b2": ct([M]),
            "mask": uniform_limbs(rng, lit.Q[:L + 1], (), lit.N), "images": [ct([Dt]) for _ in range(nimages)]}


def cnn_leg(lit, device, nimages, nlanes, rounds=4):
    """images/s of the whole inference (HoistedForm(image), Convolution, square, FC1Layer, square, FC2Layer; the model's hoisted
    forms are precomputed like in cnn_bench_test.go:43-52) through the host mirror of the reference API, independent images
    alternating over the lanes.  Returns (device-resident images/s, end-to-end images/s with the image ciphertext uploaded and the
    result downloaded per image, the results of image 0 as host arrays, inputs)"""
    from mkhe_kklss_b200 import cnn, mkckks, mkrlwe
    inp = cnn_host_inputs(lit, nimages)
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, device=device)
    ctx = dp.ctx
    for idx, arr in inp["crs"].items():
        dp.SetCRS(idx, arr)
    rl, rk = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet()
    for i in (cnn.MODEL, cnn.DATA):
        rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(ctx, i, *inp["rlk"][i]))
        for r, a in inp["rk"][i].items():
            rk.AddRotationKey(i, r, mkrlwe.SwitchingKey(ctx, a))
    up = lambda c: mkckks.Ciphertext.from_numpy(ctx, c, lit.scale)
    ev0 = mkckks.Evaluator(dp)
    evs = [ev0] + [ev0.ShallowCopy() for _ in range(nlanes - 1)]
    Es = [cnn.DeviceFacade(e, rl, rk) for e in evs]
    kernels, fc1 = [up(c) for c in inp["kernels"]], [up(c) for c in inp["fc1"]]
    fc2, b1, b2 = up(inp["fc2"]), up(inp["b1"]), up(inp["b2"])
    mask = mkrlwe.Poly.from_numpy(ctx, inp["mask"])
    kh, fh = cnn.hoist_model(Es[0], kernels, fc1)
    images = [up(c) for c in inp["images"]]

    def sync():
        for e in evs:
            e.ctx.sync()

    def run(n, image):
        return cnn.infer(Es[n % nlanes], image, kernels, kh, fc1, fh, fc2, b1, b2, mask, lit.scale)[0]

    first = run(0, images[0])                                   # warm-up (and the parity sample)
    sync()
    res0 = first.numpy()
    keep = [run(n, im) for n, im in enumerate(images)]          # one untimed round: pools and allocator reach their steady state
    sync()
    l0 = sum(e.ctx.launch_count() for e in evs)
    t0 = time.perf_counter()
    for r in range(rounds):
        keep = [run(n, im) for n, im in enumerate(images)]
    sync()
    dt = time.perf_counter() - t0
    launches = (sum(e.ctx.launch_count() for e in evs) - l0) / (rounds * nimages)
    t0 = time.perf_counter()
    for r in range(rounds):
        outs = [run(n, up(c)) for n, c in enumerate(inp["images"])]     # upload per image (2 polys x 7 limbs)
        host = [o.numpy() for o in outs]                                 # result download per image (level 0)
    dt_e2e = time.perf_counter() - t0
    del keep, outs, host
    ctx.close()
    return rounds * nimages / dt, rounds * nimages / dt_e2e, res0, inp, launches


def cnn_cpu_run(lit, inp, threads=1):
    """the oracle port of the same inference on the host (one image; checker + CPU baseline): returns (result arrays, seconds)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cnn_flow as F
    from oracle import oracle as O
    from mkhe_kklss_b200 import cnn
    cores = O.set_threads(threads)
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, crs_rots=[])
    for idx, arr in inp["crs"].items():
        op.CRS[idx] = arr
    rlk = {i: O.RelinKey(i, *inp["rlk"][i]) for i in (cnn.MODEL, cnn.DATA)}
    E = F.OracleFacade(O.CKKSEvaluator(op, lit.scale), rlk, inp["rk"])
    oc = lambda c: O.Ciphertext({k: v.copy() for k, v in c.items()}, lit.scale)
    kernels, fc1 = [oc(c) for c in inp["kernels"]], [oc(c) for c in inp["fc1"]]
    kh, fh = cnn.hoist_model(E, kernels, fc1)
    t0 = time.perf_counter()
    out = cnn.infer(E, oc(inp["images"][0]), kernels, kh, fc1, fh, oc(inp["fc2"]), oc(inp["b1"]), oc(inp["b2"]), inp["mask"], lit.scale)[0]
    return out.value, time.perf_counter() - t0, cores


def keygen_leg(lit, k, device, rots=(1, 2)):
    """the key set of k parties (secret + public + relinearisation b, d, v + rotation keys) made on the device from
    the counter-based streams -- seconds and bytes of key material that never cross PCIe"""
    from mkhe_kklss_b200 import mkckks, mkrlwe
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, device=device)
    for idx in [0, -1] + list(rots):
        dp.AddCRS(idx, 0xC125)
    kg = mkrlwe.KeyGenerator(dp, 0x5EED, 0)
    dp.ctx.sync()
    t0 = time.perf_counter()
    keep = []
    for i in range(k):
        sk, r = kg.GenSecretKey(i), kg.GenSecretKey(i)
        keep.append((kg.GenPublicKey(sk), kg.GenRelinearizationKey(sk, r), [kg.GenRotationKey(rot, sk) for rot in rots]))
    dp.ctx.sync()
    dt = time.perf_counter() - t0
    nbytes = k * (3 + len(rots)) * len(lit.Q) * (len(lit.Q) + len(lit.P)) * lit.N * 8
    del keep
    dp.ctx.close()
    return dt, nbytes


def cpu_oracle_run(lit, k, steps, warmup, threads, seed=0xB2000002):
    """times the oracle port of MulRelinNew on the host cores; returns (ops/s, cores used, seconds per op)"""
    from oracle import oracle as O
    cores = O.set_threads(threads)
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=seed, crs_rots=[])
    hk = host_keys(lit, k, seed)
    p.CRS[-1] = hk["u"]
    rl = {i: O.RelinKey(i, *hk["rlk"][i]) for i in range(k)}
    ev = O.CKKSEvaluator(p, lit.scale)
    rng = np.random.default_rng(seed + 99)
    level = len(lit.Q) - 1
    c0 = O.Ciphertext(host_ct(lit, k, level, rng), lit.scale)
    c1 = O.Ciphertext(host_ct(lit, k, level, rng), lit.scale)
    for _ in range(warmup):
        ev.mul_relin_new(c0, c1, rl)
    t0 = time.perf_counter()
    for _ in range(steps):
        ev.mul_relin_new(c0, c1, rl)
    dt = time.perf_counter() - t0
    return steps / dt, cores, dt / steps


def oracle_result(lit, k, seed=0xB2000002, pair=0, threads=0):
    """the checker: MulRelinNew of ciphertext pair `pair` of the timed workload (same seeds as DeviceWorkload) on the oracle"""
    from oracle import oracle as O
    O.set_threads(threads or (os.cpu_count() or 1))
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=seed, crs_rots=[])
    hk = host_keys(lit, k, seed)
    p.CRS[-1] = hk["u"]
    rl = {i: O.RelinKey(i, *hk["rlk"][i]) for i in range(k)}
    ev = O.CKKSEvaluator(p, lit.scale)
    rng = np.random.default_rng(seed + 99)
    level = len(lit.Q) - 1
    for _ in range(pair + 1):
        a, b = host_ct(lit, k, level, rng), host_ct(lit, k, level, rng)
    return ev.mul_relin_new(O.Ciphertext(a, lit.scale), O.Ciphertext(b, lit.scale), rl).value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--parties", type=int, default=8, help="parties of the headline workload (BASELINE config 2: n = 4 and n = 8)")
    ap.add_argument("--batch", type=int, default=16, help="ciphertext pairs (MulRelin ops) per step")
    ap.add_argument("--lanes", type=int, default=0, help="lanes (mkhe_ctx_fork) the ops of a step are spread over, per GPU (default: 2 on "
                                                         "one GPU, 3 when the ops are sharded over 4+ GPUs: a third op in flight hides the barriers)")
    ap.add_argument("--no-extras", action="store_true", help="skip the k=8 / hoisted-Rotate / other-config side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lib", default=None, help="development: load this build of the library instead of the in-tree one")
    ap.add_argument("--sweep", default=None, help="party-count sweep (BASELINE config 4), e.g. 2,4,8,16,32: prints one JSON line "
                                                  "with MulRelin and hoisted-Rotate ops/s per k instead of the headline run")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.lib:
        from mkhe_kklss_b200 import _lib
        _lib._default = _lib.Library(os.path.abspath(args.lib))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.lanes <= 0:
        args.lanes = 3 if world >= 4 else 2
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lit, k = PR.CKKS_PN15QP880, args.parties
    ell, nP, N = len(lit.Q), len(lit.P), lit.N
    config = {"workload": f"mkckks MulRelinNew (hoist + MulAndRelinHoisted + Rescale), {lit.name}, logN={lit.logN}, level {ell - 1}, "
                          f"k={k} parties, op0 != op1",
              "params": lit.name, "logN": lit.logN, "parties": k, "level": ell - 1, "ops_per_step": args.batch,
              "lanes_per_gpu": args.lanes,
              "parallelism": "one GPU" if world == 1 else f"every op limb-sharded over the {world} GPUs (rank r owns the limb slots s with s mod {world} == r; "
                                                         "peer-store exchanges over NVLink, in-stream barriers; total work fixed)",
              "l2_policy": "inputs larger than L2: every step streams the relinearisation keys "
                           f"({(3 * k + 1) * ell * (ell + nP) * 8 * N / 2**20:.0f} MiB) and cycles 4 ciphertext pairs"}

    # ---------------- CPU arm ----------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        ncpu = os.cpu_count() or 1
        config = dict(config, ops_per_step=1, lanes_per_gpu=0, parallelism=f"host CPU, {ncpu} threads (OpenMP over limbs)",
                      l2_policy="n/a (CPU arm)")            # one op per step here; the GPU arm's batch does not apply
        ops, cores, sec = cpu_oracle_run(lit, k, args.steps, args.warmup, ncpu)
        line = {"impl": "reference", "metric": METRIC, "value": ops, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": ops, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{args.steps} full MulRelinNew calls (oracle/ C restatement with OpenMP over limbs; the reference itself "
                                           "is Go + un-vendored lattigo and cannot be built here)"},
                "e2e": {"value": ops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------- party-count sweep (single GPU) -----------------------------------------------
    if args.sweep:
        res = {}
        for ks in [int(x) for x in args.sweep.split(",")]:
            w = DeviceWorkload(lit, ks, local_rank, seed=0xB2000100 + ks, batch=8, lanes=args.lanes)
            ms = w.timed(w.mul_relin_step, max(args.steps // 2, 2), warmup)
            mul = w.batch * max(args.steps // 2, 2) / (ms * 1e-3)
            w.prepare_rotate()
            ms = w.timed(w.rotate_step, max(args.steps // 2, 2), warmup)
            rot = w.batch * max(args.steps // 2, 2) / (ms * 1e-3)
            m = algorithmic_model(ks, ell, nP, N)
            res[str(ks)] = {"mulrelin_ops_s": mul, "rotate_hoisted_ops_s": rot,
                            "mulrelin_int_frac": m["butterflies"] * mul / w.ctx.butterfly_peak(),
                            "mulrelin_hbm_frac_of_compulsory": m["compulsory_bytes"] * mul / 1e9 / 6549.4}
            w.ctx.close()
            del w
        print(json.dumps({"metric": "party-count sweep, " + METRIC, "unit": UNIT, "lanes_per_gpu": args.lanes, "by_parties": res}))
        return

    # ---------------- GPU arm ----------------------------------------------------------------------
    dist = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL's communicator lines (init only: "comm ... rank r nranks N") go to stderr
        os.environ["NCCL_DEBUG"] = os.environ.get("MKHE_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    # N > 1: ONE stream of ops, every op limb-sharded over the N GPUs (strong scaling: the total work is that of N = 1); every
    # rank holds the same keys and operand ciphertexts (same seed) and ends every op with the whole result
    team = (world, rank, dist) if world > 1 else None
    SEED = 0xB2000002
    # the end-to-end parity leg needs landing buffer (0, 0) to always hold pair 0: the pool size divides 2 * lanes
    wl = DeviceWorkload(lit, k, local_rank, seed=SEED, batch=B, lanes=args.lanes, team=team, npairs=4 if (2 * args.lanes) % 4 == 0 else 2 * args.lanes)
    clocks = ClockSampler(local_rank)
    clocks.start()
    nv0 = nvlink_bytes(local_rank) if world > 1 else None
    ms = wl.timed(wl.mul_relin_step, args.steps, warmup, barrier)
    launches_timed = wl.last_launches
    nv1 = nvlink_bytes(local_rank) if world > 1 else None
    clk = clocks.stop()
    nvlink = None
    if world > 1:
        # what the op's three peer-store exchanges move: y_i of the nP special limbs of the 4k products, the k
        # re-decomposed p_id (level + 1 limbs each) and the k + 1 result components (level limbs after Rescale); every limb goes from its
        # owner to the other world - 1 ranks.  The NVLink hardware counters are not exposed on this (virtualised) platform.
        limbs = 4 * k * nP + k * ell + (k + 1) * (ell - 1)
        nvlink = {"algorithmic_peer_store_bytes_per_op_all_ranks": limbs * 8 * N * (world - 1),
                  "algorithmic_tx_bytes_per_op_per_rank": limbs * 8 * N * (world - 1) / world, "counters": None}
    if nv0 and nv1:
        # the counters also see the warm-up steps of this call: both reads bracket warm-up + timed steps
        nops = B * (args.steps + warmup)
        nvlink["counters"] = {"rank0_tx_bytes_per_op": (nv1[0] - nv0[0]) / nops, "rank0_rx_bytes_per_op": (nv1[1] - nv0[1]) / nops,
                              "ops_between_reads": nops,
                              "source": "nvidia-smi nvlink -gt d, rank 0's GPU, all links summed, before / after the warm-up + timed steps"}
    gstats = [ln.graph_stats() for ln in wl.lanes]      # the repeated ops of the timed region replay captured CUDA graphs
    cuda_graphs = {"captures": sum(g[0] for g in gstats), "replays": sum(g[1] for g in gstats),
                   "note": "an op issued again with the same operand addresses replays its captured launch sequence; gpu_launches counts "
                           "the kernels inside the replayed graphs"}
    ms = allmax(ms)
    value = B * args.steps / (ms * 1e-3)

    # end to end through host buffers
    wl.prepare_e2e()
    bytes_io = [0, 0]

    def e2e_fn(i):
        bytes_io[0], bytes_io[1] = wl.e2e_step(i)

    ms_e2e = allmax(wl.timed_wall(e2e_fn, args.steps, warmup, barrier))
    e2e_value = B * args.steps / (ms_e2e * 1e-3)

    # outputs of the TIMED workload kept for the parity check below (compared with the oracle on the same seeds): the
    # device-resident result of pair 0 (op 0 runs on lane 0) and the end-to-end path's host landing buffer of pair 0
    # (op n lands in buffer (n % lanes, (n // lanes) % 2); with 4 pairs and 2 lanes buffer (0, 0) always holds pair 0)
    parity_got, sharded_parity = None, None
    e2e_limbs = wl.own_limbs(wl.level + 1 - wl.nb) if team else None
    wl.sync()
    wl.mul_relin_op(0, lane=0)
    wl.sync()
    res_dev = wl.result(0)
    if rank == 0 and not args.no_cpu_baseline:
        parity_got = {"device": res_dev}
        if (2 * len(wl.lanes)) % len(wl.pairs) == 0:
            parity_got["e2e"] = {kk: np.array(v) for kk, v in wl.res_host[0][0].items()}
    if team:
        # the limb-sharded result of every rank against the single-GPU op (mkhe_ckks_mul_relin) of the same rank on the same operands
        wl.mul_relin_op(0, lane=0, sharded=False)
        wl.sync()
        res_one = wl.result(0)
        eq = all(np.array_equal(res_dev[kk], res_one[kk]) for kk in res_one)
        import torch
        t = torch.tensor([1 if eq else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        sharded_parity = {"n_ranks": world, "equal_on_every_rank": bool(int(t.item())), "timed_out": bool(wl.ctx.team_timed_out()),
                          "against": "mkhe_ckks_mul_relin (one GPU) on the same keys and ciphertexts, every rank compares its own whole result",
                          "words_compared_per_rank": int(sum(v.size for v in res_one.values()))}

    # per-kernel profile pass (events around every launch; separate from the timed region above)
    psteps = 8                      # MulRelin ops in the profiled pass
    wl.ctx.profile_begin()
    for i in range(psteps):
        wl.mul_relin_op(i, lane=0)         # one lane only: the per-kernel times are those of kernels running alone
    prof = wl.ctx.profile_end()
    model = algorithmic_model(k, ell, nP, N)
    if world > 1:                   # a rank computes its share of the limb slots: per-rank algorithmic work = 1 / N of the op's
        model["kernel_bytes"] = {kk: v / world for kk, v in model["kernel_bytes"].items()}
        model["kernel_butterflies"] = {kk: v / world for kk, v in model["kernel_butterflies"].items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    bfly_peak = wl.ctx.butterfly_peak()
    total_prof_ms = sum(v[1] for v in prof.values()) or 1.0
    kernels = {}
    merged = {}
    for name, (cnt, tms) in prof.items():          # "k_mac_digits<4>", "k_moddown_Q_" ... -> one entry per kernel
        base = name.split("<")[0].rstrip("_")
        c0, t0 = merged.get(base, (0, 0.0))
        merged[base] = (c0 + cnt, t0 + tms)
    for name, (cnt, tms) in sorted(merged.items(), key=lambda kv: -kv[1][1]):
        e = {"launches_per_step": cnt / psteps, "ms_per_step": tms / psteps, "share": tms / total_prof_ms}
        if name in model["kernel_bytes"]:
            e["hbm_gbs"] = model["kernel_bytes"][name] / (tms / psteps * 1e-3) / 1e9
            e["hbm_frac"] = e["hbm_gbs"] / hbm_peak
        if name in model["kernel_butterflies"]:
            e["butterflies_per_s"] = model["kernel_butterflies"][name] / (tms / psteps * 1e-3)
            e["int_frac"] = e["butterflies_per_s"] / bfly_peak
        kernels[name] = e
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    dom = next(iter(kernels)) if kernels else None
    roofline = None
    if dom:
        d = kernels[dom]
        per_launch_bytes = model["kernel_bytes"].get(dom, 0) / max(d["launches_per_step"], 1e-9)
        per_launch_s = d["ms_per_step"] * 1e-3 / max(d["launches_per_step"], 1e-9)
        ach = per_launch_bytes / per_launch_s / 1e9 if per_launch_s else 0.0
        int_frac, hbm_frac = d.get("int_frac") or 0.0, ach / hbm_peak
        int_bound = int_frac > hbm_frac           # the roofline is the larger of the two fractions
        roofline = {"kernel": dom, "bound": "int" if int_bound else "hbm",
                    "achieved": d.get("butterflies_per_s") if int_bound else ach, "peak": bfly_peak if int_bound else hbm_peak,
                    "unit": "butterflies/s" if int_bound else "GB/s", "frac": max(int_frac, hbm_frac),
                    "hbm": {"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_frac},
                    "traffic": (traffic.get(dom) or {}).get("dram_bytes_per_launch"),
                    "traffic_source": (traffic.get(dom) or {}).get("source"),
                    # the ncu capture is ONE launch (the hoisting of 4 polys); its own algorithmic bytes, for a like-for-like ratio
                    "traffic_algorithmic_bytes": (traffic.get(dom) or {}).get("algorithmic_bytes_of_that_launch"),
                    "binding_resource": "integer issue (fmaheavy pipe + ALU), not HBM: see integer_pipe; the HBM-bound kernels of the step are "
                                        "k_mac_parties / k_mac_intt (hbm_frac under kernels)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes,
                    "avg_launch_ms": per_launch_s * 1e3,
                    "integer_pipe": {"achieved_butterflies_per_s": d.get("butterflies_per_s"), "peak_butterflies_per_s": bfly_peak,
                                     "frac": d.get("int_frac"),
                                     "peak_source": "register-resident Shoup butterfly loop measured in this run (mkhe_bench_butterfly_peak)"},
                    "whole_step": {"compulsory_bytes": model["compulsory_bytes"],
                                   "hbm_frac_of_compulsory": model["compulsory_bytes"] / (ms / args.steps / B * 1e-3) / 1e9 / hbm_peak,
                                   "butterflies": model["butterflies"],
                                   "int_frac": model["butterflies"] / (ms / args.steps / B * 1e-3) / bfly_peak}}

    extra = {}
    if not args.no_extras and world == 1:
        ms_api = wl.timed(wl.api_step, args.steps, warmup)
        extra[f"mulrelin_new_reference_api_k{k}_ops_s"] = B * args.steps / (ms_api * 1e-3)      # allocation of the result included
        wl.api_keep.clear()
        wl.prepare_rotate()
        ms_rot = wl.timed(wl.rotate_step, args.steps, warmup)
        extra[f"rotate_hoisted_k{k}_ops_s"] = B * args.steps / (ms_rot * 1e-3)
        wl.ctx.close()
        del wl
        ko = 4 if k != 4 else 8        # the other party count of BASELINE config 2
        wl8 = DeviceWorkload(lit, ko, local_rank, seed=0xB2000008, batch=B, lanes=args.lanes)
        ms8 = wl8.timed(wl8.mul_relin_step, args.steps, warmup)
        extra[f"mulrelin_k{ko}_ops_s"] = B * args.steps / (ms8 * 1e-3)
        wl8.prepare_rotate()
        ms_rot8 = wl8.timed(wl8.rotate_step, args.steps, warmup)
        extra[f"rotate_hoisted_k{ko}_ops_s"] = B * args.steps / (ms_rot8 * 1e-3)
        wl8.ctx.close()
        del wl8
        # the other BASELINE configs as side measurements: config 1 (PN14QP439, k = 2, the reference's CPU-runnable case) and
        # config 3 (mkbfv MulRelinNew, PN15QP880 primes, k = 4)
        wl14 = DeviceWorkload(PR.CKKS_PN14QP439, 2, local_rank, seed=0xB2000014, batch=B, lanes=args.lanes)
        ms14 = wl14.timed(wl14.mul_relin_step, args.steps, warmup)
        extra["mulrelin_PN14QP439_k2_ops_s"] = B * args.steps / (ms14 * 1e-3)
        wl14.ctx.close()
        del wl14
        extra["bfv_mulrelin_PN15QP880_k4_ops_s"] = bfv_mul_relin_ops(PR.BFV_PN15QP880, 4, local_rank, max(args.steps // 2, 2), warmup, 8, args.lanes)
        # config 5: encrypted CNN inference, PN14QP433 (cnn/cnn_test.go:80-96), two parties, a batch of synthetic images
        cnn_ips, cnn_e2e, cnn_res0, cnn_inp, cnn_launches = cnn_leg(PR.CNN_PN14QP433, local_rank, 8, args.lanes)
        extra["cnn_PN14QP433_images_s"] = cnn_ips
        extra["cnn_PN14QP433_e2e_images_s"] = cnn_e2e
        extra["cnn_launches_per_image"] = cnn_launches
        if not args.no_cpu_baseline:
            want, sec, cores = cnn_cpu_run(PR.CNN_PN14QP433, cnn_inp, 1)
            extra["cnn_cpu_baseline"] = {"value": 1.0 / sec, "unit": "images/s", "cores": cores, "kind": "port",
                                         "sample": f"one image through the oracle port of the same op sequence ({sec:.1f} s), single thread"}
            extra["cnn_parity_check"] = {"equal": bool(set(map(str, want)) == set(map(str, cnn_res0)) and
                                                       all(np.array_equal(cnn_res0[kk], want[kk]) for kk in want)),
                                         "what": "fc2Out of image 0 (level 0, every component) against the oracle on the same inputs"}
        del cnn_inp
        # key generation on the device
        kg_s, kg_bytes = keygen_leg(lit, 32, local_rank)
        extra["keygen_on_device_k32"] = {"seconds": kg_s, "key_bytes": kg_bytes, "gb_per_s": kg_bytes / kg_s / 1e9,
                                         "what": "32 parties x (sk, r, pk, relin b/d/v, 2 rotation keys), PN15QP880, made in HBM: nothing crosses PCIe"}

    if not args.no_extras and world > 1:
        wl.ctx.close()
        del wl
        # other party counts, limb-sharded the same way (BASELINE config 4); k = 32 only on 4+ GPUs (5.4 GiB of keys per rank)
        for ks in [4, 16] + ([32] if world >= 4 else []):
            if ks == k:
                continue
            try:
                w2 = DeviceWorkload(lit, ks, local_rank, seed=0xB2000100 + ks, batch=4, lanes=args.lanes, team=team, npairs=2, rots=())
                st = max(args.steps // 4, 2)
                ms2 = allmax(w2.timed(w2.mul_relin_step, st, warmup, barrier))
                extra[f"limb_sharded_mulrelin_k{ks}_ops_s"] = w2.batch * st / (ms2 * 1e-3)
                w2.ctx.close()
                del w2
            except Exception as e:
                extra[f"limb_sharded_mulrelin_k{ks}_error"] = str(e)[:200]
        # for comparison: independent replicas, one stream of ops per GPU, no exchange (weak scaling; round 1's multi-GPU headline)
        try:
            w3 = DeviceWorkload(lit, 4, local_rank, seed=0xB2000300 + rank, batch=B, lanes=args.lanes, npairs=2, rots=())
            ms3 = allmax(w3.timed(w3.mul_relin_step, max(args.steps // 2, 2), warmup, barrier))
            extra["replicas_weak_mulrelin_k4_ops_s"] = world * B * max(args.steps // 2, 2) / (ms3 * 1e-3)
            w3.ctx.close()
            del w3
        except Exception as e:
            extra["replicas_weak_error"] = str(e)[:200]
        # ... and round 1's party sharding (fused peer-memory exchange of the partial x || y, NCCL for barriers) at the headline k
        if k >= world:
            try:
                extra[f"party_sharded_mulrelin_k{k}_ops_s"] = sharded_mul_relin(lit, k, rank, world, local_rank, max(args.steps // 4, 2), warmup,
                                                                               4, dist, barrier, allmax, nlanes=args.lanes)
            except Exception as e:
                extra["party_sharded_error"] = str(e)[:200]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ops, cores, sec = cpu_oracle_run(lit, k, 3, 1, 1, seed=SEED)
        cpu = {"value": ops, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 full MulRelinNew calls after 1 warm-up, same workload ({sec:.2f} s each), single thread like the Go reference; "
                         "oracle/ C restatement (the Go reference cannot be built here)"}

    parity_check = None
    if parity_got is not None:
        # the checker (oracle/) on the seeds of the timed workload; every limb of every component must be identical
        want = oracle_result(lit, k, seed=SEED, pair=0)
        parity_check = {"config": config["workload"], "oracle": "oracle/ C restatement (checker only)", "compared": {}}
        for leg, got in parity_got.items():
            eq = set(map(str, got)) == set(map(str, want))
            for kk in want:
                if leg == "e2e" and team:        # the rank read back the limbs it owns
                    eq = eq and all(np.array_equal(got[kk][j], want[kk][j]) for j in e2e_limbs)
                else:
                    eq = eq and np.array_equal(got[kk], want[kk])
            parity_check["compared"][leg] = bool(eq)
        parity_check["equal"] = all(parity_check["compared"].values())
        parity_check["words_compared"] = int(sum(v.size for v in want.values()) * len(parity_got))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "ms_per_op": ms / args.steps / B, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config,
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io[0], "d2h_bytes_per_step": bytes_io[1],
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches_timed, "cuda_graphs": cuda_graphs, "nvlink": nvlink,
                "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity_check, "sharded_parity": sharded_parity,
                "kernels": kernels, "extra": extra}
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

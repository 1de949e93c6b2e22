"""ctypes binding of libmkhe_b200.so (include/mkhe.h).

The product path has exactly one backend: the CUDA library built by `__graft_entry__.build()`.
If it is missing, or no CUDA device is present, creating a Context raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmkhe_b200.so")

u64p = C.POINTER(C.c_uint64)
intp = C.POINTER(C.c_int)

OK = 0
ERR_NAMES = {-1: "MKHE_ERR_INVALID", -2: "MKHE_ERR_CUDA", -3: "MKHE_ERR_NOMEM", -4: "MKHE_ERR_UNSUPPORTED", -5: "MKHE_ERR_NCCL"}


class MkheError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg


class Library:
    """One loaded shared object implementing include/mkhe.h."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "mkhe_kklss_b200 has no CPU fallback.")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.mkhe_version.restype = C.c_char_p
        d.mkhe_last_error.restype = C.c_char_p
        d.mkhe_last_error.argtypes = [C.c_void_p]
        d.mkhe_launch_count.restype = C.c_uint64
        d.mkhe_launch_count.argtypes = [C.c_void_p]
        d.mkhe_ctx_destroy.restype = None
        d.mkhe_ctx_destroy.argtypes = [C.c_void_p]

    def version(self) -> str:
        return self.dll.mkhe_version().decode()

    def device_count(self) -> int:
        return int(self.dll.mkhe_device_count())


_default = None


def default_library() -> Library:
    global _default
    if _default is None:
        _default = Library(LIB_PATH)
    return _default


def _u64arr(xs):
    return (C.c_uint64 * len(xs))(*[int(x) for x in xs])


def _intarr(xs):
    return (C.c_int * max(len(xs), 1))(*[int(x) for x in xs])


def _harr(hs):
    return (C.c_uint64 * max(len(hs), 1))(*[int(h) for h in hs])


class Context:
    """mkhe_ctx: moduli, tables, stream and scratch pools on one device."""

    def __init__(self, logN, Q, P, gamma=2, device=0, QMul=None, T=0, lib: Library = None):
        self.lib = lib or default_library()
        self.dll = self.lib.dll
        self.logN, self.N = logN, 1 << logN
        self.Q, self.P = [int(x) for x in Q], [int(x) for x in P]
        self.nQ, self.nP = len(self.Q), len(self.P)
        self.D = self.nQ + self.nP
        self.alpha = max(self.nP // gamma, 1)                       # mkrlwe/params.go:63-65
        self.beta_max = -(-self.nQ // self.alpha)
        self.QMul = [int(x) for x in (QMul or [])]
        self.T = int(T)
        ptr = C.c_void_p()
        rc = self.dll.mkhe_ctx_create(C.c_int(logN), _u64arr(self.Q), C.c_int(self.nQ), _u64arr(self.P), C.c_int(self.nP),
                                      C.c_int(gamma), C.c_int(device), C.byref(ptr))
        if rc != OK:
            raise MkheError(rc, "mkhe_ctx_create failed (is a CUDA device present? this package has no CPU fallback)")
        self.ptr = ptr
        if self.QMul:
            self.check(self.dll.mkhe_ctx_set_bfv(self.ptr, _u64arr(self.QMul), C.c_int(len(self.QMul)), C.c_uint64(self.T)))

    # -- plumbing -----------------------------------------------------------------------------
    def fork(self) -> "Context":
        """a lane of this context (mkhe_ctx_fork): same handles, own stream and scratch pools"""
        import copy
        f = copy.copy(self)
        ptr = C.c_void_p()
        self.check(self.dll.mkhe_ctx_fork(self.ptr, C.byref(ptr)))
        f.ptr = ptr
        f.parent = self
        f._forks = []
        self.__dict__.setdefault("_forks", []).append(f)
        return f

    def wait(self, other: "Context"):
        """this lane's stream waits for everything enqueued so far on `other`"""
        self.check(self.dll.mkhe_ctx_wait(self.ptr, other.ptr))

    def check(self, rc):
        if rc != OK:
            raise MkheError(rc, self.dll.mkhe_last_error(self.ptr).decode())

    def close(self):
        if getattr(self, "ptr", None):
            for f in self.__dict__.get("_forks", []):      # mkhe_ctx_destroy(root) destroys the forks
                f.ptr = None
            self.dll.mkhe_ctx_destroy(self.ptr)
            self.ptr = None
            parent = self.__dict__.get("parent")
            if parent is not None and self in parent.__dict__.get("_forks", []):
                parent._forks.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.dll.mkhe_sync(self.ptr))

    def launch_count(self) -> int:
        return int(self.dll.mkhe_launch_count(self.ptr))

    def set_graphs(self, on: bool):
        """opt into the CUDA-graph cache of repeated ops on this lane (mkhe_ctx_set_graphs)"""
        self.check(self.dll.mkhe_ctx_set_graphs(self.ptr, C.c_int(1 if on else 0)))

    def graph_stats(self):
        """(graphs captured, graph replays) of this lane"""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.check(self.dll.mkhe_graph_stats(self.ptr, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_ntt_tables(self, m, psi, psiinv, ninv):
        psi = np.ascontiguousarray(psi, dtype=np.uint64)
        psiinv = np.ascontiguousarray(psiinv, dtype=np.uint64)
        self.check(self.dll.mkhe_ctx_set_ntt_tables(self.ptr, C.c_int(m), psi.ctypes.data_as(u64p), psiinv.ctypes.data_as(u64p), C.c_uint64(ninv)))

    def timer_start(self):
        self.check(self.dll.mkhe_timer_start(self.ptr))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(self.dll.mkhe_timer_stop(self.ptr, C.byref(ms)))
        return float(ms.value)

    def profile_begin(self):
        self.check(self.dll.mkhe_profile_begin(self.ptr))

    def profile_end(self) -> dict:
        """{kernel: (launches, total_ms)} of the launches since profile_begin()"""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.dll.mkhe_profile_end(self.ptr, buf, C.c_size_t(len(buf))))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def butterfly_peak(self) -> float:
        v = C.c_double()
        self.check(self.dll.mkhe_bench_butterfly_peak(self.ptr, C.byref(v)))
        return float(v.value)

    # -- polys ----------------------------------------------------------------------------------
    def poly_alloc(self, nlimbs) -> int:
        h = C.c_uint64()
        self.check(self.dll.mkhe_poly_alloc(self.ptr, C.c_int(nlimbs), C.byref(h)))
        return int(h.value)

    def poly_free(self, h):
        self.check(self.dll.mkhe_poly_free(self.ptr, C.c_uint64(h)))

    def poly_set_nlimbs(self, h, n):
        self.check(self.dll.mkhe_poly_set_nlimbs(self.ptr, C.c_uint64(h), C.c_int(n)))

    def poly_get_nlimbs(self, h) -> int:
        n = C.c_int()
        self.check(self.dll.mkhe_poly_get_nlimbs(self.ptr, C.c_uint64(h), C.byref(n)))
        return int(n.value)

    def poly_upload(self, h, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        assert arr.ndim == 2 and arr.shape[1] == self.N
        self.check(self.dll.mkhe_poly_upload(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p), C.c_int(arr.shape[0])))

    def poly_download(self, h, nlimbs=None) -> np.ndarray:
        nlimbs = self.poly_get_nlimbs(h) if nlimbs is None else nlimbs
        out = np.empty((nlimbs, self.N), dtype=np.uint64)
        self.check(self.dll.mkhe_poly_download(self.ptr, C.c_uint64(h), out.ctypes.data_as(u64p), C.c_int(nlimbs)))
        return out

    def poly_upload_limb(self, h, limb, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        self.check(self.dll.mkhe_poly_upload_limb(self.ptr, C.c_uint64(h), C.c_int(limb), arr.ctypes.data_as(u64p)))

    def poly_download_limb(self, h, limb) -> np.ndarray:
        out = np.empty(self.N, dtype=np.uint64)
        self.check(self.dll.mkhe_poly_download_limb(self.ptr, C.c_uint64(h), C.c_int(limb), out.ctypes.data_as(u64p)))
        return out

    # -- asynchronous transfers on library-owned pinned memory -----------------------------------
    def host_alloc(self, shape) -> np.ndarray:
        """uint64 array of `shape` in page-locked host memory (mkhe_host_alloc); free with host_free(arr)"""
        n = int(np.prod(shape))
        p = C.c_void_p()
        self.check(self.dll.mkhe_host_alloc(self.ptr, C.c_size_t(n * 8), C.byref(p)))
        buf = (C.c_uint64 * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint64).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr):
        addr = self._pinned.pop(arr.ctypes.data)
        self.check(self.dll.mkhe_host_free(self.ptr, C.c_void_p(addr)))

    def poly_upload_async(self, h, arr: np.ndarray):
        """arr: a (view of a) host_alloc array; must stay untouched until sync()"""
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.ndim == 2 and arr.shape[1] == self.N
        self.check(self.dll.mkhe_poly_upload_async(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p), C.c_int(arr.shape[0])))

    def poly_download_async(self, h, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.ndim == 2 and arr.shape[1] == self.N
        self.check(self.dll.mkhe_poly_download_async(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p), C.c_int(arr.shape[0])))

    def poly_copy(self, dst, src):
        self.check(self.dll.mkhe_poly_copy(self.ptr, C.c_uint64(dst), C.c_uint64(src)))

    def poly_copy_lvl(self, level, dst, src):
        self.check(self.dll.mkhe_poly_copy_lvl(self.ptr, C.c_int(level), C.c_uint64(dst), C.c_uint64(src)))

    # -- switching keys ---------------------------------------------------------------------------
    def swk_alloc(self) -> int:
        h = C.c_uint64()
        self.check(self.dll.mkhe_swk_alloc(self.ptr, C.byref(h)))
        return int(h.value)

    def swk_free(self, h):
        self.check(self.dll.mkhe_swk_free(self.ptr, C.c_uint64(h)))

    def swk_upload(self, h, arr: np.ndarray):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        assert arr.shape == (self.beta_max, self.D, self.N), arr.shape
        self.check(self.dll.mkhe_swk_upload(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p)))

    def swk_download(self, h) -> np.ndarray:
        out = np.empty((self.beta_max, self.D, self.N), dtype=np.uint64)
        self.check(self.dll.mkhe_swk_download(self.ptr, C.c_uint64(h), out.ctypes.data_as(u64p)))
        return out

    def swk_upload_limb(self, h, digit, is_p, limb, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        self.check(self.dll.mkhe_swk_upload_limb(self.ptr, C.c_uint64(h), C.c_int(digit), C.c_int(is_p), C.c_int(limb), arr.ctypes.data_as(u64p)))

    def swk_download_limb(self, h, digit, is_p, limb):
        out = np.empty(self.N, dtype=np.uint64)
        self.check(self.dll.mkhe_swk_download_limb(self.ptr, C.c_uint64(h), C.c_int(digit), C.c_int(is_p), C.c_int(limb), out.ctypes.data_as(u64p)))
        return out

    # -- raw ops (thin; the reference-shaped API lives in mkrlwe.py / mkckks.py / mkbfv.py) ----------
    def ntt(self, level, hin, hout):
        self.check(self.dll.mkhe_ntt(self.ptr, C.c_int(level), C.c_uint64(hin), C.c_uint64(hout)))

    def intt(self, level, hin, hout):
        self.check(self.dll.mkhe_intt(self.ptr, C.c_int(level), C.c_uint64(hin), C.c_uint64(hout)))

    def decompose(self, level, hpoly, hswk):
        self.check(self.dll.mkhe_decompose(self.ptr, C.c_int(level), C.c_uint64(hpoly), C.c_uint64(hswk)))

    def moddown_qp_to_q_ntt(self, level, p1, p2q):
        self.check(self.dll.mkhe_moddown_qp_to_q_ntt(self.ptr, C.c_int(level), C.c_uint64(p1), C.c_uint64(p2q)))

    def external_product(self, level, ha, hbg, hc):
        self.check(self.dll.mkhe_external_product(self.ptr, C.c_int(level), C.c_uint64(ha), C.c_uint64(hbg), C.c_uint64(hc)))

    def external_product_hoisted(self, level, hah, hbg, hc):
        self.check(self.dll.mkhe_external_product_hoisted(self.ptr, C.c_int(level), C.c_uint64(hah), C.c_uint64(hbg), C.c_uint64(hc)))

    def mul_relin_hoisted(self, level, ids0, op0, h0, ids1, op1, h1, rlk_b, rlk_d, rlk_v, u, idsOut, out):
        self.check(self.dll.mkhe_mul_relin_hoisted(
            self.ptr, C.c_int(level),
            C.c_int(len(ids0)), _intarr(ids0), _harr(op0), None if h0 is None else _harr(h0),
            C.c_int(len(ids1)), _intarr(ids1), _harr(op1), None if h1 is None else _harr(h1),
            _harr(rlk_b), _harr(rlk_d), _harr(rlk_v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

    def ckks_mul_relin(self, level, nb_rescales, same, ids0, op0, ids1, op1, rlk_b, rlk_d, rlk_v, u, idsOut, out):
        self.check(self.dll.mkhe_ckks_mul_relin(
            self.ptr, C.c_int(level), C.c_int(nb_rescales), C.c_int(1 if same else 0),
            C.c_int(len(ids0)), _intarr(ids0), _harr(op0),
            C.c_int(len(ids1)), _intarr(ids1), _harr(op1),
            _harr(rlk_b), _harr(rlk_d), _harr(rlk_v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

    def rotate_hoisted(self, level, rot, ct_in, hoisted, rk, a, ct_out):
        self.check(self.dll.mkhe_rotate_hoisted(self.ptr, C.c_int(level), C.c_int(rot), C.c_int(len(hoisted)),
                                                _harr(ct_in), _harr(hoisted), _harr(rk), C.c_uint64(a), _harr(ct_out)))

    def rotate(self, level, rot, ct_in, rk, a, ct_out):
        self.check(self.dll.mkhe_rotate(self.ptr, C.c_int(level), C.c_int(rot), C.c_int(len(rk)),
                                        _harr(ct_in), _harr(rk), C.c_uint64(a), _harr(ct_out)))

    def conjugate(self, level, ct_in, ck, a, ct_out):
        self.check(self.dll.mkhe_conjugate(self.ptr, C.c_int(level), C.c_int(len(ck)), _harr(ct_in), _harr(ck), C.c_uint64(a), _harr(ct_out)))

    def rescale(self, level, nb, hin, hout):
        self.check(self.dll.mkhe_rescale(self.ptr, C.c_int(level), C.c_int(nb), C.c_uint64(hin), C.c_uint64(hout)))

    def poly_add(self, level, a, b, out):
        self.check(self.dll.mkhe_poly_add(self.ptr, C.c_int(level), C.c_uint64(a), C.c_uint64(b), C.c_uint64(out)))

    def poly_sub(self, level, a, b, out):
        self.check(self.dll.mkhe_poly_sub(self.ptr, C.c_int(level), C.c_uint64(a), C.c_uint64(b), C.c_uint64(out)))

    def poly_neg(self, level, a, out):
        self.check(self.dll.mkhe_poly_neg(self.ptr, C.c_int(level), C.c_uint64(a), C.c_uint64(out)))

    def ckks_mult_by_const(self, level, hin, hout, c_real, c_imag, scale):
        self.check(self.dll.mkhe_ckks_mult_by_const(self.ptr, C.c_int(level), C.c_int(len(hin)), _harr(hin), _harr(hout),
                                                    C.c_double(c_real), C.c_double(c_imag), C.c_double(scale)))

    def ckks_mul_ptxt(self, level, pt, hin, hout):
        self.check(self.dll.mkhe_ckks_mul_ptxt(self.ptr, C.c_int(level), C.c_uint64(pt), C.c_int(len(hin)), _harr(hin), _harr(hout)))

    def decrypt(self, level, ct, sk, pt):
        self.check(self.dll.mkhe_decrypt(self.ptr, C.c_int(level), C.c_int(len(sk)), _harr(ct), _harr(sk), C.c_uint64(pt)))

    # key generation and encryption on the device ("mkhe-ctr-1" streams, include/mkhe_prng.h)
    def sample_crs(self, seed, stream, hswk):
        self.check(self.dll.mkhe_sample_crs(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hswk)))

    def keygen_secret(self, seed, stream, pzero, hsk):
        self.check(self.dll.mkhe_keygen_secret(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_double(pzero), C.c_uint64(hsk)))

    def keygen_public(self, seed, stream, hsk, crs_a, pk0, pk1):
        self.check(self.dll.mkhe_keygen_public(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hsk), C.c_uint64(crs_a),
                                               C.c_uint64(pk0), C.c_uint64(pk1)))

    def keygen_switching_key(self, seed, stream, hsk, out):
        self.check(self.dll.mkhe_keygen_switching_key(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hsk), C.c_uint64(out)))

    def keygen_relin(self, seed, stream, hsk, hr, crs_a, crs_u, b, d, v):
        self.check(self.dll.mkhe_keygen_relin(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hsk), C.c_uint64(hr),
                                              C.c_uint64(crs_a), C.c_uint64(crs_u), C.c_uint64(b), C.c_uint64(d), C.c_uint64(v)))

    def keygen_rotation(self, seed, stream, rotidx, hsk, crs_rot, rk):
        self.check(self.dll.mkhe_keygen_rotation(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_int(rotidx), C.c_uint64(hsk),
                                                 C.c_uint64(crs_rot), C.c_uint64(rk)))

    def keygen_conjugation(self, seed, stream, hsk, crs_cj, ck):
        self.check(self.dll.mkhe_keygen_conjugation(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hsk), C.c_uint64(crs_cj),
                                                    C.c_uint64(ck)))

    def keygen_bfv_relin(self, seed, stream, hsk, hr, crs_a1, crs_a2, crs_u, b1, b2, d1, d2, v):
        self.check(self.dll.mkhe_keygen_bfv_relin(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(hsk), C.c_uint64(hr),
                                                  C.c_uint64(crs_a1), C.c_uint64(crs_a2), C.c_uint64(crs_u), C.c_uint64(b1),
                                                  C.c_uint64(b2), C.c_uint64(d1), C.c_uint64(d2), C.c_uint64(v)))

    def encrypt(self, seed, stream, level, pt, pk0, pk1, c0, c1):
        self.check(self.dll.mkhe_encrypt(self.ptr, C.c_uint64(seed), C.c_uint64(stream), C.c_int(level), C.c_uint64(pt or 0),
                                         C.c_uint64(pk0), C.c_uint64(pk1), C.c_uint64(c0), C.c_uint64(c1)))

    # multi-GPU (party sharding, NCCL)
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.dll.mkhe_comm_unique_id(buf)
        if rc != OK:
            raise MkheError(rc, "mkhe_comm_unique_id failed (library built without NCCL?)")
        return bytes(buf)

    def comm_init(self, nranks, rank, uid: bytes):
        buf = (C.c_uint8 * 128)(*uid)
        self.check(self.dll.mkhe_comm_init(self.ptr, C.c_int(nranks), C.c_int(rank), buf))

    def p2p_export(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self.check(self.dll.mkhe_p2p_export(self.ptr, buf))
        return bytes(buf)

    def p2p_import(self, nranks, rank, handles: list):
        """handles[r] = rank r's p2p_export() bytes"""
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.check(self.dll.mkhe_p2p_import(self.ptr, C.c_int(nranks), C.c_int(rank), arr))

    def ckks_mul_relin_sharded(self, level, nb_rescales, ids0, op0, ids1, op1, own_ids, rlk_b, rlk_d, rlk_v, u, idsOut, out):
        self.check(self.dll.mkhe_ckks_mul_relin_sharded(
            self.ptr, C.c_int(level), C.c_int(nb_rescales),
            C.c_int(len(ids0)), _intarr(ids0), _harr(op0),
            C.c_int(len(ids1)), _intarr(ids1), _harr(op1),
            C.c_int(len(own_ids)), _intarr(own_ids),
            _harr(rlk_b), _harr(rlk_d), _harr(rlk_v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

    # limb-sharded ops: the team of ranks
    def team_export(self, max_parties) -> bytes:
        buf = (C.c_uint8 * 64)()
        self.check(self.dll.mkhe_team_export(self.ptr, C.c_int(max_parties), buf))
        return bytes(buf)

    def team_import(self, nranks, rank, handles: list):
        """handles[r] = rank r's team_export() bytes"""
        assert len(handles) == nranks and all(len(h) == 64 for h in handles)
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.check(self.dll.mkhe_team_import(self.ptr, C.c_int(nranks), C.c_int(rank), arr))

    def team_join_local(self, max_parties, rank, members: list):
        """ranks inside one process: members[r] = the Context of rank r (members[rank] is self)"""
        arr = (C.c_void_p * len(members))(*[m.ptr for m in members])
        self.check(self.dll.mkhe_team_join_local(self.ptr, C.c_int(max_parties), C.c_int(len(members)), C.c_int(rank), arr))

    def team_timed_out(self) -> bool:
        v = C.c_int(0)
        self.check(self.dll.mkhe_team_status(self.ptr, C.byref(v)))
        return bool(v.value)

    def team_allgather(self, level, polys):
        self.check(self.dll.mkhe_team_allgather(self.ptr, C.c_int(level), C.c_int(len(polys)), _harr(polys)))

    def team_owns_limb(self, limb) -> bool:
        return bool(self.dll.mkhe_team_owns_limb(self.ptr, C.c_int(limb)))

    def poly_upload_limb_async(self, h, limb, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.size == self.N
        self.check(self.dll.mkhe_poly_upload_limb_async(self.ptr, C.c_uint64(h), C.c_int(limb), arr.ctypes.data_as(u64p)))

    def poly_download_limb_async(self, h, limb, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.size == self.N
        self.check(self.dll.mkhe_poly_download_limb_async(self.ptr, C.c_uint64(h), C.c_int(limb), arr.ctypes.data_as(u64p)))

    def poly_upload_owned_async(self, h, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.ndim == 2 and arr.shape[1] == self.N
        self.check(self.dll.mkhe_poly_upload_owned_async(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p), C.c_int(arr.shape[0])))

    def poly_download_owned_async(self, h, arr: np.ndarray):
        assert arr.dtype == np.uint64 and arr.flags.c_contiguous and arr.ndim == 2 and arr.shape[1] == self.N
        self.check(self.dll.mkhe_poly_download_owned_async(self.ptr, C.c_uint64(h), arr.ctypes.data_as(u64p), C.c_int(arr.shape[0])))

    def team_flags(self):
        buf = (C.c_uint64 * 17)()
        self.check(self.dll.mkhe_team_flags(self.ptr, buf))
        return list(buf)

    def ckks_mul_relin_limbs(self, level, nb_rescales, ids0, op0, ids1, op1, rlk_b, rlk_d, rlk_v, u, idsOut, out):
        self.check(self.dll.mkhe_ckks_mul_relin_limbs(
            self.ptr, C.c_int(level), C.c_int(nb_rescales),
            C.c_int(len(ids0)), _intarr(ids0), _harr(op0),
            C.c_int(len(ids1)), _intarr(ids1), _harr(op1),
            _harr(rlk_b), _harr(rlk_d), _harr(rlk_v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

    # BFV
    def bfv_modup_q_to_r(self, hq, hr):
        self.check(self.dll.mkhe_bfv_modup_q_to_r(self.ptr, C.c_uint64(hq), C.c_uint64(hr)))

    def bfv_rescale_q_to_r(self, hq, hr):
        self.check(self.dll.mkhe_bfv_rescale_q_to_r(self.ptr, C.c_uint64(hq), C.c_uint64(hr)))

    def bfv_quantize(self, hr, hq):
        self.check(self.dll.mkhe_bfv_quantize(self.ptr, C.c_uint64(hr), C.c_uint64(hq)))

    def bfv_decompose(self, level, hr, h1, h2):
        self.check(self.dll.mkhe_bfv_decompose(self.ptr, C.c_int(level), C.c_uint64(hr), C.c_uint64(h1), C.c_uint64(h2)))

    def bfv_mul_relin_hoisted(self, level, ids0, op0, h0a, h0b, ids1, op1, h1a, h1b, b1, b2, d1, d2, v, u, idsOut, out):
        self.check(self.dll.mkhe_bfv_mul_relin_hoisted(
            self.ptr, C.c_int(level),
            C.c_int(len(ids0)), _intarr(ids0), _harr(op0), None if h0a is None else _harr(h0a), None if h0b is None else _harr(h0b),
            C.c_int(len(ids1)), _intarr(ids1), _harr(op1), None if h1a is None else _harr(h1a), None if h1b is None else _harr(h1b),
            _harr(b1), _harr(b2), _harr(d1), _harr(d2), _harr(v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

    def bfv_mul_relin(self, ids0, ct0, ids1, ct1, b1, b2, d1, d2, v, u, idsOut, out):
        self.check(self.dll.mkhe_bfv_mul_relin(
            self.ptr,
            C.c_int(len(ids0)), _intarr(ids0), _harr(ct0),
            C.c_int(len(ids1)), _intarr(ids1), _harr(ct1),
            _harr(b1), _harr(b2), _harr(d1), _harr(d2), _harr(v), C.c_uint64(u),
            C.c_int(len(idsOut)), _intarr(idsOut), _harr(out)))

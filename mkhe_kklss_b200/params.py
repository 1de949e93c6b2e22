"""Parameter literals of the reference's tests / benchmarks (the reference keeps them as Go literals in
its test files).  gamma is hard-wired to 2 (mkckks/params.go:20, mkbfv/params.go:77-78),
so alpha = #P/gamma = 1 and beta(level) = level+1 for every set below.

Every prime is = 1 (mod 2^16), so each set can also be instantiated at a smaller logN (used by the
parity tests to keep the CPU oracle fast).
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace


@dataclass(frozen=True)
class ParamLiteral:
    name: str
    logN: int
    Q: tuple
    P: tuple
    scale: float = 0.0          # CKKS default scale
    QMul: tuple = ()            # BFV only
    T: int = 0                  # BFV only
    gamma: int = 2
    sigma: float = 3.2

    def at_logn(self, logN: int) -> "ParamLiteral":
        return replace(self, logN=logN, name=f"{self.name}@logN{logN}")

    @property
    def N(self):
        return 1 << self.logN


# mkckks/mkckks_test.go:51-72  (60 + 13x54 | 2x59)
CKKS_PN15QP880 = ParamLiteral(
    "PN15QP880", 15,
    Q=(0xfffffffff6a0001,
       0x3fffffffd60001, 0x3fffffffca0001, 0x3fffffff6d0001, 0x3fffffff5d0001,
       0x3fffffff550001, 0x3fffffff390001, 0x3fffffff360001, 0x3fffffff2a0001,
       0x3fffffff000001, 0x3ffffffefa0001, 0x3ffffffef40001, 0x3ffffffed70001,
       0x3ffffffed30001),
    P=(0x7ffffffffe70001, 0x7ffffffffe10001),
    scale=float(1 << 54))

# mkckks/mkckks_test.go:73-90  (59 + 5x52 | 2x60)
CKKS_PN14QP439 = ParamLiteral(
    "PN14QP439", 14,
    Q=(0x7ffffffffe70001,
       0xffffffff00001, 0xfffffffe40001, 0xfffffffe20001, 0xfffffffbe0001, 0xfffffffa60001),
    P=(0xffffffffffc0001, 0xfffffffff840001),
    scale=float(1 << 52))

# cnn/cnn_test.go:80-96  (58 + 6x48 | 2x48)
CNN_PN14QP433 = ParamLiteral(
    "PN14QP433", 14,
    Q=(0x2000000002b0001,
       0x800000020001, 0x800000280001, 0x800000520001, 0x800000770001, 0x800000aa0001, 0x800000ad0001),
    P=(0x800000df0001, 0x800000f80001),
    scale=float(1 << 47))

# mkbfv/mkbfv_test.go:28-75
BFV_PN15QP880 = ParamLiteral(
    "BFV_PN15QP880", 15,
    Q=(0x3fffffffd60001, 0x3fffffff6d0001, 0x3fffffff550001, 0x3fffffff360001, 0x3fffffff000001,
       0x3ffffffef40001, 0x3ffffffed30001, 0x3ffffffe970001, 0x3ffffffe800001, 0x3ffffffe410001,
       0x7fffffffe90001, 0x7fffffffbd0001, 0x7fffffffaa0001, 0x7fffffff9f0001),
    QMul=(0x3fffffffca0001, 0x3fffffff5d0001, 0x3fffffff390001, 0x3fffffff2a0001, 0x3ffffffefa0001,
          0x3ffffffed70001, 0x3ffffffeaa0001, 0x3ffffffe920001, 0x3ffffffe790001, 0x3ffffffe320001,
          0x7fffffffbf0001, 0x7fffffffba0001, 0x7fffffffa50001, 0x7fffffff7e0001),
    P=(0xffffffffffc0001, 0xfffffffff840001),
    T=65537)

# mkbfv/mkbfv_test.go:77-108
BFV_PN14QP439 = ParamLiteral(
    "BFV_PN14QP439", 14,
    Q=(0x1fffffffe30001, 0x1fffffffd10001, 0x1fffffffbf0001, 0x1fffffffb60001, 0x1fffffff920001,
       0x3fffffffd60001),
    QMul=(0x1fffffffd80001, 0x1fffffffc50001, 0x1fffffffb90001, 0x1fffffffa50001, 0x1fffffff900001,
          0x3fffffffca0001),
    P=(0xffffffffffc0001, 0xfffffffff840001),
    T=65537)

ALL = {p.name: p for p in (CKKS_PN15QP880, CKKS_PN14QP439, CNN_PN14QP433, BFV_PN15QP880, BFV_PN14QP439)}

"""Parameter literals of the reference's tests / benchmarks (the reference keeps them as Go literals in
its test files).  gamma is hard-wired to 2 (mkckks/params.go:20, mkbfv/params.go:77-78),
so alpha = #P/gamma = 1 and beta(level) = level+1 for every set the reference enables.  PN16QP1761 (four special
primes, alpha = 2) is the set the reference keeps commented out in mkrlwe_test.go:22-35; it exercises the general
DecomposeAndSplit branch.

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

# mkrlwe/mkrlwe_test.go:22-35 (commented out there): 55 + 33x45 | 4x55, gamma = 2 -> alpha = 2
PN16QP1761 = ParamLiteral(
    "PN16QP1761", 16,
    Q=(0x80000000080001, 0x2000000a0001, 0x2000000e0001, 0x1fffffc20001,
       0x200000440001, 0x200000500001, 0x200000620001, 0x1fffff980001,
       0x2000006a0001, 0x1fffff7e0001, 0x200000860001, 0x200000a60001,
       0x200000aa0001, 0x200000b20001, 0x200000c80001, 0x1fffff360001,
       0x200000e20001, 0x1fffff060001, 0x200000fe0001, 0x1ffffede0001,
       0x1ffffeca0001, 0x1ffffeb40001, 0x200001520001, 0x1ffffe760001,
       0x2000019a0001, 0x1ffffe640001, 0x200001a00001, 0x1ffffe520001,
       0x200001e80001, 0x1ffffe0c0001, 0x1ffffdee0001, 0x200002480001,
       0x1ffffdb60001, 0x200002560001),
    P=(0x80000000440001, 0x7fffffffba0001, 0x80000000500001, 0x7fffffffaa0001),
    scale=float(1 << 45))
# its first 7 Q limbs (odd count: the last digit is a single limb at the top level, partial digits below): test size
PN16QP1761_Q7 = replace(PN16QP1761, name="PN16QP1761[:7]", Q=PN16QP1761.Q[:7])

# same limbs with gamma = 1: alpha = 4 -> digits of 4 + 3 limbs at the top level, 4 + 2 and 4 + 1 below (partial lifts)
PN16QP1761_Q7_ALPHA4 = replace(PN16QP1761_Q7, name="PN16QP1761[:7],gamma=1", gamma=1)

ALL = {p.name: p for p in (CKKS_PN15QP880, CKKS_PN14QP439, CNN_PN14QP433, BFV_PN15QP880, BFV_PN14QP439, PN16QP1761)}

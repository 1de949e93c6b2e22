"""generates the cumulative table of include/mkhe_prng.h: |X| for X = round(N(0, sigma^2)) truncated to |X| <= bound,
sigma = 3.2 (rlwe.DefaultSigma), bound = int(6 sigma) = 19 (mkrlwe/keygen.go:35), scaled to 2^64.
    python tools/gen_cdt.py          prints the initialiser; the committed constants ARE the specification."""
from fractions import Fraction
import math

SIGMA, BOUND = 3.2, 19


def phi(x):
    return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


w = [phi((m + 0.5) / SIGMA) - phi((m - 0.5) / SIGMA) for m in range(BOUND + 1)]
mag = [Fraction(w[0])] + [2 * Fraction(x) for x in w[1:]]
tot = sum(mag)
acc, out = Fraction(0), []
for m in range(BOUND + 1):
    acc += mag[m] / tot
    out.append(min(int(acc * (1 << 64)), (1 << 64) - 1))
out[-1] = (1 << 64) - 1
print("{ " + ", ".join(f"0x{v:016x}ull" for v in out) + " }")

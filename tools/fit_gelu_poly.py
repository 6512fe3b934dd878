"""Fit of the GELU epilogue polynomial (layoutdit_b200/csrc/gemm.cuh, GeluCoef).

gelu(x) = x Phi(x) = max(x, 0) - |x| h(z),  z = |x| / sqrt 2,  h(z) = 0.5 erfc(z) = exp2(P(z)).
P = degree-7 least-squares fit of log2(0.5 erfc(z)) on Chebyshev nodes of [0, 5]; prints the
monomial coefficients (low to high) and the error of the fp32 Horner evaluation against the
exact erf-GELU in double precision.  Needs scipy (CPU only; not used at run time).
"""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

ZMAX, DEG = 5.0, 7
z = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * ZMAX / 2 + ZMAX / 2
fit = C.Chebyshev.fit(z, np.log2(erfc(z)) - 1, DEG, domain=[0, ZMAX])
mono = fit.convert(kind=np.polynomial.Polynomial).coef
print("coefficients (c0..c7):", [float(np.float32(c)) for c in mono])

c32 = mono.astype(np.float32)
x = np.linspace(-12, 12, 2400001).astype(np.float32)
zz = (np.abs(x) * np.float32(0.70710678118654752)).astype(np.float32)
p = np.full_like(zz, c32[DEG])
for k in range(DEG - 1, -1, -1):
    p = (p * zz + c32[k]).astype(np.float32)
h = np.exp2(p.astype(np.float64)).astype(np.float32)
out = (np.maximum(x, 0) - np.abs(x) * h).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
err = np.abs(out - ref)
m = np.abs(ref) > 1e-6
ulp = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 1e-38))) - 8)
print(f"max abs err {err.max():.3e}; max rel err (|gelu| > 1e-6) {(err / np.abs(ref))[m].max():.3e}; "
      f"max err in bf16 half-ulps {(err / ulp)[m].max():.4f}")
zt = np.linspace(0, 60, 600001)
print("P monotonically decreasing on [0, 60]:", bool((np.diff(np.polyval(mono[::-1], zt)) < 0).all()),
      "; P(5) =", float(np.polyval(mono[::-1], 5.0)))

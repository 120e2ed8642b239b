"""Fit of the polynomial used by gelu_erf() in aurora_b200/csrc/ptx.cuh:
R(z) ~ 0.5 * erfcx(z) on [0, 4.5], weighted by the impact on |gelu| error (exp(-z^2) * z), so that
gelu(x) = max(x, 0) - |x| * exp(-z^2) * R(z), z = |x| / sqrt(2), needs one ex2 and no reciprocal."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfcx

L, DEG = 4.5, 6
z = np.cos(np.linspace(0, np.pi, 6001)) * L / 2 + L / 2
f = 0.5 * erfcx(z)
V = C.chebvander(2 * z / L - 1, DEG)
wt = np.exp(-z * z) * (z + 0.05)
w = wt.copy()
for _ in range(60):  # Remez-like re-weighting towards a minimax fit
    coef, *_ = np.linalg.lstsq(V * w[:, None], f * w, rcond=None)
    err = np.abs((V @ coef - f) * wt)
    w = w * (1 + 3 * err / err.max())
cz = P.Polynomial(C.cheb2poly(coef))(P.Polynomial([-1, 2 / L])).coef.astype(np.float32)
x = np.linspace(-9, 9, 400001).astype(np.float32)
zz = np.minimum(np.abs(x) * np.float32(0.7071067811865476), np.float32(L))
r = np.zeros_like(zz) + cz[-1]
for c in cz[-2::-1]:
    r = r * zz + c
g = np.maximum(x, 0) - np.abs(x) * (np.exp2(-(zz * zz) * np.float32(1.4426950408889634)).astype(np.float32) * r)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("coefficients (z^0..z^6):", [float(c) for c in cz])
print("max |gelu - exact| =", np.abs(g - ref).max())

"""Coefficients of the polynomial GELU used by the GEMM epilogue (csrc/common.cuh: gelu_erf).

gelu(x) = x * Phi(x),  Phi(x) = 0.5 * (1 + erf(x / sqrt(2))) ~ sat(0.5 + x * Q(x^2))   with sat = clamp to [0, 1] (free: FFMA.SAT).
Q has degree 8 in u = x^2, fitted on |x| <= 4.25 by Lawson-weighted least squares in a Chebyshev basis with weight |x| (the error that
matters is the one of x * Phi), converted to monomials.  The leading coefficient is positive, so beyond the fitted range x * Q(x^2)
runs monotonically to +-inf and the saturation returns Phi = 1 / 0 exactly: no range clamp on u is needed (checked below on
|x| <= 12 densely and out to 3e38).  Prints the fp32 coefficients and the max |gelu error| of the fp32 Horner evaluation.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf

XMAX, DEG = 4.25, 8
x = np.linspace(1e-6, XMAX, 60001)
t = 2 * x * x / XMAX ** 2 - 1
V = C.chebvander(t, DEG) * x[:, None]                       # x * Q(u)
y = 0.5 * erf(x / np.sqrt(2))                               # Phi(x) - 0.5
w = x.copy()
for _ in range(600):
    c = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
    e = np.abs(V @ c - y) * x
    w = w * (0.5 + e / e.max())
    w /= w.mean()
ps = np.zeros(1)
for k, ck in enumerate(C.cheb2poly(c)):
    ps = P.polyadd(ps, ck * P.polypow([-1.0, 2.0], k))      # t = 2 s - 1, s = u / XMAX^2
d = [np.float32(ck / XMAX ** (2 * k)) for k, ck in enumerate(ps)]
print("coefficients of u^k:", ["%.9e" % v for v in d])


def gelu_kernel(xv):
    """fp32 mirror of csrc/common.cuh gelu_erf."""
    xv = xv.astype(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        u = (xv * xv).astype(np.float32)
        acc = np.full_like(u, d[-1])
        for ck in d[-2::-1]:
            acc = (acc * u + ck).astype(np.float32)
        phi = np.clip((xv * acc + np.float32(0.5)).astype(np.float32), np.float32(0), np.float32(1))
    return (xv * phi).astype(np.float32)


xs = np.concatenate([np.linspace(-12, 12, 2400001), np.logspace(0, 38, 40001), -np.logspace(0, 38, 40001)]).astype(np.float32)
xe = xs.astype(np.float64)
print("max |gelu error|, |x| <= 12 dense + out to 1e38:", np.nanmax(np.abs(gelu_kernel(xs) - 0.5 * xe * (1 + erf(xe / np.sqrt(2))))),
      "NaNs:", int(np.isnan(gelu_kernel(xs)).sum()))
far = np.array([-3e38, -1e19, -1e4, -100.0, -50.0, -6.0, 6.0, 50.0, 100.0, 1e4, 1e19, 3e38], dtype=np.float32)
print("far from the origin:", dict(zip(far.tolist(), gelu_kernel(far).tolist())))

"""Coefficients of the polynomial erf used by the GELU epilogue (csrc/common.cuh: gelu_erf).

erf(z) ~ z * P(z^2) on [0, 3], minimax by Lawson-weighted least squares in a Chebyshev basis, converted to monomials in u = z^2;
prints the fp32 coefficients (highest degree last) and the max |gelu error| of the fp32 Horner evaluation against scipy's erf.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf

ZMAX, DEG = 3.0, 8
z = np.linspace(1e-6, ZMAX, 40001)
t = 2 * z * z / ZMAX ** 2 - 1
V = C.chebvander(t, DEG) * z[:, None]
y = erf(z)
w = np.ones_like(z)
for _ in range(300):
    c = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
    e = np.abs(V @ c - y)
    w = w * (0.5 + e / e.max())
    w /= w.mean()
ps = np.zeros(1)
for k, ck in enumerate(C.cheb2poly(c)):
    ps = P.polyadd(ps, ck * P.polypow([-1.0, 2.0], k))          # t = 2 s - 1, s = u / ZMAX^2
d = [np.float32(ck / ZMAX ** (2 * k)) for k, ck in enumerate(ps)]
print("coefficients of u^k:", ["%.9e" % v for v in d])
def gelu_kernel(x):
    """fp32 mirror of csrc/common.cuh gelu_erf: erf = clamp(z * P(min(z^2, 9)), -1, 1)."""
    x = x.astype(np.float32)
    zf = (x * np.float32(0.7071067811865476)).astype(np.float32)
    u = np.minimum((zf * zf).astype(np.float32), np.float32(ZMAX * ZMAX))
    acc = np.full_like(u, d[-1])
    for ck in d[-2::-1]:
        acc = (acc * u + ck).astype(np.float32)
    e = np.clip((zf * acc).astype(np.float32), np.float32(-1), np.float32(1))
    hx = (x * np.float32(0.5)).astype(np.float32)
    return (hx * e + hx).astype(np.float32)


x = np.linspace(-10, 10, 800001).astype(np.float32)
xe = x.astype(np.float64)
print("max |gelu error| on [-10, 10]:", np.abs(gelu_kernel(x) - 0.5 * xe * (1 + erf(xe / np.sqrt(2)))).max())
far = np.array([-1e4, -100.0, -50.0, -6.0, 6.0, 50.0, 100.0, 1e4], dtype=np.float32)
print("far from the origin:", dict(zip(far.tolist(), gelu_kernel(far).tolist())))

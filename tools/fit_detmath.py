"""Derive the polynomial coefficients used by include/cpm_detmath.h.

Least-squares Chebyshev-node fits in float64, rounded to float32.  Run once; the
printed hex-float constants are pasted into the header.  Kept in the repo so the
constants are reproducible.
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P

def fit(f, lo, hi, deg, n=4000):
    k = np.arange(n)
    x = 0.5*(lo+hi) + 0.5*(hi-lo)*np.cos(np.pi*(k+0.5)/n)
    c = C.Chebyshev.fit(x, f(x), deg, domain=[lo, hi])
    p = c.convert(kind=P.Polynomial, domain=[lo,hi], window=[lo,hi])
    return p.coef

def show(name, coef):
    print(name)
    for i, c in enumerate(coef):
        f = np.float32(c)
        print(f"  c{i} = {float(f).hex()}  ({f!r})")

# asin(x) = x + x*z*Q(z), z = x^2 in [0, 0.25]  -> Q(z) = (asin(x)/x - 1)/z
def q_asin(z):
    x = np.sqrt(z)
    return (np.arcsin(x)/x - 1.0)/z
show("asin Q(z), z in [1e-12,0.25], deg 5", fit(q_asin, 1e-9, 0.25, 5))

# atan(t) = t + t*z*R(z), z=t^2, |t| <= tan(pi/8)
def r_atan(z):
    t = np.sqrt(z)
    return (np.arctan(t)/t - 1.0)/z
T = np.tan(np.pi/8)
show("atan R(z), z in [1e-9, tan(pi/8)^2], deg 5", fit(r_atan, 1e-9, T*T, 5))

# check errors
def horner(coef, z):
    r = np.zeros_like(z)
    for c in coef[::-1]:
        r = r*z + np.float64(np.float32(c))
    return r
z = np.linspace(1e-9, 0.25, 100001)
x = np.sqrt(z)
print("asin max rel err", np.max(np.abs((x + x*z*horner(fit(q_asin,1e-9,0.25,5), z))/np.arcsin(x) - 1)))
z = np.linspace(1e-9, T*T, 100001)
t = np.sqrt(z)
print("atan max rel err", np.max(np.abs((t + t*z*horner(fit(r_atan,1e-9,T*T,5), z))/np.arctan(t) - 1)))


# ---- cpm_native_logf: 32-interval table over m in [0.75, 1.5) ------------------------------------------------------------
# interval i = top five bits of the mantissa field after the "+0x00400000" shift (see cpm_detmath.h); rc = fl(1 / centre),
# lc = fl(-log(rc)) -- the logarithm of the EXACT reciprocal of rc, so that m * rc - 1 carries no table rounding error.
# The two intervals that touch m = 1 use rc = 1, lc = 0: log(x) near 1 keeps its relative accuracy.
def nlog_table():
    import math
    rows = []
    for i in range(32):
        if i in (15, 16):
            c = 1.0
        elif i < 15:
            c = 0.75 + (i + 0.5) / 64
        else:
            c = 1.0 + (i - 15.5) / 32
        rc = np.float32(1.0 / c)
        lc = np.float32(-math.log(float(rc)))
        rows.append((rc, lc))
    print("/* CPM_NLOG_TABLE: {rc, lc} x 32 (tools/fit_detmath.py: nlog_table) */")
    for i in range(0, 32, 2):
        print("    " + " ".join(f"{float(a).hex()}f, {float(b).hex()}f," for a, b in rows[i:i + 2]))
nlog_table()

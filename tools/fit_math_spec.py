"""Derive the fp32 polynomial coefficients of the bit-reproducible math spec (DESIGN.md "Math spec").

Run once; the printed constants are pasted (as hex floats) into csrc/nig_math.cuh and, independently,
into oracle/nig_oracle.c. Both sides evaluate them with the same fmaf Horner order, so GPU and
CPU oracle agree bit-for-bit. Accuracy targets: exp <= ~1 ulp; log/sincos ~1e-7 abs.
The log / sincos fits belonged to the Box-Muller normals of the first builds; the normals are now the table inverse CDF
of tools/fit_normal_table.py and only the exp fit is still in use (the fp64 sin/cos of the robot were fitted separately).
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P


def cheb_fit(fn, lo, hi, deg, n=4001):
    k = np.arange(n)
    x = np.cos(np.pi * (k + 0.5) / n)
    t = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
    c = C.chebfit(x, fn(t), deg)
    # convert to monomial in t
    p = C.cheb2poly(c)
    # x = (2t - (hi+lo)) / (hi-lo)
    a = 2.0 / (hi - lo)
    b = -(hi + lo) / (hi - lo)
    out = np.zeros(1)
    lin = np.array([b, a])
    powp = np.array([1.0])
    for ck in p:
        out = P.polyadd(out, ck * powp)
        powp = P.polymul(powp, lin)
    return out


def horner32(coefs, t):
    t = t.astype(np.float32)
    acc = np.full_like(t, np.float32(coefs[-1]))
    for c in coefs[-2::-1]:
        acc = (acc.astype(np.float64) * t.astype(np.float64) + np.float64(np.float32(c))).astype(np.float32)
    return acc


def show(name, coefs):
    print(name)
    for i, c in enumerate(coefs):
        print(f"  c{i} = {float(np.float32(c)).hex():>22s}  /* {np.float32(c):.9e} */")


# log(m) = f*Q(f), f = m-1 in [sqrt(.5)-1, sqrt(2)-1]
lo, hi = np.sqrt(0.5) - 1, np.sqrt(2) - 1
def q(f):
    f = np.where(np.abs(f) < 1e-12, 1e-12, f)
    return np.log1p(f) / f
for deg in (7, 8, 9):
    cq = cheb_fit(q, lo, hi, deg)
    t = np.linspace(lo, hi, 200001)
    approx = horner32(cq, t).astype(np.float64) * t.astype(np.float32)
    err = np.max(np.abs(approx - np.log1p(t.astype(np.float32).astype(np.float64))))
    print("log deg", deg, "max abs err", err)
cq = cheb_fit(q, lo, hi, 8)
show("LOG_Q (log(1+f) = f*Q(f))", cq)

# sin(pi/4 * y)/y... use phi = y*(pi/4), y in [-1,1]: sin(phi) = phi*S(phi^2), cos(phi)=Cc(phi^2)
def s(z):
    r = np.sqrt(np.maximum(z, 1e-30)); return np.sin(r) / r
def c(z):
    return np.cos(np.sqrt(np.maximum(z, 0)))
zmax = (np.pi / 4) ** 2
cs = cheb_fit(s, 0, zmax, 4)
cc = cheb_fit(c, 0, zmax, 4)
phi = np.linspace(-np.pi / 4, np.pi / 4, 200001).astype(np.float32)
z = (phi * phi).astype(np.float32)
sa = horner32(cs, z).astype(np.float64) * phi
ca = horner32(cc, z).astype(np.float64)
print("sin err", np.max(np.abs(sa - np.sin(phi.astype(np.float64)))), "cos err", np.max(np.abs(ca - np.cos(phi.astype(np.float64)))))
show("SIN_S (sin(p) = p*S(p^2))", cs)
show("COS_C (cos(p) = C(p^2))", cc)

# exp(r), r in [-ln2/2, ln2/2]: exp(r) = 1 + r*E(r)
def e(r):
    r = np.where(np.abs(r) < 1e-12, 1e-12, r)
    return np.expm1(r) / r
h = np.log(2) / 2 * 1.0001
for deg in (5, 6):
    ce = cheb_fit(e, -h, h, deg)
    r = np.linspace(-h, h, 400001).astype(np.float32)
    ea = (horner32(ce, r).astype(np.float64) * r + 1.0).astype(np.float32)
    ex = np.exp(r.astype(np.float64))
    ulp = np.abs(ea.astype(np.float64) - ex) / np.spacing(ex.astype(np.float32)).astype(np.float64)
    print("exp deg", deg, "max ulp err", ulp.max())
ce = cheb_fit(e, -h, h, 6)
show("EXP_E (exp(r) = 1 + r*E(r))", ce)
print("ln2_hi", float(np.float32(0.693145751953125)).hex(), "ln2_lo", float(np.float32(np.log(2) - 0.693145751953125)).hex(),
      "log2e", float(np.float32(1 / np.log(2))).hex(), "pi/4", float(np.float32(np.pi / 4)).hex(), "ln2", float(np.float32(np.log(2))).hex())

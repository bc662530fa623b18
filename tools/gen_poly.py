#!/usr/bin/env python
"""Generate the polynomial coefficients of csrc/b200mp_math.cuh with mpmath (Chebyshev interpolation
at 60 digits, converted to the monomial basis), and report the max error of each double-precision
evaluation scheme.  Output is pasted into the header; rerun to audit.

    python tools/gen_poly.py
"""
import struct

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, n):
    """Degree-(n-1) Chebyshev interpolant of f on [a,b] -> monomial coefficients."""
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n):
            A[i, j] = x ** j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[j] for j in range(n)]


def horner(coefs, u):
    acc = np.full_like(u, float(coefs[-1]))
    for c in coefs[-2::-1]:
        acc = acc * u + float(c)
    return acc


def show(name, coefs):
    print(f"// {name}")
    print("   {" + ", ".join(repr(float(c)) for c in coefs) + "}")


def split(x, bits):
    b = struct.unpack("<Q", struct.pack("<d", float(x)))[0]
    b &= ~((1 << (53 - bits)) - 1)
    return mp.mpf(struct.unpack("<d", struct.pack("<Q", b))[0])


def main():
    half_pi = mp.pi / 2
    umax = (half_pi * mp.mpf("1.02")) ** 2
    fs = lambda u: (mp.sin(mp.sqrt(u)) / mp.sqrt(u) - 1) / u if u > 0 else mp.mpf(-1) / 6
    for n in (7, 8):
        P = cheb_fit(fs, mp.mpf(0), umax, n)
        r = np.linspace(-float(half_pi), float(half_pi), 20001)
        u = r * r
        approx = r + r * u * horner(P, u)
        exact = np.array([float(mp.sin(mp.mpf(float(x)))) for x in r])
        print(f"sin half-pi n={n}: max abs err {np.abs(approx - exact).max():.3e}")
        if n == 8:
            show("kSinHalfPi: sin(r) = r + r*u*P(u), u = r*r, |r| <= pi/2", P)
    tmax = mp.tan(mp.pi / 8) * mp.mpf("1.01")
    fa = lambda u: (mp.atan(mp.sqrt(u)) / mp.sqrt(u) - 1) / u if u > 0 else mp.mpf(-1) / 3
    for n in (10, 11, 12):
        Q = cheb_fit(fa, mp.mpf(0), tmax ** 2, n)
        t = np.linspace(1e-6, float(tmax), 20001)
        u = t * t
        approx = t + t * u * horner(Q, u)
        exact = np.array([float(mp.atan(mp.mpf(float(x)))) for x in t])
        print(f"atan pi/8 n={n}: max rel err {(np.abs(approx - exact) / exact).max():.3e}")
        if n == 10:
            show("kAtanPi8: atan(t) = t + t*u*Q(u), u = t*t, |t| <= tan(pi/8)", Q)
    umax4 = (mp.pi / 4 * mp.mpf("1.02")) ** 2
    S = cheb_fit(fs, mp.mpf(0), umax4, 6)
    fc = lambda u: (mp.cos(mp.sqrt(u)) - 1 + u / 2) / (u * u) if u > 0 else mp.mpf(1) / 24
    Cc = cheb_fit(fc, mp.mpf(0), umax4, 6)
    r = np.linspace(-float(mp.pi / 4), float(mp.pi / 4), 20001)
    u = r * r
    s_ap = r + r * u * horner(S, u)
    c_ap = 1.0 - 0.5 * u + u * u * horner(Cc, u)
    es = np.abs(s_ap - np.array([float(mp.sin(mp.mpf(float(x)))) for x in r])).max()
    ec = np.abs(c_ap - np.array([float(mp.cos(mp.mpf(float(x)))) for x in r])).max()
    print(f"sin pi/4 n=6: max abs err {es:.3e}; cos pi/4 n=6: {ec:.3e}")
    show("kSinPi4: sin(r) = r + r*u*S(u), |r| <= pi/4", S)
    show("kCosPi4: cos(r) = 1 - u/2 + u*u*C(u), |r| <= pi/4", Cc)
    pi_hi = mp.mpf(float(mp.pi))
    pi_lo = mp.mpf(float(mp.pi - pi_hi))
    print("pi hi/lo:", repr(float(pi_hi)), repr(float(pi_lo)))
    h1 = split(half_pi, 33)
    h2 = split(half_pi - h1, 33)
    h3 = mp.mpf(float(half_pi - h1 - h2))
    print("pi/2 3-term:", [repr(float(v)) for v in (h1, h2, h3)])
    print("1/pi", repr(float(1 / mp.pi)), "2/pi", repr(float(2 / mp.pi)))
    print("tan(pi/8)", repr(float(mp.tan(mp.pi / 8))), "tan(3pi/8)", repr(float(mp.tan(3 * mp.pi / 8))))
    print("pi/4 hi lo", repr(float(mp.pi / 4)), repr(float(mp.pi / 4 - mp.mpf(float(mp.pi / 4)))))
    print("pi/2 hi lo", repr(float(half_pi)), repr(float(half_pi - mp.mpf(float(half_pi)))))


if __name__ == "__main__" and "--f32" not in __import__("sys").argv:
    main()


def main_f32():
    """FP32 schemes of Math<float>: coefficients rounded to float, evaluated in float32 arithmetic."""
    f32 = np.float32
    half_pi = mp.pi / 2
    fs = lambda u: (mp.sin(mp.sqrt(u)) / mp.sqrt(u) - 1) / u if u > 0 else mp.mpf(-1) / 6
    fa = lambda u: (mp.atan(mp.sqrt(u)) / mp.sqrt(u) - 1) / u if u > 0 else mp.mpf(-1) / 3

    def h32(coefs, u):
        acc = np.full_like(u, f32(float(coefs[-1])))
        for c in coefs[-2::-1]:
            acc = (acc * u + f32(float(c))).astype(f32)
        return acc

    for n in (4, 5, 6):
        P = cheb_fit(fs, mp.mpf(0), (half_pi * mp.mpf("1.01")) ** 2, n)
        r = np.linspace(-float(half_pi), float(half_pi), 40001).astype(f32)
        u = (r * r).astype(f32)
        approx = (r + (r * u).astype(f32) * h32(P, u)).astype(f32)
        exact = np.array([float(mp.sin(mp.mpf(float(x)))) for x in r])
        err = np.abs(approx.astype(np.float64) - exact)
        print(f"f32 sin half-pi n={n}: max abs err {err.max():.3e} ({err.max() / 2 ** -24:.2f} ulp of 1)")
        if n == 5:
            print("   kSinHalfPiF = {" + ", ".join(repr(float(f32(float(c)))) + "f" for c in P) + "}")
    tmax = mp.tan(mp.pi / 8) * mp.mpf("1.005")
    for n in (4, 5, 6):
        Q = cheb_fit(fa, mp.mpf(0), tmax ** 2, n)
        t = np.linspace(1e-4, float(tmax), 40001).astype(f32)
        u = (t * t).astype(f32)
        approx = (t + (t * u).astype(f32) * h32(Q, u)).astype(f32)
        exact = np.array([float(mp.atan(mp.mpf(float(x)))) for x in t])
        rel = np.abs(approx.astype(np.float64) - exact) / exact
        print(f"f32 atan pi/8 n={n}: max rel err {rel.max():.3e} ({rel.max() / 2 ** -24:.2f} ulp)")
        if n == 5:
            print("   kAtanPi8F = {" + ", ".join(repr(float(f32(float(c)))) + "f" for c in Q) + "}")
    pi_hi = f32(float(mp.pi))
    pi_lo = f32(float(mp.pi - mp.mpf(float(pi_hi))))
    hp_hi = f32(float(half_pi))
    hp_lo = f32(float(half_pi - mp.mpf(float(hp_hi))))
    print("f32 pi hi/lo", repr(float(pi_hi)), repr(float(pi_lo)), " pi/2 hi/lo", repr(float(hp_hi)), repr(float(hp_lo)),
          " 1/pi", repr(float(f32(float(1 / mp.pi)))), " pi/4", repr(float(f32(float(mp.pi / 4)))),
          " tan(pi/8)", repr(float(f32(float(mp.tan(mp.pi / 8))))), " tan(3pi/8)", repr(float(f32(float(mp.tan(3 * mp.pi / 8))))))


if __name__ == "__main__":
    import sys
    if "--f32" in sys.argv:
        main_f32()

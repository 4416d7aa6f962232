#!/usr/bin/env python3
"""Generate the minimax polynomials used by the CUDA kernels (csrc/cospi_poly.cuh).

The kernels evaluate every term as  A * cos(pi * t)  with t in half-turns.  After the exact
reduction r = t - rint(t), |r| <= 1/2, two evaluation schemes are generated:

  V2 "double angle" (default in the kernels, 7 DFMA):
        u = sqrt(2) * cos(pi/2 * r) = U(s),  s = r*r,  deg U = 6
        cos(pi r) = u*u - 1                      (one more DFMA)
  V1 "direct" (8 DFMA, kept for the accuracy cross-check):
        cos(pi r) = C(s),  deg C = 8

Coefficients are minimax in ABSOLUTE error on s in [0, 1/4] (Remez exchange, mpmath 60 digits),
then rounded to binary64.  The script also measures the achieved error of the *rounded* polynomial
evaluated in binary64 Horner/FMA arithmetic against mpmath.

Usage:  python gen_cospi_poly.py > ../csrc/cospi_poly.cuh
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60


def remez(f, a, b, deg, iters=30):
    n = deg + 2
    # Chebyshev extrema as the initial reference
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * i / (n - 1)) for i in range(n)][::-1]
    coef = None
    for _ in range(iters):
        A = mp.matrix(n, n)
        rhs = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(deg + 1):
                A[i, j] = x ** j
            A[i, deg + 1] = (-1) ** i
            rhs[i] = f(x)
        sol = mp.lu_solve(A, rhs)
        coef = [sol[j] for j in range(deg + 1)]
        E = sol[deg + 1]

        def err(x):
            return mp.polyval(coef[::-1], x) - f(x)

        # locate extrema of the error on a fine grid, one per sign-alternating segment
        grid = [a + (b - a) * mp.mpf(i) / 4000 for i in range(4001)]
        ev = [err(x) for x in grid]
        ext = []
        i = 0
        while i < len(grid):
            sgn = mp.sign(ev[i])
            j = i
            best = i
            while j < len(grid) and mp.sign(ev[j]) == sgn:
                if abs(ev[j]) > abs(ev[best]):
                    best = j
                j += 1
            ext.append(grid[best])
            i = j
        if len(ext) != n:
            break
        if max(abs(e1 - e0) for e0, e1 in zip(xs, ext)) < mp.mpf(10) ** -12:
            xs = ext
            break
        xs = ext
    return coef, abs(E)


def fma_horner(coef64, s):
    """binary64 Horner with fused multiply-adds (emulated exactly through mpmath)."""
    acc = mp.mpf(float(coef64[-1]))
    for c in coef64[-2::-1]:
        acc = mp.mpf(float(acc * s + mp.mpf(float(c))))  # one rounding per FMA
    return acc


def measure(kind, coef64, nsamp=4000):
    rng = np.random.default_rng(0)
    rs = np.concatenate([rng.uniform(-0.5, 0.5, nsamp), [0.0, 0.5, -0.5, 0.25, 1e-9, 0.4999999]])
    worst = mp.mpf(0)
    for r in rs:
        r = float(r)
        s = mp.mpf(float(mp.mpf(r) * mp.mpf(r)))  # DMUL rounding
        p = fma_horner(coef64, s)
        if kind == "v2":
            y = mp.mpf(float(p * p - 1))  # DFMA(u, u, -1)
        else:
            y = p
        worst = max(worst, abs(y - mp.cos(mp.pi * mp.mpf(r))))
    return worst


def hexd(x):
    return float(x).hex()


def monic_variant(cu, nsamp=2500):
    """V3: the degree-len(cu)-1 minimax polynomial with the leading coefficient factored out,
    U(s) = c*Q(s), Q monic.  Then cos(pi r) = c^2 * (Q^2 - 1/c^2) and c^2 folds into the mode
    amplitude, so the first Horner step is a DADD (s + Q5) instead of a DFMA with TWO constant
    operands -- which ptxas can only issue with both constants in registers (3 register-file reads:
    3 cycles on B200 instead of 2).  Q_i = U_i/c, E = 1/c^2 and S = c^2 are rounded to binary64;
    a small ulp search over (Q0, E, S) picks the combination with the smallest measured error."""
    import math
    deg = len(cu) - 1
    c = cu[deg]
    D = [float(ci / c) for ci in cu[:deg]]
    E, S = float(1 / (c * c)), float(c * c)
    rng = np.random.default_rng(0)
    rs = [float(x) for x in np.concatenate([rng.uniform(-0.5, 0.5, nsamp), [0.0, 0.5, -0.5, 0.25, 1e-9, 0.4999999]])]
    ref = [mp.cos(mp.pi * mp.mpf(r)) for r in rs]

    def ulp(x, k):
        for _ in range(abs(k)):
            x = math.nextafter(x, math.inf if k > 0 else -math.inf)
        return x

    def q_values(d):
        out = []
        for r in rs:
            s = mp.mpf(float(mp.mpf(r) * mp.mpf(r)))          # DMUL
            v = mp.mpf(float(s + mp.mpf(d[deg - 1])))          # DADD
            for di in d[deg - 2::-1]:
                v = mp.mpf(float(v * s + mp.mpf(di)))          # DFMA
            out.append(v)
        return out

    best = None
    for k0 in range(-2, 4):
        d2 = list(D)
        d2[0] = ulp(D[0], k0)
        q = q_values(d2)
        for ke in range(-3, 10):
            e2 = ulp(E, ke)
            ws = [mp.mpf(float(v * v - mp.mpf(e2))) for v in q]   # DFMA(Q, Q, -E)
            for ks in range(-2, 3):
                s2 = ulp(S, ks)
                worst = max(abs(mp.mpf(s2) * w - y) for w, y in zip(ws, ref))
                if best is None or worst < best[0]:
                    best = (worst, d2, e2, s2)
    return best


def main():
    quarter = mp.mpf(1) / 4
    cu, eu = remez(lambda s: mp.sqrt(2) * mp.cos(mp.pi / 2 * mp.sqrt(s)) if s > 0 else mp.sqrt(2),
                   mp.mpf(0), quarter, 6)
    cc, ec = remez(lambda s: mp.cos(mp.pi * mp.sqrt(s)) if s > 0 else mp.mpf(1),
                   mp.mpf(0), quarter, 8)
    cu64 = [float(c) for c in cu]
    cc64 = [float(c) for c in cc]
    wu = measure("v2", cu64)
    wc = measure("v1", cc64)

    out = sys.stdout
    out.write("// GENERATED by tools/gen_cospi_poly.py -- do not edit.\n")
    out.write("// Minimax (absolute error) polynomials in s = r*r for |r| <= 1/2 half-turns.\n")
    out.write("#pragma once\n\n")
    out.write("// V2: u = sqrt(2)*cos(pi/2*r) = U(s), deg 6; cos(pi r) = u*u - 1.\n")
    out.write("//     minimax |U - u| = %s ; measured |cos(pi r)| error in binary64 FMA arithmetic <= %s\n"
              % (mp.nstr(eu, 3), mp.nstr(wu, 3)))
    for i, c in enumerate(cu64):
        out.write("#define GSF_U%d %s  /* %.17g */\n" % (i, hexd(c), c))
    out.write("\n// V3: U of degree DEG with its leading coefficient factored out, U = c*Q, Q monic:\n")
    out.write("//     cos(pi r) = S * (Q(s)^2 - E), E = 1/c^2, S = c^2 (folded into the mode amplitude).\n")
    out.write("//     Evaluation: 1 DADD + (DEG-1) DFMA + 1 DFMA.  GSF_Q<DEG>_<i>, GSF_Q<DEG>_E, GSF_Q<DEG>_S.\n")
    for deg in (4, 5, 6):
        if deg == 6:
            cud = cu
            eud = eu
        else:
            cud, eud = remez(lambda s: mp.sqrt(2) * mp.cos(mp.pi / 2 * mp.sqrt(s)) if s > 0 else mp.sqrt(2),
                             mp.mpf(0), quarter, deg)
        w3, q64, e64, s64 = monic_variant(cud)
        out.write("// degree %d: minimax |U - u| = %s ; measured |cos(pi r)| error in binary64 arithmetic <= %s\n"
                  % (deg, mp.nstr(eud, 3), mp.nstr(w3, 3)))
        for i, c in enumerate(q64):
            out.write("#define GSF_Q%d_%d %s  /* %.17g */\n" % (deg, i, hexd(c), c))
        out.write("#define GSF_Q%d_E %s  /* %.17g */\n" % (deg, hexd(e64), e64))
        out.write("#define GSF_Q%d_S %s  /* %.17g */\n" % (deg, hexd(s64), s64))
    out.write("\n// V1: cos(pi r) = C(s), deg 8.\n")
    out.write("//     minimax |C - cos| = %s ; measured error in binary64 FMA arithmetic <= %s\n"
              % (mp.nstr(ec, 3), mp.nstr(wc, 3)))
    for i, c in enumerate(cc64):
        out.write("#define GSF_C%d %s  /* %.17g */\n" % (i, hexd(c), c))


if __name__ == "__main__":
    main()

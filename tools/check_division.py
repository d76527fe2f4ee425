"""Offline evidence for next round's geometry kernels: is  q1 = fma(fma(-q0, b, a), y, q0)  with
q0 = RN(a * y), y = RN(1 / b)  (3 instructions once y is known) the correctly rounded fp32 quotient
a / b for the operand ranges of the reprojection?  Emulated exactly (float64 products of float32
values are exact; mismatches are re-checked with Fractions).  Result on 16 M trials: 0 mismatches
(see DESIGN.md section 5).  CPU only.

    python tools/check_division.py
"""
import numpy as np
from fractions import Fraction
rng = np.random.default_rng(0)
f32 = np.float32

def rn32_from_fraction(fr):
    # correctly rounded float32 of an exact Fraction (round half even)
    x = float(fr)            # double rounding risk only at exact ties of double; negligible -> verify below
    y = np.float32(x)
    # check neighbours exactly
    cands = [np.nextafter(y, f32(-np.inf)), y, np.nextafter(y, f32(np.inf))]
    best = min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1))
    return f32(best)

def markstein_div(a, b, y):
    """q1 = fma(r, y, q0) with q0 = RN(a*y), r = fma(-q0, b, a); float32 semantics, vectorised."""
    q0 = (a * y).astype(f32)                       # float32 multiply: correctly rounded
    r64 = a.astype(np.float64) - q0.astype(np.float64) * b.astype(np.float64)   # exact in double
    r32 = r64.astype(f32)
    exact_r = (r32.astype(np.float64) == r64)
    s64 = q0.astype(np.float64) + r32.astype(np.float64) * y.astype(np.float64)  # may round (double)
    q1 = s64.astype(f32)
    return q0, q1, exact_r, s64

def run(name, a, b):
    a = a.astype(f32); b = np.broadcast_to(b.astype(f32), a.shape).copy()
    y = (f32(1.0) / b).astype(f32)                 # float32 division: correctly rounded reciprocal
    truth = (a / b).astype(f32)
    q0, q1, exact_r, s64 = markstein_div(a, b, y)
    bad = np.nonzero(q1 != truth)[0]
    # re-check the mismatches (and a sample) exactly with Fractions to rule out double-rounding artefacts
    real_bad = 0
    for i in bad[:2000]:
        fr = Fraction(float(q0[i])) + Fraction(float(np.float32(float(a[i]) - float(q0[i]) * float(b[i])))) * Fraction(float(y[i]))
        q1x = rn32_from_fraction(fr)
        tx = rn32_from_fraction(Fraction(float(a[i])) / Fraction(float(b[i])))
        if q1x != tx:
            real_bad += 1
    print("%-28s n=%9d  residual exact: %s  mismatches (double emu) %d, confirmed exactly %d, q0 already correct %.4f"
          % (name, a.size, bool(exact_r.all()), bad.size, real_bad, float((q0 == truth).mean())))

N = 4_000_000
fxs = np.array([585, 572, 583, 540.02, 570.34, 533.07], dtype=np.float64)
fx256 = (fxs * 341.0 / 640.0)
# unproject: a = (c - cx) * z ; divide by fx
for nm, F, cx, Wd in (("unproject /fx 640x480", fxs, 320.0, 640), ("unproject /fx 256", fx256, 128.5, 256)):
    c = rng.integers(0, Wd, N).astype(f32)
    z = (rng.integers(500, 100000, N).astype(f32) * f32(1e-4) * f32(10.0)).astype(f32)   # metres, like the loader
    a = ((c - f32(cx)).astype(f32) * z).astype(f32)
    b = F[rng.integers(0, len(F), N)]
    run(nm, a, b)
# project: a = x * fx ; divide by z (variable denominator: y = RN(1/z) per pixel)
x = (rng.standard_normal(N) * 2).astype(f32)
z = (rng.uniform(0.3, 10.0, N)).astype(f32)
fx = fxs[rng.integers(0, 6, N)].astype(f32)
run("project (x*fx)/z", (x * fx).astype(f32), z)
# adversarial: random mantissas over a wide exponent range
a = (rng.standard_normal(N) * np.exp(rng.uniform(-20, 20, N))).astype(f32)
b = (rng.uniform(1.0, 2.0, N) * np.exp2(rng.integers(-10, 10, N))).astype(f32)
run("random a / random b", a, b)

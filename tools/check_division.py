"""Evidence for the three-instruction division of the geometry kernels (csrc/geometry.cu,
div_exact):  is  q1 = fma(fma(-q0, b, a), y, q0)  with q0 = RN(a * y), y = RN(1 / b)  the correctly
rounded fp32 quotient a / b?  Emulated exactly (float64 products of float32 values are exact; sums
that land within two float64 ulps of a float32 rounding boundary, and all mismatches, are re-checked
with Fractions).  CPU only.

    python tools/check_division.py               # 16 M sampled trials over the kernels' operand ranges
    python tools/check_division.py --exhaustive  # every fp32 significand of a, for each 3DMatch focal
                                                 # length (raw and after Resize(256)+CenterCrop(256))

A fixed divisor only needs one binade of numerators: scaling a by a power of two scales q0, r and q1
exactly (no under/overflow inside the range the kernels guard), and the sign is symmetric.
Results: sampled 0 mismatches in 16 M; exhaustive 0 mismatches in 18 x 2^23 (profiles/
r1_division_exhaustive.txt).
"""
import sys
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

def exhaustive(b):
    """All 2^23 significands a in [1, 2) against the fixed divisor b.  Returns (mismatches, rechecked)."""
    b32 = f32(b)
    y32 = f32(f32(1.0) / b32)
    bad_total = recheck_total = 0
    for lo in range(1 << 23, 1 << 24, 1 << 21):
        a = (np.arange(lo, lo + (1 << 21), dtype=np.uint32)).astype(np.float64)
        a = (a / float(1 << 23)).astype(f32)                      # 1.xxx, every significand once
        bb = np.full(a.shape, b32, f32)
        yy = np.full(a.shape, y32, f32)
        truth = (a / bb).astype(f32)
        q0, q1, exact_r, s64 = markstein_div(a, bb, yy)
        assert exact_r.all(), "residual not exactly representable"
        # sums whose float64 rounding could have moved them across a float32 rounding boundary
        up = np.nextafter(q1, f32(np.inf)).astype(np.float64)
        dn = np.nextafter(q1, f32(-np.inf)).astype(np.float64)
        q64 = q1.astype(np.float64)
        ulp64 = np.spacing(np.abs(s64))
        risky = (np.abs(s64 - (q64 + up) * 0.5) <= 2 * ulp64) | (np.abs(s64 - (q64 + dn) * 0.5) <= 2 * ulp64)
        idx = np.nonzero(risky | (q1 != truth))[0]
        recheck_total += idx.size
        for i in idx:
            r = Fraction(float(a[i])) - Fraction(float(q0[i])) * Fraction(float(b32))
            q1x = rn32_from_fraction(Fraction(float(q0[i])) + r * Fraction(float(y32)))
            tx = rn32_from_fraction(Fraction(float(a[i])) / Fraction(float(b32)))
            if q1x != tx:
                bad_total += 1
                print("  MISMATCH a=%r b=%r fast=%r ieee=%r" % (float(a[i]), float(b32), float(q1x), float(tx)))
    return bad_total, recheck_total


fxs = np.array([585, 572, 583, 540.02, 570.34, 533.07], dtype=np.float64)
if "--exhaustive" in sys.argv:
    consts = [("raw 640x480", v) for v in fxs.astype(f32)]
    consts += [("fx*341/640", v) for v in (fxs * 341.0 / 640.0).astype(f32)]      # intrinsic_transform, SDD:47-119
    consts += [("fy*256/480", v) for v in (fxs * 256.0 / 480.0).astype(f32)]
    total = 0
    for nm, v in consts:
        bad, re = exhaustive(v)
        total += bad
        print("%-12s b=%-12.6f  2^23 numerators: mismatches %d  (exact re-checks %d)" % (nm, float(v), bad, re))
    print("exhaustive: %d mismatches over %d divisors x 2^23 numerators" % (total, len(consts)))
    sys.exit(1 if total else 0)

N = 4_000_000
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

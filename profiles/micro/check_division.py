"""Numerical check of dpft::div_by (csrc/dpft_device.cuh): RN(a r) with r = RN(1/b) followed by two
exact-remainder corrections equals the IEEE fp32 quotient a / b.  fp32 FMA is emulated in float64 (the product
of two fp32 values is exact there)."""
import numpy as np

rng = np.random.default_rng(1)


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def div_by(a, b):
    r = (1.0 / b.astype(np.float64)).astype(np.float32)
    q = (a * r).astype(np.float32)
    q = fma(fma(-q, b, a), r, q)
    return fma(fma(-q, b, a), r, q)


bad = tot = 0
for _ in range(40):
    n = 2_000_000
    b = rng.uniform(10, 2000, n).astype(np.float32)                                   # (row - cy) / fy
    a = (rng.integers(0, 2000, n) - rng.uniform(0, 1000, n)).astype(np.float32)
    bad += int((div_by(a, b) != (a / b).astype(np.float32)).sum()); tot += n
    b = np.exp(rng.uniform(np.log(0.05), np.log(20), n)).astype(np.float32)           # w.x / w.z, d0 / w.z
    for a in (rng.normal(0, 2, n).astype(np.float32), rng.uniform(0, 10, n).astype(np.float32)):
        bad += int((div_by(a, b) != (a / b).astype(np.float32)).sum()); tot += n
for W in list(range(2, 700)) + [1024, 1280, 1920, 2048, 4096]:                        # u / ((W-1)/2)
    c = np.full(25000, (W - 1) / 2, dtype=np.float32)
    u = rng.uniform(-2 * W, 3 * W, 25000).astype(np.float32)
    bad += int((div_by(u, c) != (u / c).astype(np.float32)).sum()); tot += 25000
print("mismatches", bad, "of", tot)

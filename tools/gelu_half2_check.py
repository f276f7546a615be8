"""Accuracy of the half2-polynomial GELU of the fused FFN epilogues (common.cuh gelu_pair_bf16) against the exact GELU, after the
bf16 rounding of the result; also fits the polynomial (degree 4-6) of log2 erfc(|x| / sqrt 2)."""
import numpy as np
from scipy.special import erfc, erf
def bf16(x):
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)
h = np.float16
def fma16(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(h)
XMAX = 5.939697
def fit(deg):
    xs = np.linspace(0, XMAX, 20001)
    y = np.log2(erfc(xs / np.sqrt(2)))
    # weight: sensitivity of gelu to P: d gelu / dP = -0.5 |x| e ln2
    w = 0.5 * xs * erfc(xs / np.sqrt(2)) * np.log(2) + 1e-6
    return np.polyfit(xs, y, deg, w=w)
def gelu_h2(x, coef, tail32=False):
    x = x.astype(np.float32)
    xh = x.astype(h)
    ax = np.minimum(np.abs(xh), h(XMAX))
    c = [h(v) for v in coef]
    p = np.full_like(ax, c[0])
    for v in c[1:]:
        p = fma16(p, ax, np.full_like(ax, v))
    e = np.exp2(p.astype(np.float64)).astype(h)
    if tail32:
        out = np.maximum(x, 0) - 0.5 * np.abs(x) * e.astype(np.float32)
        return out.astype(np.float32)
    t = (np.abs(xh).astype(np.float64) * e.astype(np.float64)).astype(h)
    out = fma16(np.full_like(t, h(-0.5)), t, np.maximum(xh, h(0)))
    return out.astype(np.float32)
rng = np.random.default_rng(0)
x = np.concatenate([rng.standard_normal(2000000) * 1.5, rng.uniform(-8, 8, 500000)]).astype(np.float32)
exact = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
ref = bf16(exact.astype(np.float32)).astype(np.float64)
def report(name, y):
    y = bf16(y).astype(np.float64)
    err = y - exact
    print(f"{name:28s} rms abs {np.sqrt(np.mean(err**2)):.3e}  max abs {np.abs(err).max():.3e}  rms/rms(gelu) {np.sqrt(np.mean(err**2))/np.sqrt(np.mean(exact**2)):.3e}  mismatch vs exact-bf16 {np.mean(y!=ref):.4f}")
report("exact -> bf16", exact.astype(np.float32))
for deg in (6, 5, 4):
    coef = fit(deg)
    xs = np.linspace(0, XMAX, 20001)
    print(deg, "fit max |dP| on [0,3]:", np.abs(np.polyval(coef, xs) - np.log2(erfc(xs/np.sqrt(2))))[xs<3].max(), list(coef))
    report(f"half2 deg {deg}", gelu_h2(x, coef))
    report(f"half2 deg {deg} fp32 tail", gelu_h2(x, coef, True))

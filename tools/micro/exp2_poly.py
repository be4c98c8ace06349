"""Design aid for round 2: exp2 on the FMA pipe for a fraction of the attention scores.

The tcgen05 attention kernel is MUFU-bound (one ex2 per score, 16 per clock and SM).  Computing some of the exponentials
as   2^x = 2^floor(x) * p(x - floor(x))   with a low-degree polynomial p on [0, 1) moves that work to the FMA pipe
(the integer part goes into the exponent field).  This script fits p by least squares on Chebyshev nodes, evaluates it in
float32 the way the kernel would (Horner with fma), and reports the relative error against the bf16 rounding of the
probabilities (2^-9) that follows in the kernel, to pick the degree.

    python tools/micro/exp2_poly.py
"""
import numpy as np


def fit(degree, n=4096):
    k = np.arange(n)
    f = 0.5 - 0.5 * np.cos((2 * k + 1) * np.pi / (2 * n))          # Chebyshev nodes on [0, 1]
    # minimise the RELATIVE error: fit p(f) / 2^f ~ 1
    V = np.vander(f, degree + 1, increasing=True) / np.exp2(f)[:, None]
    c, *_ = np.linalg.lstsq(V, np.ones_like(f), rcond=None)
    return c


def eval_f32(c, x):
    """float32 evaluation: x <= 0 (scores minus the running maximum), as in the softmax warps."""
    x = x.astype(np.float32)
    fl = np.floor(x)
    f = (x - fl).astype(np.float32)
    acc = np.full_like(f, np.float32(c[-1]))
    for a in c[-2::-1]:
        acc = (acc * f + np.float32(a)).astype(np.float32)          # one FFMA per coefficient
    # scale by 2^floor(x): add floor(x) to the exponent field (valid while the result stays normal)
    bits = acc.view(np.int32) + (fl.astype(np.int32) << 23)
    out = bits.view(np.float32)
    return np.where(x < -120.0, np.float32(0), out)


if __name__ == "__main__":
    x = -np.abs(np.random.default_rng(0).normal(0, 6, 2_000_000)).astype(np.float32)
    ref = np.exp2(x.astype(np.float64))
    print(f"{'degree':>6s} {'max rel err':>12s} {'vs bf16 ulp/2 (2^-9)':>22s}   coefficients (c0 .. cn)")
    for d in (2, 3, 4, 5):
        c = fit(d)
        got = eval_f32(c, x).astype(np.float64)
        m = ref > 1e-30
        rel = np.max(np.abs(got[m] - ref[m]) / ref[m])
        print(f"{d:6d} {rel:12.3e} {rel / 2 ** -9:22.3f}   " + " ".join(f"{v:.9g}" for v in c))

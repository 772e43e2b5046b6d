"""Round-2 design study (CPU only): how often would a split-bf16 tensor-core PRE-FILTER decide a kmeans label without
the exact fp32 chain?  For each row: approximate scores s~_j = (xh+xl).(ch+cl) - c2_j evaluated in fp32 from bf16 hi/lo parts
(4 products), rigorous bound eps = 2^-13 * |x|_2 * max_j |c_j|_2 (split residuals 2^-16 per operand, fp32 accumulation,
and the exact chain's own rounding, with a safety factor), candidates = {j : s~_j >= max s~ - 2 eps}.  A row needs the
exact chain only if it has more than one candidate."""
import numpy as np


def bf16_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_rn(x):
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def study(N, d, k, seed=0, clustered=False):
    rng = np.random.default_rng(seed)
    if clustered:
        cen0 = rng.normal(size=(k, d)).astype(np.float32)
        x = (cen0[rng.integers(0, k, N)] + 0.7 * rng.normal(size=(N, d))).astype(np.float32)
    else:
        x = rng.normal(size=(N, d)).astype(np.float32)
    c = x[rng.choice(N, k, replace=False)].copy()                      # centroids after a few iterations look like data means
    c = (0.5 * c + 0.5 * x[:k * 50].reshape(k, 50, d).mean(1)).astype(np.float32)
    c2 = 0.5 * (c.astype(np.float64) ** 2).sum(1)
    exact = x.astype(np.float64) @ c.T.astype(np.float64) - c2        # stands in for the fp32 chain (its error is inside eps)
    xh = bf16_trunc(x); xl = bf16_rn(x - xh)
    ch = bf16_trunc(c); cl = bf16_rn(c - ch)
    approx = (xh @ ch.T + xh @ cl.T + xl @ ch.T + xl @ cl.T).astype(np.float32) - c2.astype(np.float32)
    eps = 2.0 ** -13 * np.linalg.norm(x, axis=1) * np.linalg.norm(c, axis=1).max()
    err = np.abs(approx - exact).max(1)
    assert np.all(err <= eps), "the bound must hold"
    m = approx.max(1, keepdims=True)
    ncand = (approx >= m - 2 * eps[:, None]).sum(1)
    lab_ok = np.all(exact.argmax(1)[ncand == 1] == approx.argmax(1)[ncand == 1])
    return (ncand > 1).mean(), err.max() / eps.max(), lab_ok


def adversarial(seed=1):
    """Near-tied centroids, a 1e6 dynamic range between rows, and tiny coordinates: the bound must still hold and every
    label decided by the pre-filter must equal the exact argmax (ties and near-ties simply fall back)."""
    rng = np.random.default_rng(seed)
    N, d, k = 50000, 100, 20
    x = rng.normal(size=(N, d)).astype(np.float32)
    x[:N // 4] *= 1e3; x[N // 4:N // 2] *= 1e-3                       # rows of very different magnitude
    c = rng.normal(size=(k, d)).astype(np.float32)
    c[7] = c[3]; c[11] = c[3] * np.float32(1 + 2 ** -20)               # an exact duplicate and a near-duplicate centroid
    c2 = 0.5 * (c.astype(np.float64) ** 2).sum(1)
    exact = x.astype(np.float64) @ c.T.astype(np.float64) - c2
    xh = bf16_trunc(x); xl = bf16_rn(x - xh); ch = bf16_trunc(c); cl = bf16_rn(c - ch)
    approx = (xh @ ch.T + xh @ cl.T + xl @ ch.T + xl @ cl.T).astype(np.float32) - c2.astype(np.float32)
    eps = 2.0 ** -13 * np.linalg.norm(x, axis=1) * np.linalg.norm(c, axis=1).max() + 2.0 ** -20 * np.abs(c2).max()   # + the c2 subtraction's rounding
    assert np.all(np.abs(approx - exact).max(1) <= eps)
    ncand = (approx >= approx.max(1, keepdims=True) - 2 * eps[:, None]).sum(1)
    decided = ncand == 1
    assert np.all(exact.argmax(1)[decided] == approx.argmax(1)[decided])
    return (~decided).mean()


if __name__ == "__main__":
    print(f"adversarial (duplicate / near-duplicate centroids, 1e6 dynamic range): {100 * adversarial():.2f} % of rows fall back, all decided labels exact")
    for N, d, k, cl in [(200000, 100, 20, False), (200000, 100, 20, True), (200000, 32, 20, False), (100000, 256, 32, False)]:
        frac, tight, ok = study(N, d, k, clustered=cl)
        print(f"N={N} d={d} k={k} clustered={cl}: rows needing the exact chain {100 * frac:.3f} %, max err / eps {tight:.3f}, decided labels correct: {ok}")

import numpy as np

NP = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": np.finfo(np.float32).eps, "d": np.finfo(np.float64).eps,
       "c": np.finfo(np.float32).eps, "z": np.finfo(np.float64).eps}


def rand(rng, t, rows, cols, ld=None, lo=-1.0, hi=1.0, fill=np.nan):
    """Column-major storage array of shape (ld, cols); rows beyond `rows` hold `fill`
    (NaN by default, so any read of padding poisons the result)."""
    ld = max(rows, 1) if ld is None else ld
    a = np.full((ld, max(cols, 0)), fill, dtype=NP[t], order="F")
    v = rng.uniform(lo, hi, (rows, cols))
    if t in "cz":
        v = v + 1j * rng.uniform(lo, hi, (rows, cols))
    a[:rows, :] = v.astype(NP[t])
    return a


def jalla_arbitrary(rng, t, m=None, n=None):
    """The QuickCheck generator of Numeric/Jalla/Test.hs:22-41: shape in [1,100]^2 and
    entries uniform in [-1000, 1000] (both parts for complex)."""
    m = int(rng.integers(1, 101)) if m is None else m
    n = int(rng.integers(1, 101)) if n is None else n
    return rand(rng, t, m, n, lo=-1000.0, hi=1000.0)


def fro(x):
    return float(np.linalg.norm(np.asarray(x, dtype=np.complex128 if np.iscomplexobj(x) else np.float64)))


def op(a, trans):
    return a if trans == 111 else (a.T if trans == 112 else a.conj().T)


def gemm_bound(t, k, A, B, Cref):
    """north_star: relative Frobenius error <= 4*n*eps*|A||B|/|C| (n = contraction length)."""
    c = fro(Cref)
    return 4 * max(k, 1) * EPS[t] * fro(A) * fro(B) / (c if c > 0 else 1.0)

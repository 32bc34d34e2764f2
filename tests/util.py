"""Shared helpers for the parity tests."""
import numpy as np
import scipy.sparse as sp

REL_FRO_TOL = 1e-10          # north_star: relative Frobenius error of the product <= 1e-10
NEAR_THRESHOLD = 1e-6        # pattern may differ only for |v| within this relative band of thr


def rel_fro(a, b):
    a = sp.csc_matrix(a)
    b = sp.csc_matrix(b)
    d = a - b
    den = np.sqrt((abs(b).power(2)).sum())
    num = np.sqrt((abs(d).power(2)).sum())
    return float(num / den) if den > 0 else float(num)


def compare_sparse(got, ref, thr=0.0, tol=REL_FRO_TOL):
    """Values: relative Frobenius error on the common pattern <= tol.
    Pattern: entries present on one side only must have |v| within a tiny band around thr
    (thr == 0: they must be ~0, i.e. rounding-level cancellations)."""
    got = sp.csc_matrix(got)
    ref = sp.csc_matrix(ref)
    assert got.shape == ref.shape
    g = got.copy(); g.data = np.ones_like(g.data, dtype=np.float64)
    r = ref.copy(); r.data = np.ones_like(r.data, dtype=np.float64)
    g = sp.csc_matrix((np.ones(got.nnz), got.indices, got.indptr), shape=got.shape)
    r = sp.csc_matrix((np.ones(ref.nnz), ref.indices, ref.indptr), shape=ref.shape)
    common = g.multiply(r)
    only_g = g - common
    only_r = r - common
    only_g.eliminate_zeros(); only_r.eliminate_zeros()
    scale = max(float(abs(ref).max()) if ref.nnz else 0.0, 1e-300)
    for only, src, name in ((only_g, got, "product"), (only_r, ref, "oracle")):
        if only.nnz:
            vals = abs(np.asarray(src[only.nonzero()]).ravel())
            band = max(thr * NEAR_THRESHOLD, 1e-13 * scale)
            assert np.all(np.abs(vals - thr) <= band + thr * NEAR_THRESHOLD), (
                f"{only.nnz} entries only in the {name} are not near the threshold: "
                f"|v| in [{vals.min():.3e}, {vals.max():.3e}], thr={thr:.3e}")
    gc = got.multiply(common)
    rc = ref.multiply(common)
    err = rel_fro(gc, rc)
    assert err <= tol, f"relative Frobenius error {err:.3e} > {tol:.1e}"
    return err


def banded(n, half_bandwidth=82, seed=20240617, lam=20.0, scale=0.05, dtype=np.float64):
    """Synthetic banded symmetric matrix of SURVEY 8(d): a_ij = s*exp(-|i-j|/lam)/(1+|i-j|) + diag."""
    rng = np.random.default_rng(seed)
    diags, offs = [], []
    for d in range(0, half_bandwidth + 1):
        v = scale * np.exp(-d / lam) / (1.0 + d)
        if d == 0:
            diags.append(rng.uniform(-1.0, 1.0, n)); offs.append(0)
        else:
            diags.append(np.full(n - d, v)); offs.append(d)
            diags.append(np.full(n - d, v)); offs.append(-d)
    return sp.diags(diags, offs, shape=(n, n), format="csc", dtype=dtype)


def random_sparse(n, fill, seed, complex_=False, symmetric=False):
    rng = np.random.default_rng(seed)
    m = sp.random(n, n, fill, random_state=rng, format="csc")
    if complex_:
        m = m + 1j * sp.random(n, n, fill, random_state=rng, format="csc")
    if symmetric:
        m = (m + m.conj().T) * 0.5
    return sp.csc_matrix(m)

"""Parity metrics (test infrastructure).

compare_*: the reference's own printed metrics, test_driver/toolbox.F90:26-176 -- relative L2 error and
max percentage error, the 2-D versions on abs() of the entries to ignore the eigenvector sign/phase.
The gates below are the north_star's: |dlambda_i| < n*eps*||A|| and
||A x - lambda B x|| / (n*eps*||A||*||x||) < 30, plus B-orthogonality ||Z^H B Z - I|| / (n*eps).
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def compare_1d(w1, w2):
    """toolbox.F90:52-74: (relative L2 error, max % error over |w1| >= 1e-10)."""
    w1 = np.asarray(w1)
    w2 = np.asarray(w2)
    rel = np.linalg.norm(w1 - w2) / max(np.linalg.norm(w1), 1e-300)
    mask = np.abs(w1) >= 1e-10
    pct = float(np.max(np.abs((w1[mask] - w2[mask]) / w1[mask]))) * 100 if mask.any() else 0.0
    return float(rel), pct


def compare_2d_abs(z1, z2):
    """toolbox.F90:98-118,147-167: same on abs(entries) (sign / phase invariant for simple eigenvalues)."""
    a1 = np.abs(z1)
    a2 = np.abs(z2)
    rel = np.linalg.norm(a1 - a2) / max(np.linalg.norm(a1), 1e-300)
    mask = a1 >= 1e-10
    pct = float(np.max(np.abs((a1[mask] - a2[mask]) / a1[mask]))) * 100 if mask.any() else 0.0
    return float(rel), pct


def full_from_upper(a):
    u = np.triu(a)
    f = u + np.triu(a, 1).conj().T
    if np.iscomplexobj(f):
        f[np.diag_indices_from(f)] = f.diagonal().real
    return f


def eig_gates(a, b, w, z, w_ref=None):
    """Returns dict of the north_star gates for eigenpairs (w[:m], z[:, :m]) of A x = lambda B x."""
    n = a.shape[0]
    m = z.shape[1]
    an = np.linalg.norm(a, 2) if n <= 2048 else np.linalg.norm(a, 1)
    r = a @ z - (b @ z) * w[:m][None, :]
    xn = np.linalg.norm(z, axis=0)
    res = np.linalg.norm(r, axis=0) / (n * EPS * an * xn)
    g = z.conj().T @ (b @ z)
    orth = np.linalg.norm(g - np.eye(m), "fro") / (n * EPS)
    out = {"residual_max": float(res.max()), "b_orth": float(orth), "normA": float(an)}
    if w_ref is not None:
        out["dlambda_over_gate"] = float(np.max(np.abs(w[:m] - w_ref[:m])) / (n * EPS * an))
    return out


def std_gates(a, w, z):
    """Gates for the standard problem A z = w z (Hermitian A): residual and orthogonality in n*eps units."""
    n = a.shape[0]
    an = max(np.linalg.norm(a, 1), 1e-300)
    r = a @ z - z * w[None, : z.shape[1]]
    res = np.linalg.norm(r, axis=0).max() / (n * EPS * an)
    orth = np.linalg.norm(z.conj().T @ z - np.eye(z.shape[1]), "fro") / (n * EPS)
    return {"residual_max": float(res), "orth": float(orth)}

"""Seeded synthetic data sets of the shapes BASELINE.json names: scalar inputs ~ U(0,1),
functional inputs = smooth random curves f(t) = sum_j a_j / j sin(j pi t + phi_j), output = an fgp_BB3-style
analytic function (R/7_blackBoxFunctions.R:139-150).  Used by bench.py and __graft_entry__.smoke(); no oracle import.
"""
import numpy as np


def smooth_curves(rng, n, k, nharm=8):
    a = rng.uniform(-1.0, 1.0, size=(n, nharm))
    ph = rng.uniform(0.0, 2.0 * np.pi, size=(n, nharm))
    t = np.linspace(0.0, 1.0, k)
    j = np.arange(1, nharm + 1)
    return np.einsum("nj,njk->nk", a / j, np.sin(j[None, :, None] * np.pi * t[None, None, :] + ph[:, :, None]))


def make_data(seed, n, ds, f_lens):
    """-> sIn (n x ds) or None, fIn (list of n x k_i), y (n)"""
    rng = np.random.default_rng(seed)
    sIn = rng.uniform(size=(n, ds)) if ds else None
    fIn = [smooth_curves(rng, n, k) for k in f_lens]
    y = np.zeros(n)
    if ds:
        y += sIn[:, 0] + (2.0 * sIn[:, 1] if ds > 1 else 0.0)
    for i, f in enumerate(fIn):
        t = np.linspace(0.0, 1.0, f.shape[1])
        y += 4.0 * np.mean(t * f, axis=1) if i == 0 else np.mean(f, axis=1)
    return sIn, fIn, y


def fd_points(theta, lower, upper, ndeps=1e-3):
    """The 2d+1 points one finite-difference gradient of stats::optim evaluates (central differences with step
    ndeps clipped at the box; R/3_training_SF.R:125 calls optim without gr): theta, theta +- ndeps e_i."""
    theta = np.asarray(theta, dtype=np.float64)
    d = theta.size
    pts = np.repeat(theta[:, None], 2 * d + 1, axis=1)
    for i in range(d):
        pts[i, 1 + 2 * i] = min(theta[i] + ndeps, upper[i])
        pts[i, 2 + 2 * i] = max(theta[i] - ndeps, lower[i])
    return pts

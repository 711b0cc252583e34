"""Emulation of R's default RNG (Mersenne-Twister + Inversion) -- oracle/test infrastructure.

Follows R's src/main/RNG.c (not vendored under /root/reference; R >= 3.5 per
DESCRIPTION:32).  Used to regenerate the reference's example inputs
(`set.seed(100); runif(...)`, R/1_fgpm_Class.R:199-206) and the presample draw
`runif(n.ls * n.presample)` of R/3_training_SF.R:91.
Verified: set.seed(100); runif(5) -> 0.30776611 0.25767250 0.55232243 0.05638315 0.46854928
(printed in docs/reference/get_active_in.html).
"""
import numpy as np

_I2_32M1 = 2.328306437080797e-10  # R's MT_genrand scale: 2.3283064365386963e-10
_MT_SCALE = 2.3283064365386963e-10


class RRng:
    """R's `set.seed(seed)` / `runif` / `rnorm` for RNGkind("Mersenne-Twister", "Inversion")."""

    def __init__(self, seed):
        self.set_seed(seed)

    def set_seed(self, seed):
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
            words = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(np.uint32(69069) * s + np.uint32(1))
                words[j] = s
        # words[0] is dummy[0] = mti; FixupSeeds forces mti = N = 624
        self._bg = np.random.MT19937()
        st = self._bg.state
        st["state"]["key"] = words[1:].copy()
        st["state"]["pos"] = 624
        self._bg.state = st

    def _unif_raw(self, n):
        u = self._bg.random_raw(n).astype(np.float64) * _MT_SCALE
        # fixup(): keep strictly inside (0, 1)
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return u

    def runif(self, n, lo=0.0, hi=1.0):
        u = self._unif_raw(int(n))
        if lo == 0.0 and hi == 1.0:
            return u
        return lo + (hi - lo) * u

    def rnorm(self, n):
        """INVERSION: u = unif_rand(); u = (int)(BIG*u) + unif_rand(); qnorm5(u/BIG)."""
        from scipy.special import ndtri
        BIG = 134217728.0
        u = self._unif_raw(2 * int(n))
        u1 = u[0::2]
        u2 = u[1::2]
        v = np.floor(BIG * u1) + u2
        return ndtri(v / BIG)


def r_matrix(x, nrow=None, ncol=None):
    """R's matrix(x, nrow, ncol): column-major fill."""
    x = np.asarray(x, dtype=np.float64)
    if nrow is None:
        nrow = x.size // ncol
    if ncol is None:
        ncol = x.size // nrow
    return x.reshape((ncol, nrow)).T.copy()


def expand_grid(*factors):
    """R's expand.grid: first factor varies fastest; returns (prod, nfactors) matrix."""
    grids = np.meshgrid(*factors, indexing="ij")
    cols = [g.reshape(-1, order="F") for g in grids]
    return np.stack(cols, axis=1)

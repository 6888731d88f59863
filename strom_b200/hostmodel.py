"""Host-side model numbers the engine is fed (they stay on the host in the reference too):
GTR eigensystem (reference src/model.hpp:212-265) and discrete-gamma mean rates
(reference src/model.hpp:267-313).  numpy/scipy stand in for Eigen / Boost.Math."""
from __future__ import annotations

import numpy as np


def gtr_eigen(freqs, exchangeabilities):
    """(eigenvectors, inverse eigenvectors, eigenvalues) of the normalised GTR rate matrix, in the
    row-major 4x4 form Model::setBeagleEigenDecomposition uploads; eigenvalues ascending."""
    pi = np.asarray(freqs, dtype=np.float64)
    r = [float(x) for x in exchangeabilities]
    R = np.zeros((4, 4))
    R[np.triu_indices(4, 1)] = r
    R = R + R.T
    Q = R * pi[None, :]
    rows = Q.sum(axis=1)
    scale = 1.0 / float(pi @ rows)
    Q = Q * scale
    np.fill_diagonal(Q, -scale * rows)
    s = np.sqrt(pi)
    sym = (s[:, None] * Q) / s[None, :]
    lam, U = np.linalg.eigh(0.5 * (sym + sym.T))
    return np.ascontiguousarray(U / s[:, None]), np.ascontiguousarray(U.T * s[None, :]), lam


def gamma_rates(shape: float, ncateg: int):
    """Mean rate of each of `ncateg` equal-probability slices of Gamma(shape, 1/shape)."""
    from scipy.special import gammainc, gammaincinv
    probs = np.full(ncateg, 1.0 / ncateg)
    if ncateg == 1:
        return np.ones(1), probs
    cuts = gammaincinv(shape, np.arange(1, ncateg) / ncateg) / shape        # category boundaries
    upper_plus = np.concatenate([gammainc(shape + 1.0, cuts * shape), [1.0]])
    lower_plus = np.concatenate([[0.0], upper_plus[:-1]])
    return (upper_plus - lower_plus) * ncateg, probs


def invariant_gamma_rates(shape: float, ncateg_gamma: int, pinvar: float):
    """GTR+I+G expressed through the category API: one rate-0 category of weight pinvar."""
    r, w = gamma_rates(shape, ncateg_gamma)
    return np.concatenate([[0.0], r / (1.0 - pinvar)]), np.concatenate([[pinvar], w * (1.0 - pinvar)])

"""Liouville (transfer-matrix) representation of unitaries, computed on the GPU.

Mirror of the one function of the reference's ``superoperator.py`` that the hot path calls on every
``cache_control_matrix`` and ``concatenate`` (``pulse_sequence.py:675-677``, ``:1827``, ``:1854``).
"""
import numpy as np

from . import _lib
from . import basis as _basis

__all__ = ['liouville_representation']


def normalize_liouville_columns(liouville, basis):
    """Turn raw traces tr(C_i U C_j U^dagger) into expansion coefficients: column j is divided by
    tr(C_j C_j) when the basis elements are not normalised -- what ``basis.expand(...)`` of the
    reference does for ``basis.isnorm == False`` (``basis.py:650-698``); a no-op for the orthonormal
    Pauli / GGM bases."""
    if getattr(basis, 'isnorm', None) is None:
        basis = np.asarray(basis).view(_basis.Basis)
    if basis.isnorm:
        return liouville
    arr = np.asarray(basis)
    norms = np.einsum('bij,bji->b', arr, arr)
    if not np.iscomplexobj(liouville):
        norms = norms.real
    return liouville/norms


def liouville_representation(U, basis) -> np.ndarray:
    r"""U_ij = tr(C_i U C_j U^\dagger) for U of shape (..., d, d) (reference
    ``superoperator.py:51-84``).  Real for Hermitian bases, like the reference's
    ``basis.expand(..., hermitian=basis.isherm)``."""
    U = np.asarray(U)
    d = U.shape[-1]
    lead = U.shape[:-2]
    Uc = _lib.as_c128(U.reshape(-1, d, d))
    Bc = _lib.as_c128(np.asarray(basis))
    n, n_basis = Uc.shape[0], Bc.shape[0]
    out = np.empty((n, n_basis, n_basis), dtype=np.complex128)
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_liouville_representation(ctx, n, d, n_basis, _lib.ptr(Uc),
                                                            _lib.ptr(Bc), _lib.ptr(out)))
    out = out.reshape(*lead, n_basis, n_basis)
    herm = getattr(basis, 'isherm', None)
    if herm is None:
        herm = np.allclose(Bc, Bc.conj().swapaxes(-1, -2), atol=1e-14, rtol=0)
    return normalize_liouville_columns(np.ascontiguousarray(out.real) if herm else out, basis)

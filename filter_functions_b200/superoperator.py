"""Liouville (transfer-matrix) representation of unitaries, computed on the GPU, and the small host
checks the reference offers for the superoperators the path produces (error transfer matrices).

``liouville_representation`` is what the hot path calls on every ``cache_control_matrix`` and
``concatenate`` (reference ``pulse_sequence.py:675-677``, ``:1827``, ``:1854``); the Choi-matrix
conversion and the (conditional) complete-positivity tests (``superoperator.py:87-256``) are O(d^6) host
functions on single matrices.
"""
import numpy as np

from . import _lib
from . import basis as _basis

__all__ = ['liouville_representation', 'liouville_to_choi', 'liouville_is_CP', 'liouville_is_cCP']


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


def liouville_to_choi(superoperator, basis) -> np.ndarray:
    r"""Choi matrix :math:`\sum_{ij}\mathcal{S}_{ij}\,C_j^T\otimes C_i` of a superoperator given in
    Liouville representation with respect to ``basis`` (shape (..., d^2, d^2), same contract as the
    reference's ``superoperator.liouville_to_choi``)."""
    S = np.asarray(superoperator)
    C = np.asarray(basis)
    # (C_j^T (x) C_i)[(a, c), (b, e)] = C_j[b, a] C_i[c, e]
    right = np.tensordot(S, C, axes=([-1], [0]))                  # (..., i, b, a)
    choi = np.einsum('...iba,ice->...acbe', right, C)
    return choi.reshape(S.shape)


def _hermitian_eig(mat):
    """eigh, falling back to the general eigensolver when LAPACK's divide and conquer gives up."""
    try:
        return np.linalg.eigh(mat)
    except np.linalg.LinAlgError:
        vals, vecs = np.linalg.eig(mat)
        return vals.real, vecs


def liouville_is_CP(superoperator, basis, return_eig: bool = False, atol=None):
    """True where the superoperator is completely positive, i.e. its Choi matrix is positive
    semidefinite within ``atol`` (default: the basis' comparison tolerance); optionally also the
    eigenvalues / eigenvectors of the Choi matrix (reference ``superoperator.py:133-190``)."""
    vals, vecs = _hermitian_eig(liouville_to_choi(superoperator, basis))
    tol = atol or getattr(basis, '_atol', np.finfo(complex).eps*np.shape(basis)[-1]**3)
    ok = (vals >= -tol).all(axis=-1)
    return (ok, (vals, vecs)) if return_eig else ok


def liouville_is_cCP(superoperator, basis, return_eig: bool = False, atol=None):
    """True where the superoperator is CONDITIONALLY completely positive: its Choi matrix projected onto
    the complement of the maximally entangled state is positive semidefinite (generators of CP maps;
    reference ``superoperator.py:193-256``)."""
    S = np.asarray(superoperator)
    n = S.shape[-1]
    d = int(round(np.sqrt(n)))
    omega = np.zeros(n)
    omega[::d + 1] = 1/np.sqrt(d)                                  # sum_i |ii> / sqrt(d)
    proj = np.eye(n) - np.outer(omega, omega)
    vals, vecs = _hermitian_eig(proj @ liouville_to_choi(S, basis) @ proj)
    tol = atol or getattr(basis, '_atol', np.finfo(complex).eps*d**3)
    ok = (vals >= -tol).all(axis=-1)
    return (ok, (vals, vecs)) if return_eig else ok

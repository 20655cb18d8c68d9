"""Operator bases consumed by the engine as dense ``(n_basis, d, d)`` complex128 arrays.

Host-side set-up, executed once per pulse (not a component to accelerate).
What the hot path and its callers need is restated here: the ``Basis`` ndarray subclass with the
predicates the path branches on, the Pauli and generalised Gell-Mann constructors, completion of a
partial basis (``Basis.from_partial`` -- the CNOT example of the reference builds its basis that way)
and basis expansion.  The trace tensor is dense (the reference keeps it in the ``sparse`` package).
"""
import warnings
from functools import cached_property
from itertools import product
from typing import Optional, Sequence

import numpy as np

from . import util

__all__ = ['Basis', 'expand', 'ggm_expand', 'normalize']


class Basis(np.ndarray):
    """ndarray subclass of shape (n_basis, d, d) (reference ``basis.py:58-392``).

    ``Basis(array_like, traceless=None, btype=None, labels=None)`` wraps the given elements as they
    are; :meth:`pauli` and :meth:`ggm` build the two orthonormal Hermitian bases the reference
    ships.
    """

    def __new__(cls, basis_array, traceless: Optional[bool] = None, btype: Optional[str] = None,
                labels: Optional[Sequence[str]] = None) -> 'Basis':
        if not util.is_sequence_like(basis_array):
            raise TypeError('Invalid data type. Must be array_like')
        if isinstance(basis_array, cls):
            elements = basis_array
        else:
            if np.ndim(basis_array) == 2 and hasattr(basis_array, 'shape'):
                basis_array = [basis_array]                  # a single d x d element
            elements = util.parse_operators(basis_array, 'basis_array')
            n, d = elements.shape[0], elements.shape[-1]
            if n > d*d:
                raise ValueError('Given overcomplete set of basis matrices. '
                                 'Not linearly independent.')
        if labels is not None and len(labels) != len(elements):
            raise ValueError(f'Got {len(labels)} basis labels but expected {len(elements)}')
        new = elements.view(cls)
        new._describe(btype or 'Custom', labels)
        return new

    @staticmethod
    def _default_labels(n: int):
        return [f'$C_{{{i}}}$' for i in range(n)]

    def _describe(self, btype, labels) -> None:
        """Metadata of the reference's Basis (``basis.py:162-201``): type, element labels, dimension
        and the tolerances its comparisons use (eps d^3 absolute, no relative part)."""
        self.btype = btype
        self.labels = labels if labels is not None else self._default_labels(len(self))
        self.d = self.shape[-1]
        self._eps = np.finfo(complex).eps
        self._atol, self._rtol = self._eps*self.d**3, 0

    def __array_finalize__(self, parent) -> None:
        # views and slices inherit the parent's description (ndarray subclassing protocol)
        if parent is None:
            return
        shape = getattr(parent, 'shape', ())
        self.btype = getattr(parent, 'btype', 'Custom')
        self.labels = getattr(parent, 'labels', None)
        if self.labels is None:
            self.labels = self._default_labels(shape[0] if shape else 0)
        self.d = getattr(parent, 'd', shape[-1] if shape else 0)
        self._eps = np.finfo(complex).eps
        self._atol, self._rtol = self._eps*self.d**3, 0

    def __eq__(self, other) -> bool:
        try:
            if self.shape != other.shape:
                return False
        except AttributeError:
            return np.equal(self, other)
        return np.allclose(self.view(np.ndarray), np.asarray(other), atol=self._atol,
                           rtol=self._rtol)

    __hash__ = None

    def __contains__(self, item) -> bool:
        return any(np.isclose(np.asarray(item), self.view(np.ndarray), rtol=self._rtol,
                              atol=self._atol).all(axis=(1, 2)))

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        if arr.ndim == 0:
            return arr[()]
        return super().__array_wrap__(arr, context, return_scalar)

    @cached_property
    def isherm(self) -> bool:
        return self.H == self

    @cached_property
    def isnorm(self) -> bool:
        return normalize(self) == self

    @cached_property
    def isorthogonal(self) -> bool:
        if self.ndim == 2 or len(self) == 1:
            return True
        flat = self.view(np.ndarray).reshape(len(self), -1)
        gram = flat.conj() @ flat.T
        off = gram[~np.identity(len(self), dtype=bool)]
        return np.allclose(off, 0, atol=self._eps*(self.d**2)**3, rtol=self._rtol)

    @cached_property
    def isorthonorm(self) -> bool:
        return self.isorthogonal and self.isnorm

    @cached_property
    def istraceless(self) -> bool:
        """True if all elements are traceless except possibly one proportional to the identity."""
        arr = self.view(np.ndarray)
        trace = np.einsum('...jj', arr)
        trace = np.where(np.abs(trace) <= self._eps*self.d**2, 0, trace)
        nonzero = np.atleast_1d(trace).nonzero()[0]
        if nonzero.size == 0:
            return True
        if nonzero.size == 1:
            elem = arr[nonzero[0]] if arr.ndim == 3 else arr
            offdiag = elem[~np.eye(self.d, dtype=bool)]
            return bool((np.diag(elem) == elem[0, 0]).all() and not offdiag.any())
        return False

    @cached_property
    def iscomplete(self) -> bool:
        flat = self.view(np.ndarray).reshape(self.shape[0], -1)
        return np.linalg.matrix_rank(flat) == self.d**2

    @property
    def H(self) -> 'Basis':
        return self.T.conj()

    @property
    def T(self) -> 'Basis':
        return self.swapaxes(-1, -2) if self.ndim >= 2 else self

    @cached_property
    def four_element_traces(self) -> np.ndarray:
        """T_ijkl = tr(C_i C_j C_k C_l) as a dense array (reference ``basis.py:330-348`` keeps it
        sparse; dense is n_basis^4 * 16 B, fine for d <= 4).  The array answers ``todense()`` like the
        reference's sparse tensor does, so code written against either works."""
        arr = self.view(np.ndarray)
        pair = np.einsum('iab,jbc->ijac', arr, arr)
        return np.einsum('ijac,klca->ijkl', pair, pair).view(_DenseTensor)

    @cached_property
    def sparse(self) -> np.ndarray:
        """The elements as an array that answers ``todense()`` (the reference returns a ``sparse.COO``
        here, ``basis.py:325-328``; this package keeps bases dense)."""
        return self.view(np.ndarray).view(_DenseTensor)

    def tidyup(self, eps_scale: Optional[float] = None) -> None:
        """Set entries below machine precision (times ``eps_scale``, default d^3) to zero, in place."""
        arr = self.view(np.ndarray)
        atol = self._eps*(self.d**3 if eps_scale is None else eps_scale)
        arr.real[np.abs(arr.real) <= atol] = 0
        arr.imag[np.abs(arr.imag) <= atol] = 0
        self._invalidate_cached_properties()

    def _print_checks(self) -> None:
        """Debug aid of the reference's Basis (``basis.py:234-238``): the predicates, one per line."""
        for name in ('isherm', 'istraceless', 'iscomplete', 'isorthonorm'):
            print(f'{name} :\t {getattr(self, name)}')

    def _invalidate_cached_properties(self) -> None:
        """Forget the predicates computed for the previous contents (in-place changes)."""
        for name in ('isherm', 'isnorm', 'isorthogonal', 'isorthonorm', 'istraceless', 'iscomplete',
                     'four_element_traces', 'sparse'):
            self.__dict__.pop(name, None)

    def normalize(self, copy: bool = False):
        if copy:
            return normalize(self)
        self /= _norm(self)
        self._invalidate_cached_properties()
        return None

    def expand(self, M, hermitian: bool = False, traceless: bool = False, tidyup: bool = False):
        if self.btype == 'GGM' and self.iscomplete:
            return ggm_expand(M, traceless, hermitian, tidyup)
        return expand(M, self, self.isnorm, hermitian, tidyup)

    @classmethod
    def from_partial(cls, partial_basis_array, traceless: Optional[bool] = None,
                     btype: Optional[str] = None, labels: Optional[Sequence[str]] = None) -> 'Basis':
        """A complete orthonormal basis that contains the given (orthogonal) elements, same contract as
        the reference's ``Basis.from_partial`` (``basis.py:492-620``): the elements are normalised and
        kept in order, the rest of the operator space is filled with an orthonormal set spanning the
        orthogonal complement; ``traceless=True`` (default: if the given elements allow it) puts the
        identity first and makes every other element traceless.  Raises ``ValueError`` for elements that
        are not orthogonal, not traceless although a traceless basis was requested, or for a wrong
        number of labels (len(elements) or d^2; new elements are labelled ``$C_{i}$``)."""
        if labels is None:
            own = getattr(partial_basis_array, 'labels', None)
            if own is not None and len(own) == len(partial_basis_array):
                labels = own
        elems = cls(partial_basis_array)
        elems.normalize(copy=False)
        if not elems.isherm:
            warnings.warn("(Some) elems not hermitian! The resulting basis also won't be.")
        if not elems.isorthogonal:
            raise ValueError('The basis elements are not orthogonal!')
        if traceless is None:
            traceless = elems.istraceless
        elif traceless and not elems.istraceless:
            raise ValueError('The basis elements are not traceless (up to an identity element) '
                             'but a traceless basis was requested!')
        d = elems.d
        if labels is not None and len(labels) not in (len(elems), d*d):
            raise ValueError(f'Got {len(labels)} labels but expected {len(elems)} or {d*d}')

        # coordinates of the elements in the (orthonormal, Hermitian) GGM basis; a traceless basis keeps the
        # identity apart: its coordinate is dropped, and with it an element that WAS the identity
        frame = cls.ggm(d).view(np.ndarray)
        coords = ggm_expand(elems.view(np.ndarray), traceless=traceless, hermitian=elems.isherm,
                            tidyup=True)
        if traceless:
            identity, frame, coords = frame[:1], frame[1:], coords[..., 1:]
        coords = coords[np.abs(coords).sum(axis=-1) != 0]
        if coords.size:
            # rows of V^H beyond the rank span the orthogonal complement of the given coordinates
            _, sing, vh = np.linalg.svd(coords, full_matrices=True)
            rank = int((sing > sing.max()*max(coords.shape)*np.finfo(float).eps).sum())
            full = np.concatenate((coords, vh[rank:].conj()))
            new = np.einsum('ij,jkl->ikl', full, frame)
        else:
            new = frame
        if traceless:
            new = np.concatenate((identity, new))
        new = new.view(cls)
        new.tidyup()
        if labels is not None and len(labels) == len(elems):
            labels = list(labels)
            if traceless:   # the identity's label moves to the front with it
                where = next((i for i, e in enumerate(elems.view(np.ndarray))
                              if np.allclose(e, identity[0], rtol=elems._rtol, atol=elems._atol)), 0)
                labels.insert(0, labels.pop(where))
            labels += [f'$C_{{{i}}}$' for i in range(len(labels), len(new))]
        return cls(new, btype=btype or 'From partial', labels=labels)

    @classmethod
    def pauli(cls, n: int) -> 'Basis':
        """Normalised n-qubit Pauli basis {I,X,Y,Z}^n / sqrt(2^n) (reference ``basis.py:393-426``)."""
        elems = util.paulis
        for _ in range(n - 1):
            elems = np.einsum('aij,bkl->abikjl', elems, util.paulis).reshape(
                len(elems)*4, elems.shape[1]*2, elems.shape[2]*2)
        elems = elems/np.sqrt(2**n)
        labels = [''.join(tup) for tup in product(['I', 'X', 'Y', 'Z'], repeat=n)]
        return cls(elems, btype='Pauli', labels=labels)

    @classmethod
    def ggm(cls, d: int) -> 'Basis':
        """Normalised generalised Gell-Mann basis: identity, symmetric, antisymmetric, diagonal
        elements in this order (reference ``basis.py:428-489``)."""
        out = np.zeros((d*d, d, d), dtype=complex)
        out[0] = np.eye(d)/np.sqrt(d)
        pairs = [(j, k) for j in range(d) for k in range(j + 1, d)]
        n_sym = len(pairs)
        for i, (j, k) in enumerate(pairs, start=1):
            out[i, j, k] = out[i, k, j] = 1/np.sqrt(2)
            out[i + n_sym, j, k] = -1j/np.sqrt(2)
            out[i + n_sym, k, j] = 1j/np.sqrt(2)
        for l in range(1, d):
            diag = np.zeros(d)
            diag[:l] = 1
            diag[l] = -l
            out[2*n_sym + l] = np.diag(diag/np.sqrt(l*(l + 1)))
        return cls(out, btype='GGM', labels=[rf'$\Lambda_{{{i}}}$' for i in range(d*d)])


class _DenseTensor(np.ndarray):
    """Plain ndarray that also answers ``todense()`` (drop-in for the reference's sparse trace tensor)."""

    def todense(self) -> np.ndarray:
        return self.view(np.ndarray)


def _norm(b) -> np.ndarray:
    b = np.asarray(b)
    return np.linalg.norm(b, axis=(-1, -2))[..., None, None]


def normalize(b: Basis) -> Basis:
    """Copy of ``b`` normalised to unit Frobenius norm (reference ``basis.py:625-647``)."""
    arr = np.asarray(b)
    return (arr/_norm(arr)).view(Basis)


def _tidy(arr):
    eps = np.finfo(arr.dtype).eps*(arr.shape[-1] if arr.ndim else 1)
    if np.iscomplexobj(arr):
        arr.real[np.abs(arr.real) <= eps] = 0
        arr.imag[np.abs(arr.imag) <= eps] = 0
    else:
        arr[np.abs(arr) <= eps] = 0
    return arr


def expand(M, basis, normalized: bool = True, hermitian: bool = False, tidyup: bool = False):
    """Coefficients c_j = tr(M C_j) / tr(C_j^dagger C_j) (reference ``basis.py:650-698``)."""
    M = np.asarray(M)
    arr = np.asarray(basis)
    real = hermitian and bool(getattr(basis, 'isherm', False))
    coeffs = np.tensordot(M, arr, axes=[(-2, -1), (-1, -2)])
    if real:
        coeffs = coeffs.real
    if not normalized:
        norms = np.einsum('bij,bji->b', arr, arr)
        coeffs = coeffs/(norms.real if real else norms)
    return _tidy(coeffs) if tidyup else coeffs


def ggm_expand(M, traceless: bool = False, hermitian: bool = False, tidyup: bool = False):
    """Expansion in the GGM basis from its construction rule (reference ``basis.py:701-787``)."""
    M = np.asarray(M)
    if M.shape[-1] != M.shape[-2]:
        raise ValueError('M should be square in its last two axes')
    square = M.ndim < 3
    if square:
        M = M[None]
    d = M.shape[-1]
    pairs = [(j, k) for j in range(d) for k in range(j + 1, d)]
    n_sym = len(pairs)
    coeffs = np.zeros((*M.shape[:-2], d*d), dtype=float if hermitian else complex)

    def cast(x):
        return x.real if hermitian else x

    if not traceless:
        coeffs[..., 0] = cast(np.trace(M, axis1=-2, axis2=-1))/np.sqrt(d)
    if pairs:
        js, ks = (np.array(ix) for ix in zip(*pairs))
        upper, lower = M[..., js, ks], M[..., ks, js]
        coeffs[..., 1:n_sym + 1] = cast(upper + lower)/np.sqrt(2)
        coeffs[..., n_sym + 1:2*n_sym + 1] = cast(1j*(upper - lower))/np.sqrt(2)
    diag = np.diagonal(M, axis1=-2, axis2=-1)
    for l in range(1, d):
        coeffs[..., 2*n_sym + l] = cast(diag[..., :l].sum(axis=-1) - l*diag[..., l])/np.sqrt(l*(l + 1))
    if square:
        coeffs = coeffs[0]
    return _tidy(coeffs) if tidyup else coeffs

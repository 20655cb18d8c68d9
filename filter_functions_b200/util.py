"""Host-side helpers of the hot path (argument parsing, frequency grids, small matrix products).

Mirrors the part of the reference's ``util.py`` that the control-matrix / filter-function /
infidelity path touches; names, argument meaning and error behaviour are the reference's.  Heavy
arithmetic is not done here: :func:`cexp` runs on the GPU, the rest is O(d^2) glue.
"""
import functools
import inspect
from typing import Callable, Sequence

import numpy as np
from numpy import ndarray

from . import _lib

__all__ = ['paulis', 'abs2', 'cexp', 'cexpm1', 'get_sample_frequencies', 'mdot', 'adot', 'integrate',
           'parse_optional_parameters', 'parse_spectrum', 'parse_operators',
           'get_indices_from_identifiers', 'is_sequence_like', 'hash_array_along_axis',
           'all_array_equal', 'CalculationError', 'dot_HS', 'oper_equiv', 'tensor',
           'remove_float_errors', 'progressbar', 'progressbar_range']

#: Pauli matrices I, X, Y, Z (reference ``util.py:109-118``)
paulis = np.array([[[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]],
                  dtype=complex)


class CalculationError(Exception):
    """Indicates a quantity could not be computed (reference ``util.py:1146``)."""


def abs2(x):
    """|x|^2 without the square root (reference ``util.py:121-133``)."""
    x = np.asarray(x)
    return x.real**2 + x.imag**2


def _elementwise_on_device(x, out, where, kernel):
    """Shared shell of :func:`cexp` / :func:`cexpm1`: evaluate on the GPU, then honour ``out`` and the
    ufunc-style ``where`` mask (entries where the mask is False keep what ``out`` held)."""
    x = np.asarray(x, dtype=float)
    flat = _lib.as_f64(x.ravel())
    res = np.empty(flat.shape, dtype=np.complex128)
    if flat.size:
        ctx = _lib.context()
        _lib.check(ctx, kernel(ctx, flat, res))
    res = res.reshape(x.shape)
    if where is True and out is None:
        return res
    if out is None:
        out = np.empty(x.shape, dtype=np.complex128)
    np.copyto(out, res, where=np.broadcast_to(where, x.shape) if where is not True else True)
    return out


def cexp(x, out=None, where=True):
    """exp(i x) on the GPU (reference ``util.py:136-162``)."""
    return _elementwise_on_device(x, out, where, lambda ctx, flat, res: _lib.lib().ffb_cexp(
        ctx, flat.size, _lib.ptr(flat), 1.0, _lib.ptr(res)))


def cexpm1(x, out=None, where=True):
    r"""exp(i x) - 1 = -2 sin^2(x/2) + i sin(x) on the GPU, without the cancellation of
    ``cexp(x) - 1`` at small x (reference ``util.py:165-182``)."""
    return _elementwise_on_device(x, out, where, lambda ctx, flat, res: _lib.lib().ffb_cexpm1(
        ctx, flat.size, _lib.ptr(flat), _lib.ptr(res)))


def integrate(f, x=None, dx=1.0):
    """Trapezoid rule over the last axis (reference ``util.py:880-906``).  Host version for tiny
    arrays such as the smallness parameter; the infidelity integral itself runs on the GPU."""
    f = np.asarray(f)
    dx = np.diff(x) if x is not None else dx
    ret = f[..., 1:] + f[..., :-1]
    ret = ret*dx
    return ret.sum(axis=-1)/2


def parse_optional_parameters(**allowed_kwargs: Sequence) -> Callable:
    """Decorator validating keyword-like parameters against a set of legal values and raising
    ``ValueError`` otherwise (behaviour of reference ``util.py:185-211``)."""
    def decorator(func):
        params = tuple(inspect.signature(func).parameters)
        defaults = {k: v.default for k, v in inspect.signature(func).parameters.items()}

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            for name, allowed in allowed_kwargs.items():
                pos = params.index(name)
                value = args[pos] if pos < len(args) else kwargs.get(name, defaults[name])
                if value not in allowed:
                    raise ValueError(f"Invalid value for {name}: {value}. "
                                     + f"Should be one of {allowed}.")
            return func(*args, **kwargs)
        return wrapper
    return decorator


def parse_spectrum(spectrum, omega, idx) -> ndarray:
    """Broadcast and validate a noise spectrum (reference ``util.py:214-227``)."""
    spectrum = np.asarray(spectrum)
    shape = (len(idx),)*(spectrum.ndim - 1) + (len(omega),)
    try:
        spectrum = np.broadcast_to(spectrum, shape)
    except ValueError as err:
        raise ValueError(f'Spectrum should be of shape {shape}, not {spectrum.shape}.') from err
    if spectrum.ndim == 3 and not np.allclose(spectrum, spectrum.conj().swapaxes(0, 1)):
        raise ValueError('Cross-spectra given but not Hermitian along first two axes')
    elif spectrum.ndim > 3:
        raise ValueError(f'Expected spectrum to have < 4 dimensions, not {spectrum.ndim}')
    return spectrum


def is_sequence_like(obj) -> bool:
    return hasattr(obj, '__len__') and hasattr(obj, '__getitem__')


def parse_operators(opers, err_loc: str) -> ndarray:
    """Convert a sequence of operators to a (n, d, d) complex array (reference ``util.py:230-276``;
    accepts ndarrays and anything exposing ``full()`` (QuTiP), ``to_array()``, ``todense()`` (sparse) or
    ``data`` + ``dexp`` (qopt dense operators))."""
    parsed = []
    for oper in opers:
        if isinstance(oper, ndarray):
            parsed.append(oper.squeeze())
        elif hasattr(oper, 'full'):
            parsed.append(oper.full())
        elif hasattr(oper, 'to_array'):
            parsed.append(oper.to_array())
        elif hasattr(oper, 'todense'):
            parsed.append(oper.todense())
        elif hasattr(oper, 'data') and hasattr(oper, 'dexp'):
            parsed.append(oper.data)
        else:
            raise TypeError(f'Expected operators in {err_loc} to be NumPy arrays or QuTiP Qobjs!')
    parsed = np.asarray(parsed, dtype=complex)
    if parsed.ndim > 3:
        raise ValueError(f'Expected operators in {err_loc} to be two-dimensional!')
    if len(set(parsed.shape[-2:])) != 1:
        raise ValueError(f'Expected operators in {err_loc} to be square!')
    return parsed


def get_indices_from_identifiers(all_identifiers, identifiers) -> ndarray:
    """Indices of ``identifiers`` within ``all_identifiers`` (reference ``util.py:331-357``)."""
    table = {identifier: index for index, identifier in enumerate(all_identifiers)}
    if identifiers is None:
        return np.arange(len(all_identifiers))
    try:
        if isinstance(identifiers, str):
            return np.array([table[identifiers]])
        return np.array([table[identifier] for identifier in identifiers])
    except KeyError:
        raise ValueError('Invalid identifiers given. All available ones '
                         + f'are: {all_identifiers}')


def mdot(arr, axis: int = 0) -> ndarray:
    """arr[0] @ arr[1] @ ... along ``axis`` (reference ``util.py:862-865``); O(n d^3) host glue."""
    return functools.reduce(np.matmul, np.swapaxes(arr, 0, axis))


def adot(arr, axis: int = 0) -> ndarray:
    """Running products result[i] = arr[i] @ ... @ arr[0] (reference ``util.py:868-877``)."""
    arr = np.swapaxes(np.asarray(arr), 0, axis)
    out = np.empty_like(arr)
    out[0] = arr[0]
    for i in range(1, len(arr)):
        out[i] = arr[i] @ out[i - 1]
    return out.swapaxes(0, axis)


@parse_optional_parameters(spacing=('log', 'linear'))
def get_sample_frequencies(pulse, n_samples: int = 300, spacing: str = 'log',
                           include_quasistatic: bool = False, omega_min=None, omega_max=None):
    """Default frequency grid 2 pi 1e-2/tau ... 2 pi 10/dt_min (reference ``util.py:1054-1093``)."""
    xspace = np.geomspace if spacing == 'log' else np.linspace
    omega_min = 2*np.pi*1e-2/pulse.tau if omega_min is None else omega_min
    omega_max = 2*np.pi*1e+1/pulse.dt.min() if omega_max is None else omega_max
    omega = xspace(omega_min, omega_max, n_samples - include_quasistatic)
    if include_quasistatic:
        return np.insert(omega, 0, 0)
    return omega


def dot_HS(U, V, eps=None):
    """Hilbert-Schmidt inner product tr(U^dagger V), rounded to the precision ``eps`` (default: the
    rounding error of the product; reference ``util.py:1003-1051``)."""
    U, V = np.asarray(U), np.asarray(V)
    if eps is None:
        try:
            eps = np.finfo(U.dtype).eps*np.prod(U.shape)*V.shape[-1]*2
        except ValueError:
            eps = 0
    res = np.einsum('...ij,...ij', U.conj(), V)
    if eps != 0:
        res = np.around(res, decimals=abs(int(np.log10(eps))))
    return res if np.imag(res).any() else np.real(res)


def oper_equiv(psi, phi, eps=None, normalized: bool = False):
    """Are ``psi`` and ``phi`` equal up to a global phase?  Returns ``(bool, phase)`` with
    ``phi = exp(i phase) psi`` (reference ``util.py:941-1000``)."""
    psi, phi = np.atleast_2d(np.asarray(psi), np.asarray(phi))
    if eps is None:
        eps = (max(np.finfo(psi.dtype).eps, np.finfo(phi.dtype).eps)
               * np.prod(psi.shape)*phi.shape[-1]*2)
        if not normalized:
            eps *= (np.prod(psi.shape[-2:])*phi.shape[-1]*2)**2
    try:
        inner_product = dot_HS(psi, phi, eps=0)
    except ValueError as err:
        raise ValueError('psi and phi have incompatible dimensions!') from err
    norm = 1 if normalized else np.sqrt(dot_HS(psi, psi, eps=0)*dot_HS(phi, phi, eps=0))
    return abs(norm - abs(inner_product)) <= eps, np.angle(inner_product)


def hash_array_along_axis(arr, axis: int = 0):
    """Hashes of the sub-arrays along ``axis``; -0.0 is normalised first (``util.py:1096-1100``)."""
    return [hash((a + 0.0).tobytes()) for a in np.swapaxes(arr, 0, axis)]


def all_array_equal(it) -> bool:
    """True if all arrays of the iterable are byte-identical (``util.py:1103-1109``)."""
    return len(set(hash(np.asarray(i).tobytes()) for i in it)) == 1


def progressbar_range(*args, show_progressbar: bool = False, **kwargs):
    """The engine computes all segments in one launch; the flag is accepted and ignored."""
    return range(*args)


def progressbar(iterable, *args, **kwargs):
    """The reference wraps loops in ``tqdm`` (``util.py:1112-1143``); here nothing loops on the host long
    enough to need one, the iterable is returned unchanged."""
    return iterable


def tensor(*args, rank: int = 2, optimize=False) -> ndarray:
    """Tensor (Kronecker) product over the last ``rank`` axes of the operands, broadcast over the leading
    axes (same call signature and shape rules as the reference's ``util.tensor``, ``util.py:360-470``;
    a 1-d operand of a rank-2 product is a row vector).  Host helper used to set up multi-qubit operators;
    ``optimize`` is accepted for compatibility (the product is formed pairwise by broadcasting).
    """
    if not args:
        raise TypeError('tensor() needs at least one operand')

    def pair(a, b):
        a_nd, b_nd = a.ndim, b.ndim
        if a_nd < rank:
            a = a.reshape((1,)*(rank - a_nd) + a.shape)
        if b_nd < rank:
            b = b.reshape((1,)*(rank - b_nd) + b.shape)
        try:
            lead = np.broadcast_shapes(a.shape[:-rank], b.shape[:-rank])
        except ValueError:
            raise ValueError(f'Incompatible shapes {a.shape} and {b.shape} for tensor product of rank '
                             f'{rank}.') from None
        ta, tb = a.shape[a.ndim - rank:], b.shape[b.ndim - rank:]
        # a[..., i1, 1, i2, 1, ...] * b[..., 1, j1, 1, j2, ...] -> (..., i1 j1, i2 j2, ...)
        a = np.broadcast_to(a, lead + ta).reshape(lead + tuple(x for n in ta for x in (n, 1)))
        b = np.broadcast_to(b, lead + tb).reshape(lead + tuple(x for n in tb for x in (1, n)))
        return (a*b).reshape(lead + tuple(m*n for m, n in zip(ta, tb)))

    return functools.reduce(pair, (np.asanyarray(a) for a in args))


def remove_float_errors(arr: ndarray, eps_scale=None) -> ndarray:
    """Set entries (real and imaginary parts separately) that are below the dtype's machine precision
    times ``eps_scale`` (default: the length of the last axis) to exactly zero, in place
    (``util.py:909-938``).  Meant for arrays of norm ~ 1."""
    arr = np.asanyarray(arr)
    scale = eps_scale if eps_scale is not None else (arr.shape[-1] if arr.ndim else 1)
    atol = np.finfo(arr.dtype).eps*scale
    if np.iscomplexobj(arr):
        if arr.ndim:
            arr.real[np.abs(arr.real) <= atol] = 0
            arr.imag[np.abs(arr.imag) <= atol] = 0
        else:
            arr = arr.dtype.type(complex(0 if abs(arr.real) <= atol else arr.real,
                                         0 if abs(arr.imag) <= atol else arr.imag))
    elif arr.ndim:
        arr[np.abs(arr) <= atol] = 0
    elif abs(arr) <= atol:
        arr = arr.dtype.type(0)
    return arr

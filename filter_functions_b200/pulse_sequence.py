"""``PulseSequence`` and concatenation: the reference's object / cache / sequencing layer for the
control-matrix -> filter-function -> infidelity path, with every numerical body running on the GPU.

Public names, signatures, cache keys (``_data``, ``_frequency_data``, ``_intermediates``), alias
lookup, invalidation rules and exception types follow ``pulse_sequence.py`` of the reference
(v1.2.1; ``:61-1267`` for the class, ``:1340-1483`` / ``:1599-1887`` for concatenation).  Out of scope
here: ``extend``/``remap``, second-order filter functions and derivatives.
"""
import bisect
import copy
import ctypes
from itertools import accumulate, chain, compress, zip_longest
from types import MappingProxyType
from typing import Any, Iterable, Optional

import numpy as np
from numpy import ndarray

from . import _lib, numeric, util
from .basis import Basis
from .superoperator import liouville_representation, normalize_liouville_columns

__all__ = ['PulseSequence', 'SequenceBatch', 'concatenate', 'concatenate_many',
           'concatenate_periodic', 'concatenate_without_filter_function']

_DATA_ALIASES = {
    'eigenvalues': 'eigvals',
    'eigenvectors': 'eigvecs',
    'propagators': 'propagators',
    'total propagator': 'total_propagator',
    'total propagator liouville': 'total_propagator_liouville',
}
_FREQUENCY_DATA_ALIASES = {
    'frequencies': 'omega',
    'total phases': 'total_phases',
    'filter function': 'filter_function',
    'fidelity filter function': 'filter_function',
    'generalized filter function': 'filter_function_gen',
    'pulse correlation filter function': 'filter_function_pc',
    'fidelity pulse correlation filter function': 'filter_function_pc',
    'generalized pulse correlation filter function': 'filter_function_pc_gen',
    'second order filter function': 'filter_function_2',
    'control matrix': 'control_matrix',
    'pulse correlation control matrix': 'control_matrix_pc',
}


def _unpack_hamiltonian(H, n_dt: int, name: str):
    """``[[operator, coefficients, identifier?], ...]`` -> sorted (opers, identifiers, coeffs);
    validation and default identifiers as the reference's ``_parse_hamiltonian`` (``:1288-1337``)."""
    if not util.is_sequence_like(H):
        raise TypeError(f'Expected {name} to be a sequence, not of type {type(H)}!')
    if not all(util.is_sequence_like(item) for item in H):
        raise TypeError(f'Expected {name} to be a sequence of sequences but found at least one '
                        'item of H not a sequence!')
    opers, coeffs, *rest = zip_longest(*H, fillvalue=None)
    identifiers = list(rest[0]) if rest else None
    if not all(util.is_sequence_like(coeff) for coeff in coeffs):
        raise TypeError(f'Expected coefficients in {name} to be a sequence')
    prefix = 'A' if name == 'H_c' else 'B'
    if identifiers is None:
        identifiers = [f'{prefix}_{i}' for i in range(len(opers))]
    else:
        identifiers = [f'{prefix}_{i}' if ident is None else ident
                       for i, ident in enumerate(identifiers)]
        if len(set(identifiers)) != len(identifiers):
            raise ValueError(f'{name} identifiers should be unique')
    if not all(len(coeff) == n_dt for coeff in coeffs):
        raise ValueError(f'Expected all coefficients in {name} to be of len(dt) = {n_dt}!')
    order = np.argsort(identifiers)
    opers = util.parse_operators(opers, name)
    return opers[order], np.asarray(identifiers)[order], np.asarray(coeffs)[order]


def _merge_equal_segments(pulse):
    """Join consecutive segments with identical control coefficients (for ``__eq__``)."""
    same = (np.diff(pulse.c_coeffs) == 0).all(axis=0).nonzero()[0]
    if not same.size:
        return pulse.c_coeffs, pulse.n_coeffs, pulse.dt
    c_coeffs = np.delete(pulse.c_coeffs, same, axis=1)
    n_coeffs = np.delete(pulse.n_coeffs, same, axis=1)
    dt = np.delete(pulse.dt, same)
    for old, new in zip(same, same - np.arange(len(same))):
        dt[new] += pulse.dt[old]
    return c_coeffs, n_coeffs, dt


class _CacheDict(dict):
    """A cache dictionary of a ``PulseSequence`` that counts its mutations in a counter it shares with
    the pulse (``pulse._stamp``).  Whatever ``concatenate`` remembers about a set of gate pulses is valid
    exactly as long as their stamps have not moved."""
    __slots__ = ('_counter',)

    def __init__(self, counter, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._counter = counter

    def __setitem__(self, key, value):
        self._counter[0] += 1
        dict.__setitem__(self, key, value)

    def __delitem__(self, key):
        self._counter[0] += 1
        dict.__delitem__(self, key)

    def update(self, *args, **kwargs):
        self._counter[0] += 1
        dict.update(self, *args, **kwargs)

    def pop(self, *args):
        self._counter[0] += 1
        return dict.pop(self, *args)

    def popitem(self):
        self._counter[0] += 1
        return dict.popitem(self)

    def clear(self):
        self._counter[0] += 1
        dict.clear(self)

    def setdefault(self, key, default=None):
        self._counter[0] += 1
        return dict.setdefault(self, key, default)

    def __reduce__(self):       # pickles as a plain dict; PulseSequence restores the shared counter
        return (dict, (dict(self),))


_CACHE_NAMES = ('_data', '_frequency_data', '_intermediates')


class PulseSequence:
    r"""A piecewise-constant control pulse with its noise operators.

    ``PulseSequence(H_c, H_n, dt, basis=None)`` with ``H_c = [[A_i, a_i(t), 'id'], ...]`` and
    ``H_n = [[B_j, s_j(t), 'id'], ...]`` exactly as in the reference (``pulse_sequence.py:61-357``).
    Cached quantities live in three dictionaries with the reference's keys; all of them are NumPy
    arrays.
    """

    # A PulseSequence has __len__ / __getitem__ (slicing into segments), so np.array([pulse, ...]) would
    # unpack it into single-segment pulses; the array interface declares it a 0-d object instead, as the
    # reference does (``pulse_sequence.py:241-251``): np.array(list_of_pulses) is a 1-d object array.
    __array_interface__ = {'shape': (), 'typestr': '|O', 'version': 3}

    def __new__(cls, *args, **kwargs):
        new = super().__new__(cls)
        stamp = new.__dict__['_stamp'] = [0]     # bumped by every attribute / cache change
        for name in _CACHE_NAMES:
            new.__dict__[name] = _CacheDict(stamp)
        return new

    def __setattr__(self, name, value):
        if name in _CACHE_NAMES and not isinstance(value, _CacheDict):
            value = _CacheDict(self.__dict__['_stamp'], value)
        object.__setattr__(self, name, value)
        self.__dict__['_stamp'][0] += 1

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items()
                 if k not in ('_stamp', '_computed_control_matrix', '_oper_hash_cache')}
        state['_control_matrix_was_computed'] = self._control_matrix_is_computed()
        for name in _CACHE_NAMES:
            state[name] = dict(state[name])
        return state

    def __setstate__(self, state):
        computed = state.pop('_control_matrix_was_computed', False)
        stamp = self.__dict__['_stamp']
        for name in _CACHE_NAMES:
            self.__dict__[name] = _CacheDict(stamp, state.pop(name, {}))
        self.__dict__.update(state)
        if computed and 'control_matrix' in self._frequency_data:
            self._mark_computed(self._frequency_data['control_matrix'])

    def __init__(self, H_c, H_n, dt, basis: Optional[Basis] = None):
        if not util.is_sequence_like(dt):
            raise TypeError(f'Expected a sequence of time steps, not {type(dt)}')
        self.dt = np.asarray(dt)
        if not np.isreal(self.dt).all():
            raise ValueError('Times dt are not (all) real!')
        if (self.dt < 0).any():
            raise ValueError('Time steps are not (all) positive!')
        self.c_opers, self.c_oper_identifiers, self.c_coeffs = _unpack_hamiltonian(
            H_c, len(self.dt), 'H_c')
        self.n_opers, self.n_oper_identifiers, self.n_coeffs = _unpack_hamiltonian(
            H_n, len(self.dt), 'H_n')
        if self.c_opers.shape[-2:] != self.n_opers.shape[-2:]:
            raise ValueError('Control and noise Hamiltonian not same dimension!')
        self.d = self.c_opers.shape[-1]
        if basis is None:
            self.basis = Basis.ggm(self.d)
        else:
            if not isinstance(basis, Basis):
                raise ValueError("Expected basis to be an instance of the "
                                 + f"'filter_functions.basis.Basis' class, not {type(basis)}!")
            if basis.shape[1:] != (self.d, self.d):
                raise ValueError("Expected basis elements to be of shape "
                                 + f"({self.d}, {self.d}), not {basis.shape[1:]}!")
            self.basis = basis

    @classmethod
    def from_arrays(cls, c_opers, c_oper_identifiers, c_coeffs, n_opers, n_oper_identifiers,
                    n_coeffs, dt, basis=None):
        """Alternative constructor from already parsed arrays (reference ``:313-357``)."""
        new = cls.__new__(cls)
        new.c_opers = np.asanyarray(c_opers)
        new.c_oper_identifiers = np.asanyarray(c_oper_identifiers)
        new.c_coeffs = np.asanyarray(c_coeffs)
        new.n_opers = np.asanyarray(n_opers)
        new.n_oper_identifiers = np.asanyarray(n_oper_identifiers)
        new.n_coeffs = np.asanyarray(n_coeffs)
        new.dt = np.asanyarray(dt)
        new.d = new.c_opers.shape[-1]
        if basis is None:
            new.basis = Basis.ggm(new.d)
        else:       # the same Basis OBJECT is kept: its cached predicates are shared by all pulses built on it
            new.basis = basis if isinstance(basis, Basis) else np.asanyarray(basis).view(Basis)
        if not len(new.c_opers) == len(new.c_oper_identifiers) == len(new.c_coeffs):
            raise ValueError('Control Hamiltonian not same length!')
        if not len(new.n_opers) == len(new.n_oper_identifiers) == len(new.n_coeffs):
            raise ValueError('Noise Hamiltonian not same length!')
        if not len(set(new.c_opers.shape[1:] + new.n_opers.shape[1:])) == 1:
            raise ValueError('Control and/or noise Hamiltonian not same, square dimension!')
        if not new.dt.size == new.n_coeffs.shape[1] == new.c_coeffs.shape[1]:
            raise ValueError('Time steps not same length!')
        if not new.basis.d == new.d:
            raise ValueError('Basis dimension not same as Hamiltonian dimension!')
        return new

    # ---- dunder methods --------------------------------------------------------------------------
    def __str__(self):
        return f'{repr(self)}\n\tof dimension {self.d} and duration {self.duration}'

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, self.__class__):
            return NotImplemented
        atol = np.finfo(complex).eps*self.basis.shape[0]
        mine, theirs = _merge_equal_segments(self), _merge_equal_segments(other)
        if len(mine[2]) != len(theirs[2]) or not np.allclose(mine[2], theirs[2], 1e-10, atol):
            return False
        for kind, coeffs_a, coeffs_b in (('c', mine[0], theirs[0]), ('n', mine[1], theirs[1])):
            ids_a = getattr(self, f'{kind}_oper_identifiers')
            ids_b = getattr(other, f'{kind}_oper_identifiers')
            if len(ids_a) != len(ids_b):
                return False
            order_a, order_b = np.argsort(ids_a), np.argsort(ids_b)
            if not np.array_equal(ids_a[order_a], ids_b[order_b]):
                return False
            if not np.array_equal(getattr(self, f'{kind}_opers')[order_a],
                                  getattr(other, f'{kind}_opers')[order_b]):
                return False
            if not np.array_equal(coeffs_a[order_a], coeffs_b[order_b]):
                return False
        return bool(self.basis == other.basis)

    __hash__ = None

    def __len__(self) -> int:
        return len(self.dt)

    def __getitem__(self, key) -> 'PulseSequence':
        new_dt = np.atleast_1d(self.dt[key])
        if not new_dt.size:
            raise IndexError('Cannot create empty PulseSequence')
        return self.__class__.from_arrays(
            c_opers=self.c_opers, n_opers=self.n_opers,
            c_oper_identifiers=self.c_oper_identifiers,
            n_oper_identifiers=self.n_oper_identifiers,
            c_coeffs=np.atleast_2d(self.c_coeffs.T[key]).T,
            n_coeffs=np.atleast_2d(self.n_coeffs.T[key]).T,
            dt=new_dt, basis=self.basis)

    def __copy__(self) -> 'PulseSequence':
        copied = self.__class__.__new__(self.__class__)
        for key, value in self.__dict__.items():
            if key in _CACHE_NAMES:
                copied.__dict__[key].update(value)       # own dictionaries, same cached arrays
            elif key != '_stamp':
                copied.__dict__[key] = value
        return copied

    def __deepcopy__(self, memo) -> 'PulseSequence':
        copied = self.__class__.__new__(self.__class__)
        copied.__setstate__(copy.deepcopy(self.__getstate__(), memo))
        return copied

    def __matmul__(self, other: 'PulseSequence') -> 'PulseSequence':
        if not isinstance(other, self.__class__):
            raise TypeError(f'Incompatible type for concatenation: {type(other)}')
        return concatenate((self, other))

    def __imatmul__(self, other):
        raise NotImplementedError

    # ---- cache protocol --------------------------------------------------------------------------
    def is_cached(self, attr: str) -> bool:
        alias = attr.lower().replace('_', ' ')
        if alias in _DATA_ALIASES:
            return _DATA_ALIASES[alias] in self._data
        if alias in _FREQUENCY_DATA_ALIASES:
            return _FREQUENCY_DATA_ALIASES[alias] in self._frequency_data
        return (attr in self._intermediates or attr in self._frequency_data
                or attr in self._data)

    @property
    def t(self) -> ndarray:
        if 't' not in self._data:
            self._data['t'] = np.concatenate(([0], self.dt.cumsum()))
        return self._data['t']

    @t.setter
    def t(self, val):
        self._data['t'] = val

    @property
    def tau(self):
        if 'tau' not in self._data:
            self._data['tau'] = self.t[-1] if 't' in self._data else self.dt.sum()
        return self._data['tau']

    @tau.setter
    def tau(self, val):
        self._data['tau'] = val

    @property
    def duration(self):
        return self.tau

    @property
    def data(self):
        return MappingProxyType(self._data)

    @property
    def frequency_data(self):
        return MappingProxyType(self._frequency_data)

    @property
    def intermediates(self):
        return MappingProxyType(self._intermediates)

    def _cached_property(key):  # noqa: N805  (helper used while the class body executes)
        def getter(self):
            if key not in self._data:
                self.diagonalize()
            return self._data[key]

        def setter(self, value):
            self._data[key] = value
        return property(getter, setter)

    eigvals = _cached_property('eigvals')
    eigvecs = _cached_property('eigvecs')
    propagators = _cached_property('propagators')
    total_propagator = _cached_property('total_propagator')
    del _cached_property

    @property
    def total_propagator_liouville(self) -> ndarray:
        if 'total_propagator_liouville' not in self._data:
            self._data['total_propagator_liouville'] = liouville_representation(
                self.total_propagator, self.basis)
        return self._data['total_propagator_liouville']

    @total_propagator_liouville.setter
    def total_propagator_liouville(self, value) -> None:
        self._data['total_propagator_liouville'] = value

    @property
    def omega(self):
        return self._frequency_data.get('omega', None)

    @omega.setter
    def omega(self, value) -> None:
        """Setting different frequencies drops everything that depends on them (``:1158-1169``)."""
        old = self._frequency_data.get('omega', None)
        new = np.array(value, copy=True)
        if not np.array_equal(old, new):
            self.cleanup('frequency dependent')
        self._frequency_data['omega'] = new

    @property
    def nbytes(self) -> int:
        total = 0
        for val in chain(self._data.values(), self._frequency_data.values(),
                         self._intermediates.values()):
            total += getattr(val, 'nbytes', 0)
        return total

    @util.parse_optional_parameters(method=('conservative', 'greedy', 'frequency dependent', 'all'))
    def cleanup(self, method: str = 'conservative') -> None:
        """Drop cached by-products; the four modes delete what the reference deletes (``:1188``)."""
        if method == 'all':
            self._data.clear()
            self._frequency_data.clear()
            self._intermediates.clear()
            return
        if method == 'frequency dependent':
            self._frequency_data.clear()
            self._intermediates.clear()
            return
        for key in ('eigvals', 'eigvecs', 'propagators'):
            self._data.pop(key, None)
        if method == 'greedy':
            self._intermediates.clear()
            for key in ('total_propagator', 'total_propagator_liouville'):
                self._data.pop(key, None)
            for key in ('total_phases', 'control_matrix', 'control_matrix_pc'):
                self._frequency_data.pop(key, None)

    # ---- numerics --------------------------------------------------------------------------------
    def diagonalize(self) -> None:
        """Eigen-decompose the control Hamiltonian and cache eigvals / eigvecs / propagators."""
        if not all(key in self._data for key in ('eigvals', 'eigvecs', 'propagators')):
            self._data['eigvals'], self._data['eigvecs'], self._data['propagators'] = \
                numeric._diagonalize_from_coeffs(self.c_opers, self.c_coeffs, self.dt)
        self._data['total_propagator'] = self._data['propagators'][-1]

    def get_control_matrix(self, omega, show_progressbar: bool = False,
                           cache_intermediates: bool = False) -> ndarray:
        """Control matrix (n_nops, n_basis, n_omega) for ``omega``; cached per frequency grid."""
        self.omega = omega
        if 'control_matrix' in self._frequency_data:
            return self._frequency_data['control_matrix']
        if 'control_matrix_pc' in self._frequency_data:
            self._frequency_data['control_matrix'] = np.sum(
                self._frequency_data['control_matrix_pc'], axis=0)
            return self._frequency_data['control_matrix']
        self.diagonalize()
        control_matrix = numeric._control_matrix_from_scratch(
            self.eigvals, self.eigvecs, self.propagators, self.omega, self.basis, self.n_opers,
            self.n_coeffs, self.dt, self.t, cache_intermediates=cache_intermediates, keep=True)
        if cache_intermediates:
            control_matrix, intermediates = control_matrix
            self._intermediates.update(intermediates)
        self.cache_control_matrix(self.omega, control_matrix)
        self._mark_computed(control_matrix)
        return self._frequency_data['control_matrix']

    def _mark_computed(self, control_matrix) -> None:
        """Remember that the cached control matrix is what this package computes for this pulse (as
        opposed to an array the user handed to ``cache_control_matrix``): ``concatenate`` may then pick
        the cheaper of two equivalent routes."""
        import weakref
        self.__dict__['_computed_control_matrix'] = weakref.ref(control_matrix)

    def _control_matrix_is_computed(self) -> bool:
        ref = self.__dict__.get('_computed_control_matrix')
        return ref is not None and ref() is self._frequency_data.get('control_matrix')

    def cache_control_matrix(self, omega, control_matrix: Optional[ndarray] = None,
                             show_progressbar: bool = False,
                             cache_intermediates: bool = False) -> None:
        self.omega = omega
        if control_matrix is None:
            control_matrix = self.get_control_matrix(self.omega, show_progressbar,
                                                     cache_intermediates)
        key = 'control_matrix_pc' if control_matrix.ndim == 4 else 'control_matrix'
        self._frequency_data[key] = control_matrix
        self.cache_total_phases(self.omega)
        if 'total_propagator_liouville' not in self._data:
            self.total_propagator_liouville = liouville_representation(self.total_propagator,
                                                                       self.basis)

    def get_pulse_correlation_control_matrix(self) -> ndarray:
        if 'control_matrix_pc' in self._frequency_data:
            return self._frequency_data['control_matrix_pc']
        raise util.CalculationError(
            "Could not get the pulse correlation control matrix since it "
            + "was not computed during concatenation. Please run the "
            + "concatenation again with 'calc_pulse_correlation_FF' set to "
            + "True.")

    def _is_cold(self) -> bool:
        """Nothing of the path cached yet (so the fused one-call pipeline applies)."""
        return not (any(k in self._frequency_data for k in ('control_matrix', 'control_matrix_pc',
                                                            'filter_function'))
                    or any(k in self._data for k in ('eigvals', 'eigvecs', 'propagators')))

    def _cold_pipeline(self, omega, spectrum=None):
        """Cold cache + fidelity filter function: ONE library call (single upload, all kernels back
        to back, single download) instead of one round trip per stage; caches exactly what the
        step-by-step route caches.  With ``spectrum`` (already broadcast for all noise operators) the
        infidelity integral runs in the same call and is returned."""
        self.omega = omega
        omega_arr = _lib.as_f64(self.omega)
        c_opers, c_coeffs = _lib.as_c128(self.c_opers), _lib.as_f64(self.c_coeffs)
        n_opers, n_coeffs = _lib.as_c128(self.n_opers), _lib.as_f64(self.n_coeffs)
        dt = _lib.as_f64(self.dt)
        # t is only passed if it is cached already; otherwise the library forms [0, cumsum(dt)] itself
        t = _lib.as_f64(self._data['t']) if 't' in self._data else None
        basis = _lib.as_c128(np.asarray(self.basis))
        # ``self.d`` only normalises the infidelity (the reference's tests overwrite it with the size
        # of a computational subspace, tests/test_precision.py:297); matrix shapes come from the arrays
        G, d = len(dt), c_opers.shape[-1]
        n_cops, n_nops, n_basis, n_omega = len(c_opers), len(n_opers), len(basis), len(omega_arr)
        S = infid = None
        s_ndim = s_complex = 0
        # all results in ONE page-locked block, neighbours in the order the library downloads them
        specs = [((G, d), np.float64), ((G, d, d), np.complex128), ((G + 1, d, d), np.complex128),
                 ((n_omega,), np.complex128), ((n_basis, n_basis), np.complex128),
                 ((n_nops, n_basis, n_omega), np.complex128), ((n_nops, n_nops, n_omega), np.complex128)]
        if spectrum is not None:
            s_ndim, s_complex = spectrum.ndim, int(np.iscomplexobj(spectrum))
            S = _lib.as_c128(spectrum) if s_complex else _lib.as_f64(spectrum)
            specs.append(((n_nops, n_nops) if s_ndim == 3 else (n_nops,), np.float64))
        arrays = _lib.empty_many(specs)
        eigvals, eigvecs, propagators, phases, liouville, B, F = arrays[:7]
        if spectrum is not None:
            infid = arrays[7]
        ctx = _lib.context()
        p = _lib.ptr
        # the cached arrays stay mirrored on the device: a second spectrum, decay amplitudes or a
        # concatenation that is handed them later does not upload them again
        _lib.keep_on_device(ctx)
        _lib.check(ctx, _lib.lib().ffb_pulse_filter_function(
            ctx, G, d, n_cops, n_nops, n_basis, n_omega, p(c_opers), p(c_coeffs), p(n_opers),
            p(n_coeffs), p(dt), p(t), p(basis), p(omega_arr), p(S), s_ndim, s_complex, p(eigvals),
            p(eigvecs), p(propagators), p(B), p(F), p(infid), p(phases), p(liouville)))
        _lib.freeze_shadowed(ctx, eigvals, eigvecs, propagators, phases, B, F)
        self._data.update(eigvals=eigvals, eigvecs=eigvecs, propagators=propagators,
                          total_propagator=propagators[-1])
        self._data['total_propagator_liouville'] = normalize_liouville_columns(
            np.ascontiguousarray(liouville.real) if self.basis.isherm else liouville, self.basis)
        self._frequency_data.update(control_matrix=B, total_phases=phases, filter_function=F)
        self._mark_computed(B)
        if infid is None:
            return None
        # a copy: the few numbers the caller keeps must not hold the whole result block (and its device
        # mirror) alive after the pulse's caches are dropped
        infid = np.array(infid)
        if self.d != d:
            infid *= d/self.d
        return infid

    @util.parse_optional_parameters(which=('fidelity', 'generalized'), order=(1, 2))
    def get_filter_function(self, omega, which: str = 'fidelity', order: int = 1,
                            show_progressbar: bool = False, cache_intermediates: bool = False,
                            cache_second_order_cumulative: bool = False) -> ndarray:
        r"""First-order filter function :math:`F_{\alpha\beta}(\omega)` (or the generalized
        :math:`F_{\alpha\beta,kl}(\omega)`), cached per frequency grid (reference ``:691-805``)."""
        if order == 2:
            raise NotImplementedError('Second-order filter functions are out of scope of '
                                      'filter_functions_b200 (SURVEY.md section 2, row 11)')
        self.omega = omega
        key = 'filter_function' if which == 'fidelity' else 'filter_function_gen'
        if key in self._frequency_data:
            return self._frequency_data[key]
        if (which == 'fidelity' and not cache_intermediates and len(self.omega)
                and self._is_cold()):
            self._cold_pipeline(self.omega)
            return self._frequency_data[key]
        control_matrix = self.get_control_matrix(self.omega, show_progressbar, cache_intermediates)
        self.cache_filter_function(self.omega, control_matrix=control_matrix, which=which,
                                   order=order, show_progressbar=show_progressbar,
                                   cache_intermediates=cache_intermediates)
        return self._frequency_data[key]

    @util.parse_optional_parameters(which=('fidelity', 'generalized'), order=(1, 2))
    def cache_filter_function(self, omega, control_matrix: Optional[ndarray] = None,
                              filter_function: Optional[ndarray] = None, which: str = 'fidelity',
                              order: int = 1, show_progressbar: bool = False,
                              cache_intermediates: bool = False,
                              cache_second_order_cumulative: bool = False) -> None:
        """Cache the filter function; a 4-d ``control_matrix`` is a pulse-correlation control
        matrix, in which case the pulse-correlation filter function is cached too (``:807-902``)."""
        if order == 2:
            raise NotImplementedError('Second-order filter functions are out of scope of '
                                      'filter_functions_b200 (SURVEY.md section 2, row 11)')
        self.omega = omega
        if filter_function is None:
            if control_matrix is None:
                control_matrix = self.get_control_matrix(self.omega, show_progressbar,
                                                         cache_intermediates)
            self.cache_control_matrix(self.omega, control_matrix)
            if control_matrix.ndim == 4:
                F_pc = numeric.calculate_pulse_correlation_filter_function(control_matrix, which)
                if which == 'fidelity':
                    self._frequency_data['filter_function_pc'] = F_pc
                else:
                    self._frequency_data['filter_function_pc'] = F_pc.trace(axis1=4, axis2=5)
                    self._frequency_data['filter_function_pc_gen'] = F_pc
                filter_function = F_pc.sum(axis=(0, 1))
            else:
                filter_function = numeric.calculate_filter_function(control_matrix, which)
        if which == 'fidelity':
            self._frequency_data['filter_function'] = filter_function
        else:
            self._frequency_data['filter_function'] = filter_function.trace(axis1=2, axis2=3)
            self._frequency_data['filter_function_gen'] = filter_function

    @util.parse_optional_parameters(which=('fidelity', 'generalized'))
    def get_pulse_correlation_filter_function(self, which: str = 'fidelity') -> ndarray:
        key = 'filter_function_pc' if which == 'fidelity' else 'filter_function_pc_gen'
        if key in self._frequency_data:
            return self._frequency_data[key]
        if 'control_matrix_pc' in self._frequency_data:
            F_pc = numeric.calculate_pulse_correlation_filter_function(
                self._frequency_data['control_matrix_pc'], which=which)
            self._frequency_data[key] = F_pc
            return F_pc
        raise util.CalculationError(
            "Could not get the pulse correlation filter function since it "
            + "was not computed during concatenation. Please run the "
            + "concatenation again with 'calc_pulse_correlation_FF' set to True.")

    def get_total_phases(self, omega) -> ndarray:
        self.omega = omega
        if 'total_phases' not in self._frequency_data:
            self._frequency_data['total_phases'] = util.cexp(self.omega*self.tau)
        return self._frequency_data['total_phases']

    def cache_total_phases(self, omega, total_phases: Optional[ndarray] = None) -> None:
        self.omega = omega
        if total_phases is None:
            total_phases = self.get_total_phases(self.omega)
        self._frequency_data['total_phases'] = total_phases


# ------------------------------------------------------------------------------------------------
# concatenation
# ------------------------------------------------------------------------------------------------
def _oper_hashes(pulse, kind: str):
    """Hashes of a pulse's control / noise operators, computed once per operator array (the arrays
    of a ``PulseSequence`` are never modified in place by this package)."""
    opers = pulse.c_opers if kind == 'control' else pulse.n_opers
    cache = pulse.__dict__.setdefault('_oper_hash_cache', {})
    entry = cache.get(kind)
    if entry is None or entry[0] is not opers:
        entry = cache[kind] = (opers, tuple(util.hash_array_along_axis(opers, axis=0)))
    return entry[1]


def _join_hamiltonians_core(distinct, order, kind: str):
    """Merge the operator lists of the pulses of a sequence into one Hamiltonian.

    Equal operators (byte-wise) are merged; an identifier used for two different operators gets the
    pulse position appended; operators missing on some pulse get zero (control) or, if constant
    elsewhere, that constant (noise) coefficients.  Behaviour of the reference's
    ``_concatenate_hamiltonian`` (``:1340-1483``), including its error messages.  All bookkeeping runs
    over the DISTINCT pulse objects (a sequence of 100 Cliffords has at most 24).  Returns the merged
    operators and identifiers (sorted by identifier), per distinct pulse its block of the joined
    coefficient array (rows in final order, rows of operators it does not carry filled in) and its
    identifier mapping, the mapping per position, and whether an identifier clash had to be renamed
    (then the outcome depends on the positions and cannot be remembered per set of objects)."""
    attr = 'c' if kind == 'control' else 'n'
    entries = [(_oper_hashes(p, kind), getattr(p, f'{attr}_oper_identifiers').tolist(),
                getattr(p, f'{attr}_opers'), getattr(p, f'{attr}_coeffs')) for p in distinct]
    first_pos = [None]*len(distinct)
    for pos, i in enumerate(order):
        if first_pos[i] is None:
            first_pos[i] = pos
    first = {}              # hash -> (pulse position, index within that pulse) of first occurrence
    ids_of_oper, opers_of_id = {}, {}
    for pos, (hashes, idents, _, _) in sorted(zip(first_pos, entries), key=lambda t: t[0]):
        for loc, (h, ident) in enumerate(zip(hashes, idents)):
            if h not in first:
                first[h] = (pos, loc)
                ids_of_oper[h] = {ident}
            else:
                ids_of_oper[h].add(ident)
            opers_of_id.setdefault(ident, set()).add(h)
    if any(len(v) > 1 for v in ids_of_oper.values()):
        raise ValueError(f'Trying to concatenate pulses with equal {kind} operators but '
                         + f'different identifiers. Please choose unique {kind} identifiers!')

    row_of = {h: i for i, h in enumerate(first)}
    new_ids = [entries[order[pos]][1][loc] for pos, loc in first.values()]
    maps = [dict(zip(entry[1], entry[1])) for entry in entries]
    mapping = {pos: maps[i] for pos, i in enumerate(order)}
    clash = False
    for ident, hashes in opers_of_id.items():
        if len(hashes) > 1:
            clash = True
            for h in hashes:
                pulse_pos = first[h][0]
                new_ids[row_of[h]] = f'{ident}_{pulse_pos}'
                mapping[pulse_pos] = dict(mapping[pulse_pos])      # this position only
                mapping[pulse_pos][ident] = new_ids[row_of[h]]

    by_id = np.argsort(new_ids)
    final_row = np.empty(len(by_id), dtype=int)
    final_row[by_id] = np.arange(len(by_id))            # unsorted row -> row after sorting by identifier
    n_rows = len(new_ids)
    # per distinct pulse its block of the joined array in FINAL row order, NaN where it lacks an operator
    blocks, lacking = [], np.zeros(n_rows, dtype=bool)
    for entry in entries:
        rows = final_row[[row_of[h] for h in entry[0]]]
        if len(rows) == n_rows and (rows == np.arange(n_rows)).all():
            block = np.asarray(entry[3], dtype=float)
        else:
            block = np.full((n_rows, entry[3].shape[1]), np.nan)
            block[rows] = entry[3]
            present = np.zeros(n_rows, dtype=bool)
            present[rows] = True
            lacking |= ~present
        blocks.append(block)
    if lacking.any():
        # every pulse of the call appears in `distinct`, so what is inferred from the distinct blocks is
        # what the reference infers from the whole joined array
        for row in lacking.nonzero()[0]:
            if kind == 'noise':
                present = np.concatenate([blk[row][~np.isnan(blk[row])] for blk in blocks])
                if not (present == present[0]).all():
                    raise ValueError('Not all pulses have the same noise operators and '
                                     + 'non-trivial noise sensitivities so I cannot infer them.')
                fill = present[0]
            else:
                fill = 0.0
            for i, blk in enumerate(blocks):
                if np.isnan(blk[row]).any():
                    if blk is entries[i][3]:
                        blk = blocks[i] = blk.copy()
                    blk[row, np.isnan(blk[row])] = fill
    opers = np.array([entries[order[pos]][2][loc] for pos, loc in first.values()])[by_id]
    identifiers = np.array([new_ids[i] for i in by_id])
    return opers, identifiers, blocks, maps, mapping, clash


def _distinct_pulses(pulses):
    """Distinct pulse objects of a sequence in order of first occurrence, and for every position the
    index of its object in that list."""
    slot, distinct, order = {}, [], []
    for p in pulses:
        i = slot.get(id(p))
        if i is None:
            i = slot[id(p)] = len(distinct)
            distinct.append(p)
        order.append(i)
    return distinct, order


class _SequencePlan:
    """Everything ``concatenate`` derives from a SET of gate objects and not from their order in a
    sequence: joined operators and identifiers, per gate its block of the joined [control; noise; dt]
    rows, identifier mappings, which joined noise operators each gate carries, durations -- and, once a
    sequence has been evaluated, the gates' stacked control matrices / phases / propagators mirrored in
    device memory (``library``).  Valid while the gates' stamps have not moved."""
    __slots__ = ('refs', 'stamps', 'control', 'noise', 'blocks', 'maps', 'n_control', 'n_noise',
                 'taus', 'present', 'present_c', 'basis', 'library')


_SEQUENCE_PLANS = {}
_SEQUENCE_PLAN_LIMIT = 16
_SEQUENCE_PLAN_MAX_BYTES = 1 << 20


def _plan_for(distinct):
    key = tuple(sorted(id(p) for p in distinct))
    plan = _SEQUENCE_PLANS.get(key)
    if plan is not None:
        for p in distinct:
            i = id(p)
            if plan.refs[i]() is not p or plan.stamps[i] != p._stamp[0]:
                del _SEQUENCE_PLANS[key]
                return key, None
        return key, plan
    # A sequence that happens to miss some gates of a library (29 % of 100-Clifford sequences miss at least
    # one of the 24 Cliffords) can use the plan of a superset, provided its own gates together still carry
    # every control and noise operator of that plan's join -- then its join has the same operators, (sorted)
    # identifiers and mappings, and the gates' coefficient blocks are the same.
    for other in _SEQUENCE_PLANS.values():
        if len(other.refs) > len(key) and all(i in other.refs for i in key):
            if (all(other.refs[id(p)]() is p and other.stamps[id(p)] == p._stamp[0] for p in distinct)
                    and np.logical_or.reduce([other.present[i] for i in key]).all()
                    and np.logical_or.reduce([other.present_c[i] for i in key]).all()):
                return key, other
    return key, None


def _unique_by_identity(items):
    seen, out = set(), []
    for item in items:
        if id(item) not in seen:
            seen.add(id(item))
            out.append(item)
    return out


def _join_pulses(pulses, want_maps: bool = True):
    """Host part of a concatenation: validation, joined Hamiltonians, new PulseSequence.  Returns
    ``(newpulse, control mapping, noise mapping, distinct pulse objects, position -> distinct index,
    plan or None)``.  Written for long sequences of recurring gates (randomized benchmarking): the outcome
    of the bookkeeping is remembered per set of gate objects (:class:`_SequencePlan`), and a call then
    costs ONE concatenation of prepared coefficient blocks."""
    import weakref
    try:
        pulses = tuple(pulses)
    except TypeError:
        raise TypeError(f'Expected pulses to be iterable, not {type(pulses)}')
    distinct, order = _distinct_pulses(pulses)
    if not all(isinstance(pulse, PulseSequence) for pulse in distinct):
        raise TypeError('Can only concatenate PulseSequences!')
    key, plan = _plan_for(distinct)
    if plan is not None:
        ids = [id(p) for p in distinct]                     # per distinct object, then plain indexing
        blocks = [plan.blocks[i] for i in ids]
        joined = np.concatenate([blocks[i] for i in order], axis=1)
        n_c, n_n = plan.n_control, plan.n_noise
        newpulse = PulseSequence.from_arrays(plan.control[0], plan.control[1], joined[:n_c],
                                             plan.noise[0], plan.noise[1], joined[n_c:n_c + n_n],
                                             joined[-1], pulses[0].basis)
        taus = [plan.taus[i] for i in ids]
        newpulse.tau = sum([taus[i] for i in order])
        if not want_maps:      # concatenate() reads the plan's presence rows instead
            return newpulse, None, None, distinct, order, plan
        c_maps, n_maps = [plan.maps[0][i] for i in ids], [plan.maps[1][i] for i in ids]
        c_map = {pos: c_maps[i] for pos, i in enumerate(order)}
        n_map = {pos: n_maps[i] for pos, i in enumerate(order)}
        return newpulse, c_map, n_map, distinct, order, plan

    if len(set(pulse.d for pulse in distinct)) != 1:
        raise ValueError('Trying to concatenate PulseSequence instances with different dimension!')
    bases = _unique_by_identity(pulse.basis for pulse in distinct)
    if len(bases) > 1 and not util.all_array_equal(bases):
        raise ValueError('Trying to concatenate PulseSequence instances with different bases!')

    c_opers, c_ids, c_blocks, c_maps, c_map, c_clash = _join_hamiltonians_core(distinct, order,
                                                                               'control')
    n_opers, n_ids, n_blocks, n_maps, n_map, n_clash = _join_hamiltonians_core(distinct, order,
                                                                               'noise')
    dts = [pulse.dt for pulse in distinct]
    taus = [pulse.tau for pulse in distinct]
    c_coeffs = np.concatenate([c_blocks[i] for i in order], axis=1)
    n_coeffs = np.concatenate([n_blocks[i] for i in order], axis=1)
    dt = np.concatenate([dts[i] for i in order])
    newpulse = PulseSequence.from_arrays(c_opers, c_ids, c_coeffs, n_opers, n_ids, n_coeffs, dt,
                                         pulses[0].basis)
    newpulse.tau = sum([taus[i] for i in order])      # same left-to-right sum as over the pulses

    small = sum(b.nbytes for b in c_blocks) + sum(b.nbytes for b in n_blocks) <= _SEQUENCE_PLAN_MAX_BYTES
    if (not c_clash and not n_clash and small
            and all(np.asarray(x).dtype == np.float64 for x in dts)):
        plan = _SequencePlan()
        plan.refs = {id(p): weakref.ref(p) for p in distinct}
        plan.control, plan.noise = (c_opers, c_ids), (n_opers, n_ids)
        plan.n_control, plan.n_noise = len(c_ids), len(n_ids)
        plan.blocks = {id(p): np.concatenate([cb, nb, np.asarray(t, dtype=float)[None]])
                       for p, cb, nb, t in zip(distinct, c_blocks, n_blocks, dts)}
        plan.maps = ({id(p): m for p, m in zip(distinct, c_maps)},
                     {id(p): m for p, m in zip(distinct, n_maps)})
        plan.taus = {id(p): tau for p, tau in zip(distinct, taus)}
        column = {ident: i for i, ident in enumerate(n_ids.tolist())}
        plan.present = {}
        for p, m in zip(distinct, n_maps):
            row = np.zeros(len(n_ids), dtype=bool)
            row[[column[ident] for ident in m.values()]] = True
            plan.present[id(p)] = row
        column_c = {ident: i for i, ident in enumerate(c_ids.tolist())}
        plan.present_c = {}
        for p, m in zip(distinct, c_maps):
            row = np.zeros(len(c_ids), dtype=bool)
            row[[column_c[ident] for ident in m.values()]] = True
            plan.present_c[id(p)] = row
        plan.basis = pulses[0].basis
        plan.library = None
        plan.stamps = {id(p): p._stamp[0] for p in distinct}    # last: reading tau may have cached it
        while len(_SEQUENCE_PLANS) >= _SEQUENCE_PLAN_LIMIT:
            _SEQUENCE_PLANS.pop(next(iter(_SEQUENCE_PLANS)))
        _SEQUENCE_PLANS[key] = plan
    else:
        plan = None
    return newpulse, c_map, n_map, distinct, order, plan


def concatenate_without_filter_function(pulses: Iterable[PulseSequence],
                                        return_identifier_mappings: bool = False) -> Any:
    """Concatenate the Hamiltonians only (reference ``:1599-1665``)."""
    newpulse, c_map, n_map, _, _, _ = _join_pulses(pulses)
    if return_identifier_mappings:
        return newpulse, c_map, n_map
    return newpulse


def _gate_library(ctx, plan, distinct, omega, ctrl, phases, liouville, basis):
    """Stacks for the fused concatenation call (control matrices, total phases, Liouville and Hilbert
    propagators of the distinct gates, basis).  With a plan, sequences drawn from the same set of cached
    gate objects (randomized benchmarking: 1000 sequences over 24 Cliffords) share ONE set of stacks
    that stays mirrored in device memory: no per-call stacking, no per-call upload.  Returns
    ``(lib_B, lib_ph, lib_L, lib_U, basis, rank)``; ``rank[i]`` is the row of ``distinct[i]``."""
    if plan is not None and plan.library is not None and _same_grid(plan.library[0], omega):
        stacks, rows = plan.library[1], plan.library[2]
        return (*stacks, [rows[id(p)] for p in distinct])
    n = len(distinct)
    props = [pls.total_propagator for pls in distinct]
    shapes = [((n,) + np.shape(ctrl[0]), np.complex128), ((n,) + np.shape(phases[0]), np.complex128),
              ((n,) + np.shape(liouville[0]), np.float64), ((n,) + np.shape(props[0]), np.complex128),
              (np.shape(basis), np.complex128)]
    total = sum(int(np.prod(sh))*np.dtype(dt).itemsize for sh, dt in shapes)
    # (a plan borrowed from a superset of these gates keeps only a library that covers all of ITS gates)
    keep = plan is not None and total <= (16 << 20) and len(plan.refs) == n
    stacks = _lib.empty_many(shapes, ctx) if keep else [np.empty(sh, dtype=dt) for sh, dt in shapes]
    for i in range(n):
        for stack, part in zip(stacks[:4], (ctrl[i], phases[i], liouville[i], props[i])):
            stack[i] = part
    stacks[4][...] = np.asarray(basis)
    if keep:
        for stack in stacks:
            _lib.mirror_input(ctx, stack)
        plan.library = (np.array(omega, copy=True), tuple(stacks), {id(p): i for i, p in enumerate(distinct)})
        for p in distinct:                                     # gathering the parts may have cached some
            plan.stamps[id(p)] = p._stamp[0]
    return (*stacks, list(range(n)))


_SAME_GRID = {}


def _same_grid(cached, omega) -> bool:
    """``np.array_equal(cached, omega)``, remembered per pair of array OBJECTS: a sequence of gates whose
    pulses each hold their own copy of the same frequency grid would otherwise compare the grids once
    per gate and call."""
    if cached is omega:
        return True
    key = (id(cached), id(omega))
    memo = _SAME_GRID.get(key)
    if memo is not None and memo[0]() is cached and memo[1]() is omega:
        return memo[2]
    equal = bool(np.shape(cached) == np.shape(omega) and np.array_equal(cached, omega))
    try:
        import weakref
        if len(_SAME_GRID) > 4096:
            _SAME_GRID.clear()
        _SAME_GRID[key] = (weakref.ref(cached), weakref.ref(omega), equal)
    except TypeError:
        pass
    return equal


def _frequency_entries(pls, keys, omega):
    """Cached frequency-dependent arrays ``keys`` of ``pls`` if they belong to the grid ``omega``
    (one comparison for all of them, and without the copy + comparison the ``omega`` setter makes on
    every access); ``None`` for what is not cached or belongs to another grid."""
    cached = pls._frequency_data.get('omega')
    if cached is None or not _same_grid(cached, omega):
        return (None,)*len(keys)
    return tuple(pls._frequency_data.get(key) for key in keys)


@util.parse_optional_parameters(which=('fidelity', 'generalized'))
def concatenate(pulses: Iterable[PulseSequence], calc_pulse_correlation_FF: bool = False,
                calc_filter_function: Optional[bool] = None,
                calc_second_order_FF: Optional[bool] = None, which: str = 'fidelity',
                omega=None, show_progressbar: bool = False) -> PulseSequence:
    r"""Concatenate pulses left to right (``concatenate((A, B))`` is :math:`B\circ A`) and, when the
    constituents have cached control matrices (or ``omega`` is given), compute the control matrix of
    the sequence from them.  Decision logic, caching and errors follow the reference (``:1668-1887``);
    the host side is written for long sequences of recurring gates (per-object work is done once per
    distinct pulse object, cached arrays are read without re-validating the frequency grid).
    """
    if calc_second_order_FF:
        raise NotImplementedError('Second-order filter functions are out of scope of '
                                  'filter_functions_b200 (SURVEY.md section 2, row 11)')
    pulses = tuple(pulses)
    if len(pulses) == 1:
        return copy.deepcopy(pulses[0])

    newpulse, _, n_oper_mapping, distinct, inverse, plan = _join_pulses(pulses, want_maps=False)
    have_propagators = all('total_propagator' in pls._data for pls in distinct)

    def host_total_propagator():
        if have_propagators and 'total_propagator' not in newpulse._data:
            props = [pls.total_propagator for pls in distinct]
            newpulse.total_propagator = util.mdot([props[i] for i in inverse][::-1])

    if calc_pulse_correlation_FF:
        calc_filter_function = True
    if calc_filter_function is False:
        host_total_propagator()
        return newpulse

    # which (renamed) noise operators does each pulse carry?  One row per distinct object and mapping
    # (positions share their object's mapping unless an identifier clash renamed something there)
    new_ids = newpulse.n_oper_identifiers.tolist()
    if plan is not None:
        present = np.array([plan.present[id(p)] for p in distinct])
        # (every gate carries every operator -- the randomized-benchmarking case: the rows of all positions
        # are then one broadcast row, not a gather over the sequence)
        n_opers_present = (np.broadcast_to(present[0], (len(pulses), present.shape[1])) if present.all()
                           else present[inverse])
    else:
        column = {ident: i for i, ident in enumerate(new_ids)}
        rows_by_mapping, present_rows = {}, []
        for pos in range(len(pulses)):
            mapping = n_oper_mapping[pos]
            row = rows_by_mapping.get(id(mapping))
            if row is None:
                row = [False]*len(new_ids)
                for ident in mapping.values():
                    row[column[ident]] = True
                rows_by_mapping[id(mapping)] = row
            present_rows.append(row)
        n_opers_present = np.array(present_rows, dtype=bool)

    equal_n_opers = (n_opers_present.sum(axis=0) > 1).any()
    if omega is None:
        # the reference's predicate is is_cached('control_matrix') -- a pulse that only holds a
        # pulse-correlation control matrix does not count (pulse_sequence.py:1783)
        with_ctrl = [pls for pls in distinct if 'control_matrix' in pls._frequency_data]
        with_omega = with_ctrl or [pls for pls in distinct if 'omega' in pls._frequency_data]
        grids = _unique_by_identity(pls.omega for pls in with_omega)
        equal_omega = bool(grids) and all(_same_grid(grid, grids[0]) for grid in grids[1:])
        if not equal_omega:
            host_total_propagator()
            if calc_filter_function:
                raise ValueError("Calculation of filter function forced but not all pulses "
                                 + "have the same frequencies cached and none were supplied!")
            if calc_pulse_correlation_FF:
                raise ValueError("Cannot compute the pulse correlation filter functions; do not "
                                 + "have the frequencies at which to evaluate.")
            return newpulse
        if calc_filter_function is None and (not equal_n_opers or not with_ctrl):
            host_total_propagator()
            return newpulse
        omega = with_omega[0].omega

    if not equal_n_opers:
        host_total_propagator()
        newpulse.cache_filter_function(omega, which=which)
        return newpulse

    if (not calc_pulse_correlation_FF and len(omega) and not n_opers_present.all()
            and _from_scratch_is_cheaper(pulses, distinct, inverse, n_opers_present, newpulse, omega)):
        # Few of the (pulse, noise operator) rows are cached (config 5: 22 of 162): filling the atomic
        # stack from scratch and contracting it costs more FP64 work than the control matrix of the
        # joined pulse from scratch, which is the same quantity (reference tests/test_core.py:685-743)
        newpulse.get_filter_function(omega, which=which, show_progressbar=show_progressbar)
        return newpulse

    # per distinct pulse object: total phases, Liouville propagator, control matrix on this grid
    # (not needed when the gates' stacks of an earlier sequence over the same gates are still valid)
    lib_ready = (plan is not None and plan.library is not None and not calc_pulse_correlation_FF
                 and which == 'fidelity' and _same_grid(plan.library[0], omega))
    lib_phases, lib_liouville, lib_ctrl = [], [], []
    if not lib_ready:
        for pls in distinct:
            ph, B = _frequency_entries(pls, ('total_phases', 'control_matrix'), omega)
            lib_phases.append(pls.get_total_phases(omega) if ph is None else ph)
            lib_ctrl.append(pls.get_control_matrix(omega, show_progressbar) if B is None else B)
            lib_liouville.append(pls.total_propagator_liouville)

    n_basis, n_omega = len(newpulse.basis), len(omega)
    if lib_ready or (
            n_opers_present.all() and not calc_pulse_correlation_FF and which == 'fidelity'
            and n_basis <= 64 and n_omega and newpulse.basis.isherm
            and all(np.isrealobj(liou) for liou in lib_liouville)):
        # every gate carries every noise operator: the whole tail of this function (running
        # propagator / phase products, from_atomic, Liouville representation, filter function) is
        # ONE library call on the distinct pulses plus an index list, as in concatenate_many
        d = newpulse.c_opers.shape[-1]
        ctx = _lib.context()
        lib_B, lib_ph, lib_L, lib_U, basis, rank = _gate_library(
            ctx, plan, distinct, omega, lib_ctrl, lib_phases, lib_liouville, newpulse.basis)
        index = np.array([rank[i] for i in inverse], dtype=np.int32)
        n_nops = lib_B.shape[1]
        U, liouville, total_phases, B, F = _lib.empty_many([   # one block, one download
            ((1, d, d), np.complex128), ((1, n_basis, n_basis), np.complex128),
            ((1, n_omega), np.complex128), ((1, n_nops, n_basis, n_omega), np.complex128),
            ((1, n_nops, n_nops, n_omega), np.complex128)])
        omega_arr = _lib.as_f64(omega)
        tau = np.array([newpulse.tau], dtype=np.float64)
        p = _lib.ptr
        _lib.keep_on_device(ctx)
        _lib.check(ctx, _lib.lib().ffb_concatenate_many(
            ctx, 1, len(index), lib_B.shape[0], d, n_nops, n_basis, n_omega, p(index), p(lib_B),
            p(lib_ph), p(lib_L), p(lib_U), p(basis), None, 0, 0, p(omega_arr), p(U), p(liouville),
            p(B), p(F), None, p(tau), p(total_phases)))
        _lib.freeze_shadowed(ctx, total_phases, B, F)
        if 'total_propagator' not in newpulse._data:
            newpulse.total_propagator = U[0]
        newpulse._frequency_data['omega'] = np.array(omega, copy=True)
        newpulse._frequency_data['total_phases'] = total_phases[0]
        newpulse.total_propagator_liouville = normalize_liouville_columns(
            np.ascontiguousarray(liouville[0].real), newpulse.basis)
        newpulse._frequency_data['control_matrix'] = B[0]
        newpulse._frequency_data['filter_function'] = F[0]
        return newpulse

    phases = np.array([lib_phases[i] for i in inverse[:-1]]).cumprod(axis=0)
    propagators_liouville = util.adot([lib_liouville[i] for i in inverse[:-1]])

    if 'total_propagator' not in newpulse._data:
        props = [pls.total_propagator for pls in distinct]
        newpulse.total_propagator = util.mdot([props[i] for i in inverse][::-1])

    if n_omega and np.isrealobj(propagators_liouville):
        # general case on the device: the atomic stack (cached rows + from-scratch rows of the noise
        # operators a pulse does not carry) never exists on the host
        control_matrix, filter_function = _concatenate_on_device(
            pulses, newpulse, n_opers_present, [lib_ctrl[i] for i in inverse], omega, phases,
            propagators_liouville, calc_pulse_correlation_FF, which)
        newpulse.cache_total_phases(omega)
        newpulse.total_propagator_liouville = liouville_representation(newpulse.total_propagator,
                                                                       newpulse.basis)
        newpulse.cache_control_matrix(omega, control_matrix)
        if calc_pulse_correlation_FF:
            if which == 'fidelity':
                newpulse._frequency_data['filter_function_pc'] = filter_function
            else:
                newpulse._frequency_data['filter_function_pc'] = filter_function.trace(axis1=4,
                                                                                       axis2=5)
                newpulse._frequency_data['filter_function_pc_gen'] = filter_function
            filter_function = filter_function.sum(axis=(0, 1))
        newpulse.cache_filter_function(omega, filter_function=filter_function, which=which)
        return newpulse

    if n_opers_present.all():
        control_matrix_atomic = np.array(lib_ctrl)[inverse]
    else:
        control_matrix_atomic = np.empty((len(pulses), len(new_ids), n_basis, n_omega),
                                         dtype=complex)
        seg_edges = [0] + list(accumulate(len(pls.dt) for pls in pulses))
        for i, (pls, present) in enumerate(zip(pulses, n_opers_present)):
            control_matrix_atomic[i, present] = lib_ctrl[inverse[i]]
            if not present.all():
                control_matrix_atomic[i, ~present] = numeric.calculate_control_matrix_from_scratch(
                    pls.eigvals, pls.eigvecs, pls.propagators, omega, pls.basis,
                    newpulse.n_opers[~present],
                    newpulse.n_coeffs[~present, seg_edges[i]:seg_edges[i + 1]],
                    pls.dt, t=pls.t, show_progressbar=show_progressbar, cache_intermediates=False)

    newpulse.cache_total_phases(omega)
    newpulse.total_propagator_liouville = liouville_representation(newpulse.total_propagator,
                                                                   newpulse.basis)
    control_matrix = numeric.calculate_control_matrix_from_atomic(
        phases, control_matrix_atomic, propagators_liouville, show_progressbar,
        which='correlations' if calc_pulse_correlation_FF else 'total')
    newpulse.cache_filter_function(omega, control_matrix, which=which)
    return newpulse


def _from_scratch_is_cheaper(pulses, distinct, inverse, n_opers_present, newpulse, omega) -> bool:
    """Cost model (real FP64 multiply-adds per frequency) of the two equivalent routes to the control
    matrix of a concatenated pulse.  Atomic route (reference ``pulse_sequence.py:1822-1866``): rows a
    pulse has not cached are computed from scratch on its segments, then the stack is contracted with
    the cumulative Liouville propagators (``2 n_nops n_basis^2`` per pulse).  Scratch route: all rows on
    all segments.  The scratch route is only admissible if every cached control matrix is one this
    package computed (a user-supplied matrix must enter the result), and is taken only with a clear
    margin.  ``FFB_CONCAT_ROUTE=atomic|scratch`` forces a route (measurements)."""
    import os
    forced = os.environ.get('FFB_CONCAT_ROUTE')
    cached = []
    for pls in distinct:
        B, = _frequency_entries(pls, ('control_matrix',), omega)
        if B is not None and not pls._control_matrix_is_computed():
            return False
        cached.append(B is not None)
    if forced in ('atomic', 'scratch'):
        return forced == 'scratch'
    d = newpulse.c_opers.shape[-1]
    n_basis, n_nops = len(newpulse.basis), n_opers_present.shape[1]
    per_row_segment = n_basis*(2 + 2*d*(d - 1))
    segments = [len(pls.dt) for pls in distinct]
    atomic = 0
    for pos, i in enumerate(inverse):
        to_compute = n_nops - (int(n_opers_present[pos].sum()) if cached[i] else 0)
        atomic += segments[i]*to_compute*per_row_segment
    atomic += (len(pulses) - 1)*n_nops*n_basis*n_basis*2
    scratch = len(newpulse.dt)*n_nops*per_row_segment
    return scratch < 0.8*atomic


def _concatenate_on_device(pulses, newpulse, n_opers_present, ctrl, omega, phases, liouville,
                           correlations: bool, which: str):
    """Tail of :func:`concatenate` for pulses that do not all carry the same noise operators
    (reference ``:1822-1866``) through ``ffb_concatenate_pulses``: returns the (pulse-correlation)
    control matrix and the matching filter function.  ``ctrl[i]`` is pulse i's own control matrix on
    ``omega``; its rows are, in order, the merged operators ``n_opers_present[i]`` marks."""
    P, n_nops = n_opers_present.shape
    n_basis, n_omega, d = len(newpulse.basis), len(omega), newpulse.c_opers.shape[-1]
    row_of = np.where(n_opers_present, np.cumsum(n_opers_present, axis=1) - 1, -1).astype(np.int32)
    seg_edges = [0] + list(accumulate(len(pls.dt) for pls in pulses))
    keep = []   # arrays the pointer tables refer to

    def table(arrays):
        tab = (ctypes.c_void_p*P)()
        for i, arr in enumerate(arrays):
            if arr is not None:
                keep.append(arr)
                tab[i] = arr.ctypes.data
        return tab

    cached = table([_lib.as_c128(B) for B in ctrl])
    need = [not present.all() for present in n_opers_present]
    for pls, needed in zip(pulses, need):
        if needed:
            pls.diagonalize()
    G = np.array([len(pls.dt) for pls in pulses], dtype=np.int32)
    eigvals = table([_lib.as_f64(pls.eigvals) if nd else None for pls, nd in zip(pulses, need)])
    eigvecs = table([_lib.as_c128(pls.eigvecs) if nd else None for pls, nd in zip(pulses, need)])
    props = table([_lib.as_c128(pls.propagators) if nd else None for pls, nd in zip(pulses, need)])
    dts = table([_lib.as_f64(pls.dt) if nd else None for pls, nd in zip(pulses, need)])
    ts = table([_lib.as_f64(pls.t) if nd else None for pls, nd in zip(pulses, need)])
    coeffs = table([_lib.as_f64(newpulse.n_coeffs[:, seg_edges[i]:seg_edges[i + 1]]) if nd else None
                    for i, nd in enumerate(need)])
    n_opers = _lib.as_c128(newpulse.n_opers)
    basis = _lib.as_c128(np.asarray(newpulse.basis))
    omega_arr = _lib.as_f64(omega)
    phases = _lib.as_c128(phases)
    liouville = _lib.as_f64(liouville)
    lead = (P,) if correlations else ()
    B = _lib.empty(lead + (n_nops, n_basis, n_omega))
    gen = which == 'generalized'
    F = _lib.empty(lead*2 + (n_nops, n_nops) + ((n_basis, n_basis) if gen else ()) + (n_omega,))
    ctx = _lib.context()
    p = _lib.ptr
    _lib.keep_on_device(ctx)
    _lib.check(ctx, _lib.lib().ffb_concatenate_pulses(
        ctx, P, d, n_nops, n_basis, n_omega, p(row_of), cached, p(G), eigvals, eigvecs, props, dts,
        ts, coeffs, p(n_opers), p(basis), p(omega_arr), p(phases), p(liouville), int(correlations),
        2 if gen else 1, p(B), p(F)))
    _lib.freeze_shadowed(ctx, B, F)
    return B, F


def concatenate_periodic(pulse: PulseSequence, repeats: int,
                         check_invertible: bool = True) -> PulseSequence:
    r"""Concatenate ``repeats`` copies of ``pulse``; with a cached control matrix the control matrix of
    the repeated pulse follows from the geometric series
    :math:`\tilde{\mathcal B}^{(1)}\sum_{g<G}(e^{i\omega T}\mathcal Q^{(1)})^g` in
    O(log ``repeats``) GPU passes (reference ``pulse_sequence.py:1890-1977``)."""
    if not isinstance(pulse, PulseSequence):
        raise TypeError('Can only concatenate PulseSequences!')
    repeats = int(repeats)
    newpulse = PulseSequence.from_arrays(
        c_opers=pulse.c_opers, c_oper_identifiers=pulse.c_oper_identifiers,
        c_coeffs=np.tile(pulse.c_coeffs, (1, repeats)), n_opers=pulse.n_opers,
        n_oper_identifiers=pulse.n_oper_identifiers,
        n_coeffs=np.tile(pulse.n_coeffs, (1, repeats)), dt=np.tile(pulse.dt, repeats),
        basis=pulse.basis)
    newpulse.tau = repeats*pulse.tau
    if not pulse.is_cached('control_matrix'):
        return newpulse
    phases_at = pulse.get_total_phases(pulse.omega)
    control_matrix_at = pulse.get_control_matrix(pulse.omega)
    newpulse.total_propagator = np.linalg.matrix_power(pulse.total_propagator, repeats)
    newpulse.cache_total_phases(pulse.omega)
    control_matrix_tot = numeric.calculate_control_matrix_periodic(
        phases_at, control_matrix_at, pulse.total_propagator_liouville, repeats, check_invertible)
    newpulse.cache_filter_function(pulse.omega, control_matrix_tot)
    return newpulse


# ------------------------------------------------------------------------------------------------
# batched concatenation
# ------------------------------------------------------------------------------------------------
class SequenceBatch:
    """Result of :func:`concatenate_many`: ``n_seq`` gate sequences over one library of pulses, held
    as stacked arrays (leading axis = sequence) instead of ``n_seq`` ``PulseSequence`` objects.

    Attributes: ``pulses`` (the library), ``indices`` (n_seq, L), ``omega``, ``tau`` (n_seq,),
    ``total_propagator`` (n_seq, d, d), ``total_propagator_liouville`` (n_seq, n_basis, n_basis),
    ``control_matrix`` (n_seq, n_nops, n_basis, n_omega) or ``None``, ``filter_function``
    (n_seq, n_nops, n_nops, n_omega) or ``None``, ``infidelities`` (if a spectrum was given).
    """

    def __init__(self, pulses, indices, omega):
        self.pulses = pulses
        self.indices = indices
        self.omega = omega
        self.tau = None
        self.total_propagator = None
        self.total_propagator_liouville = None
        self.control_matrix = None
        self.filter_function = None
        self.infidelities = None

    def __len__(self) -> int:
        return len(self.indices)

    @property
    def n_oper_identifiers(self):
        return self.pulses[0].n_oper_identifiers

    def infidelity(self, spectrum, n_oper_identifiers=None) -> ndarray:
        """``ff.infidelity`` for every sequence: (n_seq, n_sel) or (n_seq, n_sel, n_sel)."""
        if self.filter_function is None:
            raise util.CalculationError('concatenate_many was run with calc_filter_function=False')
        idx = util.get_indices_from_identifiers(self.n_oper_identifiers, n_oper_identifiers)
        return numeric._integrate_against_spectrum(self.filter_function, np.asarray(spectrum),
                                                   self.omega, idx, self.pulses[0].d)

    def pulse(self, i: int) -> PulseSequence:
        """Sequence ``i`` as a full ``PulseSequence`` with everything computed here in its cache."""
        row = [k for k in self.indices[i] if k >= 0]
        new = concatenate_without_filter_function([self.pulses[k] for k in row])
        new.total_propagator = self.total_propagator[i]
        new.total_propagator_liouville = self.total_propagator_liouville[i]
        new._frequency_data['omega'] = np.array(self.omega, copy=True)
        new.cache_total_phases(new.omega)
        if self.control_matrix is not None:
            new._frequency_data['control_matrix'] = self.control_matrix[i]
        if self.filter_function is not None:
            new._frequency_data['filter_function'] = self.filter_function[i]
        return new


def concatenate_many(pulses, indices, spectrum=None, omega=None, calc_control_matrix: bool = True,
                     calc_filter_function: bool = True) -> SequenceBatch:
    r"""Concatenate many sequences of pulses drawn from one library in a single GPU call.

    ``indices[s]`` lists the positions in ``pulses`` of the gates of sequence ``s`` in temporal order
    (negative entries are padding), i.e. sequence ``s`` is ``concatenate(pulses[indices[s]])``.  All
    library pulses must carry the same noise operators and have (or be able to compute) control
    matrices on one frequency grid -- the situation of randomized benchmarking
    (``examples/randomized_benchmarking.py:70-91`` of the reference loops over ``ff.concatenate``).
    With ``spectrum`` the infidelities of all sequences are integrated in the same call
    (``batch.infidelities``).  ``calc_control_matrix=False`` / ``calc_filter_function=False`` skip the
    download of the respective (large) arrays.
    """
    pulses = tuple(pulses)
    if not pulses or not all(isinstance(pls, PulseSequence) for pls in pulses):
        raise TypeError('Can only concatenate PulseSequences!')
    indices = np.ascontiguousarray(np.atleast_2d(np.asarray(indices)), dtype=np.int32)
    if indices.ndim != 2 or indices.shape[1] == 0:
        raise ValueError(f'Expected indices of shape (n_seq, L), not {indices.shape}')
    if indices.max(initial=-1) >= len(pulses):
        raise ValueError(f'indices refer to pulse {indices.max()} but only {len(pulses)} given')
    if (indices < 0).all(axis=1).any():
        raise ValueError('Every sequence needs at least one gate')
    first = pulses[0]
    if len(set(pls.d for pls in pulses)) != 1:
        raise ValueError('Trying to concatenate PulseSequence instances with different dimension!')
    bases = _unique_by_identity(pls.basis for pls in pulses)
    if len(bases) > 1 and not util.all_array_equal(bases):
        raise ValueError('Trying to concatenate PulseSequence instances with different bases!')
    if not first.basis.isherm:
        raise ValueError('concatenate_many needs a Hermitian basis (real Liouville propagators)')
    for pls in pulses[1:]:
        if (not np.array_equal(pls.n_oper_identifiers, first.n_oper_identifiers)
                or _oper_hashes(pls, 'noise') != _oper_hashes(first, 'noise')):
            raise ValueError('concatenate_many requires all pulses to have the same noise operators; '
                             + 'use concatenate for the general case')
    if omega is None:
        grids = _unique_by_identity(pls.omega for pls in pulses)
        if any(g is None for g in grids) or (len(grids) > 1 and not util.all_array_equal(grids)):
            raise ValueError('Not all pulses have the same frequencies cached and none were '
                             + 'supplied!')
        omega = first.omega
    omega = _lib.as_f64(omega)

    lib_B = _lib.as_c128(np.array([pls.get_control_matrix(omega) for pls in pulses]))
    lib_phase = _lib.as_c128(np.array([pls.get_total_phases(omega) for pls in pulses]))
    lib_liouville = _lib.as_f64(np.array([np.real(pls.total_propagator_liouville)
                                          for pls in pulses]))
    lib_U = _lib.as_c128(np.array([pls.total_propagator for pls in pulses]))
    basis = _lib.as_c128(np.asarray(first.basis))
    n_lib, n_nops, n_basis, n_omega = lib_B.shape
    d = first.c_opers.shape[-1]   # matrix size; first.d only normalises the infidelity
    n_seq, L = indices.shape

    S = None
    s_ndim = s_complex = 0
    if spectrum is not None:
        spectrum = util.parse_spectrum(np.asarray(spectrum), omega, np.arange(n_nops))
        s_ndim, s_complex = spectrum.ndim, int(np.iscomplexobj(spectrum))
        S = _lib.as_c128(spectrum) if s_complex else _lib.as_f64(spectrum)

    batch = SequenceBatch(pulses, indices, omega)
    taus = np.array([pls.tau for pls in pulses] + [0.0])
    batch.tau = taus[indices].sum(axis=1)
    batch.total_propagator = np.empty((n_seq, d, d), dtype=np.complex128)
    liouville = np.empty((n_seq, n_basis, n_basis), dtype=np.complex128)
    if calc_control_matrix:
        batch.control_matrix = _lib.empty((n_seq, n_nops, n_basis, n_omega))
    if calc_filter_function:
        batch.filter_function = _lib.empty((n_seq, n_nops, n_nops, n_omega))
    if S is not None:
        batch.infidelities = np.empty((n_seq,) + ((n_nops, n_nops) if s_ndim == 3 else (n_nops,)))

    ctx = _lib.context()
    p = _lib.ptr
    rows_per_call = max(1, 65535//n_nops)
    for lo in range(0, n_seq, rows_per_call):
        hi = min(n_seq, lo + rows_per_call)
        sub = lambda a: None if a is None else p(a[lo:hi])  # noqa: E731
        _lib.check(ctx, _lib.lib().ffb_concatenate_many(
            ctx, hi - lo, L, n_lib, d, n_nops, n_basis, n_omega, p(indices[lo:hi]), p(lib_B),
            p(lib_phase), p(lib_liouville), p(lib_U), p(basis), p(S), s_ndim, s_complex, p(omega),
            sub(batch.total_propagator), sub(liouville), sub(batch.control_matrix),
            sub(batch.filter_function), sub(batch.infidelities), None, None))
    batch.total_propagator_liouville = normalize_liouville_columns(
        np.ascontiguousarray(liouville.real), first.basis)
    if batch.infidelities is not None and first.d != d:
        batch.infidelities *= d/first.d
    return batch

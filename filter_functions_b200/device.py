"""Device-resident variant of the path: inputs and results stay in HBM between calls.

``DevicePulse`` uploads a pulse once and then runs diagonalize -> control matrix -> filter function
-> infidelity with the ``ffb_dev_*`` entry points of the C ABI on torch's current CUDA stream.  torch
is used for what it is good at -- device memory, streams, events and (in ``distributed.py``)
NCCL -- not for arithmetic.  This is what ``bench.py`` times for the HBM-resident ``value`` and what
the frequency-sharded multi-GPU path is built from.
"""
import numpy as np

from . import _lib

__all__ = ['DevicePulse']


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.FFBError('DevicePulse needs a CUDA device; filter_functions_b200 has no CPU '
                            'fallback')
    return torch


class DevicePulse:
    """One pulse (control + noise Hamiltonian, basis) and one frequency grid resident on a GPU.

    Parameters mirror ``PulseSequence.from_arrays`` plus ``omega`` and an optional ``spectrum`` of
    shape (n_omega,), (n_nops, n_omega) or (n_nops, n_nops, n_omega).
    """

    def __init__(self, c_opers, c_coeffs, n_opers, n_coeffs, dt, basis, omega, spectrum=None,
                 device=None):
        torch = _torch()
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.ctx = _lib.context(self.device.index)
        c_opers = np.ascontiguousarray(c_opers, dtype=np.complex128)
        n_opers = np.ascontiguousarray(n_opers, dtype=np.complex128)
        basis = np.ascontiguousarray(np.asarray(basis), dtype=np.complex128)
        dt = np.ascontiguousarray(dt, dtype=np.float64)
        self.G, self.d = len(dt), c_opers.shape[-1]
        self.n_cops, self.n_nops, self.n_basis = len(c_opers), len(n_opers), len(basis)
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        self.n_omega = len(omega)

        def herm(m):
            return bool((m == m.conj().swapaxes(-1, -2)).all())
        b0 = np.asarray(basis[0])
        ident0 = bool((b0 == b0[0, 0].real*np.eye(self.d)).all())   # exactly c * identity, c real
        self.herm_flags = (1 if herm(n_opers) else 0) | (2 if herm(basis) else 0) | (4 if ident0 else 0)

        up = self._upload
        self.c_opers, self.n_opers, self.basis = up(c_opers), up(n_opers), up(basis)
        self.c_coeffs = up(np.ascontiguousarray(c_coeffs, dtype=np.float64))
        self.n_coeffs = up(np.ascontiguousarray(n_coeffs, dtype=np.float64))
        self.dt = up(dt)
        self.t = up(np.concatenate(([0.0], dt.cumsum())))
        self.omega = up(omega)
        self.spectrum = None
        self.spectrum_ndim = 0
        self.spectrum_complex = False
        if spectrum is not None:
            spectrum = np.asarray(spectrum)
            self.spectrum_ndim = spectrum.ndim
            self.spectrum_complex = np.iscomplexobj(spectrum)
            self.spectrum = up(np.ascontiguousarray(
                spectrum, dtype=np.complex128 if self.spectrum_complex else np.float64))
        G, d = self.G, self.d
        empty = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype,  # noqa: E731
                                                                device=self.device)
        self.eigvals = empty(G, d)
        self.eigvecs = empty(G, d, d, dtype=torch.complex128)
        self.propagators = empty(G + 1, d, d, dtype=torch.complex128)
        self.control_matrix = empty(self.n_nops, self.n_basis, self.n_omega, dtype=torch.complex128)
        self.filter_function = empty(self.n_nops, self.n_nops, self.n_omega, dtype=torch.complex128)
        n_inf = self.n_nops**2 if self.spectrum_ndim == 3 else self.n_nops
        self.infidelity = empty(n_inf)
        self.idx = torch.arange(self.n_nops, dtype=torch.int32, device=self.device)

    def _upload(self, arr):
        return self.torch.from_numpy(arr).to(self.device)

    def bind_stream(self):
        """Make the library enqueue on torch's current stream (so torch events time it)."""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.ctx, _lib.lib().ffb_set_stream(self.ctx, stream, 1))

    def diagonalize(self):
        L, p = _lib.lib(), lambda t: t.data_ptr()  # noqa: E731
        _lib.check(self.ctx, L.ffb_dev_diagonalize(
            self.ctx, self.G, self.d, self.n_cops, p(self.c_opers), p(self.c_coeffs), p(self.dt),
            p(self.eigvals), p(self.eigvecs), p(self.propagators)))

    def calculate_control_matrix(self):
        L, p = _lib.lib(), lambda t: t.data_ptr()  # noqa: E731
        _lib.check(self.ctx, L.ffb_dev_control_matrix_from_scratch(
            self.ctx, self.G, self.d, self.n_nops, self.n_basis, self.n_omega, p(self.eigvals),
            p(self.eigvecs), p(self.propagators), p(self.omega), p(self.basis), p(self.n_opers),
            p(self.n_coeffs), p(self.dt), p(self.t), self.herm_flags, p(self.control_matrix)))

    def calculate_filter_function(self):
        L, p = _lib.lib(), lambda t: t.data_ptr()  # noqa: E731
        _lib.check(self.ctx, L.ffb_dev_filter_function(
            self.ctx, 1, self.n_nops, self.n_basis, self.n_omega, p(self.control_matrix), 0,
            p(self.filter_function)))

    def calculate_infidelity(self):
        if self.spectrum is None:
            raise ValueError('no spectrum was given')
        L, p = _lib.lib(), lambda t: t.data_ptr()  # noqa: E731
        _lib.check(self.ctx, L.ffb_dev_infidelity(
            self.ctx, 1, self.n_nops, self.n_nops, p(self.idx), self.n_omega,
            p(self.filter_function), p(self.spectrum), self.spectrum_ndim,
            int(self.spectrum_complex), p(self.omega), self.d, p(self.infidelity)))

    def step(self):
        """One pass of the hot path, all on device, asynchronous on the bound stream."""
        self.diagonalize()
        self.calculate_control_matrix()
        self.calculate_filter_function()
        if self.spectrum is not None:
            self.calculate_infidelity()

"""Stand-in for the ``opt_einsum`` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY. It exists so that the *unmodified* reference at /root/reference can be
imported in the build container by ``oracle/gen_golden.py`` and ``oracle/validate_oracle.py``.
Nothing in ``filter_functions_b200`` imports it.

opt_einsum with its default NumPy backend lowers a contraction to a sequence of
``numpy.tensordot`` / ``numpy.einsum`` calls along a searched path; ``numpy.einsum(...,
optimize=...)`` does the same thing with NumPy's own path search, so results agree to summation
order (1e-16 level) and timing is representative.
"""
import numpy as np

from . import contract as _contract_module  # noqa: F401  (gradient.py imports the submodule)
from .contract import ContractExpression, contract, contract_expression  # noqa: F401

__all__ = ['contract', 'contract_expression', 'ContractExpression']

"""``opt_einsum.contract`` stand-in backed by ``numpy.einsum`` (see package docstring)."""
import numpy as np


def _dense(x):
    return x.todense() if hasattr(x, 'todense') else x


def contract(subscripts, *operands, optimize=True, backend=None, out=None, **_):
    operands = [_dense(op) for op in operands]
    if isinstance(optimize, (list, tuple)) and optimize and isinstance(optimize[0], tuple):
        # explicit pairwise path in opt_einsum format -> numpy format
        optimize = ['einsum_path', *optimize]
    elif optimize is None or optimize is False:
        optimize = False
    elif not isinstance(optimize, list):
        optimize = 'optimal' if len(operands) <= 5 else 'greedy'
    kwargs = {} if out is None else {'out': out}
    result = np.einsum(subscripts, *operands, optimize=optimize, **kwargs)
    if backend == 'sparse':
        import sparse
        return sparse.COO.from_numpy(np.asarray(result))
    return result


class ContractExpression:
    """Callable returned by :func:`contract_expression`; caches the contraction path."""

    def __init__(self, subscripts, *shapes, optimize=True, **_):
        self.subscripts = subscripts
        self.shapes = shapes
        dummies = [np.empty(shape) for shape in shapes]
        self.path = np.einsum_path(subscripts, *dummies, optimize='optimal')[0]

    def __call__(self, *operands, out=None, backend=None, **_):
        operands = [_dense(op) for op in operands]
        kwargs = {} if out is None else {'out': out}
        return np.einsum(self.subscripts, *operands, optimize=self.path, **kwargs)


def contract_expression(subscripts, *shapes, optimize=True, **kwargs):
    return ContractExpression(subscripts, *shapes, optimize=optimize, **kwargs)

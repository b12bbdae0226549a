import numpy


def einsum(*args, **kwargs):
    """pyscf.lib.einsum contracts pairwise through tensordot/dgemm."""
    kwargs.setdefault("optimize", True)
    return numpy.einsum(*args, **kwargs)

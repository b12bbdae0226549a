"""Fermi factors and zeroth-order thermodynamics (host scalars/vectors).

Restates the cqcpy.ft_utils / cqcpy.utils helpers the reference calls at
kelvin/ccsd.py:628-638,730-738, kelvin/cc_utils.py:571-572.
These are O(n) host quantities; D1/D2 return device tensors because they feed
the integration kernels.
"""
import numpy
import torch


def ff(beta, eps, mu):
    return 1.0/(numpy.exp(beta*(eps - mu)) + 1.0)


def ffv(beta, eps, mu):
    return 1.0/(numpy.exp(-beta*(eps - mu)) + 1.0)


def GP0(beta, eps, mu):
    return -numpy.log(1.0 + numpy.exp(-beta*(eps - mu)))/beta


def uGP0(beta, ea, eb, mu):
    return GP0(beta, ea, mu), GP0(beta, eb, mu)


def dGP0(beta, eps, mu):
    x = beta*(eps - mu)
    return numpy.log(1.0 + numpy.exp(-x))/(beta*beta) + (eps - mu)/(beta*(numpy.exp(x) + 1.0))


def HtoK(T):
    return T*315775.128914


def _t(e, dev):
    return torch.as_tensor(numpy.asarray(e, dtype=numpy.float64)).to(dev)


def D1(ev, eo, device=None):
    """D1[a,i] = ev[a] - eo[i]."""
    from . import _lib
    dev = device or _lib.device()
    ev, eo = _t(ev, dev), _t(eo, dev)
    return (ev[:, None] - eo[None, :]).contiguous()


def D2(ev, eo, device=None):
    """D2[a,b,i,j] = ev[a] + ev[b] - eo[i] - eo[j]."""
    return D2u(ev, ev, eo, eo, device)


def D2u(eva, evb, eoa, eob, device=None):
    """D2u[a,B,i,J] = eva[a] + evb[B] - eoa[i] - eob[J]."""
    from . import _lib
    dev = device or _lib.device()
    eva, evb, eoa, eob = _t(eva, dev), _t(evb, dev), _t(eoa, dev), _t(eob, dev)
    return (eva[:, None, None, None] + evb[None, :, None, None]
            - eoa[None, None, :, None] - eob[None, None, None, :]).contiguous()

"""Imaginary-time grids and the exp-weighted time integration.

Drop-in for kelvin/quadrature.py (same function names, argument order and
(ti, g, G) conventions).  The grids are O(ng^2) host scalars; the integration
(`int_tbar1/2`, `int_L1/2`, reference: a Python loop over grid points building
two ng*N temporaries per point, kelvin/quadrature.py:292-345) is one fused CUDA
kernel that reads each amplitude once and builds the propagator weights
exp(D*(tau_x - tau_y)) on the fly (kb200_int_tbar / kb200_int_L).
"""
import ctypes
import math

import numpy
import torch

from . import _lib

# mode 1: ng-1 exponentials per element, weights as running products;
# mode 0: one exponential per (y, x) pair, the reference's literal formula.
INT_MODE = 1


# ---------------------------------------------------------------------------
# weights on a uniform grid in the integration variable s
# ---------------------------------------------------------------------------
def get_G(ng, delta):
    """Running-Simpson weight matrix, row y integrates 0..s_y
    (kelvin/quadrature.py:32-42): row 1 trapezoid, row y = row y-2 + Simpson panel."""
    G = numpy.zeros((ng, ng))
    if ng > 1:
        G[1, 0] = G[1, 1] = 0.5*delta
    for y in range(2, ng):
        G[y] = G[y - 2]
        G[y, y - 2] += delta/3.0
        G[y, y - 1] += 4.0*delta/3.0
        G[y, y] += delta/3.0
    return G


def get_gint(ng, delta):
    """Composite Simpson weights for 0..s_max; trapezoid on the first panel
    when ng is even (kelvin/quadrature.py:45-60)."""
    g = numpy.zeros(ng)
    start = 2
    if ng % 2 == 0:
        g[0] += 0.5*delta
        g[1] += 0.5*delta
        start = 3
    for y in range(start, ng, 2):
        g[y - 2] += delta/3.0
        g[y - 1] += 4.0*delta/3.0
        g[y] += delta/3.0
    return g


def _stretch(quad, ng):
    """(s grid, delta, tau(s)/beta, dtau/ds / beta) of each Simpson-type rule
    (kelvin/quadrature.py:102-201)."""
    idx = numpy.arange(ng, dtype=float)
    if quad == 'ln':
        delta = (math.e - 1.0)/(ng - 1.0)
        s = numpy.asarray([float(i)*delta + 1.0 for i in range(ng)])
        return delta, numpy.log(s), 1.0/s
    if quad == 'sin':
        delta = numpy.pi/(ng - 1.0)
        s = numpy.asarray([float(i)*delta - numpy.pi/2 for i in range(ng)])
        return delta, (numpy.sin(s) + 1.0)/2.0, numpy.cos(s)/2.0
    if quad == 'exp':
        delta = numpy.log(2.0)/(ng - 1.0)
        s = idx*delta
        return delta, numpy.exp(s) - 1.0, numpy.exp(s)
    if quad in ('quad', 'cub', 'quar'):
        n = {'quad': 2, 'cub': 3, 'quar': 4}[quad]
        delta = 1.0/(ng - 1.0)
        s = idx*delta
        return delta, (numpy.power(s, n) + s)/2.0, (numpy.power(s, n - 1)*float(n) + 1.0)/2.0
    raise Exception("Unrecognized quadrature rule: {}".format(quad))


def simpsons(ng, beta):
    """kelvin/quadrature.py:102-107."""
    delta = beta/(ng - 1.0)
    ti = numpy.asarray([float(i)*delta for i in range(ng)])
    return ti, get_gint(ng, delta), get_G(ng, delta)


def d_simpsons(ng, beta):
    """beta-derivative of the 'lin' weights (kelvin/quadrature.py:108-113)."""
    delta = beta/(ng - 1.0)
    ddelta = delta/beta
    return get_gint(ng, ddelta), get_G(ng, ddelta)


def midpoint(ng, beta):
    """kelvin/quadrature.py:16-29,94-99."""
    delta = beta/ng
    ti = numpy.asarray([float(i)*delta + delta/2 for i in range(ng)])
    G = numpy.zeros((ng, ng))
    for y in range(ng):
        G[y, :y + 1] = delta
    g = numpy.full(ng, delta)
    return ti, g, G


def _left_weights(ng, delta):
    g = numpy.zeros(ng)
    g[:ng - 1] = delta
    G = numpy.zeros((ng, ng))
    for i in range(ng):
        for j in range(i - 1):
            G[i, j] = delta
    return g, G


def left(ng, beta):
    """kelvin/quadrature.py:63-83 (note G[i, :i-1], as in the reference)."""
    delta = beta/(ng - 1.)
    ti = numpy.asarray([float(i)*delta for i in range(ng)])
    g, G = _left_weights(ng, delta)
    return ti, g, G


def ft_quad(ng, beta, quad):
    """(ti, g, G) -- kelvin/quadrature.py:215-235."""
    if quad == 'lin':
        return simpsons(ng, beta)
    if quad == 'mid':
        return midpoint(ng, beta)
    if quad == 'L':
        return left(ng, beta)
    delta, tau, jac = _stretch(quad, ng)
    g = beta*get_gint(ng, delta)*jac
    G = get_G(ng, delta)
    for i in range(ng):
        G[i] = beta*G[i]*jac
    return beta*tau, g, G


def d_ft_quad(ng, beta, quad):
    """(gd, Gd) = d(g, G)/d(beta) -- kelvin/quadrature.py:238-256."""
    if quad == 'lin':
        return d_simpsons(ng, beta)
    if quad == 'L':
        delta = beta/(ng - 1.0)
        return _left_weights(ng, delta/beta)
    if quad == 'mid':
        raise Exception("Unrecognized quadrature rule: {}".format(quad))
    delta, tau, jac = _stretch(quad, ng)
    g = get_gint(ng, delta)*jac
    G = get_G(ng, delta)
    for i in range(ng):
        G[i] = G[i]*jac
    return g, G


# ---------------------------------------------------------------------------
# device integration
# ---------------------------------------------------------------------------
def _small(x, dev):
    return _lib.const_dev(x, dev)


def _host(x):
    """Host float64 copy of a small grid array (None for device tensors: the kernels then take
    their generic path) as a ctypes pointer, kept alive by the returned array."""
    if x is None or isinstance(x, torch.Tensor):
        return None, None
    a = numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.float64))
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _mode_bits(G, mode):
    """Kernel mode word: bit 0 = product-form weights, bit 1 = G is lower triangular."""
    m = INT_MODE if mode is None else mode
    if not isinstance(G, torch.Tensor):
        Gn = numpy.asarray(G)
        if not numpy.any(numpy.triu(Gn, 1)):
            m |= 2
    return m


def int_tbar(ng, tbar, ti, D, G, out=None, mode=None, rows=None):
    """out[y] = sum_x G[y,x] * exp(D*(ti[x]-ti[y]))_{x<y} * tbar[x]
    for any amplitude rank (kelvin/quadrature.py:292-317).  rows=(y0, y1)
    restricts the output to those grid points.  The grid points of tbar / out may be rows of a
    wider buffer (any leading stride)."""
    lib = _lib.load()
    dev = _lib.device()
    tbar = _lib.as_dev_rows(tbar, dev)
    D = _lib.as_dev(D, dev)
    if tbar.shape[0] != ng or tuple(tbar.shape[1:]) != tuple(D.shape):
        raise Exception("int_tbar: shape mismatch {} vs ng={} D{}".format(
            tuple(tbar.shape), ng, tuple(D.shape)))
    y0, y1 = (0, ng) if rows is None else rows
    if out is None:
        out = torch.empty((y1 - y0,) + tuple(D.shape), dtype=torch.float64, device=dev)
    n = D.numel()
    if n == 0 or y1 == y0:
        return out
    tid, Gd = _small(ti, dev), _small(G, dev)
    (ka, th), (kb, Gh) = _host(ti), _host(G)
    rc = lib.kb200_int_tbar_strided_h(ng, n, _lib.ptr(tbar), _lib.row_stride(tbar), _lib.ptr(D),
                                      _lib.ptr(tid), _lib.ptr(Gd), _lib.ptr(out),
                                      _lib.row_stride(out), y0, y1, _mode_bits(G, mode), th, Gh,
                                      _lib.stream_ptr())
    _lib.check(rc, "kb200_int_tbar")
    return out


def int_tbar_update(ng, tbar, ti, D, G, amp, alpha, out4, g=None, W=None, T1x=None, T1y=None,
                    c2=1.0, c11=0.0, mode=None, rows=None):
    """The fused update of one amplitude block (kb200_int_tbar_update): integrates tbar
    (kelvin/quadrature.py:292-317), measures ||new - amp||^2 and ||amp||^2, damps amp in place
    and measures its new norm (kelvin/cc_utils.py:278-295), and accumulates the block's term of
    the energy functional sum_y g_y (c2 amp + c11 T1x T1y).W (kelvin/ft_cc_energy.py:35-72) --
    all in one pass; the integrated amplitudes are never stored.  out4: 4 device doubles."""
    lib = _lib.load()
    dev = _lib.device()
    tbar = _lib.as_dev_rows(tbar, dev)
    D = _lib.as_dev(D, dev)
    if tbar.shape[0] != ng or tuple(tbar.shape[1:]) != tuple(D.shape) \
            or tuple(amp.shape) != tuple(tbar.shape) or not amp[0].is_contiguous():
        raise Exception("int_tbar_update: shape mismatch")
    y0, y1 = (0, ng) if rows is None else rows
    n = D.numel()
    tid, Gd = _small(ti, dev), _small(G, dev)
    gd = _small(g, dev) if g is not None else None
    nvb = noa = nob = 1
    if T1x is not None:
        nvb, noa, nob = int(D.shape[1]), int(D.shape[2]), int(D.shape[3])
    (ka, th), (kb, gh), (kc, Gh) = _host(ti), _host(g), _host(G)
    rc = lib.kb200_int_tbar_update_h(
        ng, n, _lib.ptr(tbar), _lib.row_stride(tbar), _lib.ptr(D), _lib.ptr(tid), _lib.ptr(Gd),
        _lib.ptr(amp), _lib.row_stride(amp), y0, y1, alpha,
        _lib.ptr(W) if W is not None else None,
        _lib.ptr(T1x) if T1x is not None else None, _lib.ptr(T1y) if T1y is not None else None,
        _lib.row_stride(T1x) if T1x is not None else 0,
        _lib.row_stride(T1y) if T1y is not None else 0, nvb, noa, nob,
        _lib.ptr(gd) if gd is not None else None, c2, c11,
        out4 if isinstance(out4, int) else _lib.ptr(out4), _lib.ptr(_lib.reduce_scratch(dev)),
        _mode_bits(G, mode), th, gh, Gh, _lib.stream_ptr())
    _lib.check(rc, "kb200_int_tbar_update")


def int_tbar1(ng, t1bar, ti, D1, G):
    """Integrate t1bar with exponential factor (kelvin/quadrature.py:292)."""
    return int_tbar(ng, t1bar, ti, D1, G)


def int_tbar2(ng, t2bar, ti, D2, G):
    """Integrate t2bar with exponential factor (kelvin/quadrature.py:306)."""
    return int_tbar(ng, t2bar, ti, D2, G)


def int_tbar1_single(ng, ig, t1bar, ti, D1, G):
    """Integrate t1bar with exponential factor at the single grid point ig
    (kelvin/quadrature.py:348-357) = row ig of int_tbar1."""
    return int_tbar(ng, t1bar, ti, D1, G, rows=(ig, ig + 1))[0]


def int_tbar2_single(ng, ig, t2bar, ti, D2, G):
    """Integrate t2bar with exponential factor at the single grid point ig
    (kelvin/quadrature.py:360-369) = row ig of int_tbar2."""
    return int_tbar(ng, t2bar, ti, D2, G, rows=(ig, ig + 1))[0]


def int_L(ng, Lold, ti, D, g, G, out=None, mode=None, rows=None):
    """Lbar[s] = (1/g[s]) sum_y g[y] G[y,s] exp(D^T*(ti[s]-ti[y]))_{y>=s} L[y]
    with D indexed (v..,o..) and L indexed (o..,v..) (kelvin/quadrature.py:320-345)."""
    lib = _lib.load()
    dev = _lib.device()
    Lold = _lib.as_dev_rows(Lold, dev)
    D = _lib.as_dev(D, dev)
    r = D.dim()
    h = r // 2
    # L axes (o..., v...) <- D axes (v..., o...)
    perm = list(range(h, r)) + list(range(h))
    dshape = [D.shape[p] for p in perm]
    if Lold.shape[0] != ng or list(Lold.shape[1:]) != dshape:
        raise Exception("int_L: shape mismatch")
    dstr = [D.stride(p) for p in perm]
    dims = [1]*(4 - r) + dshape
    strs = [0]*(4 - r) + dstr
    cd = (ctypes.c_int32*4)(*dims)
    cs = (ctypes.c_int64*4)(*strs)
    s0, s1 = (0, ng) if rows is None else rows
    if out is None:
        out = torch.empty((s1 - s0,) + tuple(Lold.shape[1:]), dtype=torch.float64, device=dev)
    if out.numel() == 0:
        return out
    tid, gd, Gd = _small(ti, dev), _small(g, dev), _small(G, dev)
    (ka, th), (kb, gh), (kc, Gh) = _host(ti), _host(g), _host(G)
    rc = lib.kb200_int_L_strided_h(ng, cd, cs, _lib.ptr(Lold), _lib.row_stride(Lold), _lib.ptr(D),
                                   _lib.ptr(tid), _lib.ptr(gd), _lib.ptr(Gd), _lib.ptr(out),
                                   _lib.row_stride(out), s0, s1, _mode_bits(G, mode), th, gh, Gh,
                                   _lib.stream_ptr())
    _lib.check(rc, "kb200_int_L")
    return out


def int_L1(ng, L1old, ti, D1, g, G):
    """Return L1bar (kelvin/quadrature.py:320)."""
    return int_L(ng, L1old, ti, D1, g, G)


def int_L2(ng, L2old, ti, D2, g, G):
    """Return L2bar (kelvin/quadrature.py:334)."""
    return int_L(ng, L2old, ti, D2, g, G)

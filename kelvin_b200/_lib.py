"""ctypes binding of libkb200.so (include/kelvin_b200.h).

There is deliberately NO fallback: if the shared object is missing, or CUDA is
unavailable, every compute entry point raises.  The library is built in-tree by
``__graft_entry__.build()`` (nvcc, sm_100a).
"""
import ctypes
import os

import torch

from .plan import kb200_op

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KB200_LIB", os.path.join(_HERE, "libkb200.so"))

_lib = None

EXPORTS = (
    "kb200_version", "kb200_last_error", "kb200_launch_count", "kb200_launch_count_reset",
    "kb200_plan_workspace_bytes", "kb200_plan_run", "kb200_plan_run_timed", "kb200_int_tbar", "kb200_int_L", "kb200_int_tbar_rows", "kb200_int_L_rows",
    "kb200_reduce_scratch_doubles", "kb200_energy_pair", "kb200_dot_g", "kb200_damp_norms",
    "kb200_dress4", "kb200_dress2", "kb200_gather4", "kb200_scatter4_add", "kb200_gsum", "kb200_scale_by", "kb200_dot_keep",
    "kb200_max_absdiff", "kb200_set_plan_streams",
    "kb200_int_tbar_strided", "kb200_int_tbar_update", "kb200_int_L_strided",
    "kb200_damp_norms_rows", "kb200_launch_count_add",
    "kb200_int_tbar_strided_h", "kb200_int_tbar_update_h", "kb200_int_L_strided_h",
)


class KB200Error(RuntimeError):
    pass


def load():
    """Load the C-ABI library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KB200Error(
            "libkb200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "kelvin_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.kb200_version.restype = ctypes.c_int
    lib.kb200_last_error.restype = ctypes.c_char_p
    lib.kb200_launch_count.restype = i64
    lib.kb200_launch_count_reset.restype = None
    lib.kb200_launch_count_add.restype = None
    lib.kb200_launch_count_add.argtypes = [i64]
    lib.kb200_reduce_scratch_doubles.restype = i64
    lib.kb200_plan_workspace_bytes.restype = i64
    lib.kb200_plan_workspace_bytes.argtypes = [ctypes.POINTER(kb200_op), ctypes.c_int]
    lib.kb200_plan_run.argtypes = [ctypes.POINTER(kb200_op), ctypes.c_int, vp,
                                   ctypes.POINTER(vp), ctypes.c_int, vp, i64, vp]
    lib.kb200_plan_run_timed.argtypes = [ctypes.POINTER(kb200_op), ctypes.c_int, vp,
                                         ctypes.POINTER(vp), ctypes.c_int, vp, i64, vp,
                                         ctypes.POINTER(ctypes.c_float)]
    lib.kb200_int_tbar.argtypes = [ctypes.c_int, i64, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    lib.kb200_int_L.argtypes = [ctypes.c_int, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    lib.kb200_int_tbar_rows.argtypes = [ctypes.c_int, i64, vp, vp, vp, vp, vp, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, vp]
    lib.kb200_int_L_rows.argtypes = [ctypes.c_int, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                     vp, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, vp]
    lib.kb200_int_tbar_strided.argtypes = [ctypes.c_int, i64, vp, i64, vp, vp, vp, vp, i64,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.kb200_int_tbar_update.argtypes = [ctypes.c_int, i64, vp, i64, vp, vp, vp, vp, i64,
                                          ctypes.c_int, ctypes.c_int, dbl, vp, vp, vp, i64, i64,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, dbl, dbl,
                                          vp, vp, ctypes.c_int, vp]
    lib.kb200_int_L_strided.argtypes = [ctypes.c_int, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                        vp, i64, vp, vp, vp, vp, vp, i64, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, vp]
    lib.kb200_int_tbar_strided_h.argtypes = [ctypes.c_int, i64, vp, i64, vp, vp, vp, vp, i64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]
    lib.kb200_int_tbar_update_h.argtypes = [ctypes.c_int, i64, vp, i64, vp, vp, vp, vp, i64,
                                            ctypes.c_int, ctypes.c_int, dbl, vp, vp, vp, i64, i64,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, dbl, dbl,
                                            vp, vp, ctypes.c_int, vp, vp, vp, vp]
    lib.kb200_int_L_strided_h.argtypes = [ctypes.c_int, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                          vp, i64, vp, vp, vp, vp, vp, i64, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, vp, vp, vp, vp]
    lib.kb200_energy_pair.argtypes = [ctypes.c_int] * 5 + [vp, vp, vp, vp, vp, dbl, dbl, vp, vp, vp]
    lib.kb200_dot_g.argtypes = [ctypes.c_int, i64, vp, vp, vp, vp, vp, vp]
    lib.kb200_damp_norms.argtypes = [i64, vp, vp, dbl, vp, vp, vp]
    lib.kb200_damp_norms_rows.argtypes = [ctypes.c_int, i64, vp, i64, vp, i64, dbl, vp, vp, vp]
    lib.kb200_dress4.argtypes = [ctypes.POINTER(i32), vp, vp, vp, vp, vp, vp, vp]
    lib.kb200_dress2.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    lib.kb200_gather4.argtypes = [ctypes.POINTER(i32), ctypes.POINTER(i64), vp,
                                  ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
    lib.kb200_scatter4_add.argtypes = [ctypes.POINTER(i32), ctypes.POINTER(i64), vp,
                                       ctypes.POINTER(vp), ctypes.POINTER(vp), dbl, vp, vp]
    lib.kb200_dot_keep.argtypes = [ctypes.c_int, ctypes.POINTER(i32), ctypes.POINTER(i64),
                                   ctypes.POINTER(i64), vp, vp, dbl, dbl, vp, vp, vp]
    lib.kb200_gsum.argtypes = [ctypes.c_int, i64, vp, vp, vp, vp]
    lib.kb200_scale_by.argtypes = [ctypes.c_int, i64, vp, vp, vp]
    lib.kb200_set_plan_streams.argtypes = [ctypes.c_int]
    lib.kb200_max_absdiff.argtypes = [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(i64),
                                      ctypes.POINTER(i64), vp, vp, vp, vp]
    for nm in EXPORTS:
        getattr(lib, nm)
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise KB200Error("%s failed (%d): %s" % (what, rc, load().kb200_last_error().decode()))


def device():
    if not torch.cuda.is_available():
        raise KB200Error("kelvin_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def as_dev(x, dev=None):
    """NumPy / torch (any device) -> contiguous float64 CUDA tensor."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(x, dtype=torch.float64).to(dev).contiguous()


def as_dev_rows(x, dev=None):
    """As as_dev, but a device tensor whose grid points (leading axis) are rows of a wider
    buffer is taken as it is: every kernel that walks the grid takes a row stride."""
    dev = dev or device()
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.dim() >= 1 \
            and x.shape[0] > 0 and x[0].is_contiguous() \
            and (x.shape[0] == 1 or x.stride(0) >= x[0].numel()):
        return x
    return as_dev(x, dev)


def row_stride(x):
    return int(x.stride(0)) if x.shape[0] > 1 else int(x[0].numel())


_const_cache = {}


def const_dev(x, dev=None):
    """Device copy of a small host array (grids, weights), cached by value so that the
    iteration loops issue no synchronous host->device copies."""
    import numpy
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    a = numpy.ascontiguousarray(numpy.asarray(x, dtype=numpy.float64))
    if a.nbytes > (1 << 20):
        return torch.as_tensor(a).to(dev)
    key = (dev.index, a.shape, a.tobytes())
    t = _const_cache.get(key)
    if t is None:
        if len(_const_cache) > 256:
            _const_cache.clear()
        t = torch.as_tensor(a).to(dev)
        _const_cache[key] = t
    return t


_scratch = {}


def reduce_scratch(dev):
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    if key not in _scratch:
        n = load().kb200_reduce_scratch_doubles()
        _scratch[key] = torch.empty(int(n), dtype=torch.float64, device=dev)
    return _scratch[key]


def max_absdiff(X, Y):
    """(max |X - Y|, max |X|) of two CUDA float64 tensors of one shape (rank <= 5; strided /
    permuted views are read in place).  Synchronises: use it for one-off checks."""
    lib = load()
    if tuple(X.shape) != tuple(Y.shape) or X.dim() > 5:
        raise KB200Error("max_absdiff: shapes %s vs %s" % (tuple(X.shape), tuple(Y.shape)))
    if X.numel() == 0:
        return 0.0, 0.0
    pad = 5 - X.dim()
    dims = (ctypes.c_int32*5)(*([1]*pad + list(X.shape)))
    sx = (ctypes.c_int64*5)(*([0]*pad + list(X.stride())))
    sy = (ctypes.c_int64*5)(*([0]*pad + list(Y.stride())))
    out = torch.empty(2, dtype=torch.float64, device=X.device)
    rc = lib.kb200_max_absdiff(dims, sx, sy, ptr(X), ptr(Y), ptr(out), stream_ptr())
    check(rc, "kb200_max_absdiff")
    d, m = out.cpu().tolist()
    return float(d), float(m)


def dot_keep(A, la, B, lb, keep, alpha=1.0, out=None, beta=0.0):
    """out[keep] (+)= alpha * einsum('<la>,<lb>-><keep>', A, B) for tensors of rank <= 5
    sharing the same index letters (any order); `keep` is one letter."""
    lib = load()
    dev = A.device
    assert sorted(la) == sorted(lb) and keep in la and len(la) <= 5
    da = dict(zip(la, A.shape))
    sa = dict(zip(la, A.stride()))
    sb = dict(zip(lb, B.stride()))
    for l, n in zip(lb, B.shape):
        assert da[l] == n, "dot_keep: shape mismatch on %s" % l
    rest = [l for l in la if l != keep]
    pad = 4 - len(rest)
    dims = (ctypes.c_int32*4)(*([1]*pad + [da[l] for l in rest]))
    cA = (ctypes.c_int64*5)(*([sa[keep]] + [0]*pad + [sa[l] for l in rest]))
    cB = (ctypes.c_int64*5)(*([sb[keep]] + [0]*pad + [sb[l] for l in rest]))
    nk = da[keep]
    if out is None:
        out = torch.zeros(nk, dtype=torch.float64, device=dev)
    rc = lib.kb200_dot_keep(nk, dims, cA, cB, ptr(A), ptr(B), alpha, beta, ptr(out),
                            ptr(reduce_scratch(dev)), stream_ptr())
    check(rc, "kb200_dot_keep")
    return out


def _ptr_array(xs):
    return (ctypes.c_void_p*4)(*[None if x is None else x.data_ptr() for x in xs])


def index_dev(idx, dev):
    """int32 device copy of a host index list."""
    import numpy
    return torch.as_tensor(numpy.asarray(idx, dtype=numpy.int32)).to(dev)


def gather4(src, idx, scale, out_shape):
    """out[i0..i3] = src[idx0[i0], .., idx3[i3]] * prod scale_k[i_k]; idx_k / scale_k are device
    vectors or None (identity / 1).  kb200_gather4."""
    lib = load()
    out = torch.empty(tuple(out_shape), dtype=torch.float64, device=src.device)
    if out.numel() == 0:
        return out
    d = (ctypes.c_int32*4)(*out_shape)
    st = (ctypes.c_int64*4)(*src.stride())
    rc = lib.kb200_gather4(d, st, ptr(src), _ptr_array(idx), _ptr_array(scale), ptr(out), stream_ptr())
    check(rc, "kb200_gather4")
    return out


def scatter4_add(dst, src, idx, scale=(None,)*4, alpha=1.0, perm=(0, 1, 2, 3)):
    """dst.permute(perm)[idx0[i0], .., idx3[i3]] += alpha * src[i0..i3] * prod scale_k[i_k].
    kb200_scatter4_add."""
    lib = load()
    if src.numel() == 0:
        return dst
    src = src.contiguous()
    d = (ctypes.c_int32*4)(*src.shape)
    dstride = dst.stride()
    st = (ctypes.c_int64*4)(*[dstride[p] for p in perm])
    rc = lib.kb200_scatter4_add(d, st, ptr(src), _ptr_array(idx), _ptr_array(scale), alpha,
                                ptr(dst), stream_ptr())
    check(rc, "kb200_scatter4_add")
    return dst

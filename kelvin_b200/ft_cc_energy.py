"""FT-CC grand-potential functional on the GPU.

Drop-in for kelvin/ft_cc_energy.py:7-32 (``ft_cc_energy``) and :35-72
(``ft_ucc_energy``): Omega_cc = (1/beta) sum_y g_y [ t1.f_ov +
(c2*t2 + c11*t1 t1).<ij||ab> ].  The reference materialises an (ng, n^4)
temporary; here each term is one streaming reduction kernel
(kb200_energy_pair / kb200_dot_g) with a deterministic two-pass sum.
"""
import torch

from . import _lib, engine, plan as _plan


def _permute_plan(shape, src_letters, dst_letters):
    """Plan with the single op dst[dst_letters] = src[src_letters]."""
    key = ("permute", tuple(shape), src_letters, dst_letters)

    def build():
        op = _plan.ROp(("dst", dst_letters), 1.0, [("src", src_letters)])
        dims = dict(zip(src_letters, shape))
        shapes = {"src": tuple(shape), "dst": tuple(dims[l] for l in dst_letters)}
        return engine.Plan([op], "g", None, ["src"], ["dst"], name="permute",
                           shapes=shapes, batched={"src": False, "dst": False})
    return engine.cached(key, build)


def permute4(src, src_letters, dst_letters):
    """Index-permuted copy of a 4-index device tensor with the kb200 permute kernel."""
    dev = _lib.device()
    src = _lib.as_dev(src, dev)
    p = _permute_plan(tuple(src.shape), src_letters, dst_letters)
    dst = torch.empty(p.shapes["dst"], dtype=torch.float64, device=dev)
    p.run({"src": src, "dst": dst}, 1)
    return dst


def oovv_to_abij(eri):
    """<ij||ab> stored [i,j,a,b] -> [a,b,i,j] (the order T2 is stored in)."""
    return permute4(eri, "ijab", "abij")


def _gvec(g, dev):
    return _lib.const_dev(g, dev)


def energy_terms(terms1, terms2, g, dev):
    """terms1: list of (T1[ng,a,i], f_ai);  terms2: list of
    (T2[ng,a,b,i,j], T1x or None, T1y or None, I_abij, c2, c11).
    Returns the device vector of the individual contributions (unscaled)."""
    lib = _lib.load()
    gd = _gvec(g, dev)
    ng = gd.numel()
    out = torch.zeros(len(terms1) + len(terms2), dtype=torch.float64, device=dev)
    scratch = _lib.reduce_scratch(dev)
    k = 0
    for T1, fai in terms1:
        n = fai.numel()
        rc = lib.kb200_dot_g(ng, n, _lib.ptr(T1), _lib.ptr(fai), _lib.ptr(gd),
                             out.data_ptr() + 8*k, _lib.ptr(scratch), _lib.stream_ptr())
        _lib.check(rc, "kb200_dot_g")
        k += 1
    for T2, T1x, T1y, Iab, c2, c11 in terms2:
        nva, nvb, noa, nob = T2.shape[1:]
        rc = lib.kb200_energy_pair(
            ng, nva, nvb, noa, nob, _lib.ptr(T2),
            _lib.ptr(T1x) if T1x is not None else None,
            _lib.ptr(T1y) if T1y is not None else None,
            _lib.ptr(Iab), _lib.ptr(gd), c2, c11, out.data_ptr() + 8*k,
            _lib.ptr(scratch), _lib.stream_ptr())
        _lib.check(rc, "kb200_energy_pair")
        k += 1
    return out


def ft_cc_energy(T1, T2, f, eri, g, beta, Qterm=True, eri_abij=None):
    """Return the FT-CC free energy (kelvin/ft_cc_energy.py:7-32).
    f: F.ov [i,a]; eri: I.oovv [i,j,a,b] (or pass eri_abij pre-permuted)."""
    dev = _lib.device()
    T1 = _lib.as_dev(T1, dev)
    T2 = _lib.as_dev(T2, dev)
    fai = _lib.as_dev(f, dev).t().contiguous()
    Iab = eri_abij if eri_abij is not None else oovv_to_abij(eri)
    c11 = 0.5 if Qterm else 0.0
    parts = energy_terms([(T1, fai)],
                         [(T2, T1 if Qterm else None, T1 if Qterm else None, Iab, 0.25, c11)],
                         g, dev)
    return float(parts.sum().item())/beta


def ft_ucc_energy(T1a, T1b, T2aa, T2ab, T2bb, fa, fb, Ia, Ib, Iabab, g, beta, Qterm=True,
                  abij=None):
    """Return the unrestricted FT-CC free energy (kelvin/ft_cc_energy.py:35-72).
    fa/fb: F.ov blocks; Ia/Ib/Iabab: the oovv blocks.  With Qterm=False the
    reference contracts 0.25*T2aa (not T2bb) with Ib (quirk Q1, :58); kept."""
    dev = _lib.device()
    T1a, T1b, T2aa, T2ab, T2bb = [_lib.as_dev(x, dev) for x in (T1a, T1b, T2aa, T2ab, T2bb)]
    faT = _lib.as_dev(fa, dev).t().contiguous()
    fbT = _lib.as_dev(fb, dev).t().contiguous()
    if abij is None:
        abij = (oovv_to_abij(Ia), oovv_to_abij(Iabab), oovv_to_abij(Ib))
    Iaa, Iab, Ibb = abij
    if Qterm:
        t2 = [(T2aa, T1a, T1a, Iaa, 0.25, 0.5),
              (T2ab, T1a, T1b, Iab, 1.0, 1.0),
              (T2bb, T1b, T1b, Ibb, 0.25, 0.5)]
    else:
        t2 = [(T2aa, None, None, Iaa, 0.25, 0.0),
              (T2ab, None, None, Iab, 1.0, 0.0),
              (T2aa, None, None, Ibb, 0.25, 0.0)]
    parts = energy_terms([(T1a, faT), (T1b, fbT)], t2, g, dev)
    return float(parts.sum().item())/beta

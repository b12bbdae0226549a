"""Iteration loops and integral dressing for FT-CCSD on the GPU.

Drop-in for the on-path part of kelvin/cc_utils.py: ``ft_integrals`` (:569),
``uft_integrals`` (:696), ``form_new_ampl`` (:17), ``form_new_ampl_u`` (:52),
``ft_cc_iter`` (:111), ``ft_ucc_iter`` (:245), ``ft_lambda_iter`` (:414),
``ft_ulambda_iter`` (:483).  Same arguments, same log lines, same convergence
predicates (including the g/u differences in the residual norms, quirk Q5).
Damping and all norms run in one fused pass per tensor (kb200_damp_norms).
"""
import ctypes
import logging
import math
import time

import numpy
import torch

from . import _lib, _trace, ft_cc_energy, ft_cc_equations, ft_utils, quadrature
from .ov_blocks import one_e_blocks, two_e_blocks, two_e_blocks_full


# ---------------------------------------------------------------------------
# dressing
# ---------------------------------------------------------------------------
def _vec(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(numpy.asarray(x, dtype=numpy.float64)).to(dev)


def _dress4(eri, s0, s1, s2, s3):
    lib = _lib.load()
    out = torch.empty_like(eri)
    d = (ctypes.c_int32*4)(*eri.shape)
    rc = lib.kb200_dress4(d, _lib.ptr(eri), _lib.ptr(s0), _lib.ptr(s1), _lib.ptr(s2),
                          _lib.ptr(s3), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "kb200_dress4")
    return out


def _dress2(f, e, s0, s1):
    lib = _lib.load()
    out = torch.empty_like(f)
    rc = lib.kb200_dress2(f.shape[0], f.shape[1], _lib.ptr(f), _lib.ptr(e), _lib.ptr(s0),
                          _lib.ptr(s1), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "kb200_dress2")
    return out


def _dress_F(fmo, en, so, sv):
    """F - diag(e), each index scaled by sqrt(f) (o) / sqrt(1-f) (v)
    (kelvin/cc_utils.py:577-588)."""
    s = {"o": so, "v": sv}
    return one_e_blocks(*[_dress2(fmo, en, s[p[0]], s[p[1]]) for p in ("oo", "ov", "vo", "vv")])


def ft_integrals(sys, en, beta, mu):
    """Return one and two-electron integrals in the general spin orbital basis,
    pre-contracted with the Fermi factors (kelvin/cc_utils.py:569-602)."""
    dev = _lib.device()
    fo = ft_utils.ff(beta, en, mu)
    fv = ft_utils.ffv(beta, en, mu)
    so, sv = _vec(numpy.sqrt(fo), dev), _vec(numpy.sqrt(fv), dev)
    fmo = _lib.as_dev(sys.g_fock_tot(), dev)
    F = _dress_F(fmo, _vec(en, dev), so, sv)
    eri = _lib.as_dev(sys.g_aint_tot(), dev)
    s = {"o": so, "v": sv}
    I = two_e_blocks(**{p: _dress4(eri, s[p[0]], s[p[1]], s[p[2]], s[p[3]])
                        for p in two_e_blocks.names})
    return F, I


def uft_integrals(sys, ea, eb, beta, mu):
    """Unrestricted dressed integrals Fa, Fb, Ia, Ib, Iabab
    (kelvin/cc_utils.py:696-777)."""
    dev = _lib.device()
    sa = {"o": _vec(numpy.sqrt(ft_utils.ff(beta, ea, mu)), dev),
          "v": _vec(numpy.sqrt(ft_utils.ffv(beta, ea, mu)), dev)}
    sb = {"o": _vec(numpy.sqrt(ft_utils.ff(beta, eb, mu)), dev),
          "v": _vec(numpy.sqrt(ft_utils.ffv(beta, eb, mu)), dev)}
    fa, fb = sys.u_fock_tot()
    Fa = _dress_F(_lib.as_dev(fa, dev), _vec(ea, dev), sa["o"], sa["v"])
    Fb = _dress_F(_lib.as_dev(fb, dev), _vec(eb, dev), sb["o"], sb["v"])
    eriA, eriB, eriAB = sys.u_aint_tot()
    same = eriB is eriA
    eriA = _lib.as_dev(eriA, dev)
    eriB = eriA if same else _lib.as_dev(eriB, dev)
    eriAB = _lib.as_dev(eriAB, dev)
    Ia = two_e_blocks(**{p: _dress4(eriA, sa[p[0]], sa[p[1]], sa[p[2]], sa[p[3]])
                         for p in two_e_blocks.names})
    Ib = two_e_blocks(**{p: _dress4(eriB, sb[p[0]], sb[p[1]], sb[p[2]], sb[p[3]])
                         for p in two_e_blocks.names})
    Iabab = two_e_blocks_full(**{p: _dress4(eriAB, sa[p[0]], sb[p[1]], sa[p[2]], sb[p[3]])
                                 for p in two_e_blocks_full.names})
    return Fa, Fb, Ia, Ib, Iabab


def _sub2(X, rows, cols, dev):
    r = torch.as_tensor(numpy.asarray(rows, dtype=numpy.int64)).to(dev)
    c = torch.as_tensor(numpy.asarray(cols, dtype=numpy.int64)).to(dev)
    return X.index_select(0, r).index_select(1, c).contiguous()


class _Space(object):
    """Index list + sqrt(occupation) dressing vector of one (o|v, spin) active set."""
    def __init__(self, idx, f, dev):
        self.host = list(idx)
        self.idx = _lib.index_dev(idx, dev)
        self.s = _vec(numpy.sqrt(numpy.asarray(f, dtype=numpy.float64)), dev)
        self.n = len(self.host)


def _active_F(fmo, en, so, sv, dev):
    """Dressed F - diag(e) restricted to the active sets (kelvin/cc_utils.py:836-840)."""
    n = fmo.shape[0]
    one = torch.ones(n, dtype=torch.float64, device=dev)
    base = _dress2(fmo, _vec(en, dev), one, one)
    sp = {"o": so, "v": sv}
    blocks = []
    for p in ("oo", "ov", "vo", "vv"):
        a, b = sp[p[0]], sp[p[1]]
        blocks.append((_sub2(base, a.host, b.host, dev)*a.s[:, None]*b.s[None, :]).contiguous())
    return one_e_blocks(*blocks)


def _active_I(eri, spaces, names, cls):
    """spaces: per index position a dict {'o': _Space, 'v': _Space}."""
    out = {}
    for p in names:
        sel = [spaces[k][p[k]] for k in range(4)]
        out[p] = _lib.gather4(eri, [x.idx for x in sel], [x.s for x in sel], [x.n for x in sel])
    return cls(**out)


def ft_active_integrals(sys, en, focc, fvir, iocc, ivir):
    """Dressed integrals with small occupations excluded (kelvin/cc_utils.py:825-859):
    block[p,q,r,s] = <P Q||R S> * sqrt(f) ..., P = iocc[p] / ivir[p]."""
    dev = _lib.device()
    so, sv = _Space(iocc, focc, dev), _Space(ivir, fvir, dev)
    F = _active_F(_lib.as_dev(sys.g_fock_tot(), dev), en, so, sv, dev)
    eri = _lib.as_dev(sys.g_aint_tot(), dev)
    sp = {"o": so, "v": sv}
    return F, _active_I(eri, [sp]*4, two_e_blocks.names, two_e_blocks)


def uft_active_integrals(sys, ea, eb, foa, fva, fob, fvb, iocca, ivira, ioccb, ivirb):
    """Unrestricted version (kelvin/cc_utils.py:862-938)."""
    dev = _lib.device()
    sa = {"o": _Space(iocca, foa, dev), "v": _Space(ivira, fva, dev)}
    sb = {"o": _Space(ioccb, fob, dev), "v": _Space(ivirb, fvb, dev)}
    fa, fb = sys.u_fock_tot()
    Fa = _active_F(_lib.as_dev(fa, dev), ea, sa["o"], sa["v"], dev)
    Fb = _active_F(_lib.as_dev(fb, dev), eb, sb["o"], sb["v"], dev)
    eriA, eriB, eriAB = sys.u_aint_tot()
    same = eriB is eriA
    eriA = _lib.as_dev(eriA, dev)
    eriB = eriA if same else _lib.as_dev(eriB, dev)
    eriAB = _lib.as_dev(eriAB, dev)
    Ia = _active_I(eriA, [sa]*4, two_e_blocks.names, two_e_blocks)
    Ib = _active_I(eriB, [sb]*4, two_e_blocks.names, two_e_blocks)
    Iabab = _active_I(eriAB, [sa, sb, sa, sb], two_e_blocks_full.names, two_e_blocks_full)
    return Fa, Fb, Ia, Ib, Iabab


# ---------------------------------------------------------------------------
# amplitude updates
# ---------------------------------------------------------------------------
def form_new_ampl(method, F, I, T1old, T2old, D1, D2, ti, ng, G, t0_zero=False):
    """Form new amplitudes (kelvin/cc_utils.py:17-49)."""
    if method == "CCSD":
        return ft_cc_equations.ccsd_stanton(F, I, T1old, T2old, D1, D2, ti, ng, G,
                                            t0_zero=t0_zero)
    if method == "CCD":
        return T1old, ft_cc_equations.ccd_simple(F, I, T2old, D2, ti, ng, G)
    if method == "LCCSD":
        return ft_cc_equations.lccsd_simple(F, I, T1old, T2old, D1, D2, ti, ng, G)
    if method == "LCCD":
        return T1old, ft_cc_equations.lccd_simple(F, I, T2old, D2, ti, ng, G)
    raise Exception("Unrecognized method keyword")


def form_new_ampl_u(method, Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold,
                    D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G, t0_zero=False, closed_shell=False):
    """Form new amplitudes, unrestricted (kelvin/cc_utils.py:52-86)."""
    if method == "CCSD":
        return ft_cc_equations.uccsd_stanton(
            Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold,
            D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G, t0_zero=t0_zero, closed_shell=closed_shell)
    raise Exception("Unrecognized method keyword for unrestricted calc")


def _closed_shell(Fa, Fb, Ia, Ib, Iabab, Ds, *amp_sets):
    """Integrals, denominators (D1a, D1b, D2aa, D2ab, D2bb) and every amplitude set mirror
    symmetric?  (The denominators enter through the time integration that follows the
    residual: with D1a != D1b equal residuals would still integrate to different amplitudes.)"""
    if not ft_cc_equations.CLOSED_SHELL:
        return False
    dev = _lib.device()
    D1a, D1b, D2aa, D2ab, D2bb = [_lib.as_dev(d, dev) for d in Ds]
    same = ft_cc_equations._same
    if not (same(D1a, D1b) and same(D2aa, D2bb) and same(D2ab, D2ab.permute(1, 0, 3, 2))):
        return False
    if not ft_cc_equations.closed_shell_integrals(Fa, Fb, Ia, Ib, Iabab):
        return False
    return all(ft_cc_equations.closed_shell_amplitudes(*a) for a in amp_sets)


class _Stats(object):
    """One device buffer for the per-iteration scalars; read back once."""
    def __init__(self, n, dev):
        self.buf = torch.zeros(3*n, dtype=torch.float64, device=dev)
        self.dev = dev

    def damp(self, k, old, new, alpha):
        """old <- alpha*old + (1-alpha)*new with the three norms; the grid points (leading axis)
        of either tensor may be rows of a wider buffer."""
        lib = _lib.load()
        out = self.buf.data_ptr() + 24*k
        scr = _lib.ptr(_lib.reduce_scratch(self.dev))
        if old.is_contiguous() and new.is_contiguous():
            rc = lib.kb200_damp_norms(old.numel(), _lib.ptr(old), _lib.ptr(new), alpha, out, scr,
                                      _lib.stream_ptr())
        else:
            if old.dim() < 2 or tuple(old.shape) != tuple(new.shape) \
                    or not old[0].is_contiguous() or not new[0].is_contiguous():
                raise Exception("damp: blocks must be contiguous within a grid point")
            rc = lib.kb200_damp_norms_rows(old.shape[0], old[0].numel(), _lib.ptr(old),
                                           _lib.row_stride(old), _lib.ptr(new),
                                           _lib.row_stride(new), alpha, out, scr,
                                           _lib.stream_ptr())
        _lib.check(rc, "kb200_damp_norms")

    def read(self):
        return self.buf.cpu().numpy().reshape(-1, 3)


def _norm(x):
    """||x||_2 of a device tensor through the fused norm kernel (alpha=1: no update)."""
    st = _Stats(1, x.device)
    st.damp(0, x, x, 1.0)
    return math.sqrt(st.read()[0, 1])


class GccStep(object):
    """State and ONE damped iteration of the FT-CCSD fixed-point loop in general spin orbitals:
    the loop body of kelvin/cc_utils.py:131-160.  Residual plan (sharded over the ranks when
    torch.distributed is up), then per block ONE fused pass: integration, residual norm,
    damping, new norm, energy term (kb200_int_tbar_update)."""

    def __init__(self, T1old, T2old, F, I, D1, D2, g, G, beta, ng, ti):
        dev = self.dev = _lib.device()
        self.F, self.I, self.g, self.G, self.beta, self.ng, self.ti = F, I, g, G, beta, ng, ti
        self.D1, self.D2 = _lib.as_dev(D1, dev), _lib.as_dev(D2, dev)
        self.T1 = _lib.as_dev(T1old, dev).clone()
        self.T2 = _lib.as_dev(T2old, dev).clone()
        self.nl1 = _norm(self.T1) + 0.1
        self.nl2 = _norm(self.T2) + 0.1
        self.Iabij = ft_cc_energy.oovv_to_abij(I.oovv)
        self.fai = _lib.as_dev(F.ov, dev).t().contiguous()
        self.stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self.work = {}            # residual buffers kept across iterations (stable addresses)
        # T[0] == 0 is preserved by the update (row 0 of G vanishes): skip that grid point
        self.t0 = ft_cc_equations.t0_is_zero(G, (self.T1, self.T2))
        # a caller-supplied guess need not be antisymmetric: then the full sums are evaluated
        self.antisym = ft_cc_equations.is_antisymmetric(self.T2)

    def step(self, alpha):
        """-> (E, res1 + res2) as logged by the reference."""
        ng = self.ng
        b1, b2 = ft_cc_equations.ccsd_stanton_bar(self.F, self.I, self.T1, self.T2,
                                                  t0_zero=self.t0, antisym=self.antisym,
                                                  work=self.work)
        sp = self.stats.data_ptr()
        quadrature.int_tbar_update(ng, b1, self.ti, self.D1, self.G, self.T1, alpha, sp,
                                   g=self.g, W=self.fai, c2=1.0)
        quadrature.int_tbar_update(ng, b2, self.ti, self.D2, self.G, self.T2, alpha, sp + 32,
                                   g=self.g, W=self.Iabij, T1x=self.T1, T1y=self.T1,
                                   c2=0.25, c11=0.5)
        s = self.stats.cpu().numpy().reshape(2, 4)
        res1 = math.sqrt(s[0, 0])/self.nl1
        res2 = math.sqrt(s[1, 0])/self.nl2
        self.nl1 = math.sqrt(s[0, 2]) + 0.1
        self.nl2 = math.sqrt(s[1, 2]) + 0.1
        E = float(s[0, 3] + s[1, 3])/self.beta
        return E, res1 + res2


def ft_cc_iter(method, T1old, T2old, F, I, D1, D2, g, G, beta, ng, ti, iprint, conv_options,
               flags_out=None):
    """Fixed-point FT-CCSD loop, general spin orbitals (kelvin/cc_utils.py:111-173)."""
    tbeg = time.time()
    dev = _lib.device()
    converged = False
    ethresh = conv_options["econv"]
    tthresh = conv_options["tconv"]
    max_iter = conv_options["max_iter"]
    alpha = conv_options["damp"]
    i = 0
    Eold = 888888888.888888888
    if method == "CCSD":
        st = GccStep(T1old, T2old, F, I, D1, D2, g, G, beta, ng, ti)
        if flags_out is not None:
            flags_out.update({"t0": st.t0, "antisym": st.antisym})
        while i < max_iter and not converged:
            E, res = st.step(alpha)
            logging.info(' %2d  %.10f   %.4E' % (i + 1, E, res))
            i = i + 1
            if numpy.abs(E - Eold) < ethresh and res < tthresh:
                converged = True
            Eold = E
        T1old, T2old = st.T1, st.T2
    else:
        T1old = _lib.as_dev(T1old, dev).clone()
        T2old = _lib.as_dev(T2old, dev).clone()
        nl1 = _norm(T1old) + 0.1
        nl2 = _norm(T2old) + 0.1
        Iabij = ft_cc_energy.oovv_to_abij(I.oovv)
        st = _Stats(2, dev)
        while i < max_iter and not converged:
            T1, T2 = form_new_ampl(method, F, I, T1old, T2old, D1, D2, ti, ng, G)
            # residuals, damping and new norms in one pass per tensor
            st.damp(0, T1old, T1, alpha)
            st.damp(1, T2old, T2, alpha)
            E = ft_cc_energy.ft_cc_energy(T1old, T2old, F.ov, I.oovv, g, beta, eri_abij=Iabij)
            s = st.read()
            res1 = math.sqrt(s[0, 0])/nl1
            res2 = math.sqrt(s[1, 0])/nl2
            nl1 = math.sqrt(s[0, 2]) + 0.1
            nl2 = math.sqrt(s[1, 2]) + 0.1
            logging.info(' %2d  %.10f   %.4E' % (i + 1, E, res1 + res2))
            i = i + 1
            if numpy.abs(E - Eold) < ethresh and res1 + res2 < tthresh:
                converged = True
            Eold = E
    if not converged:
        logging.warning("{} did not converge!".format(method))
    tend = time.time()
    logging.info("Total {} time: {:.4f} s".format(method, (tend - tbeg)))
    return Eold, T1old, T2old


class UccStep(object):
    """State and ONE damped iteration of the unrestricted FT-CCSD fixed-point loop: the loop body
    of kelvin/cc_utils.py:274-305.  Decides once per solve (device checks of the inputs) which
    reductions of the program apply: tau_0 shortcut, closed shell (alpha == beta), singlet
    (T2aa = T2ab - T2ab(a<->b)), permutational antisymmetry of the same-spin doubles."""

    def __init__(self, amps, Fa, Fb, Ia, Ib, Iabab, Ds, g, G, beta, ng, ti, known=None):
        dev = self.dev = _lib.device()
        self.ints = (Fa, Fb, Ia, Ib, Iabab)
        self.Ds = [_lib.as_dev(d, dev) for d in Ds]
        self.g, self.G, self.beta, self.ng, self.ti = g, G, beta, ng, ti
        self.old = [_lib.as_dev(x, dev).clone() for x in amps]
        self.abij = (ft_cc_energy.oovv_to_abij(Ia.oovv), ft_cc_energy.oovv_to_abij(Iabab.oovv),
                     ft_cc_energy.oovv_to_abij(Ib.oovv))
        self.fT = (_lib.as_dev(Fa.ov, dev).t().contiguous(), _lib.as_dev(Fb.ov, dev).t().contiguous())
        self.stats = torch.zeros(20, dtype=torch.float64, device=dev)
        self.work = {}            # residual buffers kept across iterations (stable addresses)
        self.set_flags(known)

    def set_flags(self, known=None):
        """known: dict with any of t0, closed_shell, singlet, antisym decided by the caller
        (e.g. a copy of this solver's own state); everything else is checked on the device."""
        known = known or {}
        self.finish()
        old = self.old
        Fa, Fb, Ia, Ib, Iabab = self.ints
        self.t0 = known["t0"] if "t0" in known else ft_cc_equations.t0_is_zero(self.G, old)
        self.cs = known["closed_shell"] if "closed_shell" in known else \
            _closed_shell(Fa, Fb, Ia, Ib, Iabab, self.Ds, old)
        self.singlet = known["singlet"] if "singlet" in known else \
            bool(self.cs and ft_cc_equations.is_singlet(old[2], old[3]))
        if "antisym" in known:
            self.antisym = known["antisym"]
        else:
            self.antisym = bool(self.singlet) or (
                ft_cc_equations.is_antisymmetric(old[2])
                and (self.cs or ft_cc_equations.is_antisymmetric(old[4])))

    def flags(self):
        return {"t0": self.t0, "closed_shell": self.cs, "singlet": self.singlet,
                "antisym": self.antisym}

    def finish(self):
        """Closed-shell runs: copy the alpha blocks into the beta ones (T1b, T2bb)."""
        if getattr(self, "_beta_stale", False):
            self.old[1].copy_(self.old[0])
            self.old[4].copy_(self.old[2])
            self._beta_stale = False

    def step(self, alpha):
        """-> (E, res1 + res2) as logged by the reference (kelvin/cc_utils.py:297-305)."""
        ng = self.ng
        old = self.old
        _trace.mark("step")
        bars = ft_cc_equations.uccsd_stanton_bar(
            *self.ints, *old, t0_zero=self.t0, closed_shell=self.cs, beta_copies=False,
            singlet=self.singlet, antisym=self.antisym, work=self.work)
        _trace.mark("bars")
        live = (0, 2, 3) if self.cs else (0, 1, 2, 3, 4)
        # energy terms: singles with F.ov; doubles with <ij||ab> and the (already updated) singles
        T1a, T1b = old[0], (old[0] if self.cs else old[1])
        eterm = {0: dict(W=self.fT[0], c2=1.0), 1: dict(W=self.fT[1], c2=1.0),
                 2: dict(W=self.abij[0], T1x=T1a, T1y=T1a, c2=0.25, c11=0.5),
                 3: dict(W=self.abij[1], T1x=T1a, T1y=T1b, c2=1.0, c11=1.0),
                 4: dict(W=self.abij[2], T1x=T1b, T1y=T1b, c2=0.25, c11=0.5)}
        sp = self.stats.data_ptr()
        for k in live:
            quadrature.int_tbar_update(ng, bars[k], self.ti, self.Ds[k], self.G, old[k], alpha,
                                       sp + 32*k, g=self.g, **eterm[k])
        # (alpha == beta: the beta blocks are copies of the alpha ones; nothing reads them during
        # the iterations, so they are brought up to date once, in finish())
        self._beta_stale = bool(self.cs)
        _trace.mark("updated")
        s = self.stats.cpu().numpy().reshape(5, 4).copy()
        _trace.mark("stats")
        _trace.report()
        if self.cs:
            s[1] = s[0]
            s[4] = s[2]
        n = numpy.sqrt(s[:, :3])
        nl1 = n[0, 1] + 0.1 + n[1, 1]
        nl2 = n[2, 1] + 0.1 + n[3, 1] + n[4, 1]
        res1 = n[0, 0]/nl1 + n[1, 0]/nl1
        res2 = n[2, 0]/nl2 + n[3, 0]/nl2 + n[4, 0]/nl2
        E = float(s[0, 3] + s[1, 3] + s[2, 3] + s[3, 3] + s[4, 3])/self.beta
        return E, float(res1 + res2)


def ft_ucc_iter(method, T1aold, T1bold, T2aaold, T2abold, T2bbold, Fa, Fb, Ia, Ib, Iabab,
                D1a, D1b, D2aa, D2ab, D2bb, g, G, beta, ng, ti, iprint, conv_options,
                flags_out=None):
    """Fixed-point FT-UCCSD loop (kelvin/cc_utils.py:245-317).  The norms nl1/nl2
    are those of the amplitudes *before* the update, as in the reference.
    flags_out: dict that receives the symmetry verdicts of the solve (UccStep.flags)."""
    if method != "CCSD":
        raise Exception("Unrecognized method keyword for unrestricted calc")
    tbeg = time.time()
    converged = False
    ethresh = conv_options["econv"]
    tthresh = conv_options["tconv"]
    max_iter = conv_options["max_iter"]
    alpha = conv_options["damp"]
    i = 0
    Eold = 888888888.888888888
    st = UccStep((T1aold, T1bold, T2aaold, T2abold, T2bbold), Fa, Fb, Ia, Ib, Iabab,
                 (D1a, D1b, D2aa, D2ab, D2bb), g, G, beta, ng, ti)
    if flags_out is not None:
        flags_out.update(st.flags())
    while i < max_iter and not converged:
        E, res = st.step(alpha)
        logging.info(' %2d  %.10f   %.4E' % (i + 1, E, res))
        i = i + 1
        if numpy.abs(E - Eold) < ethresh and res < tthresh:
            converged = True
        Eold = E
    if not converged:
        logging.warning("{} did not converge!".format(method))
    st.finish()
    tend = time.time()
    logging.info("Total {} time: {:.4f} s".format(method, (tend - tbeg)))
    old = st.old
    return Eold, (old[0], old[1]), (old[2], old[3], old[4])


def form_new_ampl_extrap(ig, method, F, I, T1, T2, T1bar, T2bar, D1, D2, ti, ng, G):
    """kelvin/cc_utils.py:89-96."""
    if method == "CCSD":
        return ft_cc_equations.ccsd_stanton_single(
            ig, F, I, T1, T2, T1bar, T2bar, D1, D2, ti, ng, G)
    raise Exception("Unrecognized method keyword")


def form_new_ampl_extrap_u(ig, method, Fa, Fb, Ia, Ib, Iabab,
                           T1a, T1b, T2aa, T2ab, T2bb, T1bara, T1barb,
                           T2baraa, T2barab, T2barbb, D1a, D1b,
                           D2aa, D2ab, D2bb, ti, ng, G):
    """kelvin/cc_utils.py:98-108."""
    if method == "CCSD":
        return ft_cc_equations.uccsd_stanton_single(
            ig, Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb, T1bara, T1barb,
            T2baraa, T2barab, T2barbb, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G)
    raise Exception("Unrecognized method keyword")


def _pointwise(method, drivers, Ds, single, G, ng, ti, conv_options, nres1):
    """Shared body of ft_cc_iter_extrap / ft_ucc_iter_extrap (kelvin/cc_utils.py:176-242,
    320-411): march through the grid points; at each one start from the linear extrapolation
    of the two previous points and iterate the single-point update to convergence.
    drivers: F.vo / I.vvoo blocks; single(ig, amps_at_ig, bars) -> new amplitudes at ig;
    nres1: how many of the blocks are singles (their residuals use nl1)."""
    thresh = conv_options["tconv"]
    max_iter = conv_options["max_iter"]
    alpha = conv_options["damp"]
    dev = _lib.device()
    drivers = [_lib.as_dev(x, dev) for x in drivers]
    nb = len(drivers)
    bars = [torch.zeros((ng,) + tuple(d.shape), dtype=torch.float64, device=dev) for d in drivers]
    news = [torch.zeros_like(b) for b in bars]
    st = _Stats(nb, dev)
    for ig in range(ng):
        if ig == 0:
            for b, d in zip(bars, drivers):
                b[0].copy_(d).neg_()
            continue  # don't bother computing at T = inf
        elif ig == 1:
            for k in range(nb):
                bars[k][ig].copy_(drivers[k]).neg_()
                news[k][ig].copy_(quadrature.int_tbar(ng, bars[k], ti, Ds[k], G,
                                                        rows=(ig, ig + 1))[0])
        else:
            # linear extrapolation: T[ig] = T[ig-1] + (T[ig-2] - T[ig-1])*fac
            fac = (ti[ig] - ti[ig - 1])/(ti[ig - 2] - ti[ig - 1])
            for k in range(nb):
                news[k][ig].copy_(news[k][ig - 1])
                st.damp(k, news[k][ig], news[k][ig - 2], 1.0 - fac)
        converged = False
        nl1 = math.sqrt(float(news[0][ig].numel()))
        nl2 = math.sqrt(float(news[nres1][ig].numel()))
        logging.info("Time point {}".format(ig))
        i = 0
        while i < max_iter and not converged:
            out = single(ig, [x[ig] for x in news], bars)
            for k in range(nb):
                st.damp(k, news[k][ig], out[k], alpha)
            s = numpy.sqrt(st.read()[:, 0])
            res1 = float(sum(s[:nres1]))/nl1
            res2 = float(sum(s[nres1:]))/nl2
            logging.info(' %2d  %.4E' % (i + 1, res1 + res2))
            i = i + 1
            if res1 + res2 < thresh:
                converged = True
    return news


def ft_cc_iter_extrap(method, F, I, D1, D2, g, G, beta, ng, ti, iprint, conv_options):
    """Pointwise-extrapolated FT-CCSD solver, general spin orbitals
    (kelvin/cc_utils.py:176-242).  Returns (T1, T2)."""
    def single(ig, amps, bars):
        return form_new_ampl_extrap(ig, method, F, I, amps[0], amps[1], bars[0], bars[1],
                                    D1, D2, ti, ng, G)
    T1, T2 = _pointwise(method, (F.vo, I.vvoo), (D1, D2), single, G, ng, ti, conv_options, 1)
    return T1, T2


def ft_ucc_iter_extrap(method, Fa, Fb, Ia, Ib, Iabab, D1a, D1b, D2aa, D2ab, D2bb,
                       g, G, beta, ng, ti, iprint, conv_options):
    """Pointwise-extrapolated FT-UCCSD solver (kelvin/cc_utils.py:320-411).
    Returns ((T1a, T1b), (T2aa, T2ab, T2bb))."""
    def single(ig, amps, bars):
        t1, t2 = form_new_ampl_extrap_u(ig, method, Fa, Fb, Ia, Ib, Iabab, *amps, *bars,
                                        D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G)
        return (t1[0], t1[1], t2[0], t2[1], t2[2])
    out = _pointwise(method, (Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo),
                     (D1a, D1b, D2aa, D2ab, D2bb), single, G, ng, ti, conv_options, 2)
    return (out[0], out[1]), (out[2], out[3], out[4])


def ft_lambda_iter(method, L1old, L2old, T1, T2, F, I, D1, D2, g, G, beta, ng, ti, iprint,
                   conv_options):
    """Fixed-point FT-CCSD Lambda loop (kelvin/cc_utils.py:414-480); method = CCSD, CCD,
    LCCSD or LCCD (:441-456)."""
    from .programs import METHODS
    if method not in METHODS:
        raise Exception("Unrecognized method keyword")
    tbeg = time.time()
    dev = _lib.device()
    converged = False
    thresh = conv_options["tconv"]
    max_iter = conv_options["max_iter"]
    alpha = conv_options["damp"]
    i = 0
    L1old = _lib.as_dev(L1old, dev).clone()
    L2old = _lib.as_dev(L2old, dev).clone()
    nl1 = _norm(L1old) + 0.1
    nl2 = _norm(L2old) + 0.1
    st = _Stats(2, dev)
    T1, T2 = _lib.as_dev(T1, dev), _lib.as_dev(T2, dev)
    asym = ft_cc_equations.is_antisymmetric(T2) and ft_cc_equations.is_antisymmetric(L2old)
    work = {}
    while i < max_iter and not converged:
        if method == "LCCSD":
            L1, L2 = ft_cc_equations.lccsd_lambda_simple(
                F, I, T1, T2, L1old, L2old, D1, D2, ti, ng, g, G, beta)
        elif method == "LCCD":
            L1 = L1old
            L2 = ft_cc_equations.lccd_lambda_simple(F, I, T2, L2old, D2, ti, ng, g, G, beta)
        elif method == "CCSD":
            L1, L2 = ft_cc_equations.ccsd_lambda_opt(
                F, I, T1, T2, L1old, L2old, D1, D2, ti, ng, g, G, beta, antisym=asym, work=work)
        else:
            L1 = L1old
            L2 = ft_cc_equations.ccd_lambda_simple(F, I, T2, L2old, D2, ti, ng, g, G, beta)
        st.damp(0, L1old, L1, alpha)
        st.damp(1, L2old, L2, alpha)
        s = st.read()
        res1 = math.sqrt(s[0, 0])/nl1
        res2 = math.sqrt(s[1, 0])/nl2
        nl1 = math.sqrt(s[0, 2]) + 0.1
        nl2 = math.sqrt(s[1, 2]) + 0.1
        L1 = None
        L2 = None
        logging.info(' %2d  %.10f' % (i + 1, res1 + res2))
        i = i + 1
        if res1 + res2 < thresh:
            converged = True
    if not converged:
        logging.warning("CCSD Lambda-equations did not converge!")
    tend = time.time()
    logging.info("Total CCSD Lambda time: %f s" % (tend - tbeg))
    return L1old, L2old


def ft_ulambda_iter(method, L1ain, L1bin, L2aain, L2abin, L2bbin, T1aold, T1bold,
                    T2aaold, T2abold, T2bbold, Fa, Fb, Ia, Ib, Iabab, D1a, D1b, D2aa, D2ab, D2bb,
                    g, G, beta, ng, ti, iprint, conv_options):
    """Fixed-point FT-UCCSD Lambda loop (kelvin/cc_utils.py:483-566); the
    doubles norm weights the ab block by 4 (:514-517)."""
    if method != "CCSD":
        raise Exception("Unrecognized method keyword")
    tbeg = time.time()
    dev = _lib.device()
    converged = False
    thresh = conv_options["tconv"]
    max_iter = conv_options["max_iter"]
    alpha = conv_options["damp"]
    i = 0
    old = [_lib.as_dev(x, dev).clone() for x in (L1ain, L1bin, L2aain, L2abin, L2bbin)]
    nrm = [_norm(x) for x in old]
    nl1 = nrm[0] + nrm[1] + 0.1
    nl2 = nrm[2] + 0.1 + nrm[4] + 4*nrm[3]
    st = _Stats(5, dev)
    Ts = [_lib.as_dev(x, dev) for x in (T1aold, T1bold, T2aaold, T2abold, T2bbold)]
    cs = _closed_shell(Fa, Fb, Ia, Ib, Iabab, (D1a, D1b, D2aa, D2ab, D2bb), Ts, old)
    # caller-supplied guesses need not be antisymmetric: then the full sums are evaluated
    asym = all(ft_cc_equations.is_antisymmetric(x) for x in
               ([Ts[2], old[2]] + ([] if cs else [Ts[4], old[4]])))
    live = (0, 2, 3) if cs else (0, 1, 2, 3, 4)
    work = {}
    while i < max_iter and not converged:
        new = ft_cc_equations.uccsd_lambda_opt(
            Fa, Fb, Ia, Ib, Iabab, *Ts,
            old[0], old[1], old[2], old[3], old[4], D1a, D1b, D2aa, D2ab, D2bb,
            ti, ng, g, G, beta, closed_shell=cs, antisym=asym, work=work)
        for k in live:
            st.damp(k, old[k], new[k], alpha)
        if cs:
            old[1].copy_(old[0])
            old[4].copy_(old[2])
        new = None
        s = numpy.sqrt(st.read())
        if cs:
            s[1] = s[0]
            s[4] = s[2]
        res1 = s[0, 0]/nl1 + s[1, 0]/nl1
        res2 = s[2, 0]/nl2 + s[3, 0]/nl2 + s[4, 0]/nl2
        nl1 = s[0, 2] + s[1, 2] + 0.1
        nl2 = s[2, 2] + 0.1 + s[4, 2] + 4*s[3, 2]
        logging.info(' %2d  %.10f' % (i + 1, res1 + res2))
        i = i + 1
        if res1 + res2 < thresh:
            converged = True
    if not converged:
        logging.warning("CCSD Lambda-equations did not converge!")
    tend = time.time()
    logging.info("Total CCSD Lambda time: %f s" % (tend - tbeg))
    return old[0], old[1], old[2], old[3], old[4]


# ---------------------------------------------------------------------------
# normal-ordered 2-RDM assembly and occupation-number response
# ---------------------------------------------------------------------------
def _scaled(P, facs, beta):
    return _dress4(P.contiguous(), *facs)/beta


def g_n2rdm_full(beta, sfo, sfv, P2):
    """Dress the nine P blocks with sqrt(f)/sqrt(1-f) and add their
    antisymmetric images (kelvin/cc_utils.py:1449-1466)."""
    dev = P2[0].device
    s = {"o": _vec(sfo, dev), "v": _vec(sfv, dev)}
    pats = ("vvvv", "vovv", "vvvo", "oovv", "vovo", "vvoo", "oovo", "ovoo", "oooo")
    blocks = [_scaled(P, [s[c] for c in pat], beta) for P, pat in zip(P2, pats)]
    n2 = blocks[0].clone()
    for k, (swap01, swap23) in ((1, (True, False)), (2, (False, True)), (6, (False, True)),
                                (7, (True, False))):
        n2 += blocks[k]
        n2 -= blocks[k].permute(1, 0, 2, 3) if swap01 else blocks[k].permute(0, 1, 3, 2)
    for k in (3, 5, 8):
        n2 += blocks[k]
    b = blocks[4]
    n2 += b - b.permute(0, 1, 3, 2) - b.permute(1, 0, 2, 3) + b.permute(1, 0, 3, 2)
    return n2


def u_n2rdm_full(beta, sfoa, sfva, sfob, sfvb, P2):
    """(P2aa, P2bb, P2ab) (kelvin/cc_utils.py:1506-1565)."""
    dev = P2[0][0].device
    sa = {"o": _vec(sfoa, dev), "v": _vec(sfva, dev)}
    sb = {"o": _vec(sfob, dev), "v": _vec(sfvb, dev)}
    pats = ("vvvv", "vovv", "vvvo", "oovv", "vovo", "vvoo", "oovo", "ovoo", "oooo")
    out = []
    for k, s in ((0, sa), (1, sb)):
        out.append(g_n2rdm_full(beta, s["o"], s["v"], [P2[b][k] for b in range(9)]))

    def sc(P, pat, spins):
        return _scaled(P, [(sa if sp == "a" else sb)[c] for c, sp in zip(pat, spins)], beta)
    ab = sc(P2[0][2], pats[0], "abab").clone()
    for b in range(1, 9):
        ab += sc(P2[b][2], pats[b], "abab")
    for b in (1, 2, 6, 7):
        ab += sc(P2[b][3], pats[b], "baba").permute(1, 0, 3, 2)
    ab -= sc(P2[4][3], pats[4], "abba").permute(0, 1, 3, 2)
    ab -= sc(P2[4][4], pats[4], "baab").permute(1, 0, 2, 3)
    ab += sc(P2[4][5], pats[4], "baba").permute(1, 0, 3, 2)
    return (out[0], out[1], ab)


_P2_PATS = ("vvvv", "vovv", "vvvo", "oovv", "vovo", "vvoo", "oovo", "ovoo", "oooo")
_ID, _S01, _S23, _S0123 = (0, 1, 2, 3), (1, 0, 2, 3), (0, 1, 3, 2), (1, 0, 3, 2)
# antisymmetric images of each same-spin block: (sign, destination axis of each source axis)
_P2_IMAGES = {0: ((1, _ID),), 1: ((1, _ID), (-1, _S01)), 2: ((1, _ID), (-1, _S23)), 3: ((1, _ID),),
              4: ((1, _ID), (-1, _S23), (-1, _S01), (1, _S0123)), 5: ((1, _ID),),
              6: ((1, _ID), (-1, _S23)), 7: ((1, _ID), (-1, _S01)), 8: ((1, _ID),)}


def _scatter_block(dst, P, pat, spaces, beta, sign, perm):
    sel = [spaces[k][pat[k]] for k in range(4)]
    _lib.scatter4_add(dst, P, [x.idx for x in sel], [x.s for x in sel], alpha=sign/beta, perm=perm)


def _n2rdm_same_spin_active(beta, n, sp, blocks, dev):
    n2 = torch.zeros((n,)*4, dtype=torch.float64, device=dev)
    for k, P in enumerate(blocks):
        for sign, perm in _P2_IMAGES[k]:
            _scatter_block(n2, P, _P2_PATS[k], [sp]*4, beta, sign, perm)
    return n2


def g_n2rdm_full_active(beta, n, iocc, ivir, sfo, sfv, P2):
    """Normal-ordered 2-RDM in the full orbital space from active-space P blocks
    (kelvin/cc_utils.py:1469-1503); sfo/sfv are the sqrt-occupations of the active sets."""
    dev = P2[0].device
    sp = {"o": _Space(iocc, numpy.asarray(sfo)**2, dev), "v": _Space(ivir, numpy.asarray(sfv)**2, dev)}
    sp["o"].s, sp["v"].s = _vec(sfo, dev), _vec(sfv, dev)
    return _n2rdm_same_spin_active(beta, n, sp, P2, dev)


def u_n2rdm_full_active(beta, na, nb, iocca, ivira, ioccb, ivirb, sfoa, sfva, sfob, sfvb, P2):
    """kelvin/cc_utils.py:1568-1625."""
    dev = P2[0][0].device

    def space(io, iv, so, sv):
        sp = {"o": _Space(io, numpy.asarray(so)**2, dev), "v": _Space(iv, numpy.asarray(sv)**2, dev)}
        sp["o"].s, sp["v"].s = _vec(so, dev), _vec(sv, dev)
        return sp
    sa, sb = space(iocca, ivira, sfoa, sfva), space(ioccb, ivirb, sfob, sfvb)
    aa = _n2rdm_same_spin_active(beta, na, sa, [P2[b][0] for b in range(9)], dev)
    bb = _n2rdm_same_spin_active(beta, nb, sb, [P2[b][1] for b in range(9)], dev)
    ab = torch.zeros((na, nb, na, nb), dtype=torch.float64, device=dev)
    abab, baba = [sa, sb, sa, sb], [sb, sa, sb, sa]
    for b in range(9):
        _scatter_block(ab, P2[b][2], _P2_PATS[b], abab, beta, 1.0, _ID)
    for b in (1, 2, 6, 7):
        _scatter_block(ab, P2[b][3], _P2_PATS[b], baba, beta, 1.0, _S0123)
    _scatter_block(ab, P2[4][3], _P2_PATS[4], [sa, sb, sb, sa], beta, -1.0, _S23)
    _scatter_block(ab, P2[4][4], _P2_PATS[4], [sb, sa, sa, sb], beta, -1.0, _S01)
    _scatter_block(ab, P2[4][5], _P2_PATS[4], baba, beta, 1.0, _S0123)
    return (aa, bb, ab)


def g_Fd_on(Fd, ndia, ndba, ndji, ndai):
    """Perturbed-occupation contribution of the Fock matrix, host O(n^3)
    (kelvin/cc_utils.py:1628-1633)."""
    e = numpy.einsum
    return -(e('ia,aik->k', ndia, Fd) + e('ba,abk->k', ndba, Fd)
             + e('ji,ijk->k', ndji, Fd) + e('ai,iak->k', ndai, Fd))


def u_Fd_on(Fdaa, Fdab, Fdba, Fdbb, ndia, ndba, ndji, ndai):
    """kelvin/cc_utils.py:1688-1706."""
    nd = lambda k: (ndia[k], ndba[k], ndji[k], ndai[k])  # noqa: E731
    tempA = g_Fd_on(Fdaa, *nd(0)) + g_Fd_on(Fdba, *nd(1))
    tempB = g_Fd_on(Fdbb, *nd(1)) + g_Fd_on(Fdab, *nd(0))
    return tempA, tempB


def on_response(leaves, ints, spin_of_leaf):
    """Partial traces 'sum over all indices but one' of every Lagrangian
    derivative block with its dressed integral block, accumulated per
    (space, spin) of the kept index.  This is the generic form of
    kelvin/cc_utils.py:1648-1685 (g_d_on_oo / g_d_on_vv) and :1746-1895
    (u_d_on_oo / u_d_on_vv), whose term lists are exactly these traces with the
    symmetry-equivalent positions merged.

    leaves: list of (adjoint tensor, its letters, dressed block, its letters,
                     pattern string, leaf-name prefix, weight)
    Returns dict (space, spin) -> device vector."""
    acc = {}
    for A, la, B, lb, pat, pre, w in leaves:
        for pos, l in enumerate(lb):
            key = (pat[pos], spin_of_leaf(pre, pos))
            out = acc.get(key)
            res = _lib.dot_keep(A, la, B, lb, l, alpha=w, out=out, beta=0.0 if out is None else 1.0)
            acc[key] = res
    return acc


def _add_pair(rdm2, A, B, fac, exchange):
    """rdm2[p,q,r,s] += fac*A[p,r]*B[q,s]; with exchange also rdm2[p,q,s,r] -= fac*A[p,r]*B[q,s]."""
    X = fac*A[:, None, :, None]*B[None, :, None, :]
    rdm2 += X
    if exchange:
        rdm2 -= X.transpose(2, 3)


def g_full_rdm2(fo, n1rdm, rdm2):
    """Mean-field and mixed parts of the full 2-RDM, added in place (kelvin/cc_utils.py:2018-2024)."""
    dev = rdm2.device
    d = torch.diag(torch.as_tensor(fo, dtype=torch.float64)).to(dev)
    _add_pair(rdm2, d, d, 1.0, True)
    _add_pair(rdm2, d, n1rdm, 0.5, True)
    _add_pair(rdm2, n1rdm, d, 0.5, True)


def u_full_rdm2(foa, fob, n1rdm, rdm2):
    """kelvin/cc_utils.py:2036-2053.  The reference's bb exchange term pairs diag(fob) with
    diag(foa) (:2045); kept as is, identical for spin-symmetric occupations."""
    dev = rdm2[0].device
    da = torch.diag(torch.as_tensor(foa, dtype=torch.float64)).to(dev)
    db = torch.diag(torch.as_tensor(fob, dtype=torch.float64)).to(dev)
    _add_pair(rdm2[0], da, da, 1.0, True)
    _add_pair(rdm2[0], da, n1rdm[0], 0.5, True)
    _add_pair(rdm2[0], n1rdm[0], da, 0.5, True)
    _add_pair(rdm2[1], db, db, 1.0, False)
    rdm2[1] -= (db[:, None, :, None]*da[None, :, None, :]).transpose(2, 3)
    _add_pair(rdm2[1], db, n1rdm[1], 0.5, True)
    _add_pair(rdm2[1], n1rdm[1], db, 0.5, True)
    _add_pair(rdm2[2], da, db, 1.0, False)
    _add_pair(rdm2[2], da, n1rdm[1], 0.5, False)
    _add_pair(rdm2[2], n1rdm[0], db, 0.5, False)

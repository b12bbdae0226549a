"""Finite-temperature CCSD driver (GPU).

Drop-in for the finite-temperature path of ``kelvin.ccsd.ccsd``
(kelvin/ccsd.py:20-165): same constructor keywords and defaults, ``run()`` ->
(Omega_tot, Omega_cc), ``compute_ESN()`` setting E/S/N and their pieces, the
same saved attributes (T1, T2, L1, L2, G0, G1, Gcc, Gtot, dia, dba, dji, dai,
P2, n1rdm, n2rdm, rono, ronv, ron1) and the same log lines.  Amplitudes are
CUDA float64 tensors.  Zero-temperature CCSD is outside this path and raises.
``rt_iter='point'`` (pointwise-extrapolated solver, kelvin/cc_utils.py:176-242,320-411) and
``singles=False`` (FT-CCD, general spin orbitals as in the reference) run on the same
kernels; ``athresh>0`` (active-space truncation, kelvin/ccsd.py:642-660,745-785) runs them
on rectangular no != nv blocks.
"""
import logging
import time

import numpy
import torch

from . import _lib, cc_utils, ft_cc_energy, ft_cc_equations, ft_mp, ft_utils, quadrature


class ccsd(object):
    """Coupled cluster singles and doubles (CCSD) driver at finite temperature.

    Attributes mirror kelvin/ccsd.py:23-42 (sys, T, mu, iprint, singles, econv,
    tconv, max_iter, damp, ngrid, realtime, athresh, quad, rt_iter, T1, T2, L1, L2).
    """
    def __init__(self, sys, T=0.0, mu=0.0, iprint=0, singles=True, econv=1e-8, tconv=None,
                 max_iter=40, damp=0.0, ngrid=10, realtime=False, athresh=0.0, quad='lin',
                 rt_iter="all"):
        self.T = T
        self.mu = mu
        self.finite_T = False if T == 0 else True
        self.iprint = iprint
        self.singles = singles
        self.econv = econv
        self.tconv = tconv if tconv is not None else 1000.0*econv
        self.max_iter = max_iter
        self.damp = damp
        self.ngrid = ngrid
        self.realtime = realtime
        self.athresh = athresh
        self.quad = quad
        self.rt_iter = rt_iter
        if not self.finite_T:
            raise Exception("kelvin_b200.ccsd implements the finite-temperature path only (T > 0)")
        self.realtime = True
        if not sys.verify(self.T, self.mu):
            raise Exception("Sytem temperature inconsistent with CC temp")
        self.beta = 1.0/T
        self.beta_max = self.beta
        self.ti, self.g, self.G = quadrature.ft_quad(self.ngrid, self.beta_max, self.quad)
        self.sys = sys
        # amplitudes
        self.T1 = None
        self.T2 = None
        self.L1 = None
        self.L2 = None
        # pieces of normal-ordered 1-rdm
        self.dia = None
        self.dba = None
        self.dji = None
        self.dai = None
        # occupation number response
        self.rono = None
        self.ronv = None
        self.ron1 = None
        # pieces of 1-rdm with ONs
        self.ndia = None
        self.ndba = None
        self.ndji = None
        self.ndai = None
        # pieces of normal-ordered 2-rdm
        self.P2 = None
        self.n1rdm = None
        self.n2rdm = None
        self.r1rdm = None
        self.rorbo = None
        self.rorbv = None
        self._ints = None
        self._act = None
        # symmetry verdicts of the amplitude solve (tau_0 shortcut, closed shell, singlet,
        # antisymmetry; cc_utils.UccStep.flags), reused by the later residual passes
        self._flags = {}

    # ------------------------------------------------------------------
    def run(self, T1=None, T2=None):
        """Run CCSD calculation (kelvin/ccsd.py:109-126)."""
        logging.info('Running CCSD at an electronic temperature of %f K' % ft_utils.HtoK(self.T))
        if self.sys.has_u():
            return self._ft_uccsd(T1in=T1, T2in=T2)
        return self._ft_ccsd(T1in=T1, T2in=T2)

    def _u_flags(self):
        """Keywords for further residual evaluations at the converged amplitudes (the verdicts of
        the solve; amplitudes set from outside are re-examined by the callee)."""
        f = self._flags
        if not f:
            return {}
        return dict(t0_zero=bool(f.get("t0", False)), closed_shell=bool(f.get("closed_shell", False)),
                    singlet=bool(f.get("singlet", False)), antisym=f.get("antisym"))

    def _g_flags(self):
        f = self._flags
        if not f:
            return {}
        return dict(t0_zero=bool(f.get("t0", False)), antisym=f.get("antisym"))

    def _conv_options(self):
        return {"econv": self.econv, "tconv": self.tconv,
                "max_iter": self.max_iter, "damp": self.damp}

    # -- dressed integrals, built once per (T, mu) and kept on the device --
    def _active(self):
        """Active sets of the occupation-threshold truncation, one dict per spin with
        fo, fv (all orbitals), focc, fvir, iocc, ivir (entries with f > athresh); with
        athresh == 0 every orbital is in both sets (kelvin/ccsd.py:642-660, 745-785)."""
        if self._act is None:
            beta, mu = self.beta, self.mu
            ens = self.sys.u_energies_tot() if self.sys.has_u() else (self.sys.g_energies_tot(),)
            act = []
            for e in ens:
                fo, fv = ft_utils.ff(beta, e, mu), ft_utils.ffv(beta, e, mu)
                iocc = [i for i, x in enumerate(fo) if x > self.athresh]
                ivir = [i for i, x in enumerate(fv) if x > self.athresh]
                act.append(dict(e=e, n=e.shape[0], fo=fo, fv=fv, iocc=iocc, ivir=ivir,
                                focc=fo[iocc], fvir=fv[ivir]))
            self._act = act
        return self._act

    def _log_active(self):
        act = self._active()
        if len(act) == 1:
            a = act[0]
            logging.info("FT-CCSD orbital info:")
            for nm, v in (("nocc", len(a["iocc"])), ("nvir", len(a["ivir"])),
                          ("nact", len(a["iocc"]) + len(a["ivir"]) - a["n"])):
                logging.info('  {}: {:d}'.format(nm, v))
        else:
            logging.info("FT-UCCSD orbital info:")
            for sp, a in zip("ab", act):
                for nm, v in (("nocc", len(a["iocc"])), ("nvir", len(a["ivir"])),
                              ("nact", len(a["iocc"]) + len(a["ivir"]) - a["n"])):
                    logging.info('  {}{}: {:d}'.format(nm, sp, v))

    def _g_setup(self):
        if self._ints is None or self._ints[0] != "g":
            en = self.sys.g_energies_tot()
            if self.athresh > 0.0:
                a, = self._active()
                ev, eo = en[a["ivir"]], en[a["iocc"]]
                D1 = ft_utils.D1(ev, eo)
                D2 = ft_utils.D2(ev, eo)
                F, I = cc_utils.ft_active_integrals(
                    self.sys, en, a["focc"], a["fvir"], a["iocc"], a["ivir"])
            else:
                D1 = ft_utils.D1(en, en)
                D2 = ft_utils.D2(en, en)
                F, I = cc_utils.ft_integrals(self.sys, en, self.beta, self.mu)
            self._ints = ("g", en, D1, D2, F, I)
        return self._ints[1:]

    def _u_setup(self):
        if self._ints is None or self._ints[0] != "u":
            ea, eb = self.sys.u_energies_tot()
            if self.athresh > 0.0:
                a, b = self._active()
                eva, eoa = ea[a["ivir"]], ea[a["iocc"]]
                evb, eob = eb[b["ivir"]], eb[b["iocc"]]
                Fa, Fb, Ia, Ib, Iabab = cc_utils.uft_active_integrals(
                    self.sys, ea, eb, a["focc"], a["fvir"], b["focc"], b["fvir"],
                    a["iocc"], a["ivir"], b["iocc"], b["ivir"])
            else:
                eva = eoa = ea
                evb = eob = eb
                Fa, Fb, Ia, Ib, Iabab = cc_utils.uft_integrals(self.sys, ea, eb, self.beta, self.mu)
            D1a = ft_utils.D1(eva, eoa)
            D1b = ft_utils.D1(evb, eob)
            D2aa = ft_utils.D2(eva, eoa)
            D2ab = ft_utils.D2u(eva, evb, eoa, eob)
            D2bb = ft_utils.D2(evb, eob)
            self._ints = ("u", ea, eb, (D1a, D1b, D2aa, D2ab, D2bb), (Fa, Fb, Ia, Ib, Iabab))
        return self._ints[1:]

    def _ft_ccsd(self, T1in=None, T2in=None):
        """Solve finite temperature coupled cluster equations, general spin
        orbitals (kelvin/ccsd.py:595-712)."""
        ng, ti, G, g = self.ngrid, self.ti, self.G, self.g
        beta, mu = self.beta, self.mu
        en, D1, D2, F, I = self._g_setup()

        # 0th and 1st order contributions
        En = self.sys.const_energy()
        g0 = ft_utils.GP0(beta, en, mu)
        E0 = ft_mp.mp0(g0) + En
        E1 = self.sys.get_mp1()
        E01 = E0 + E1
        if self.athresh > 0.0:
            self._log_active()

        method = "CCSD" if self.singles else "CCD"
        if self.rt_iter[0] == 'a' or T2in is not None:
            if self.rt_iter[0] != 'a':
                logging.warning("Converngece scheme ({}) is being ignored.".format(self.rt_iter))
            if T1in is not None and T2in is not None:
                T1old = T1in if self.singles else torch.zeros_like(_lib.as_dev(T1in))
                T2old = T2in
            else:
                # MP2 guess: integrate the bare drivers (kelvin/ccsd.py:676-688)
                if self.singles:
                    T1old = quadrature.int_tbar1(
                        ng, (-F.vo).expand(ng, -1, -1).contiguous(), ti, D1, G)
                else:
                    T1old = torch.zeros((ng,) + tuple(F.vo.shape), dtype=torch.float64,
                                        device=F.vo.device)
                T2old = quadrature.int_tbar2(
                    ng, (-I.vvoo).expand(ng, -1, -1, -1, -1).contiguous(), ti, D2, G)
            E2 = ft_cc_energy.ft_cc_energy(T1old, T2old, F.ov, I.oovv, g, self.beta_max,
                                           Qterm=False)
            logging.info('MP2 Energy: {:.10f}'.format(E2))

            self._flags = {}
            Eccn, T1, T2 = cc_utils.ft_cc_iter(
                method, T1old, T2old, F, I, D1, D2, g, G, self.beta_max, ng, ti, self.iprint,
                self._conv_options(), flags_out=self._flags)
        else:
            T1, T2 = cc_utils.ft_cc_iter_extrap(
                method, F, I, D1, D2, g, G, self.beta_max, ng, ti, self.iprint,
                self._conv_options())
            Eccn = ft_cc_energy.ft_cc_energy(T1, T2, F.ov, I.oovv, g, self.beta_max)

        self.T1 = T1
        self.T2 = T2
        self.G0 = E0
        self.G1 = E1
        self.Gcc = Eccn
        self.Gtot = E0 + E1 + Eccn
        return (Eccn + E01, Eccn)

    def _ft_uccsd(self, T1in=None, T2in=None):
        """Solve finite temperature coupled cluster equations, unrestricted
        (kelvin/ccsd.py:714-880)."""
        ng, ti, G, g = self.ngrid, self.ti, self.G, self.g
        beta, mu = self.beta, self.mu
        ea, eb, (D1a, D1b, D2aa, D2ab, D2bb), (Fa, Fb, Ia, Ib, Iabab) = self._u_setup()

        En = self.sys.const_energy()
        g0 = ft_utils.uGP0(beta, ea, eb, mu)
        E0 = ft_mp.ump0(g0[0], g0[1]) + En
        E1 = self.sys.get_mp1()
        E01 = E0 + E1
        if self.athresh > 0.0:
            self._log_active()

        # as in the reference, the unrestricted loops know CCSD only: singles=False raises
        # "Unrecognized method keyword for unrestricted calc" (kelvin/cc_utils.py:84)
        method = "CCSD" if self.singles else "CCD"
        if self.rt_iter[0] == 'a' or T2in is not None:
            if self.rt_iter[0] != 'a':
                logging.warning("Converngece scheme ({}) is being ignored.".format(self.rt_iter))
            if T1in is not None and T2in is not None:
                T1aold, T1bold = T1in
                if not self.singles:
                    T1aold = torch.zeros_like(_lib.as_dev(T1aold))
                    T1bold = torch.zeros_like(_lib.as_dev(T1bold))
                T2aaold, T2abold, T2bbold = T2in
            else:
                def rep(x):
                    return (-x).expand(*((ng,) + (-1,)*x.dim())).contiguous()
                if self.singles:
                    T1aold = quadrature.int_tbar1(ng, rep(Fa.vo), ti, D1a, G)
                    T1bold = quadrature.int_tbar1(ng, rep(Fb.vo), ti, D1b, G)
                else:
                    T1aold = torch.zeros_like(rep(Fa.vo))
                    T1bold = torch.zeros_like(rep(Fb.vo))
                T2aaold = quadrature.int_tbar2(ng, rep(Ia.vvoo), ti, D2aa, G)
                T2abold = quadrature.int_tbar2(ng, rep(Iabab.vvoo), ti, D2ab, G)
                T2bbold = quadrature.int_tbar2(ng, rep(Ib.vvoo), ti, D2bb, G)

            E2 = ft_cc_energy.ft_ucc_energy(
                T1aold, T1bold, T2aaold, T2abold, T2bbold, Fa.ov, Fb.ov,
                Ia.oovv, Ib.oovv, Iabab.oovv, g, self.beta_max, Qterm=False)
            logging.info('MP2 Energy: {:.10f}'.format(E2))

            self._flags = {}
            Eccn, T1, T2 = cc_utils.ft_ucc_iter(
                method, T1aold, T1bold, T2aaold, T2abold, T2bbold, Fa, Fb, Ia, Ib, Iabab,
                D1a, D1b, D2aa, D2ab, D2bb, g, G, self.beta_max, ng, ti, self.iprint,
                self._conv_options(), flags_out=self._flags)
        else:
            T1, T2 = cc_utils.ft_ucc_iter_extrap(
                method, Fa, Fb, Ia, Ib, Iabab, D1a, D1b, D2aa, D2ab, D2bb,
                g, G, self.beta_max, ng, ti, self.iprint, self._conv_options())
            Eccn = ft_cc_energy.ft_ucc_energy(
                T1[0], T1[1], T2[0], T2[1], T2[2], Fa.ov, Fb.ov,
                Ia.oovv, Ib.oovv, Iabab.oovv, g, self.beta_max)

        self.T1 = T1
        self.T2 = T2
        self.G0 = E0
        self.G1 = E1
        self.Gcc = Eccn
        self.Gtot = E0 + E1 + Eccn
        return (Eccn + E01, Eccn)

    # ------------------------------------------------------------------
    def _ft_ccsd_lambda(self, L1=None, L2=None):
        """Solve FT-CCSD Lambda equations (kelvin/ccsd.py:882-961)."""
        ng, ti, G, g = self.ngrid, self.ti, self.G, self.g
        en, D1, D2, F, I = self._g_setup()
        if L2 is None and L1 is None:
            if self.singles:
                L1old, L2old = ft_cc_equations.ccsd_lambda_guess(F, I, self.T1, self.beta_max, ng)
            else:
                # the reference's zero singles guess has no grid axis (kelvin/ccsd.py:930); here
                # it is (ng, no, nv) so that the response densities accept it
                L2old = ft_cc_equations.ccd_lambda_guess(I, self.beta_max, ng)
                L1old = torch.zeros((ng,) + tuple(F.ov.shape), dtype=torch.float64,
                                    device=L2old.device)
        elif L1 is not None and L2 is not None:
            L1old, L2old = L1, L2
        elif L2 is not None:
            # L2 only: the reference starts the singles from zero (kelvin/ccsd.py:932-936)
            L2old = L2
            L1old = torch.zeros((ng,) + tuple(F.ov.shape), dtype=torch.float64,
                                device=_lib.device())
        else:
            # L1 only: the reference allocates a mis-shaped zero L2 here (quirk Q8); refuse
            raise Exception("provide L2 (or L1 and L2, or neither) as Lambda guess")
        L1, L2 = cc_utils.ft_lambda_iter(
            "CCSD" if self.singles else "CCD", L1old, L2old, self.T1, self.T2, F, I, D1, D2, g, G,
            self.beta_max, ng, ti,
            self.iprint, self._conv_options())
        self.L1 = L1
        self.L2 = L2

    def _ft_uccsd_lambda(self, L1=None, L2=None):
        """Solve FT-UCCSD Lambda equations (kelvin/ccsd.py:963-1073)."""
        ng, ti, G, g = self.ngrid, self.ti, self.G, self.g
        ea, eb, (D1a, D1b, D2aa, D2ab, D2bb), (Fa, Fb, Ia, Ib, Iabab) = self._u_setup()
        T1aold, T1bold = self.T1
        T2aaold, T2abold, T2bbold = self.T2
        if L2 is None and L1 is None:
            L1aold, L1bold, L2aaold, L2abold, L2bbold = ft_cc_equations.uccsd_lambda_guess(
                Fa, Fb, Ia, Ib, Iabab, self.T1[0], self.T1[1], self.beta_max, ng)
        elif L1 is not None and L2 is not None:
            L1aold, L1bold = L1
            L2aaold, L2abold, L2bbold = L2
        else:
            raise Exception("provide both L1 and L2 (or neither) as Lambda guess")
        L1a, L1b, L2aa, L2ab, L2bb = cc_utils.ft_ulambda_iter(
            "CCSD", L1aold, L1bold, L2aaold, L2abold, L2bbold, T1aold, T1bold,
            T2aaold, T2abold, T2bbold, Fa, Fb, Ia, Ib, Iabab, D1a, D1b, D2aa, D2ab, D2bb,
            g, G, self.beta_max, ng, ti, self.iprint, self._conv_options())
        self.L1 = (L1a, L1b)
        self.L2 = (L2aa, L2ab, L2bb)

    # ------------------------------------------------------------------
    # response densities
    # ------------------------------------------------------------------
    def _occ(self):
        """Fermi factors (host vectors) of the orbitals, per spin in the u case."""
        beta, mu = self.beta, self.mu
        if self.sys.has_u():
            ea, eb = self.sys.u_energies_tot()
            return ((ft_utils.ff(beta, ea, mu), ft_utils.ffv(beta, ea, mu)),
                    (ft_utils.ff(beta, eb, mu), ft_utils.ffv(beta, eb, mu)))
        en = self.sys.g_energies_tot()
        return ((ft_utils.ff(beta, en, mu), ft_utils.ffv(beta, en, mu)),)

    @staticmethod
    def _nd(d, s0, s1):
        dev = d.device
        a = torch.as_tensor(s0, dtype=torch.float64).to(dev)
        b = torch.as_tensor(s1, dtype=torch.float64).to(dev)
        return d*a[:, None]*b[None, :]

    def _pad2(self, X, rows, cols, n):
        """X scattered into an n x n zero matrix at (rows, cols) (the numpy.ix_ updates of
        kelvin/ccsd.py:1381-1385)."""
        if len(rows) == n and len(cols) == n:
            return X
        dev = X.device
        out = torch.zeros((n, n), dtype=torch.float64, device=dev)
        r = torch.as_tensor(numpy.asarray(rows, dtype=numpy.int64)).to(dev)
        c = torch.as_tensor(numpy.asarray(cols, dtype=numpy.int64)).to(dev)
        out[r[:, None], c[None, :]] = X
        return out

    def _n1rdm_total(self, a, ndia, ndba, ndji, ndai):
        n, io, iv = a["n"], a["iocc"], a["ivir"]
        return (self._pad2(ndia, io, iv, n) + self._pad2(ndba, iv, iv, n)
                + self._pad2(ndji, io, io, n) + self._pad2(ndai, iv, io, n))/self.beta

    def _g_ft_1rdm(self):
        """kelvin/ccsd.py:1333-1387."""
        if self.L2 is None:
            self._ft_ccsd_lambda()
        en, D1, D2, F, I = self._g_setup()
        a, = self._active()
        sfo, sfv = numpy.sqrt(a["focc"]), numpy.sqrt(a["fvir"])
        pia, pba, pji, pai = ft_cc_equations.ccsd_1rdm(
            self.T1, self.T2, self.L1, self.L2, D1, D2, self.ti, self.ngrid, self.g, self.G)
        self.dia, self.dba, self.dji, self.dai = pia, pba, pji, pai
        self.ndia = self._nd(pia, sfo, sfv)
        self.ndba = self._nd(pba, sfv, sfv)
        self.ndji = self._nd(pji, sfo, sfo)
        self.ndai = self._nd(pai, sfv, sfo)
        self.n1rdm = self._n1rdm_total(a, self.ndia, self.ndba, self.ndji, self.ndai)

    def _g_ft_2rdm(self):
        """kelvin/ccsd.py:1389-1432."""
        if self.L2 is None:
            self._ft_ccsd_lambda()
        en, D1, D2, F, I = self._g_setup()
        a, = self._active()
        sfo, sfv = numpy.sqrt(a["focc"]), numpy.sqrt(a["fvir"])
        self.P2 = ft_cc_equations.ccsd_2rdm(
            self.T1, self.T2, self.L1, self.L2, D1, D2, self.ti, self.ngrid, self.g, self.G)
        if self.athresh > 0.0:
            self.n2rdm = cc_utils.g_n2rdm_full_active(
                self.beta, a["n"], a["iocc"], a["ivir"], sfo, sfv, self.P2)
        else:
            self.n2rdm = cc_utils.g_n2rdm_full(self.beta, sfo, sfv, self.P2)

    def _u_ft_1rdm(self):
        """kelvin/ccsd.py:1579-1665."""
        if self.L2 is None:
            self._ft_uccsd_lambda()
        ea, eb, Ds, ints = self._u_setup()
        act = self._active()
        sq = numpy.sqrt
        pia, pba, pji, pai = ft_cc_equations.uccsd_1rdm(
            *self.T1, *self.T2, *self.L1, *self.L2, *Ds, self.ti, self.ngrid, self.g, self.G)
        self.dia, self.dba, self.dji, self.dai = pia, pba, pji, pai
        so, sv = [sq(a["focc"]) for a in act], [sq(a["fvir"]) for a in act]
        self.ndia = tuple(self._nd(pia[k], so[k], sv[k]) for k in (0, 1))
        self.ndba = tuple(self._nd(pba[k], sv[k], sv[k]) for k in (0, 1))
        self.ndji = tuple(self._nd(pji[k], so[k], so[k]) for k in (0, 1))
        self.ndai = tuple(self._nd(pai[k], sv[k], so[k]) for k in (0, 1))
        self.n1rdm = [self._n1rdm_total(act[k], self.ndia[k], self.ndba[k], self.ndji[k],
                                        self.ndai[k]) for k in (0, 1)]

    def _u_ft_2rdm(self):
        """kelvin/ccsd.py:1667-1729."""
        ea, eb, Ds, ints = self._u_setup()
        a, b = self._active()
        sq = numpy.sqrt
        self.P2 = ft_cc_equations.uccsd_2rdm(
            *self.T1, *self.T2, *self.L1, *self.L2, *Ds, self.ti, self.ngrid, self.g, self.G)
        if self.athresh > 0.0:
            self.n2rdm = cc_utils.u_n2rdm_full_active(
                self.beta, a["n"], b["n"], a["iocc"], a["ivir"], b["iocc"], b["ivir"],
                sq(a["focc"]), sq(a["fvir"]), sq(b["focc"]), sq(b["fvir"]), self.P2)
        else:
            self.n2rdm = cc_utils.u_n2rdm_full(self.beta, sq(a["focc"]), sq(a["fvir"]),
                                               sq(b["focc"]), sq(b["fvir"]), self.P2)

    # ------------------------------------------------------------------
    # occupation-number response (kelvin/ccsd.py:1434-1492, 1731-1826)
    # ------------------------------------------------------------------
    def _leaf_list(self):
        """(adjoint, dressed block) pairs of every integral block, drivers included."""
        u = self.sys.has_u()
        if u:
            ea, eb, Ds, (Fa, Fb, Ia, Ib, Iabab) = self._u_setup()
            src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
            summed, lsum, _ = ft_cc_equations._rdm_leaves(
                "u", list(self.T1) + list(self.T2), list(self.L1) + list(self.L2), Ds,
                self.ti, self.ngrid, self.g, self.G)
            drivers = [(lsum[0], "ia", Fa.vo, "ai", "vo", "Fa", 1.0),
                       (lsum[1], "ia", Fb.vo, "ai", "vo", "Fb", 1.0),
                       (lsum[2], "ijab", Ia.vvoo, "abij", "vvoo", "Ia", 0.25),
                       (lsum[4], "ijab", Ib.vvoo, "abij", "vvoo", "Ib", 0.25),
                       (lsum[3], "ijab", Iabab.vvoo, "abij", "vvoo", "Iabab", 1.0)]
        else:
            en, D1, D2, F, I = self._g_setup()
            src = {"F": F, "I": I}
            summed, lsum, _ = ft_cc_equations._rdm_leaves(
                "g", [self.T1, self.T2], [self.L1, self.L2], (D1, D2),
                self.ti, self.ngrid, self.g, self.G)
            drivers = [(lsum[0], "ia", F.vo, "ai", "vo", "F", 1.0),
                       (lsum[1], "ijab", I.vvoo, "abij", "vvoo", "I", 0.25)]
        leaves = list(drivers)
        for name, A in summed.items():
            pre, pat = name[:-1].split(".")
            letters = "pqrs"[:len(pat)]
            leaves.append((A, letters, _lib.as_dev(getattr(src[pre], pat)), letters, pat, pre, 1.0))
        return leaves

    @staticmethod
    def _spin_of_leaf(pre, pos):
        if pre in ("F", "I"):
            return "g"
        if pre in ("Fa", "Ia"):
            return "a"
        if pre in ("Fb", "Ib"):
            return "b"
        return "a" if pos % 2 == 0 else "b"

    def _padded_nd(self, a, k):
        """(ndia, ndba, ndji, ndai) of one spin as host arrays in the full orbital space;
        zero rows/columns outside the active sets make g_Fd_on equal to the reference's
        g_Fd_on_active / u_Fd_on_active (kelvin/cc_utils.py:1636-1645, 1709-1743)."""
        n, io, iv = a["n"], a["iocc"], a["ivir"]
        pick = (lambda x: x) if k is None else (lambda x: x[k])
        return (self._pad2(pick(self.ndia), io, iv, n).cpu().numpy(),
                self._pad2(pick(self.ndba), iv, iv, n).cpu().numpy(),
                self._pad2(pick(self.ndji), io, io, n).cpu().numpy(),
                self._pad2(pick(self.ndai), iv, io, n).cpu().numpy())

    def _g_ft_ron(self):
        """kelvin/ccsd.py:1434-1492."""
        a, = self._active()
        self.ron1 = self.sys.g_mp1_den()
        Fd = self.sys.g_fock_d_den()
        c = lambda x: x.cpu().numpy()  # noqa: E731
        rono = cc_utils.g_Fd_on(Fd, *self._padded_nd(a, None))
        acc = cc_utils.on_response(self._leaf_list(), None, self._spin_of_leaf)
        ronv = numpy.zeros_like(rono)
        io, iv = a["iocc"], a["ivir"]
        rono[io] -= 0.5*c(acc[("o", "g")])*a["fv"][io]
        ronv[iv] += 0.5*c(acc[("v", "g")])*a["fo"][iv]
        self.rono, self.ronv = rono, ronv

    def _u_ft_ron(self):
        """kelvin/ccsd.py:1731-1826."""
        act = self._active()
        mp1da, mp1db = self.sys.u_mp1_den()
        self.ron1 = [mp1da, mp1db]
        Fdaa, Fdab, Fdbb, Fdba = self.sys.u_fock_d_den()
        c = lambda x: x.cpu().numpy()  # noqa: E731
        nda, ndb = self._padded_nd(act[0], 0), self._padded_nd(act[1], 1)
        tA, tB = cc_utils.u_Fd_on(Fdaa, Fdab, Fdba, Fdbb, *zip(nda, ndb))
        acc = cc_utils.on_response(self._leaf_list(), None, self._spin_of_leaf)
        rono, ronv = [tA, tB], [numpy.zeros_like(tA), numpy.zeros_like(tB)]
        for k, sp in enumerate("ab"):
            io, iv = act[k]["iocc"], act[k]["ivir"]
            rono[k][io] -= 0.5*c(acc[("o", sp)])*act[k]["fv"][io]
            ronv[k][iv] += 0.5*c(acc[("v", sp)])*act[k]["fo"][iv]
        self.rono, self.ronv = rono, ronv

    # ------------------------------------------------------------------
    # orbital-energy response and relaxed 1-RDM (kelvin/ccsd.py:1494-1577, 1828-1959)
    # ------------------------------------------------------------------
    def _rorb_traces(self, pairs, no, nv):
        """sum_y g_y <L_y, T_y> with one index kept, for the (L, T, weight, occ letters,
        vir letters) tuples of one spin; returns (occ vector, vir vector) on the host."""
        dev = pairs[0][0].device
        o = torch.zeros(no, dtype=torch.float64, device=dev)
        v = torch.zeros(nv, dtype=torch.float64, device=dev)
        for L, Tg, w, okeep, vkeep in pairs:
            ll, lt = ("yia", "yai") if L.dim() == 3 else ("yijab", "yabij")
            for k in okeep:
                _lib.dot_keep(L, ll, Tg, lt, k, alpha=w, out=o, beta=1.0)
            for k in vkeep:
                _lib.dot_keep(L, ll, Tg, lt, k, alpha=w, out=v, beta=1.0)
        return o.cpu().numpy(), v.cpu().numpy()

    def _tau_weighted_stanton(self):
        """Residual pass with G[i,j]*(tau_j - tau_i) and the result scaled by g_y
        (kelvin/ccsd.py:1519-1534, 1867-1888)."""
        ti, ng = self.ti, self.ngrid
        Gnew = self.G.copy()
        for i in range(ng):
            for j in range(ng):
                Gnew[i, j] *= (ti[j] - ti[i])
        if self.sys.has_u():
            ea, eb, Ds, (Fa, Fb, Ia, Ib, Iabab) = self._u_setup()
            t1, t2 = ft_cc_equations.uccsd_stanton(
                Fa, Fb, Ia, Ib, Iabab, *self.T1, *self.T2, *Ds, ti, ng, Gnew, **self._u_flags())
            Tt = list(t1) + list(t2)
        else:
            en, D1, D2, F, I = self._g_setup()
            Tt = list(ft_cc_equations.ccsd_stanton(F, I, self.T1, self.T2, D1, D2, ti, ng, Gnew,
                                                   **self._g_flags()))
        gdev = _lib.const_dev(self.g, Tt[0].device)
        for X in Tt:
            X.mul_(gdev.view((ng,) + (1,)*(X.dim() - 1)))
        return Tt

    def _g_ft_rorb(self):
        """kelvin/ccsd.py:1494-1545."""
        Tt = self._tau_weighted_stanton()
        L1, L2 = _lib.as_dev(self.L1), _lib.as_dev(self.L2)
        a, = self._active()
        s = -1.0/self.beta
        o, v = self._rorb_traces([(L1, Tt[0], s, "i", "a"), (L2, Tt[1], 0.25*s, "ij", "ab")],
                                 L1.shape[1], L1.shape[2])
        self.rorbo, self.rorbv = numpy.zeros(a["n"]), numpy.zeros(a["n"])
        self.rorbo[a["iocc"]] -= o
        self.rorbv[a["ivir"]] += v

    def _u_ft_rorb(self):
        """kelvin/ccsd.py:1828-1918."""
        Tt = self._tau_weighted_stanton()
        L1a, L1b = (_lib.as_dev(x) for x in self.L1)
        L2aa, L2ab, L2bb = (_lib.as_dev(x) for x in self.L2)
        act = self._active()
        s = -1.0/self.beta
        oa, va = self._rorb_traces([(L1a, Tt[0], s, "i", "a"), (L2aa, Tt[2], 0.25*s, "ij", "ab"),
                                    (L2ab, Tt[3], s, "i", "a")], L1a.shape[1], L1a.shape[2])
        ob, vb = self._rorb_traces([(L1b, Tt[1], s, "i", "a"), (L2bb, Tt[4], 0.25*s, "ij", "ab"),
                                    (L2ab, Tt[3], s, "j", "b")], L1b.shape[1], L1b.shape[2])
        self.rorbo = [numpy.zeros(a["n"]) for a in act]
        self.rorbv = [numpy.zeros(a["n"]) for a in act]
        for k, (o, v) in enumerate(((oa, va), (ob, vb))):
            self.rorbo[k][act[k]["iocc"]] -= o
            self.rorbv[k][act[k]["ivir"]] += v

    def _grel_ft_1rdm(self):
        """kelvin/ccsd.py:1547-1577."""
        if self.dia is None:
            self._g_ft_1rdm()
        if self.P2 is None:
            self._g_ft_2rdm()
        (fo, fv), = self._occ()
        self._g_ft_ron()
        self._g_ft_rorb()
        rdji = numpy.diag(self.rono) + numpy.diag(self.ron1) + numpy.diag(self.rorbo) + numpy.diag(fo)
        rdba = numpy.diag(self.ronv) + numpy.diag(self.rorbv)
        self.r1rdm = rdji + rdba

    def _urel_ft_1rdm(self):
        """kelvin/ccsd.py:1920-1959."""
        if self.dia is None:
            self._u_ft_1rdm()
        if self.P2 is None:
            self._u_ft_2rdm()
        self._u_ft_ron()
        self._u_ft_rorb()
        (foa, fva), (fob, fvb) = self._occ()
        fo = (foa, fob)
        self.r1rdm = [numpy.diag(self.rono[k] + self.ron1[k] + self.rorbo[k] + fo[k]
                                 + self.ronv[k] + self.rorbv[k]) for k in (0, 1)]

    def _focc_padded(self):
        """Occupations with the entries outside the active occupied set zeroed: adding
        diag() of these over all orbitals equals the reference's numpy.ix_(iocc, iocc)
        updates (kelvin/ccsd.py:1975-1983, cc_utils.py:2027-2073).  The reference's g
        branch adds the untruncated diag(fo) into the iocc block (ccsd.py:2001), which only
        works when nothing is truncated; the truncated vector is used here."""
        return [numpy.where(a["fo"] > self.athresh, a["fo"], 0.0) for a in self._active()]

    def full_1rdm(self, relax=False):
        """Full (HF + correlation) 1-RDM as NumPy arrays (kelvin/ccsd.py:1961-2006)."""
        c = lambda x: x.cpu().numpy()  # noqa: E731
        if self.sys.orbtype == 'u':
            if relax:
                if self.r1rdm is None:
                    self._urel_ft_1rdm()
                n1 = [c(x) for x in self.n1rdm]
                return [self.r1rdm[k] + (n1[k] - numpy.diag(n1[k].diagonal())) for k in (0, 1)]
            if self.n1rdm is None:
                self._u_ft_1rdm()
            foa, fob = self._focc_padded()
            return [c(self.n1rdm[0]) + numpy.diag(foa), c(self.n1rdm[1]) + numpy.diag(fob)]
        elif self.sys.orbtype == 'g':
            if relax:
                if self.r1rdm is None:
                    self._grel_ft_1rdm()
                n1 = c(self.n1rdm)
                return self.r1rdm + (n1 - numpy.diag(n1.diagonal()))
            if self.n1rdm is None:
                self._g_ft_1rdm()
            fo, = self._focc_padded()
            return c(self.n1rdm) + numpy.diag(fo)
        raise Exception("orbital type " + self.sys.orbtype + " is not implemented for 1rdm")

    def full_2rdm(self, relax=False):
        """Full 2-RDM (device tensors; kelvin/ccsd.py:2008-2053)."""
        if relax:
            raise Exception("Rexalex 2-RDM is not implemented")
        if self.sys.orbtype == 'u':
            if self.n1rdm is None:
                self._u_ft_1rdm()
            if self.n2rdm is None:
                self._u_ft_2rdm()
            foa, fob = self._focc_padded()
            rdm2 = [x.clone() for x in self.n2rdm]
            cc_utils.u_full_rdm2(foa, fob, self.n1rdm, rdm2)
            return rdm2
        elif self.sys.orbtype == 'g':
            if self.n1rdm is None:
                self._g_ft_1rdm()
            if self.n2rdm is None:
                self._g_ft_2rdm()
            fo, = self._focc_padded()
            rdm2 = self.n2rdm.clone()
            cc_utils.g_full_rdm2(fo, self.n1rdm, rdm2)
            return rdm2
        raise Exception("orbital type " + self.sys.orbtype + " is not implemented for 1rdm")

    # ------------------------------------------------------------------
    # derivative through the quadrature weights (kelvin/ccsd.py:1075-1148, 1150-1260)
    # ------------------------------------------------------------------
    def _pair_LT(self, Ls, Ts, weights):
        """-(1/beta) sum_y g_y [ sum_blocks w <L_y, T_y> ] with L in o..v.. and T in v..o.. order."""
        dev = Ts[0].device
        tot = torch.zeros(self.ngrid, dtype=torch.float64, device=dev)
        first = True
        for L, Tt, w in zip(Ls, Ts, weights):
            if L.dim() == 3:
                _lib.dot_keep(L, "yia", Tt, "yai", "y", alpha=w, out=tot, beta=0.0 if first else 1.0)
            else:
                _lib.dot_keep(L, "yijab", Tt, "yabij", "y", alpha=w, out=tot,
                              beta=0.0 if first else 1.0)
            first = False
        return -float(numpy.dot(tot.cpu().numpy(), self.g))/self.beta

    def _nocc_gderiv(self):
        from . import _lib as lb
        beta, ng, ti, G, g = self.beta, self.ngrid, self.ti, self.G, self.g
        gd, Gd = quadrature.d_ft_quad(ng, beta, self.quad)
        Gnew = G.copy()
        for i in range(G.shape[0]):
            for j in range(G.shape[1]):
                Gnew[i, j] *= (ti[j] - ti[i])/beta
        lib = lb.load()
        if self.sys.has_u():
            ea, eb, Ds, (Fa, Fb, Ia, Ib, Iabab) = self._u_setup()
            Ts = list(self.T1) + list(self.T2)
            Ls = list(self.L1) + list(self.L2)
            w = (1.0, 1.0, 0.25, 1.0, 0.25)
            dg = ft_cc_energy.ft_ucc_energy(*Ts, Fa.ov, Fb.ov, Ia.oovv, Ib.oovv, Iabab.oovv, gd, beta)

            def stanton(Gx):
                t1, t2 = ft_cc_equations.uccsd_stanton(Fa, Fb, Ia, Ib, Iabab, *Ts, *Ds, ti, ng, Gx,
                                                       **self._u_flags())
                return list(t1) + list(t2)
        else:
            en, D1, D2, F, I = self._g_setup()
            Ds = (D1, D2)
            Ts = [self.T1, self.T2]
            Ls = [self.L1, self.L2]
            w = (1.0, 0.25)
            dg = ft_cc_energy.ft_cc_energy(self.T1, self.T2, F.ov, I.oovv, gd, beta)

            def stanton(Gx):
                return list(ft_cc_equations.ccsd_stanton(F, I, self.T1, self.T2, D1, D2, ti, ng, Gx,
                                                         **self._g_flags()))
        Ls = [lb.as_dev(x) for x in Ls]
        dG = self._pair_LT(Ls, stanton(Gd), w)
        Tn = stanton(Gnew)
        for Tt, D in zip(Tn, Ds):
            rc = lib.kb200_scale_by(ng, D.numel(), lb.ptr(Tt), lb.ptr(D), lb.stream_ptr())
            lb.check(rc, "kb200_scale_by")
        dG += self._pair_LT(Ls, Tn, w)
        return dg, dG

    _g_nocc_gderiv = _nocc_gderiv
    _u_nocc_gderiv = _nocc_gderiv

    # ------------------------------------------------------------------
    def compute_ESN(self, L1=None, L2=None, gderiv=True):
        """Compute energy, entropy, particle number (kelvin/ccsd.py:128-165)."""
        if not gderiv:
            raise Exception("gderiv=False (approximate quadrature derivative) is outside the B200 path")
        if self.T1 is None:
            raise Exception("run() must be called before compute_ESN()")
        if self.L1 is None:
            if self.sys.has_u():
                self._ft_uccsd_lambda(L1=L1, L2=L2)
                ti = time.time()
                self._u_ft_1rdm()
                self._u_ft_2rdm()
                tf = time.time()
                logging.info("RDM construction time: {} s".format(tf - ti))
            else:
                self._ft_ccsd_lambda(L1=L1, L2=L2)
                self._g_ft_1rdm()
                self._g_ft_2rdm()
        if self.sys.has_u():
            ti = time.time()
            self._u_ft_ESN(L1, L2, gderiv=gderiv)
            tf = time.time()
            logging.info("Total derivative time: {} s".format(tf - ti))
        else:
            self._g_ft_ESN(L1, L2, gderiv=gderiv)

    def _finish_ESN(self, N0, E0, N1, Ncc, B1, Bcc):
        beta, mu = self.beta, self.mu
        Bcc -= self.Gcc/beta               # derivative from the explicit factors of 1/beta
        dg, dG = self._nocc_gderiv()
        Bcc += dG + dg
        E1 = beta*B1 + mu*N1 + self.G1
        Ecc = beta*Bcc + mu*Ncc + self.Gcc
        self.N0, self.N1, self.Ncc = N0, N1, Ncc
        self.N = Ncc + N0 + N1
        self.E0, self.E1, self.Ecc = E0, E1, Ecc
        self.E = E0 + E1 + Ecc
        self.S = -beta*(self.Gtot - self.E + mu*self.N)
        self.S0 = -beta*(self.G0 - self.E0 + mu*self.N0)
        self.S1 = -beta*(self.G1 - self.E1 + mu*self.N1)
        self.Scc = self.S - self.S0 - self.S1

    def _g_ft_ESN(self, L1=None, L2=None, gderiv=True):
        """kelvin/ccsd.py:167-214."""
        beta, mu = self.beta, self.mu
        if self.dia is None:
            self._g_ft_1rdm()
            self._g_ft_2rdm()
        self._g_ft_ron()
        en = self.sys.g_energies_tot()
        fo = ft_utils.ff(beta, en, mu)
        B0 = ft_utils.dGP0(beta, en, mu)
        N0 = fo.sum()
        E0 = beta*B0.sum() + mu*N0 + self.G0
        dvec = -numpy.ones(en.shape)
        N1 = -numpy.einsum('i,i->', dvec, self.ron1)
        Ncc = -numpy.einsum('i,i->', dvec, self.rono + self.ronv)
        dvec = (en - mu)/beta
        B1 = numpy.einsum('i,i->', dvec, self.ron1)
        Bcc = numpy.einsum('i,i->', dvec, self.rono + self.ronv)
        self._finish_ESN(N0, E0, N1, Ncc, B1, Bcc)

    def _u_ft_ESN(self, L1=None, L2=None, gderiv=True):
        """kelvin/ccsd.py:216-271, including the reference's use of rono[0] + ronv[1]
        for both spins (quirk Q2, :238-239,246-247)."""
        beta, mu = self.beta, self.mu
        if self.dia is None:
            self._u_ft_1rdm()
            self._u_ft_2rdm()
        self._u_ft_ron()
        ea, eb = self.sys.u_energies_tot()
        foa = ft_utils.ff(beta, ea, mu)
        fob = ft_utils.ff(beta, eb, mu)
        B0a = ft_utils.dGP0(beta, ea, mu)
        B0b = ft_utils.dGP0(beta, eb, mu)
        N0 = foa.sum() + fob.sum()
        E0 = beta*(B0a.sum() + B0b.sum()) + mu*N0 + self.G0
        dveca = -numpy.ones(ea.shape)
        dvecb = -numpy.ones(eb.shape)
        mixed = self.rono[0] + self.ronv[1]
        N1 = -(numpy.einsum('i,i->', dveca, self.ron1[0]) + numpy.einsum('i,i->', dvecb, self.ron1[1]))
        Ncc = -(numpy.einsum('i,i->', dveca, mixed) + numpy.einsum('i,i->', dvecb, mixed))
        dveca = (ea - mu)/beta
        dvecb = (eb - mu)/beta
        B1 = numpy.einsum('i,i->', dveca, self.ron1[0]) + numpy.einsum('i,i->', dvecb, self.ron1[1])
        Bcc = numpy.einsum('i,i->', dveca, mixed) + numpy.einsum('i,i->', dvecb, mixed)
        self._finish_ESN(N0, E0, N1, Ncc, B1, Bcc)

"""CPU: the oracle (restated cqcpy layer + restated drivers) against the numbers the
reference publishes and against fixtures produced by the reference's own drivers."""
import os

import numpy
import pytest

from golden import published as pub
from kelvin_oracle import cqc, cc_equations as ocq, driver as odrv, spin_blocked as sb
import util

HERE = os.path.dirname(os.path.abspath(__file__))


def _ueg(T, mu, norb, orb, L=2*numpy.pi, cut=1.2):
    from kelvin_b200.ueg_system import UEGSystem      # host-side input generator only
    return UEGSystem(T, L, cut, mu=mu, norb=norb, orbtype=orb)


def _conv(damp, econv=1e-8, tconv=None, max_iter=50):
    return {"econv": econv, "tconv": 1000*econv if tconv is None else tconv,
            "max_iter": max_iter, "damp": damp}


def test_ueg7_omega_cc_published_g():
    """kelvin/tests/test_ft_ccsd.py:23: -0.010750238811 (1e-9), general orbitals."""
    T, mu, ng = 0.1, 0.1, 10
    s = _ueg(T, mu, 7, "g")
    beta = 1.0/T
    en = s.g_energies_tot()
    F, I = odrv.ft_integrals(s, en, beta, mu)
    D1, D2 = cqc.D1(en, en), cqc.D2(en, en)
    ti, g, G = odrv.simpsons(ng, beta)
    T1, T2 = odrv.mp2_guess_g(F, I, D1, D2, ti, ng, G)
    E, T1, T2, hist = odrv.ft_cc_iter(T1, T2, F, I, D1, D2, g, G, beta, ng, ti, _conv(0.2))
    assert abs(E - pub.UEG7_OMEGA_CC) < 1e-9
    assert len(hist) == 11


def test_ueg7_omega_cc_published_u_and_fixture():
    """Same system, unrestricted; also the converged amplitudes of the reference
    drivers (tests/golden/ueg7_u.npz)."""
    T, mu = 0.1, 0.1
    s = _ueg(T, mu, 7, "u")
    beta = 1.0/T
    ea, eb = s.u_energies_tot()
    ints = odrv.uft_integrals(s, ea, eb, beta, mu)
    Ds = (cqc.D1(ea, ea), cqc.D1(eb, eb), cqc.D2(ea, ea), cqc.D2u(ea, eb, ea, eb), cqc.D2(eb, eb))
    ng = 10
    ti, g, G = odrv.simpsons(ng, beta)
    amps = odrv.mp2_guess_u(*ints, *Ds, ti, ng, G)
    E, T1, T2, hist = odrv.ft_ucc_iter(*amps, *ints, *Ds, g, G, beta, ng, ti, _conv(0.2))
    assert abs(E - pub.UEG7_OMEGA_CC) < 1e-9
    ref = numpy.load(os.path.join(HERE, "golden", "ueg7_u.npz"))
    ng = 6
    ti, g, G = odrv.simpsons(ng, beta)
    amps = odrv.mp2_guess_u(*ints, *Ds, ti, ng, G)
    E, T1, T2, hist = odrv.ft_ucc_iter(*amps, *ints, *Ds, g, G, beta, ng, ti,
                                       _conv(0.2, 1e-11, 1e-9, 80))
    assert abs(E - float(ref["Ecc"])) < 1e-12
    assert numpy.abs(T2[1] - ref["T2ab"]).max() < 1e-12
    assert numpy.abs(T1[0] - ref["T1a"]).max() < 1e-12
    # one Lambda map application at the reference's converged Lambda is a fixed point
    Ls = [ref[k] for k in ("L1a", "L1b", "L2aa", "L2ab", "L2bb")]
    Ts = [ref[k] for k in ("T1a", "T1b", "T2aa", "T2ab", "T2bb")]
    new = odrv.uccsd_lambda_opt(*ints, *Ts, *Ls, *Ds, ti, ng, g, G, beta)
    for a, b in zip(new, Ls):
        assert numpy.abs(a - b).max() < 2e-8*max(1.0, numpy.abs(b).max())


def test_esn19_header_and_mp2():
    """bench/ueg_ft_ccsd_ESN19/ulambda_19_04_17.out:3-9: N0, density, r_s, T_F, MP2."""
    T, mu, ng = 0.5, 7.0, 10
    s = _ueg(T, mu, 19, "u", L=1.942, cut=30.0)
    assert abs(s.N - pub.ESN19["N0"]) < 1e-9
    assert abs(s.den - pub.ESN19["density"]) < 1e-9
    assert abs(s.rs - pub.ESN19["rs"]) < 1e-9
    assert abs(s.Tf - pub.ESN19["Tf"]) < 1e-9
    beta = 1.0/T
    ea, eb = s.u_energies_tot()
    ints = odrv.uft_integrals(s, ea, eb, beta, mu)
    Fa, Fb, Ia, Ib, Iabab = ints
    Ds = (cqc.D1(ea, ea), cqc.D1(eb, eb), cqc.D2(ea, ea), cqc.D2u(ea, eb, ea, eb), cqc.D2(eb, eb))
    ti, g, G = odrv.simpsons(ng, beta)
    amps = odrv.mp2_guess_u(*ints, *Ds, ti, ng, G)
    E2 = odrv.ft_ucc_energy(*amps, Fa.ov, Fb.ov, Ia.oovv, Ib.oovv, Iabab.oovv, g, beta, Qterm=False)
    assert "%.10f" % E2 == "%.10f" % pub.ESN19["MP2"]


def test_spin_blocked_equals_embedding():
    """u == g block by block (kelvin/tests/test_ft_cc_ampl.py:41-110) for the CPU port
    used as the timed baseline."""
    ints, amps = util.random_u(5, 4, 1, seed=3)
    T1 = (amps[0][0], amps[1][0])
    T2 = (amps[2][0], amps[3][0], amps[4][0])
    ref = ocq.u_stanton_terms(*ints, T1, T2)
    got = sb.u_stanton_terms(*ints, T1, T2)
    for a, b in zip(ref, got):
        assert numpy.abs(a - b).max() < 1e-12


def test_lambda_is_lagrangian_derivative():
    """The restated Lambda map equals L - (beta/g_y) dL/dT_y by central differences on a
    few elements (the content of kelvin/tests/test_ft_lambda.py:212-285)."""
    n, ng, beta = 3, 3, 1.7
    F, I, t1, t2 = util.random_g(n, ng, seed=5, scale=0.1)
    _, _, l1, l2 = util.random_g(n, ng, seed=6, scale=0.1)
    l1 = numpy.ascontiguousarray(l1.transpose(0, 2, 1))
    l2 = numpy.ascontiguousarray(l2.transpose(0, 3, 4, 1, 2))
    e = util.random_D(n)
    D1, D2 = cqc.D1(e, e), cqc.D2(e, e)
    ti, g, G = odrv.simpsons(ng, beta)

    def lagr(T1, T2):
        Eterm = odrv.ft_cc_energy(T1, T2, F.ov, I.oovv, g, beta)
        N1, N2 = odrv.ccsd_stanton(F, I, T1, T2, D1, D2, ti, ng, G)
        A1 = numpy.einsum('via,vai->v', l1, N1 - T1)
        A2 = 0.25*numpy.einsum('vijab,vabij->v', l2, N2 - T2)
        return Eterm - numpy.dot(g, A1 + A2)/beta
    L1n, L2n = odrv.ccsd_lambda_opt(F, I, t1, t2, l1, l2, D1, D2, ti, ng, g, G, beta)
    d = 1e-5
    for (y, a, i) in ((0, 0, 1), (1, 2, 0), (2, 1, 1)):
        tp, tm = t1.copy(), t1.copy()
        tp[y, a, i] += d
        tm[y, a, i] -= d
        fd = (lagr(tp, t2) - lagr(tm, t2))/(2*d)
        got = -(L1n[y, i, a] - l1[y, i, a])*g[y]/beta
        assert abs(fd - got) < 1e-7, (y, a, i, fd, got)

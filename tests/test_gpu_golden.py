"""End-to-end golden numbers published by the reference, replayed on the GPU."""
import logging

import numpy
import pytest

from golden import published as pub

pytestmark = pytest.mark.gpu


def _iteration_lines(records):
    out = []
    for r in records:
        parts = r.getMessage().split()
        if len(parts) == 3 and parts[0].isdigit():
            out.append((int(parts[0]), parts[1], parts[2]))
    return out


def _check_log(caplog, gold, nit=None):
    lines = [r.getMessage() for r in caplog.records]
    mp2 = [l for l in lines if l.startswith("MP2 Energy")][0]
    assert mp2 == "MP2 Energy: {:.10f}".format(gold["MP2"])
    its = _iteration_lines(caplog.records)
    ref = gold["iters"] if nit is None else gold["iters"][:nit]
    assert len(its) == len(ref)
    for (k, e, r), (eg, rg) in zip(its, ref):
        assert abs(float(e) - eg) <= 1.01e-10, (k, e, eg)          # all 10 printed digits
        assert abs(float(r) - rg) <= 2e-4*rg, (k, r, rg)            # 4 printed digits


def test_ueg7_omega_cc_g_and_u(built):
    """kelvin/tests/test_ft_ccsd.py:23,119-149."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.1, 0.1
    for orb in ("g", "u"):
        ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype=orb)
        cc = ccsd(ueg, T=T, mu=mu, iprint=0, max_iter=50, damp=0.2, ngrid=10)
        Etot, Ecc = cc.run()
        assert abs(Ecc - pub.UEG7_OMEGA_CC) < 1e-9, (orb, Ecc)


def test_ueg7_ng40_trajectory(built, caplog):
    """examples/ueg_ft_cc_compare.out:9-26 (u path, 14 iterations)."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.1, 0.1182968
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='u')
    assert abs(ueg.N - pub.UEG7_NG40["N0"]) < 1e-9
    cc = ccsd(ueg, T=T, mu=mu, iprint=1, max_iter=50, damp=0.1, ngrid=40, tconv=1e-8)
    with caplog.at_level(logging.INFO):
        Etot, Ecc = cc.run()
    _check_log(caplog, pub.UEG7_NG40)
    assert abs(Etot - pub.UEG7_NG40["Omega"]) < 1e-9
    assert abs(Ecc - pub.UEG7_NG40["OmegaC"]) < 1e-10


def test_esn19_trajectory(built, caplog):
    """bench/ueg_ft_ccsd_ESN19/ulambda_19_04_17.out:2-25: BASELINE config 0."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.5, 7.0
    ueg = UEGSystem(T, 1.942, 30.0, mu=mu, norb=19, orbtype='u')
    assert abs(ueg.N - pub.ESN19["N0"]) < 1e-9
    assert abs(ueg.den - pub.ESN19["density"]) < 1e-9
    assert abs(ueg.rs - pub.ESN19["rs"]) < 1e-9
    assert abs(ueg.Tf - pub.ESN19["Tf"]) < 1e-9
    cc = ccsd(ueg, T=T, mu=mu, iprint=1, max_iter=50, damp=0.0, ngrid=10)
    with caplog.at_level(logging.INFO):
        Etot, Ecc = cc.run()
    _check_log(caplog, pub.ESN19)
    assert abs(Etot - pub.ESN19["Omega"]) < 1e-9
    assert abs(Ecc - pub.ESN19["OmegaC"]) < 1e-10


def test_esn33_trajectory(built, caplog):
    """bench/ueg_ft_ccsd_ESN33/overview_19_05_11.out:1-27: the north-star config."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.5, 7.0
    ueg = UEGSystem(T, 1.942, 30.0, mu=mu, norb=33, orbtype='u')
    assert abs(ueg.N - pub.ESN33["N0"]) < 1e-9
    cc = ccsd(ueg, T=T, mu=mu, iprint=1, max_iter=50, damp=0.0, ngrid=10)
    with caplog.at_level(logging.INFO):
        Etot, Ecc = cc.run()
    _check_log(caplog, pub.ESN33)
    assert abs(Etot - pub.ESN33["Omega"]) < 1e-9
    assert abs(Ecc - pub.ESN33["OmegaC"]) < 1e-10


def test_hubbard6(built, caplog):
    """examples/hubbard1d.out:2-16."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.hubbard_system import HubbardSystem, Hubbard1D
    L, U, T, mu = 6, 1.0, 1.0, 0.0
    hub = Hubbard1D(L, 1.0, U, boundary='p')
    Oa = numpy.zeros(L)
    Ob = numpy.zeros(L)
    Oa[0::2] = 1.0
    Ob[1::2] = 1.0
    Pa = numpy.einsum('i,j->ij', Oa, Oa)
    Pb = numpy.einsum('i,j->ij', Ob, Ob)
    sys_ = HubbardSystem(T, hub, Pa, Pb, mu=mu)
    cc = ccsd(sys_, iprint=1, max_iter=80, econv=1e-11, T=T, mu=mu)
    with caplog.at_level(logging.INFO):
        Eout, Ecc = cc.run()
    assert abs((Eout - Ecc) - pub.HUBBARD6["E01"]) < 1e-11
    assert abs(Ecc - pub.HUBBARD6["OmegaC"]) < 1e-11
    assert len(_iteration_lines(caplog.records)) == pub.HUBBARD6["niter"]

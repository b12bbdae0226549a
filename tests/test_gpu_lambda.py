"""GPU parity of the Lambda map (mechanically derived VJP plan) against the oracle."""
import logging

import numpy
import pytest

from golden import published as pub

pytestmark = pytest.mark.gpu

from kelvin_oracle import cqc, driver as odrv  # noqa: E402
import util  # noqa: E402


def _relerr(got, ref):
    return numpy.abs(got.cpu().numpy() - ref).max()/numpy.abs(ref).max()


@pytest.mark.parametrize("n,ng", [(4, 3), (8, 2)])
def test_ccsd_lambda_opt_g(built, n, ng):
    from kelvin_b200 import ft_cc_equations
    F, I, t1, t2 = util.random_g(n, ng, seed=n)
    _, _, l1, l2 = util.random_g(n, ng, seed=n + 50)
    l1 = numpy.ascontiguousarray(l1.transpose(0, 2, 1))
    e = util.random_D(n)
    D1, D2 = cqc.D1(e, e), cqc.D2(e, e)
    beta = 1.3
    ti, g, G = odrv.simpsons(ng, beta)
    r1, r2 = odrv.ccsd_lambda_opt(F, I, t1, t2, l1, l2, D1, D2, ti, ng, g, G, beta)
    o1, o2 = ft_cc_equations.ccsd_lambda_opt(F, I, t1, t2, l1, l2, D1, D2, ti, ng, g, G, beta)
    assert _relerr(o1, r1) < 1e-11
    assert _relerr(o2, r2) < 1e-11
    g1, g2 = odrv.ccsd_lambda_guess(F, I, t1, beta, ng)
    q1, q2 = ft_cc_equations.ccsd_lambda_guess(F, I, t1, beta, ng)
    assert _relerr(q1, g1) < 1e-12
    assert _relerr(q2, g2) < 1e-12


@pytest.mark.parametrize("na,nb,ng", [(4, 3, 2), (6, 6, 3)])
def test_uccsd_lambda_opt(built, na, nb, ng):
    from kelvin_b200 import ft_cc_equations
    ints, amps = util.random_u(na, nb, ng, seed=na + nb)
    _, lam = util.random_u(na, nb, ng, seed=na + nb + 100)
    lam = (numpy.ascontiguousarray(lam[0].transpose(0, 2, 1)),
           numpy.ascontiguousarray(lam[1].transpose(0, 2, 1)), lam[2], lam[3], lam[4])
    ea, eb = util.random_D(na, 1), util.random_D(nb, 2)
    Ds = (cqc.D1(ea, ea), cqc.D1(eb, eb), cqc.D2(ea, ea), cqc.D2u(ea, eb, ea, eb), cqc.D2(eb, eb))
    beta = 0.9
    ti, g, G = odrv.simpsons(ng, beta)
    ref = odrv.uccsd_lambda_opt(*ints, *amps, *lam, *Ds, ti, ng, g, G, beta)
    got = ft_cc_equations.uccsd_lambda_opt(*ints, *amps, *lam, *Ds, ti, ng, g, G, beta)
    for a, b in zip(got, ref):
        assert _relerr(a, b) < 1e-11
    ref = odrv.uccsd_lambda_guess(*ints, amps[0], amps[1], beta, ng)
    got = ft_cc_equations.uccsd_lambda_guess(*ints, amps[0], amps[1], beta, ng)
    for a, b in zip(got, ref):
        assert _relerr(a, b) < 1e-12


def _lambda_lines(records):
    out = []
    for r in records:
        parts = r.getMessage().split()
        if len(parts) == 2 and parts[0].isdigit():
            out.append(float(parts[1]))
    return out


def test_esn19_lambda_converges(built, caplog):
    """ESN19 Lambda solve converges; the residual trajectory printed in the 2019
    bench log (bench/ueg_ft_ccsd_ESN19/ulambda_19_04_17.out:26-38) is NOT reproduced
    by today's reference drivers ('Lambda trajectories'), the fixed
    point is (tests/test_gpu_rdm.py::test_esn19_ESN)."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.5, 7.0
    ueg = UEGSystem(T, 1.942, 30.0, mu=mu, norb=19, orbtype='u')
    cc = ccsd(ueg, T=T, mu=mu, iprint=1, max_iter=50, damp=0.0, ngrid=10)
    cc.run()
    with caplog.at_level(logging.INFO):
        cc._ft_uccsd_lambda()
    res = _lambda_lines(caplog.records)
    assert res[-1] < 1e-5 and len(res) < 50
    # same trajectory as the restated reference path (oracle run recorded in tests/golden)
    ref = [1210.3672128695, 4.2006910027, 0.7661313809, 1.0017663926, 0.2359343579, 0.1062881637]
    for a, b in zip(res, ref):
        assert abs(a - b) <= 2e-9*max(1.0, abs(b)), (a, b)


def test_ueg7_ng40_lambda_residuals(built, caplog):
    """examples/ueg_ft_cc_compare.out:27-42: 16 Lambda residuals (damp 0.1, tconv 1e-8)."""
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.1, 0.1182968
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='u')
    cc = ccsd(ueg, T=T, mu=mu, iprint=1, max_iter=50, damp=0.1, ngrid=40, tconv=1e-8)
    cc.run()
    with caplog.at_level(logging.INFO):
        cc._ft_uccsd_lambda()
    res = _lambda_lines(caplog.records)
    ref = pub.UEG7_NG40["lambda_res"]
    assert len(res) == len(ref)
    for a, b in zip(res, ref):
        assert abs(a - b) <= 1.01e-10, (a, b)

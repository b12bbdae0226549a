"""GPU path vs fixtures produced by the unmodified reference drivers
(tests/golden/make_golden.py): amplitudes to 1e-9 relative, energies to 1e-10."""
import os

import numpy
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    return numpy.load(os.path.join(HERE, "golden", name + ".npz"))


def _rel(got, ref):
    got = got.cpu().numpy() if hasattr(got, "cpu") else numpy.asarray(got)
    return numpy.abs(got - ref).max()/max(1e-300, numpy.abs(ref).max())


def _hubbard4(T=1.0):
    from kelvin_b200.hubbard_system import HubbardSystem, Hubbard1D
    L = 4
    hub = Hubbard1D(L, 1.0, 2.0, boundary='p')
    Oa, Ob = numpy.zeros(L), numpy.zeros(L)
    Oa[0::2] = 1.0
    Ob[1::2] = 1.0
    return HubbardSystem(T, hub, numpy.einsum('i,j->ij', Oa, Oa), numpy.einsum('i,j->ij', Ob, Ob),
                         mu=0.3, orbtype='u')


def _check_scalars(cc, Etot, Ecc, ref):
    assert abs(Etot - float(ref["Etot"])) < 1e-10
    assert abs(Ecc - float(ref["Ecc"])) < 1e-10
    for nm, key in (("E", "E"), ("S", "S"), ("N", "N"), ("E0", "E0"), ("E1", "E1"), ("Ecc", "Ecc_"),
                    ("N0", "N0"), ("N1", "N1"), ("Ncc", "Ncc"), ("S0", "S0"), ("S1", "S1"),
                    ("Scc", "Scc")):
        assert abs(getattr(cc, nm) - float(ref[key])) < 1e-9, (nm, getattr(cc, nm), float(ref[key]))


def _run_u(sysm, ref, **kw):
    from kelvin_b200.ccsd import ccsd
    cc = ccsd(sysm, **kw)
    Etot, Ecc = cc.run()
    cc.compute_ESN()
    _check_scalars(cc, Etot, Ecc, ref)
    for k, nm in enumerate(("T1a", "T1b")):
        assert _rel(cc.T1[k], ref[nm]) < 1e-9
        assert _rel(cc.L1[k], ref["L1" + nm[-1]]) < 1e-8
    for k, nm in enumerate(("aa", "ab", "bb")):
        assert _rel(cc.T2[k], ref["T2" + nm]) < 1e-9
        assert _rel(cc.L2[k], ref["L2" + nm]) < 1e-8
    for k in (0, 1):
        assert _rel(cc.n1rdm[k], ref["n1rdm%d" % k]) < 1e-8
        assert numpy.abs(cc.rono[k] - ref["rono%d" % k]).max() < 1e-9
        assert numpy.abs(cc.ronv[k] - ref["ronv%d" % k]).max() < 1e-9
        assert numpy.abs(cc.ron1[k] - ref["ron1%d" % k]).max() < 1e-12
        for nm in ("dia", "dba", "dji", "dai"):
            assert _rel(getattr(cc, nm)[k], ref["%s%d" % (nm, k)]) < 1e-8
    for k in (0, 1, 2):
        assert _rel(cc.n2rdm[k], ref["n2rdm%d" % k]) < 1e-8
    for b, tup in enumerate(cc.P2):
        for k, P in enumerate(tup):
            assert _rel(P, ref["P2_%d_%d" % (b, k)]) < 1e-8, (b, k)
    # relaxed / full density matrices (kelvin/ccsd.py:1828-2053)
    full1 = cc.full_1rdm()
    rel1 = cc.full_1rdm(relax=True)
    full2 = cc.full_2rdm()
    for k in (0, 1):
        assert _rel(full1[k], ref["full1rdm%d" % k]) < 1e-8
        assert _rel(rel1[k], ref["rel1rdm%d" % k]) < 1e-8
        assert numpy.abs(cc.rorbo[k] - ref["rorbo%d" % k]).max() < 1e-9
        assert numpy.abs(cc.rorbv[k] - ref["rorbv%d" % k]).max() < 1e-9
    for k in (0, 1, 2):
        assert _rel(full2[k], ref["full2rdm%d" % k]) < 1e-8
    # trace of the relaxed 1-RDM is the particle number (kelvin/tests/test_ft_cc_relden.py:36-43)
    assert abs(numpy.trace(cc.r1rdm[0]) + numpy.trace(cc.r1rdm[1]) - cc.N) < 1e-10


def test_ueg7_u_full_path(built):
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.1, 0.1
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='u')
    _run_u(ueg, _load("ueg7_u"), T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6,
           econv=1e-11, tconv=1e-9)


def test_hubbard4_u_full_path(built):
    _run_u(_hubbard4(), _load("hubbard4_u"), T=1.0, mu=0.3, iprint=0, max_iter=80, ngrid=8,
           quad='quad', econv=1e-11, tconv=1e-9)


def _run_g(sysm, ref, **kw):
    from kelvin_b200.ccsd import ccsd
    cc = ccsd(sysm, **kw)
    Etot, Ecc = cc.run()
    cc.compute_ESN()
    _check_scalars(cc, Etot, Ecc, ref)
    assert tuple(cc.T2.shape) == ref["T2"].shape
    assert _rel(cc.T1, ref["T1"]) < 1e-9
    assert _rel(cc.T2, ref["T2"]) < 1e-9
    assert _rel(cc.L1, ref["L1"]) < 1e-8
    assert _rel(cc.L2, ref["L2"]) < 1e-8
    assert _rel(cc.n1rdm, ref["n1rdm"]) < 1e-8
    assert _rel(cc.n2rdm, ref["n2rdm"]) < 1e-8
    assert numpy.abs(cc.rono - ref["rono"]).max() < 1e-9
    assert numpy.abs(cc.ronv - ref["ronv"]).max() < 1e-9
    for b, P in enumerate(cc.P2):
        assert _rel(P, ref["P2_%d" % b]) < 1e-8, b
    if "full1rdm" in ref:
        assert _rel(cc.full_1rdm(), ref["full1rdm"]) < 1e-8
    assert _rel(cc.full_1rdm(relax=True), ref["rel1rdm"]) < 1e-8
    assert numpy.abs(cc.rorbo - ref["rorbo"]).max() < 1e-9
    assert numpy.abs(cc.rorbv - ref["rorbv"]).max() < 1e-9
    assert _rel(cc.full_2rdm(), ref["full2rdm"]) < 1e-8
    # kelvin/tests/test_ft_cc_relden.py:36-43: tr(relaxed 1-RDM) == N
    assert abs(numpy.trace(cc.r1rdm) - cc.N) < 1e-12


def test_ueg7_g_full_path(built):
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.1, 0.1
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='g')
    _run_g(ueg, _load("ueg7_g"), T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6,
           econv=1e-11, tconv=1e-9)


def test_ueg7_g_active_space(built):
    """athresh > 0: nocc=14, nvir=12 rectangular blocks (kelvin/ccsd.py:642-660)."""
    from kelvin_b200.ueg_system import UEGSystem
    T, mu = 0.05, 0.3
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='g')
    _run_g(ueg, _load("ueg7_g_active"), T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6,
           econv=1e-11, tconv=1e-9, athresh=0.01)


def test_hubbard4_u_active_space(built):
    """athresh > 0, unrestricted: nocc=2, nvir=3 per spin (kelvin/ccsd.py:745-785)."""
    ref = _load("hubbard4_u_active")
    assert ref["T1a"].shape == (8, 3, 2)
    _run_u(_hubbard4(0.5), ref, T=0.5, mu=0.3, iprint=0, max_iter=150, damp=0.2, ngrid=8,
           econv=1e-11, tconv=1e-9, athresh=0.05)


# ---------------------------------------------------------------------------
# sibling solvers on the same kernels; fixtures: make_golden.py variants
# ---------------------------------------------------------------------------
def _ueg7(orb):
    from kelvin_b200.ueg_system import UEGSystem
    return UEGSystem(0.1, 2*numpy.pi, 1.2, mu=0.1, norb=7, orbtype=orb)


@pytest.mark.parametrize("orb", ["u", "g"])
def test_pointwise_solver_fixture(built, orb):
    """rt_iter='point' (kelvin/cc_utils.py:176-242,320-411): grid points converged one after
    the other from linearly extrapolated guesses."""
    from kelvin_b200.ccsd import ccsd
    ref = _load("ueg7_variants")
    cc = ccsd(_ueg7(orb), T=0.1, mu=0.1, iprint=0, max_iter=50, damp=0.2, tconv=1e-8, ngrid=6,
              rt_iter="point")
    Etot, Ecc = cc.run()
    assert abs(Etot - float(ref["point_%s_Etot" % orb])) < 1e-10
    assert abs(Ecc - float(ref["point_%s_Ecc" % orb])) < 1e-10
    if orb == "u":
        for k, nm in enumerate(("T1a", "T1b")):
            assert _rel(cc.T1[k], ref["point_u_" + nm]) < 1e-9
        for k, nm in enumerate(("T2aa", "T2ab", "T2bb")):
            assert _rel(cc.T2[k], ref["point_u_" + nm]) < 1e-9
    else:
        assert _rel(cc.T1, ref["point_g_T1"]) < 1e-9
        assert _rel(cc.T2, ref["point_g_T2"]) < 1e-9


def test_pointwise_agrees_with_all_at_once(built):
    """kelvin/tests/test_ft_ccsd.py:166-201: both convergence schemes reach the same Omega_cc."""
    from kelvin_b200.ccsd import ccsd
    kw = dict(T=0.1, mu=0.1, iprint=0, max_iter=50, damp=0.2, tconv=1e-8, ngrid=10)
    E1 = ccsd(_ueg7("u"), **kw).run()[1]
    E2 = ccsd(_ueg7("u"), rt_iter="point", **kw).run()[1]
    assert abs(E1 - E2) < 1e-8


def test_ccd_fixture(built):
    """singles=False: FT-CCD amplitudes and Lambda (general spin orbitals); the unrestricted
    loops refuse it with the reference's message (kelvin/cc_utils.py:84)."""
    from kelvin_b200.ccsd import ccsd
    ref = _load("ueg7_variants")
    kw = dict(T=0.1, mu=0.1, iprint=0, max_iter=80, damp=0.2, ngrid=6, econv=1e-11, tconv=1e-9,
              singles=False)
    cc = ccsd(_ueg7("g"), **kw)
    Etot, Ecc = cc.run()
    assert abs(Etot - float(ref["ccd_Etot"])) < 1e-10
    assert abs(Ecc - float(ref["ccd_Ecc"])) < 1e-10
    assert float(cc.T1.abs().max()) == 0.0 and numpy.abs(ref["ccd_T1"]).max() == 0.0
    assert _rel(cc.T2, ref["ccd_T2"]) < 1e-9
    cc._ft_ccsd_lambda()
    assert _rel(cc.L2, ref["ccd_L2"]) < 1e-8
    assert float(cc.L1.abs().max()) == 0.0
    with pytest.raises(Exception, match="Unrecognized method keyword for unrestricted calc"):
        ccsd(_ueg7("u"), **kw).run()


@pytest.mark.parametrize("method", ["CCD", "LCCSD", "LCCD"])
def test_method_switches_vs_oracle(built, method):
    """cc_utils.form_new_ampl / the Lambda maps for the CCD, LCCSD and LCCD keywords
    (kelvin/cc_utils.py:32-47,441-454) against the oracle's term classes."""
    import sys
    sys.path.insert(0, os.path.join(HERE))
    import util
    from kelvin_oracle import cc_equations as ocq, cqc, driver as odrv
    from kelvin_b200 import cc_utils, ft_cc_equations
    no, nv, ng, beta = 5, 7, 3, 1.3
    F, I, t1, t2, l1, l2 = util.random_g_rect(no, nv, ng, seed=31)
    ev, eo = util.random_D(nv, 1), util.random_D(no, 2)
    D1, D2 = cqc.D1(ev, eo), cqc.D2(ev, eo)
    ti, g, G = odrv.simpsons(ng, beta)
    T1, T2 = cc_utils.form_new_ampl(method, F, I, t1, t2, D1, D2, ti, ng, G)
    r1 = numpy.stack([-F.vo]*ng)
    r2 = numpy.stack([-I.vvoo]*ng)
    L1i, L2i = odrv.int_L(ng, l1, ti, D1, g, G), odrv.int_L(ng, l2, ti, D2, g, G)
    o1 = numpy.zeros_like(l1)
    o2 = numpy.zeros_like(l2)
    for y in range(ng):
        if method in ("CCD", "LCCD"):
            ocq._D_D(r2[y], F, I, t2[y], fac=-1.0)
            ocq._LD_LD(o2[y], F, I, L2i[y], fac=-1.0)
            if method == "CCD":
                ocq._D_DD(r2[y], F, I, t2[y], fac=-1.0)
                ocq._LD_LDTD(o2[y], I, L2i[y], t2[y], fac=-1.0)
        else:
            ocq._S_S(r1[y], F, I, t1[y], fac=-1.0)
            ocq._S_D(r1[y], F, I, t2[y], fac=-1.0)
            ocq._D_S(r2[y], F, I, t1[y], fac=-1.0)
            ocq._D_D(r2[y], F, I, t2[y], fac=-1.0)
            ocq._LS_LS(o1[y], F, I, L1i[y], fac=-1.0)
            ocq._LS_LD(o1[y], F, I, L2i[y], fac=-1.0)
            ocq._LD_LS(o2[y], F, I, L1i[y], fac=-1.0)
            ocq._LD_LD(o2[y], F, I, L2i[y], fac=-1.0)
            ocq._LS_TS(o1[y], I, t1[y], fac=-1.0)
            o1[y] -= F.ov
        o2[y] -= I.oovv/beta if method == "LCCD" else I.oovv
    assert _rel(T2, odrv.int_tbar(ng, r2, ti, D2, G)) < 1e-11
    if method == "LCCSD":
        assert _rel(T1, odrv.int_tbar(ng, r1, ti, D1, G)) < 1e-11
        L1, L2 = ft_cc_equations.lccsd_lambda_simple(F, I, t1, t2, l1, l2, D1, D2, ti, ng, g, G, beta)
        assert _rel(L1, o1) < 1e-11
    elif method == "CCD":
        assert T1 is t1
        L2 = ft_cc_equations.ccd_lambda_simple(F, I, t2, l2, D2, ti, ng, g, G, beta)
    else:
        assert T1 is t1
        L2 = ft_cc_equations.lccd_lambda_simple(F, I, t2, l2, D2, ti, ng, g, G, beta)
    assert _rel(L2, o2) < 1e-11

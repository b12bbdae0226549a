"""GPU parity of the per-grid-point residual + integration against the oracle."""
import numpy
import pytest

pytestmark = pytest.mark.gpu

from kelvin_oracle import cqc, driver as odrv  # noqa: E402
import util  # noqa: E402


def _relerr(got, ref):
    return numpy.abs(got.cpu().numpy() - ref).max()/numpy.abs(ref).max()


@pytest.mark.parametrize("n,ng", [(4, 3), (9, 2), (14, 2)])
def test_ccsd_stanton_g(built, n, ng):
    from kelvin_b200 import ft_cc_equations
    F, I, t1, t2 = util.random_g(n, ng, seed=n)
    e = util.random_D(n)
    D1, D2 = cqc.D1(e, e), cqc.D2(e, e)
    ti, g, G = odrv.simpsons(ng, 1.3)
    r1, r2 = odrv.ccsd_stanton(F, I, t1, t2, D1, D2, ti, ng, G)
    o1, o2 = ft_cc_equations.ccsd_stanton(F, I, t1, t2, D1, D2, ti, ng, G)
    assert _relerr(o1, r1) < 1e-11
    assert _relerr(o2, r2) < 1e-11


@pytest.mark.parametrize("na,nb,ng", [(4, 3, 2), (7, 7, 3), (10, 9, 2)])
def test_uccsd_stanton(built, na, nb, ng):
    from kelvin_b200 import ft_cc_equations
    ints, amps = util.random_u(na, nb, ng, seed=na + nb)
    ea, eb = util.random_D(na, 1), util.random_D(nb, 2)
    D1a, D1b = cqc.D1(ea, ea), cqc.D1(eb, eb)
    D2aa, D2ab, D2bb = cqc.D2(ea, ea), cqc.D2u(ea, eb, ea, eb), cqc.D2(eb, eb)
    ti, g, G = odrv.simpsons(ng, 0.9)
    r1, r2 = odrv.uccsd_stanton(*ints, *amps, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G)
    o1, o2 = ft_cc_equations.uccsd_stanton(*ints, *amps, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G)
    for got, ref in zip(list(o1) + list(o2), list(r1) + list(r2)):
        assert _relerr(got, ref) < 1e-11


def test_u_equals_g_on_spin_packed_inputs(built):
    """The reference's own u==g check (kelvin/tests/test_ft_cc_ampl.py:41-110)
    applied to the CUDA kernels."""
    from kelvin_b200 import ft_cc_equations
    na = nb = 6
    ng = 2
    ints, amps = util.random_u(na, nb, ng, seed=21)
    Fa, Fb, Ia, Ib, Iabab = ints
    F, I = cqc.F_to_spin(Fa, Fb), cqc.I_to_spin(Ia, Ib, Iabab)
    t1 = numpy.stack([cqc.T1_to_spin(amps[0][y], amps[1][y], na, na, nb, nb) for y in range(ng)])
    t2 = numpy.stack([cqc.T2_to_spin(amps[2][y], amps[3][y], amps[4][y], na, na, nb, nb)
                      for y in range(ng)])
    g1, g2 = ft_cc_equations.ccsd_stanton_bar(F, I, t1, t2)
    u = ft_cc_equations.uccsd_stanton_bar(*ints, *amps)
    g1, g2 = g1.cpu().numpy(), g2.cpu().numpy()
    sc = numpy.abs(g2).max()
    assert numpy.abs(u[0].cpu().numpy() - g1[:, :na, :na]).max() < 1e-11*sc
    assert numpy.abs(u[1].cpu().numpy() - g1[:, na:, na:]).max() < 1e-11*sc
    assert numpy.abs(u[2].cpu().numpy() - g2[:, :na, :na, :na, :na]).max() < 1e-11*sc
    assert numpy.abs(u[3].cpu().numpy() - g2[:, :na, na:, :na, na:]).max() < 1e-11*sc
    assert numpy.abs(u[4].cpu().numpy() - g2[:, na:, na:, na:, na:]).max() < 1e-11*sc


def test_tau_sharded_step_equals_reference_loop(built):
    """The bench / multi-GPU step object (world size 1) reproduces one iteration of the
    reference loop (kelvin/cc_utils.py:274-305): energy and residual as logged."""
    from kelvin_b200 import cc_utils, ft_utils, parallel, quadrature
    from kelvin_b200.ueg_system import UEGSystem
    T, mu, ng = 0.5, 7.0, 6
    s = UEGSystem(T, 1.942, 30.0, mu=mu, norb=7, orbtype='u')
    beta = 1.0/T
    ea, eb = s.u_energies_tot()
    ti, g, G = quadrature.ft_quad(ng, beta, 'lin')
    ints = cc_utils.uft_integrals(s, ea, eb, beta, mu)
    Ds = (ft_utils.D1(ea, ea), ft_utils.D1(eb, eb), ft_utils.D2(ea, ea),
          ft_utils.D2u(ea, eb, ea, eb), ft_utils.D2(eb, eb))
    oints = odrv.uft_integrals(s, ea, eb, beta, mu)
    oD = [d.cpu().numpy() for d in Ds]
    amps = odrv.mp2_guess_u(*oints, *oD, ti, ng, G)
    solver = parallel.TauShardedUCCSD(*ints, Ds, g, G, beta, ng, ti)
    solver.set_amplitudes(*amps)
    E1, r1 = solver.step(0.3)
    E2, r2 = solver.step(0.3)
    conv = {"econv": 0.0, "tconv": 0.0, "max_iter": 2, "damp": 0.3}
    _, _, _, hist = odrv.ft_ucc_iter(*amps, *oints, *oD, g, G, beta, ng, ti, conv)
    assert abs(E1 - hist[0][0]) < 1e-12 and abs(E2 - hist[1][0]) < 1e-12
    assert abs(r1 - hist[0][1]) < 1e-10*max(1.0, hist[0][1])
    assert abs(r2 - hist[1][1]) < 1e-10*max(1.0, hist[1][1])


def test_tau0_shortcut(built):
    """With T[0] == 0 the residual at the first grid point is the bare driver; the
    shortcut (t0_zero=True: that point is not evaluated) gives what the full evaluation gives."""
    from kelvin_b200 import ft_cc_equations
    ng = 4
    ints, amps = util.random_u(6, 5, ng, seed=5)
    amps = [a.copy() for a in amps]
    for a in amps:
        a[0] = 0.0
    ti, g, G = odrv.simpsons(ng, 0.9)
    assert ft_cc_equations.t0_is_zero(G, [__import__("torch").as_tensor(a) for a in amps])
    full = ft_cc_equations.uccsd_stanton_bar(*ints, *amps)
    fast = ft_cc_equations.uccsd_stanton_bar(*ints, *amps, t0_zero=True)
    Fa, Fb, Ia, Ib, Iabab = ints
    for got, ref, drv in zip(fast, full, (Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo)):
        ref = ref.cpu().numpy()
        assert numpy.array_equal(got[0].cpu().numpy(), -drv)
        assert numpy.abs(ref[0] + drv).max() < 1e-13*numpy.abs(drv).max()
        assert _relerr(got, ref) < 1e-12
    F, I, t1, t2 = util.random_g(7, 3, seed=3)
    t1[0] = 0.0
    t2[0] = 0.0
    f1, f2 = ft_cc_equations.ccsd_stanton_bar(F, I, t1, t2)
    s1, s2 = ft_cc_equations.ccsd_stanton_bar(F, I, t1, t2, t0_zero=True)
    assert _relerr(s1, f1.cpu().numpy()) < 1e-12
    assert _relerr(s2, f2.cpu().numpy()) < 1e-12
    assert numpy.array_equal(s2[0].cpu().numpy(), -I.vvoo)


def test_closed_shell_reduction(built):
    """Mirror-symmetric (alpha == beta) unrestricted inputs: the reduced program (only the
    alpha-leading block of each spin-flip pair is evaluated) gives the full program's residual
    and Lambda map; the detection accepts these inputs and rejects generic ones."""
    import torch
    from kelvin_b200 import ft_cc_equations as fe
    n, ng = 7, 3
    ints, amps, lam = util.random_u_closed(n, ng, seed=9)
    lam = [numpy.ascontiguousarray(x) for x in
           (lam[0].transpose(0, 2, 1), lam[1].transpose(0, 2, 1), lam[2].transpose(0, 3, 4, 1, 2),
            lam[3].transpose(0, 3, 4, 1, 2), lam[4].transpose(0, 3, 4, 1, 2))]
    assert fe.closed_shell_integrals(*ints)
    assert fe.closed_shell_amplitudes(*amps)
    assert fe.closed_shell_amplitudes(*lam)
    gints, gamps = util.random_u(n, n, ng, seed=10)
    assert not fe.closed_shell_integrals(*gints)
    assert not fe.closed_shell_amplitudes(*gamps)
    bad = [a.copy() for a in amps]
    bad[3][1, 2, 3, 1, 0] += 1e-6
    assert not fe.closed_shell_amplitudes(*bad)
    full = fe.uccsd_stanton_bar(*ints, *amps)
    red = fe.uccsd_stanton_bar(*ints, *amps, closed_shell=True)
    for got, ref in zip(red, full):
        assert _relerr(got, ref.cpu().numpy()) < 1e-12
    e = util.random_D(n, 3)
    D1, D2, D2ab = cqc.D1(e, e), cqc.D2(e, e), cqc.D2u(e, e, e, e)
    ti, g, G = odrv.simpsons(ng, 1.1)
    args = (*ints, *amps, *lam, D1, D1, D2, D2ab, D2, ti, ng, g, G, 1.1)
    lfull = fe.uccsd_lambda_opt(*args)
    lred = fe.uccsd_lambda_opt(*args, closed_shell=True)
    for got, ref in zip(lred, lfull):
        assert _relerr(got, ref.cpu().numpy()) < 1e-12
    # the loop takes the reduced path on its own and lands on the same amplitudes
    from kelvin_b200 import cc_utils
    conv = {"econv": 1e-10, "tconv": 1e-8, "max_iter": 4, "damp": 0.3}
    small = [0.05*a for a in amps]
    E1, T1s, T2s = cc_utils.ft_ucc_iter("CCSD", *small, *ints, D1, D1, D2, D2ab, D2, g, G, 1.1, ng,
                                        ti, 0, conv)
    keep = fe.CLOSED_SHELL
    fe.CLOSED_SHELL = 0
    try:
        E2, U1s, U2s = cc_utils.ft_ucc_iter("CCSD", *small, *ints, D1, D1, D2, D2ab, D2, g, G, 1.1,
                                            ng, ti, 0, conv)
    finally:
        fe.CLOSED_SHELL = keep
    assert abs(E1 - E2) < 1e-12*max(1.0, abs(E2))
    for got, ref in zip(list(T1s) + list(T2s), list(U1s) + list(U2s)):
        assert _relerr(got, ref.cpu().numpy()) < 1e-11
    assert torch.equal(T2s[0], T2s[2])


@pytest.mark.parametrize("closed", [False, True])
def test_multi_stream_plan_is_bit_identical(built, closed):
    """kb200_plan_run overlaps independent launches on side streams; every per-slot read/write
    order is kept by events, so the results equal the single-stream run bit for bit (repeated
    to give a missing dependency the chance to show)."""
    import torch
    from kelvin_b200 import _lib, ft_cc_equations as fe
    lib = _lib.load()
    n, ng = 12, 2
    if closed:
        ints, amps, _ = util.random_u_closed(n, ng, seed=77)
    else:
        ints, amps = util.random_u(n, n - 1, ng, seed=78)
    old = lib.kb200_set_plan_streams(1)
    try:
        ref = [x.clone() for x in fe.uccsd_stanton_bar(*ints, *amps, closed_shell=closed)]
        lib.kb200_set_plan_streams(3)
        for _ in range(8):
            got = fe.uccsd_stanton_bar(*ints, *amps, closed_shell=closed)
            torch.cuda.synchronize()
            for a, b in zip(got, ref):
                assert torch.equal(a, b)
    finally:
        lib.kb200_set_plan_streams(old)


def test_closed_shell_mirror_rows(built):
    """From MIRROR_ROWS_MIN_BATCH grid points on, the closed-shell program also contracts its
    mirror-symmetric opposite-spin ladder terms on the rows p <= q only (plan.mirror_outputs),
    and the same-spin ladder terms on the a<b, i<j triangle (plan.antisym_outputs)."""
    from kelvin_b200 import ft_cc_equations as fe, plan as _plan
    n, ng = 9, 5
    assert ng >= fe.MIRROR_ROWS_MIN_BATCH
    ints, amps, _ = util.random_u_closed(n, ng, seed=19)
    full = fe.uccsd_stanton_bar(*ints, *amps)
    red = fe.uccsd_stanton_bar(*ints, *amps, closed_shell=True)
    p = fe.stanton_plan("u", fe._u_sizes(ints[0], ints[1]), -1.0, mirror=True, mirror_rows=True)
    assert sum(1 for op in p.low.rops if op.tri is not None) == 12
    assert any(s.startswith(_plan.TRI_PREFIX + "m") for s in p.shapes)
    for got, ref in zip(red, full):
        assert _relerr(got, ref.cpu().numpy()) < 1e-12

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


def test_step_object_equals_reference_loop(built):
    """The step object the loops and the bench share (fused integrate / damp / norm / energy
    pass per block, closed-shell + singlet program) reproduces the iterations of the reference
    loop (kelvin/cc_utils.py:274-305): energy and residual as logged."""
    from kelvin_b200 import cc_utils, ft_utils, quadrature
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
    solver = cc_utils.UccStep(amps, *ints, Ds, g, G, beta, ng, ti)
    assert solver.cs and solver.singlet and solver.antisym and solver.t0
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


def test_closed_shell_singlet_and_sumdiff(built):
    """Closed-shell inputs with T2aa = T2ab - T2ab(a<->b): the singlet-reduced program with the
    sum/difference ring contractions (what the loops run for the UEG benchmark) equals the
    general program; the detection accepts these amplitudes and rejects perturbed ones."""
    from kelvin_b200 import ft_cc_equations as fe
    n, ng = 9, 5
    ints, amps, _ = util.random_u_closed(n, ng, seed=23)
    assert fe.is_singlet(amps[2], amps[3]) and fe.is_antisymmetric(amps[2])
    bad = amps[2].copy()
    bad[1, 2, 3, 1, 0] += 1e-6
    assert not fe.is_singlet(bad, amps[3]) and not fe.is_antisymmetric(bad)
    full = fe.uccsd_stanton_bar(*ints, *amps)
    for sd in (0, 1):
        keep = fe.SUMDIFF
        fe.SUMDIFF = sd
        try:
            red = fe.uccsd_stanton_bar(*ints, *amps, closed_shell=True, singlet=True)
        finally:
            fe.SUMDIFF = keep
        for got, ref in zip(red, full):
            assert _relerr(got, ref.cpu().numpy()) < 1e-12
    p = fe.stanton_plan("u", fe._u_sizes(ints[0], ints[1]), -1.0, mirror=True, mirror_rows=True,
                        singlet=True)
    assert not any(s.startswith(("rg.aaaa", "Woooo.aa", "Wvvvv.aa")) for s in p.shapes)


def test_non_antisymmetric_guess_takes_the_full_sums(built):
    """A doubles guess that is NOT antisymmetric (the reference accepts any array): the half
    sums over contracted pairs and the triangular outputs would silently compute something
    else; the guard detects it and the plans without those rewrites reproduce the reference's
    full double sums (oracle: the Stanton equations on the array as given)."""
    from kelvin_b200 import ft_cc_equations as fe
    from kelvin_oracle import cc_equations as ocq
    n, ng = 6, 2
    F, I, t1, t2 = util.random_g(n, ng, seed=31)
    rng = numpy.random.default_rng(5)
    t2 = t2 + 0.05*rng.standard_normal(t2.shape)
    assert not fe.is_antisymmetric(t2)
    b1, b2 = fe.ccsd_stanton_bar(F, I, t1, t2)
    for y in range(ng):
        R1, R2 = ocq.stanton_terms(F, I, t1[y], t2[y])
        assert numpy.abs(b1[y].cpu().numpy() - (-F.vo - R1)).max() < 1e-11*numpy.abs(R1).max()
        assert numpy.abs(b2[y].cpu().numpy() - (-I.vvoo - R2)).max() < 1e-11*numpy.abs(R2).max()
    # with the antisymmetry assumed, the result differs visibly: the guard matters
    w1, w2 = fe.ccsd_stanton_bar(F, I, t1, t2, antisym=True)
    assert _relerr(w2, b2.cpu().numpy()) > 1e-6


def test_general_program_large_tiles(built):
    """The GENERAL unrestricted program (alpha != beta, na != nb) at a size where the hot kernel
    configuration engages -- 128x128 full tiles, launch groups of 8, half sums over antisymmetric
    contracted pairs, triangular outputs -- against the oracle's Sz-blocked port
    (kelvin/ft_cc_equations.py:130-164)."""
    from kelvin_b200 import ft_cc_equations as fe, plan as _plan
    from kelvin_oracle import spin_blocked as sb
    na, nb, ng = 24, 23, 2
    ints, amps = util.random_u(na, nb, ng, seed=41, scale=0.05)
    got = fe.uccsd_stanton_bar(*ints, *amps)
    p = fe.stanton_plan("u", fe._u_sizes(ints[0], ints[1]), -1.0)
    ops = p.low.finalize(ng)
    big = [o for o in ops if o.kind == 0 and o.tile == _plan.BIG_TILE and o.K >= 250]
    assert len(big) >= 32 and any(o.group >= 4 for o in big)
    assert sum(1 for op in p.low.rops if op.tri is not None) == 8
    w = sb.wrap_integrals(*ints)
    Fa, Fb, Ia, Ib, Iabab = ints
    drv = (Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo)
    for y in range(ng):
        rs = sb.u_stanton_terms(*ints, (amps[0][y], amps[1][y]),
                                (amps[2][y], amps[3][y], amps[4][y]), wrapped=w)
        for k, r in enumerate(rs):
            ref = -drv[k] - r
            assert numpy.abs(got[k][y].cpu().numpy() - ref).max() < 1e-11*numpy.abs(ref).max(), k


@pytest.mark.parametrize("prog", ["stanton-closed", "stanton", "lambda-closed"])
def test_hybrid_partition_two_ranks_on_one_gpu(built, prog):
    """plan.hybrid_phases / engine.PhasedPlan on the GPU: two 'ranks' (two PhasedPlan instances
    driven in lockstep on one device, the exchanges done as sums of their buffers) evaluate the
    same grid points together, each contracting its slab of the rows of every large contraction;
    both end up with the single-rank result."""
    import torch
    from kelvin_b200 import _lib, engine, ft_cc_equations as fe, plan as _plan
    dev = _lib.device()
    n, ng = 12, 2
    world = 2
    if prog == "stanton":
        ints, amps = util.random_u(n, n - 1, ng, seed=61, scale=0.1)
        ref = fe.uccsd_stanton_bar(*ints, *amps)
        rops, ins, outs = fe._stanton_rops("u", -1.0, False, False, False, False, True)
        inputs = dict(zip(fe._U_TIN, amps))
        live = range(5)
    elif prog == "stanton-closed":
        ints, amps, _ = util.random_u_closed(n, ng, seed=62, scale=0.1)
        ref = fe.uccsd_stanton_bar(*ints, *amps, closed_shell=True, singlet=True)
        rops, ins, outs = fe._stanton_rops("u", -1.0, True, False, True, True, True)
        inputs = {k: v for k, v in zip(fe._U_TIN, amps) if k in ins}
        live = (0, 2, 3)
    else:
        ints, amps, lam = util.random_u_closed(n, ng, seed=63, scale=0.1)
        lam = [numpy.ascontiguousarray(x) for x in
               (lam[0].transpose(0, 2, 1), lam[1].transpose(0, 2, 1), lam[2].transpose(0, 3, 4, 1, 2),
                lam[3].transpose(0, 3, 4, 1, 2), lam[4].transpose(0, 3, 4, 1, 2))]
        inter, rest = fe._lambda_rops("u", -1.0, True, True)
        rops = inter + rest
        ins, outs = fe._reps(fe._U_T + fe._U_L), fe._reps(fe._U_LO)
        inputs = {k: v for k, v in zip(fe._U_T + fe._U_L, list(amps) + lam) if k in ins}
        p1 = engine.Plan(rops, "u", fe._u_sizes(ints[0], ints[1]), ins, outs, name="ref")
        t = fe._u_integral_slots(*ints, dev, [s for s in p1.inputs if _plan.is_integral_slot(s)])
        t.update({k: _lib.as_dev(v, dev) for k, v in inputs.items()})
        ref = [None]*5
        for k, nm in enumerate(fe._U_LO):
            if nm in outs:
                ref[k] = t[nm] = torch.empty((ng,) + tuple(p1.shapes[nm]), dtype=torch.float64, device=dev)
        p1.run(t, ng)
        live = (0, 2, 3)
    sizes = fe._u_sizes(ints[0], ints[1])
    ranks = [engine.PhasedPlan(rops, "u", sizes, ins, outs, world, name="hyb%d" % r,
                               min_work=n**5) for r in range(world)]
    assert sum(1 for ph in ranks[0].hp.phases for op in ph if op.slab) >= 8
    assert sum(1 for e in ranks[0].hp.exchange if e) >= 2
    tens = []
    for r, hp in enumerate(ranks):
        t = fe._u_integral_slots(*ints, dev, [s for s in hp.inputs if _plan.is_integral_slot(s)])
        t.update({k: _lib.as_dev(v, dev) for k, v in inputs.items()})
        for nm in hp.outputs:
            t[nm] = torch.empty((ng,) + tuple(hp.shapes[nm]), dtype=torch.float64, device=dev)
        tens.append(t)
        hp.begin(t, ng, r)
    for p in range(len(ranks[0].plans)):
        for hp in ranks:
            hp.run_phase(p)
        bufs = [hp.exchange_buffer(p) for hp in ranks]
        if bufs[0] is not None:
            tot = bufs[0] + bufs[1]
            for b in bufs:
                b.copy_(tot)
    names = fe._U_LO if prog.startswith("lambda") else fe._U_TOUT
    for r in range(world):
        for k in live:
            assert _relerr(tens[r][names[k]], ref[k].cpu().numpy()) < 1e-12, (r, k)

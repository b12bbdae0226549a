"""GPU parity tests of the individual kernels through the C ABI (-m gpu)."""
import math

import numpy
import pytest
import torch

pytestmark = pytest.mark.gpu

from kelvin_oracle import cqc, driver as odrv  # noqa: E402
import util  # noqa: E402


def _dev(x):
    return torch.as_tensor(numpy.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("shape", [
    # letters out / A / B, dims
    ("abij", "aeim", "mbej", dict(a=7, b=5, i=6, j=4, e=9, m=3)),      # ring-like, permuted
    ("abij", "abef", "efij", dict(a=13, b=11, i=9, j=10, e=12, f=7)),  # ladder, natural
    ("ae", "mnef", "afmn", dict(a=9, e=8, m=7, n=6, f=5)),             # long K, tiny output (split-K)
    ("abef", "amef", "bm", dict(a=6, b=7, e=5, f=8, m=9)),             # skinny N (tile 1)
    ("abij", "ai", "bj", dict(a=5, b=6, i=7, j=8)),                    # outer product (K=1)
    ("ai", "me", "aeim", dict(a=33, i=33, m=33, e=33)),                # N=1 (matrix-vector)
])
def test_contraction_vs_einsum(built, shape):
    from kelvin_b200 import engine, plan
    lc, la, lb, dims = shape
    rng = numpy.random.default_rng(3)
    nb = 3
    A = rng.standard_normal((nb,) + tuple(dims[l] for l in la))
    B = rng.standard_normal(tuple(dims[l] for l in lb))       # batch-invariant operand
    C0 = rng.standard_normal((nb,) + tuple(dims[l] for l in lc))
    ops = [plan.ROp(("C", lc), 0.75, [("A", la), ("B", lb)]),
           plan.ROp(("C", lc), -1.25, [("B", lb), ("A", la)])]
    shapes = {"C": C0.shape[1:], "A": A.shape[1:], "B": B.shape}
    batched = {"C": True, "A": True, "B": False}
    p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], preset_outputs=["C"], shapes=shapes,
                    batched=batched)
    t = {"A": _dev(A), "B": _dev(B), "C": _dev(C0)}
    p.run(t, nb)
    ref = C0 + (0.75 - 1.25)*numpy.einsum("y%s,%s->y%s" % (la, lb, lc), A, B)
    got = t["C"].cpu().numpy()
    assert numpy.abs(got - ref).max() < 1e-11*max(1.0, numpy.abs(ref).max())


def test_contraction_large_tiles(built):
    """Multiple 128x128 tiles, ragged edges, k not a multiple of 16, both operand modes."""
    from kelvin_b200 import engine, plan
    rng = numpy.random.default_rng(4)
    dims = dict(a=19, b=17, i=19, j=18, e=19, f=13)
    for la, lb in (("abef", "efij"), ("efab", "ijef"), ("abef", "ijef"), ("efab", "efij")):
        A = rng.standard_normal(tuple(dims[l] for l in la))
        B = rng.standard_normal(tuple(dims[l] for l in lb))
        ops = [plan.ROp(("C", "abij"), 1.0, [("A", la), ("B", lb)])]
        shapes = {"C": tuple(dims[l] for l in "abij"), "A": A.shape, "B": B.shape}
        p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], shapes=shapes,
                        batched={"C": False, "A": False, "B": False})
        t = {"A": _dev(A), "B": _dev(B),
             "C": torch.full(shapes["C"], float("nan"), dtype=torch.float64, device="cuda")}
        p.run(t, 1)
        ref = numpy.einsum("%s,%s->abij" % (la, lb), A, B)
        assert numpy.abs(t["C"].cpu().numpy() - ref).max() < 1e-11*numpy.abs(ref).max()


@pytest.mark.parametrize("ng,n,mode", [(10, 7, 0), (10, 7, 1), (5, 33, 1), (40, 5, 1), (2, 4, 0), (17, 6, 1),
                                        (16, 5, 1), (16, 5, 0)])
def test_int_tbar(built, ng, n, mode):
    from kelvin_b200 import quadrature
    rng = numpy.random.default_rng(ng*100 + n)
    e = util.random_D(n)
    D2 = cqc.D2(e, e)
    beta = 2.0
    ti, g, G = odrv.simpsons(ng, beta)
    for Guse in (G, rng.standard_normal((ng, ng))):            # incl. non-triangular G (quirk Q4)
        tb = rng.standard_normal((ng,) + D2.shape)
        ref = odrv.int_tbar(ng, tb, ti, D2, Guse)
        got = quadrature.int_tbar(ng, tb, ti, D2, Guse, mode=mode).cpu().numpy()
        assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()
    D1 = cqc.D1(e, e)
    tb = rng.standard_normal((ng,) + D1.shape)
    ref = odrv.int_tbar(ng, tb, ti, D1, G)
    got = quadrature.int_tbar1(ng, tb, ti, D1, G).cpu().numpy()
    assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()


@pytest.mark.parametrize("ng,n,mode", [(10, 6, 0), (10, 6, 1), (9, 7, 1), (40, 4, 1), (16, 4, 1), (16, 4, 0)])
def test_int_L(built, ng, n, mode):
    from kelvin_b200 import quadrature
    rng = numpy.random.default_rng(ng*10 + n)
    ea, eb = util.random_D(n, 1), util.random_D(n + 1, 2)
    D2 = cqc.D2u(ea, eb, ea, eb)                     # (a,B,i,J), rectangular
    ti, g, G = odrv.simpsons(ng, 2.0)
    for Guse in (G, rng.standard_normal((ng, ng))):
        L2 = rng.standard_normal((ng, n, n + 1, n, n + 1))
        ref = odrv.int_L(ng, L2, ti, D2, g, Guse)
        got = quadrature.int_L(ng, L2, ti, D2, g, Guse, mode=mode).cpu().numpy()
        assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()
    D1 = cqc.D1(ea, eb)                              # (a, i) with different sizes
    L1 = rng.standard_normal((ng, n + 1, n))
    ref = odrv.int_L(ng, L1, ti, D1, g, G)
    got = quadrature.int_L1(ng, L1, ti, D1, g, G).cpu().numpy()
    assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()


def test_energy_and_norms(built):
    from kelvin_b200 import cc_utils, ft_cc_energy
    ng, n = 6, 7
    F, I, t1, t2 = util.random_g(n, ng, seed=11)
    ti, g, G = odrv.simpsons(ng, 1.5)
    for q in (True, False):
        ref = odrv.ft_cc_energy(t1, t2, F.ov, I.oovv, g, 1.5, Qterm=q)
        got = ft_cc_energy.ft_cc_energy(t1, t2, F.ov, I.oovv, g, 1.5, Qterm=q)
        assert abs(got - ref) < 1e-12*abs(ref)
    (Fa, Fb, Ia, Ib, Iabab), (T1a, T1b, T2aa, T2ab, T2bb) = util.random_u(5, 5, ng, seed=12)
    for q in (True, False):
        ref = odrv.ft_ucc_energy(T1a, T1b, T2aa, T2ab, T2bb, Fa.ov, Fb.ov, Ia.oovv, Ib.oovv,
                                 Iabab.oovv, g, 1.5, Qterm=q)
        got = ft_cc_energy.ft_ucc_energy(T1a, T1b, T2aa, T2ab, T2bb, Fa.ov, Fb.ov, Ia.oovv,
                                         Ib.oovv, Iabab.oovv, g, 1.5, Qterm=q)
        assert abs(got - ref) < 1e-12*abs(ref)
    # damping + norms
    old = _dev(t2.copy())
    new = _dev(t2[::-1].copy())
    st = cc_utils._Stats(1, old.device)
    st.damp(0, old, new, 0.3)
    s = st.read()[0]
    assert abs(s[0] - numpy.sum((t2[::-1] - t2)**2)) < 1e-12*s[0]
    assert abs(s[1] - numpy.sum(t2**2)) < 1e-12*s[1]
    upd = 0.3*t2 + 0.7*t2[::-1]
    assert numpy.abs(old.cpu().numpy() - upd).max() < 1e-15
    assert abs(s[2] - numpy.sum(upd**2)) < 1e-12*s[2]


def test_empty_and_bad_inputs(built):
    from kelvin_b200 import quadrature, _lib
    ti, g, G = odrv.simpsons(4, 1.0)
    with pytest.raises(Exception):
        quadrature.int_tbar(4, numpy.zeros((3, 2, 2)), ti, numpy.zeros((2, 2)), G)
    with pytest.raises(Exception):
        quadrature.ft_quad(4, 1.0, 'nope')
    out = quadrature.int_tbar(4, numpy.zeros((4, 0, 3)), ti, numpy.zeros((0, 3)), G)
    assert out.shape == (4, 0, 3)


def test_contraction_wave_split_k(built):
    """Few-CTA big-tile contraction: the wave-quantisation split-K path (deterministic
    workspace reduction) gives the same result as einsum."""
    from kelvin_b200 import engine, plan
    rng = numpy.random.default_rng(8)
    dims = dict(a=12, b=13, i=12, j=13, e=24, f=25)
    A = rng.standard_normal((2,) + tuple(dims[l] for l in "abef"))
    B = rng.standard_normal((2,) + tuple(dims[l] for l in "efij"))
    C0 = rng.standard_normal((2,) + tuple(dims[l] for l in "abij"))
    ops = [plan.ROp(("C", "abij"), 0.5, [("A", "abef"), ("B", "efij")])]
    shapes = {"C": C0.shape[1:], "A": A.shape[1:], "B": B.shape[1:]}
    p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], preset_outputs=["C"], shapes=shapes,
                    batched={"C": True, "A": True, "B": True})
    arr = p.low.finalize(2)
    assert arr[0].splitk > 1
    t = {"A": _dev(A), "B": _dev(B), "C": _dev(C0)}
    p.run(t, 2)
    ref = C0 + 0.5*numpy.einsum("yabef,yefij->yabij", A, B)
    assert numpy.abs(t["C"].cpu().numpy() - ref).max() < 1e-11*numpy.abs(ref).max()


@pytest.mark.parametrize("dims,nb", [
    (dict(e=33, a=31, m=9, n=8, f=33), 3),        # T=5 tile, K=2376: one K chunk per batch
    (dict(e=19, a=17, m=11, n=10, f=19), 3),      # T=3 tile
    (dict(e=33, a=33, m=33, n=33, f=33), 2),      # ESN33 shape, split over K
    (dict(e=40, a=1, m=16, n=16, f=9), 2),        # N = 1 column
])
def test_contraction_long_k_tiny_output(built, dims, nb):
    """Tile 6 (longk_kernel): whole <= 40x40 output per CTA, K split across CTAs, all four
    operand-contiguity modes; F_vv/F_oo-build shapes of kelvin/ft_cc_equations.py:96-164."""
    from kelvin_b200 import engine, plan
    rng = numpy.random.default_rng(12)
    for la, lb in (("emnf", "afmn"), ("mnfe", "afmn"), ("emnf", "fmna"), ("mnfe", "fmna")):
        A = rng.standard_normal(tuple(dims[l] for l in la))             # integral-like, unbatched
        B = rng.standard_normal((nb,) + tuple(dims[l] for l in lb))
        C0 = rng.standard_normal((nb, dims["a"], dims["e"]))
        ops = [plan.ROp(("C", "ae"), -0.5, [("A", la), ("B", lb)])]
        shapes = {"C": C0.shape[1:], "A": A.shape, "B": B.shape[1:]}
        p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], preset_outputs=["C"], shapes=shapes,
                        batched={"C": True, "A": False, "B": True})
        arr = p.low.finalize(nb)
        assert arr[0].tile == 6
        if dims["m"] == 33:
            assert arr[0].splitk > 1
        t = {"A": _dev(A), "B": _dev(B), "C": _dev(C0)}
        p.run(t, nb)
        ref = C0 - 0.5*numpy.einsum("%s,y%s->yae" % (la, lb), A, B)
        assert numpy.abs(t["C"].cpu().numpy() - ref).max() < 1e-11*numpy.abs(ref).max(), (la, lb)


@pytest.mark.parametrize("k,n", [(33, 33), (19, 19), (40, 40), (36, 7), (5, 40)])
def test_contraction_skinny_streaming(built, k, n):
    """Tile 7 (skinny_kernel): n^5 dressing terms, M ~ n^3 rows against a small K x N matrix,
    every operand/output contiguity combination, accumulate and overwrite."""
    from kelvin_b200 import engine, plan
    rng = numpy.random.default_rng(21)
    dims = dict(a=17, e=16, f=17, m=k, b=n)
    nb = 2
    for la, lb, lc in (("amef", "bm", "abef"), ("aefm", "mb", "abef"), ("amef", "mb", "aefb"),
                       ("aefm", "bm", "aefb")):
        A = rng.standard_normal((nb,) + tuple(dims[l] for l in la))
        B = rng.standard_normal((nb,) + tuple(dims[l] for l in lb))
        C0 = rng.standard_normal((nb,) + tuple(dims[l] for l in lc))
        for preset in (True, False):
            ops = [plan.ROp(("C", lc), -1.5, [("A", la), ("B", lb)])]
            shapes = {"C": C0.shape[1:], "A": A.shape[1:], "B": B.shape[1:]}
            p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], preset_outputs=["C"] if preset else [],
                            shapes=shapes, batched={"C": True, "A": True, "B": True})
            assert p.low.finalize(nb)[0].tile == 7
            t = {"A": _dev(A), "B": _dev(B), "C": _dev(C0)}
            p.run(t, nb)
            ref = -1.5*numpy.einsum("y%s,y%s->y%s" % (la, lb, lc), A, B) + (C0 if preset else 0.0)
            assert numpy.abs(t["C"].cpu().numpy() - ref).max() < 1e-11*numpy.abs(ref).max(), (la, lb, lc)


@pytest.mark.parametrize("ng", [113, 320])
def test_integration_large_grids(built, ng):
    """No limit on the grid size (round 1 staged all rows in shared memory and refused
    ngrid > ~113; the reference's own tests and examples use 200, 280, 320, 400 and 2400 points,
    e.g. kelvin/tests/test_quadrature.py): output rows are processed 16 at a time in
    registers, the grid arrays come from global memory beyond 72 points."""
    from kelvin_b200 import quadrature
    rng = numpy.random.default_rng(ng)
    n = 3
    e = util.random_D(n)
    D2 = cqc.D2(e, e)
    ti, g, G = odrv.simpsons(ng, 1.0)
    tb = rng.standard_normal((ng,) + D2.shape)
    for mode in (0, 1):
        ref = odrv.int_tbar(ng, tb, ti, D2, G)
        got = quadrature.int_tbar(ng, tb, ti, D2, G, mode=mode).cpu().numpy()
        assert numpy.abs(got - ref).max() < 1e-11*numpy.abs(ref).max()
        L2 = rng.standard_normal((ng, n, n, n, n))
        ref = odrv.int_L(ng, L2, ti, D2, g, G)
        got = quadrature.int_L(ng, L2, ti, D2, g, G, mode=mode).cpu().numpy()
        assert numpy.abs(got - ref).max() < 1e-11*numpy.abs(ref).max()
    # the driver runs on such a grid (MP2 guess + two iterations)
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    s = UEGSystem(0.5, 2*numpy.pi, 1.2, mu=0.2, norb=7, orbtype='g')
    cc = ccsd(s, T=0.5, mu=0.2, ngrid=ng, max_iter=2)
    cc.run()


def test_strided_rows_and_fused_update(built):
    """Blocks that are column ranges of one wide (ng, Ntot) buffer: integration, int_L and the
    damping kernel take row strides; the fused update (integration + residual norm + damping +
    new norm + energy term in one pass, kb200_int_tbar_update) equals the separate kernels."""
    import torch
    from kelvin_b200 import cc_utils, ft_cc_energy, ft_cc_equations as fe, quadrature
    rng = numpy.random.default_rng(3)
    ng, n = 10, 6
    F, I, t1, t2 = util.random_g(n, ng, seed=21)
    e = util.random_D(n)
    D1, D2 = cqc.D1(e, e), cqc.D2(e, e)
    ti, g, G = odrv.simpsons(ng, 2.0)
    flat, (b1, b2) = fe.flat_rows(ng, ((n, n), (n, n, n, n)), _dev(t1).device)
    tb1, tb2 = rng.standard_normal(t1.shape), rng.standard_normal(t2.shape)
    b1.copy_(_dev(tb1))
    b2.copy_(_dev(tb2))
    assert not b2.is_contiguous()
    for rows in (None, (3, 7)):
        ref = odrv.int_tbar(ng, tb2, ti, D2, G)
        got = quadrature.int_tbar(ng, b2, ti, D2, G, rows=rows).cpu().numpy()
        ref = ref if rows is None else ref[rows[0]:rows[1]]
        assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()
    L2 = numpy.ascontiguousarray(tb2.transpose(0, 3, 4, 1, 2))
    lflat, (l2,) = fe.flat_rows(ng, ((n, n, n, n),), flat.device)
    lflat[:, :] = 0.0
    l2.copy_(_dev(L2))
    wide = torch.zeros((ng, 3*n**4), dtype=torch.float64, device=flat.device)
    l2w = wide[:, n**4:2*n**4].unflatten(1, (n, n, n, n))
    l2w.copy_(l2)
    ref = odrv.int_L(ng, L2, ti, D2, g, G)
    got = quadrature.int_L(ng, l2w, ti, D2, g, G).cpu().numpy()
    assert numpy.abs(got - ref).max() < 1e-12*numpy.abs(ref).max()
    # fused update vs integrate -> damp_norms -> energy
    alpha, beta = 0.3, 2.0
    T1a, T2a = _dev(t1.copy()), _dev(t2.copy())
    T1b, T2b = _dev(t1.copy()), _dev(t2.copy())
    n1 = quadrature.int_tbar(ng, b1, ti, D1, G)
    n2 = quadrature.int_tbar(ng, b2, ti, D2, G)
    st = cc_utils._Stats(2, flat.device)
    st.damp(0, T1a, n1, alpha)
    st.damp(1, T2a, n2, alpha)
    Eref = ft_cc_energy.ft_cc_energy(T1a, T2a, F.ov, I.oovv, g, beta)
    sref = st.read()
    stats = torch.zeros(8, dtype=torch.float64, device=flat.device)
    fai = _dev(F.ov).t().contiguous()
    Iabij = ft_cc_energy.oovv_to_abij(I.oovv)
    quadrature.int_tbar_update(ng, b1, ti, D1, G, T1b, alpha, stats.data_ptr(), g=g, W=fai, c2=1.0)
    quadrature.int_tbar_update(ng, b2, ti, D2, G, T2b, alpha, stats.data_ptr() + 32, g=g, W=Iabij,
                               T1x=T1b, T1y=T1b, c2=0.25, c11=0.5)
    s = stats.cpu().numpy().reshape(2, 4)
    # (same arithmetic; the compiler contracts multiply-adds differently in the two kernels)
    assert float((T1a - T1b).abs().max()) < 1e-14*float(T1a.abs().max())
    assert float((T2a - T2b).abs().max()) < 1e-14*float(T2a.abs().max())
    assert numpy.abs(s[:, :3] - sref).max() < 1e-12*numpy.abs(sref).max()
    assert abs((s[0, 3] + s[1, 3])/beta - Eref) < 1e-12*abs(Eref)
    # strided damping
    T2c = _dev(t2.copy())
    st2 = cc_utils._Stats(1, flat.device)
    st2.damp(0, T2c, b2, alpha)
    ref = alpha*t2 + (1 - alpha)*tb2
    assert numpy.abs(T2c.cpu().numpy() - ref).max() < 1e-15
    assert abs(st2.read()[0, 0] - numpy.sum((tb2 - t2)**2)) < 1e-12*numpy.sum((tb2 - t2)**2)

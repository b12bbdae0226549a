"""CPU: the contraction-plan compiler (statement parsing, spin expansion, merging,
reverse mode, lowering to kb200_op + offset tables) executed with NumPy gathers
(tests/plan_exec.py) against the oracle."""
import numpy
import pytest

from kelvin_b200 import plan, programs
from kelvin_oracle import cqc, cc_equations as ocq
from plan_exec import run_lowered
import util


def _run(rops, mode, sizes, inputs, src, ng):
    shapes = plan.slot_shapes(rops, mode, sizes)
    batched = {s: not plan.is_integral_slot(s) for s in shapes}
    preset = [s for s in shapes if plan.is_integral_slot(s)] + list(inputs)
    # scratch blocks are stored with padded strides, as the engine allocates them
    outs = ("o1", "o2", "lo1", "lo2")
    pad = [s for s in shapes if s not in preset and len(shapes[s]) == 4
           and not s.startswith(outs) and "~" not in s]
    low = plan.Lowered(rops, shapes, batched, preset, pad=pad)
    arrays = {}
    for s, shp in shapes.items():
        if plan.is_integral_slot(s):
            pre, pat = s.split(".")
            arrays[s] = numpy.ascontiguousarray(getattr(src[pre], pat))
        elif s in inputs:
            arrays[s] = numpy.ascontiguousarray(inputs[s])
        elif s.startswith(plan.TRI_PREFIX):
            arrays[s] = numpy.zeros((ng,) + shp)       # the engine allocates these zeroed
        else:
            arrays[s] = numpy.full((ng,) + shp, numpy.nan)
    run_lowered(low, arrays, ng)
    return arrays, low


def test_stanton_plan_g():
    n, ng = 5, 3
    F, I, t1, t2 = util.random_g(n, ng, seed=0)
    rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "g")
    arr, low = _run(rops, "g", {"o": n, "v": n}, {"t1": t1, "t2": t2}, {"F": F, "I": I}, ng)
    for y in range(ng):
        R1, R2 = ocq.stanton_terms(F, I, t1[y], t2[y])
        assert numpy.abs(arr["o1"][y] - (-F.vo - R1)).max() < 1e-12
        assert numpy.abs(arr["o2"][y] - (-I.vvoo - R2)).max() < 1e-12
    # executed flops: 12 n^6 + lower order
    assert 12*n**6 <= low.flops <= 12*n**6 + 50*n**5


def test_stanton_plan_u_rectangular():
    na, nb, ng = 4, 3, 2
    ints, amps = util.random_u(na, nb, ng, seed=1)
    Fa, Fb, Ia, Ib, Iabab = ints
    sizes = {("v", "a"): na, ("o", "a"): na, ("v", "b"): nb, ("o", "b"): nb}
    names = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u")
    arr, low = _run(rops, "u", sizes, dict(zip(names, amps)),
                    {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}, ng)
    for y in range(ng):
        r = ocq.u_stanton_terms(*ints, (amps[0][y], amps[1][y]), (amps[2][y], amps[3][y], amps[4][y]))
        ref = (-Fa.vo - r[0], -Fb.vo - r[1], -Ia.vvoo - r[2], -Iabab.vvoo - r[3], -Ib.vvoo - r[4])
        for nm, rr in zip(("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb"), ref):
            assert numpy.abs(arr[nm][y] - rr).max() < 1e-12


def test_plans_g_active_space_shapes():
    """no != nv (athresh > 0, kelvin/ccsd.py:642-660): residual and Lambda plans on
    rectangular blocks against the oracle."""
    no, nv, ng = 3, 5, 2
    F, I, t1, t2, l1, l2 = util.random_g_rect(no, nv, ng, seed=11)
    sizes = {"o": no, "v": nv}
    rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "g")
    arr, _ = _run(rops, "g", sizes, {"t1": t1, "t2": t2}, {"F": F, "I": I}, ng)
    inter, rest = programs.lambda_rops("g", -1.0)
    lar, _ = _run(inter + rest, "g", sizes, {"t1": t1, "t2": t2, "l1": l1, "l2": l2},
                  {"F": F, "I": I}, ng)
    for y in range(ng):
        R1, R2 = ocq.stanton_terms(F, I, t1[y], t2[y])
        assert numpy.abs(arr["o1"][y] - (-F.vo - R1)).max() < 1e-12
        assert numpy.abs(arr["o2"][y] - (-I.vvoo - R2)).max() < 1e-12
        d1, d2 = ocq.lambda_terms(F, I, l1[y], l2[y], t1[y], t2[y])
        r1 = -d1 - F.ov - numpy.einsum('jiba,bj->ia', I.oovv, t1[y])
        assert numpy.abs(lar["lo1"][y] - r1).max() < 1e-12
        assert numpy.abs(lar["lo2"][y] - (-d2 - I.oovv)).max() < 1e-12


def test_u_plan_has_32_block_gemms():
    """32 Sz-allowed m^6 block GEMMs per residual (64 m^6 flops)."""
    m = 6
    sizes = {("v", "a"): m, ("o", "a"): m, ("v", "b"): m, ("o", "b"): m}
    rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u")
    shapes = plan.slot_shapes(rops, "u", sizes)
    low = plan.Lowered(rops, shapes, {s: not plan.is_integral_slot(s) for s in shapes},
                       [s for s in shapes if plan.is_integral_slot(s)] + ["t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb"])
    big = [d for d in low.descs if d.kind == 0 and d.M*d.N == m**4 and d.K >= m*(m - 1)//2]
    assert len(big) == 32
    # the 8 same-spin ladder contractions sum their antisymmetric pair over x < y only
    assert sum(1 for d in big if d.K == m*(m - 1)//2) == 8
    assert low.flops <= 64*m**6 + 200*m**5


@pytest.mark.parametrize("mode", ["g", "u"])
def test_lambda_plan(mode):
    ng = 2
    if mode == "g":
        n = 4
        F, I, t1, t2 = util.random_g(n, ng, seed=3)
        rng = numpy.random.default_rng(9)
        l1 = rng.standard_normal((ng, n, n))
        l2 = numpy.ascontiguousarray(util.asym(rng.standard_normal((ng, n, n, n, n))))
        inter, rest = programs.lambda_rops("g", -1.0)
        arr, _ = _run(inter + rest, "g", {"o": n, "v": n},
                      {"t1": t1, "t2": t2, "l1": l1, "l2": l2}, {"F": F, "I": I}, ng)
        for y in range(ng):
            d1, d2 = ocq.lambda_terms(F, I, l1[y], l2[y], t1[y], t2[y])
            r1 = -d1 - F.ov - numpy.einsum('jiba,bj->ia', I.oovv, t1[y])
            assert numpy.abs(arr["lo1"][y] - r1).max() < 1e-12
            assert numpy.abs(arr["lo2"][y] - (-d2 - I.oovv)).max() < 1e-12
        return
    na, nb = 4, 3
    ints, amps = util.random_u(na, nb, ng, seed=4)
    Fa, Fb, Ia, Ib, Iabab = ints
    _, lam = util.random_u(na, nb, ng, seed=5)
    lam = [numpy.ascontiguousarray(lam[0].transpose(0, 2, 1)),
           numpy.ascontiguousarray(lam[1].transpose(0, 2, 1)), lam[2], lam[3], lam[4]]
    sizes = {("v", "a"): na, ("o", "a"): na, ("v", "b"): nb, ("o", "b"): nb}
    inputs = dict(zip(("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb"), amps))
    inputs.update(zip(("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb"), lam))
    inter, rest = programs.lambda_rops("u", -1.0)
    arr, _ = _run(inter + rest, "u", sizes, inputs,
                  {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}, ng)
    for y in range(ng):
        out = [numpy.zeros_like(x[y]) for x in lam]
        ocq._uccsd_Lambda_opt(*out, *ints, (lam[0][y], lam[1][y]), (lam[2][y], lam[3][y], lam[4][y]),
                              (amps[0][y], amps[1][y]), (amps[2][y], amps[3][y], amps[4][y]), fac=-1.0)
        out[0] -= Fa.ov
        out[1] -= Fb.ov
        out[2] -= Ia.oovv
        out[3] -= Iabab.oovv
        out[4] -= Ib.oovv
        ocq._u_LS_TS(out[0], out[1], Ia, Ib, Iabab, amps[0][y], amps[1][y], fac=-1.0)
        for nm, r in zip(("lo1.a", "lo1.b", "lo2.aa", "lo2.ab", "lo2.bb"), out):
            assert numpy.abs(arr[nm][y] - r).max() < 1e-12


def test_rdm_plan_u_all_spin_blocks():
    """All 31 non-trivial unrestricted 2-RDM spin blocks + 1-RDM blocks equal the spin
    blocks of the derivative-defined g RDMs (kelvin/tests/test_ft_ccsd_rdm.py:495-752)."""
    na, nb, ng = 4, 3, 2
    gw = numpy.array([0.3, 0.7])
    ints, amps = util.random_u(na, nb, ng, seed=4)
    _, lam = util.random_u(na, nb, ng, seed=5)
    lam = [numpy.ascontiguousarray(lam[0].transpose(0, 2, 1)),
           numpy.ascontiguousarray(lam[1].transpose(0, 2, 1)), lam[2], lam[3], lam[4]]
    sizes = {("v", "a"): na, ("o", "a"): na, ("v", "b"): nb, ("o", "b"): nb}
    inputs = dict(zip(("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb"), amps))
    inputs.update(zip(("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb"), lam))
    inter, rest = programs.rdm_rops("u")
    arr, _ = _run(inter + rest, "u", sizes, inputs, {}, ng)
    summed = {s: numpy.ascontiguousarray(numpy.tensordot(gw, a, axes=(0, 0)))
              for s, a in arr.items() if s.endswith("~") and plan._INT_SLOT.match(s)}
    aops, outs = programs.rdm2_assembly_rops("u")
    ashapes = plan.slot_shapes(aops, "u", sizes)
    alow = plan.Lowered(aops, ashapes, {s: False for s in ashapes},
                        [s for s in ashapes if s.endswith("~")])
    aarr = {s: (summed[s] if s in summed else numpy.full(ashapes[s], numpy.nan)) for s in ashapes}
    run_lowered(alow, aarr, 1)
    args = [[a[y] for a in amps] + [l[y] for l in lam] for y in range(ng)]
    for slot, pname, sp in outs:
        k = programs.RDM2_USPINS[pname].index(sp)
        ref = sum(gw[y]*getattr(ocq, "uccsd_2rdm_" + pname)(*args[y])[k] for y in range(ng))
        assert numpy.abs(aarr[slot] - ref).max() < 1e-12, slot
    for nm, leaf in (("ba", "vv"), ("ji", "oo"), ("ai", "ov")):
        ref = [sum(gw[y]*getattr(ocq, "uccsd_1rdm_" + nm)(*args[y])[k] for y in range(ng)) for k in (0, 1)]
        assert numpy.abs(summed["Fa.%s~" % leaf].T - ref[0]).max() < 1e-12
        assert numpy.abs(summed["Fb.%s~" % leaf].T - ref[1]).max() < 1e-12


def test_small_shape_paths_are_exercised(monkeypatch):
    """Lower the size thresholds so that the derived-layout (long-K), skinny-streaming,
    rank-K and batch-folding code paths of the lowering are exercised at test sizes; results
    must not change."""
    monkeypatch.setattr(plan, "DERIVE", 1)
    monkeypatch.setattr(plan, "LONGK_MIN", 8)
    monkeypatch.setattr(plan, "RANKK_MIN_M", 8)
    _small_paths(monkeypatch, skinny=1)
    _small_paths(monkeypatch, skinny=0)


def _small_paths(monkeypatch, skinny):
    monkeypatch.setattr(plan, "SKINNY_TILE", skinny)
    na, nb, ng = 4, 3, 2
    ints, amps = util.random_u(na, nb, ng, seed=11)
    Fa, Fb, Ia, Ib, Iabab = ints
    sizes = {("v", "a"): na, ("o", "a"): na, ("v", "b"): nb, ("o", "b"): nb}
    names = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u")
    arr, low = _run(rops, "u", sizes, dict(zip(names, amps)),
                    {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}, ng)
    assert len(low.derived) > 0
    if skinny:
        assert any(d.tile == 7 for d in low.descs)
    else:
        assert any(d.kind == 2 for d in low.descs)
    assert any(o.batch == 1 and o.N == ng*d.N for o, d in zip(low.finalize(ng), low.descs)
               if d.kind == 0 and d.bsB != 0 and d.bsA == 0)       # tau batch folded into N
    for y in range(ng):
        r = ocq.u_stanton_terms(*ints, (amps[0][y], amps[1][y]), (amps[2][y], amps[3][y], amps[4][y]))
        ref = (-Fa.vo - r[0], -Fb.vo - r[1], -Ia.vvoo - r[2], -Iabab.vvoo - r[3], -Ib.vvoo - r[4])
        for nm, rr in zip(("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb"), ref):
            assert numpy.abs(arr[nm][y] - rr).max() < 1e-12


@pytest.mark.parametrize("method", ["CCD", "LCCSD", "LCCD"])
def test_method_variants_g(method):
    """CCD / LCCSD / LCCD residual and Lambda programs (kelvin/ft_cc_equations.py:11-62,
    292-340, 682-701) against the oracle's term classes, rectangular no != nv."""
    no, nv, ng, beta = 3, 4, 2, 1.7
    F, I, t1, t2, l1, l2 = util.random_g_rect(no, nv, ng, seed=17)
    sizes = {"o": no, "v": nv}
    s1 = programs.has_singles(method)
    ins = {"t1": t1, "t2": t2} if s1 else {"t2": t2}
    rops = plan.expand(programs.residual_program(method, -1.0), programs.tensor_defs(), "g")
    assert s1 or all(slot != "t1" for op in rops for slot, _ in op.ins)
    arr, _ = _run(rops, "g", sizes, ins, {"F": F, "I": I}, ng)
    lins = dict(ins)
    lins["l2"] = l2
    if s1:
        lins["l1"] = l1
    inter, rest = programs.lambda_rops("g", -1.0, method=method, beta=beta)
    lar, _ = _run(inter + rest, "g", sizes, lins, {"F": F, "I": I}, ng)
    for y in range(ng):
        r1, r2 = -F.vo.copy(), -I.vvoo.copy()
        o1, o2 = numpy.zeros((no, nv)), numpy.zeros((no, no, nv, nv))
        if method == "CCD":
            ocq._D_D(r2, F, I, t2[y], fac=-1.0)
            ocq._D_DD(r2, F, I, t2[y], fac=-1.0)
            ocq._LD_LD(o2, F, I, l2[y], fac=-1.0)
            ocq._LD_LDTD(o2, I, l2[y], t2[y], fac=-1.0)
            o2 -= I.oovv
        elif method == "LCCD":
            ocq._D_D(r2, F, I, t2[y], fac=-1.0)
            ocq._LD_LD(o2, F, I, l2[y], fac=-1.0)
            o2 -= I.oovv/beta
        else:
            ocq._S_S(r1, F, I, t1[y], fac=-1.0)
            ocq._S_D(r1, F, I, t2[y], fac=-1.0)
            ocq._D_S(r2, F, I, t1[y], fac=-1.0)
            ocq._D_D(r2, F, I, t2[y], fac=-1.0)
            ocq._LS_LS(o1, F, I, l1[y], fac=-1.0)
            ocq._LS_LD(o1, F, I, l2[y], fac=-1.0)
            ocq._LD_LS(o2, F, I, l1[y], fac=-1.0)
            ocq._LD_LD(o2, F, I, l2[y], fac=-1.0)
            o1 -= F.ov
            o2 -= I.oovv
            ocq._LS_TS(o1, I, t1[y], fac=-1.0)
        sc = numpy.abs(r2).max()
        assert numpy.abs(arr["o2"][y] - r2).max() < 1e-11*sc
        assert numpy.abs(lar["lo2"][y] - o2).max() < 1e-11*numpy.abs(o2).max()
        if s1:
            assert numpy.abs(arr["o1"][y] - r1).max() < 1e-11*sc
            assert numpy.abs(lar["lo1"][y] - o1).max() < 1e-11*numpy.abs(o1).max()
        else:
            assert "o1" not in arr and "lo1" not in lar


def test_ccd_is_ccsd_with_zero_singles():
    n, ng = 4, 2
    F, I, t1, t2 = util.random_g(n, ng, seed=23)
    rops = plan.expand(programs.residual_program("CCD", -1.0), programs.tensor_defs(), "g")
    arr, _ = _run(rops, "g", {"o": n, "v": n}, {"t2": t2}, {"F": F, "I": I}, ng)
    for y in range(ng):
        R1, R2 = ocq.stanton_terms(F, I, numpy.zeros((n, n)), t2[y])
        assert numpy.abs(arr["o2"][y] - (-I.vvoo - R2)).max() < 1e-12*numpy.abs(R2).max()


def _closed_inputs(n, ng, seed):
    ints, amps, lam = util.random_u_closed(n, ng, seed=seed)
    lam = [numpy.ascontiguousarray(lam[0].transpose(0, 2, 1)),
           numpy.ascontiguousarray(lam[1].transpose(0, 2, 1)),
           numpy.ascontiguousarray(lam[2].transpose(0, 3, 4, 1, 2)),
           numpy.ascontiguousarray(lam[3].transpose(0, 3, 4, 1, 2)),
           numpy.ascontiguousarray(lam[4].transpose(0, 3, 4, 1, 2))]
    return ints, amps, lam


def test_mirror_reduced_plans_on_closed_shell_inputs():
    """plan.mirror_reduce: on alpha == beta inputs the reduced unrestricted residual and
    Lambda programs (beta-leading blocks aliased to their alpha images) reproduce the
    alpha and mixed blocks of the full programs."""
    n, ng = 4, 2
    ints, amps, lam = _closed_inputs(n, ng, 41)
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    tn = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    ln = ("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb")
    full = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u")
    red = plan.mirror_reduce(full)
    assert all(s == plan.mirror_rep(s) for op in red for s in [op.out[0]] + [x for x, _ in op.ins])
    a_full, lf = _run(full, "u", sizes, dict(zip(tn, amps)), src, ng)
    keep = {k: v for k, v in zip(tn, amps) if plan.mirror_rep(k) == k}
    a_red, lr = _run(red, "u", sizes, keep, src, ng)
    assert lr.flops < 0.62*lf.flops
    for nm in ("o1.a", "o2.aa", "o2.ab"):
        assert numpy.abs(a_red[nm] - a_full[nm]).max() < 1e-12*numpy.abs(a_full[nm]).max()
    # what the caller copies is what the full program computes for the beta blocks
    assert numpy.abs(a_full["o2.bb"] - a_full["o2.aa"]).max() < 1e-12*numpy.abs(a_full["o2.aa"]).max()
    assert numpy.abs(a_full["o1.b"] - a_full["o1.a"]).max() < 1e-12*numpy.abs(a_full["o1.a"]).max()
    ab = a_full["o2.ab"]
    assert numpy.abs(ab - ab.transpose(0, 2, 1, 4, 3)).max() < 1e-12*numpy.abs(ab).max()
    inter, rest = programs.lambda_rops("u", -1.0)
    inputs = dict(zip(tn, amps))
    inputs.update(zip(ln, lam))
    l_full, _ = _run(inter + rest, "u", sizes, inputs, src, ng)
    keep = {k: v for k, v in inputs.items() if plan.mirror_rep(k) == k}
    l_red, _ = _run(plan.mirror_reduce(inter) + plan.mirror_reduce(rest), "u", sizes, keep, src, ng)
    for nm in ("lo1.a", "lo2.aa", "lo2.ab"):
        assert numpy.abs(l_red[nm] - l_full[nm]).max() < 1e-12*numpy.abs(l_full[nm]).max()
    assert numpy.abs(l_full["lo2.bb"] - l_full["lo2.aa"]).max() < 1e-12*numpy.abs(l_full["lo2.aa"]).max()


@pytest.mark.parametrize("mode", ["g", "u", "closed"])
def test_antisymmetric_outputs_triangle(mode):
    """plan.antisym_outputs: the same-spin ladder terms are computed on the a<b, i<j triangle
    only and expanded into their four images; residual unchanged."""
    ng = 2
    if mode == "g":
        n = 5
        F, I, t1, t2 = util.random_g(n, ng, seed=51)
        sizes, src, ins = {"o": n, "v": n}, {"F": F, "I": I}, {"t1": t1, "t2": t2}
        rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "g")
        outs = ("o1", "o2")
    else:
        n = 4
        if mode == "u":
            ints, amps = util.random_u(n, n - 1, ng, seed=52)
        else:
            ints, amps, _ = util.random_u_closed(n, ng, seed=53)
        Fa, Fb, Ia, Ib, Iabab = ints
        nb = ints[1].ov.shape[0]
        sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): nb, ("o", "b"): nb}
        src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
        ins = dict(zip(("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb"), amps))
        rops = plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u")
        outs = ("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb")
        if mode == "closed":
            rops = plan.mirror_reduce(rops)
            ins = {k: v for k, v in ins.items() if plan.mirror_rep(k) == k}
            outs = tuple(s for s in outs if plan.mirror_rep(s) == s)
    tri = plan.antisym_outputs(rops)
    ntri = sum(1 for op in tri if op.tri is not None)
    assert ntri == {"g": 4, "u": 8, "closed": 4}[mode]
    if mode == "closed":
        # opposite-spin ladder terms on the rows p <= q only (p < q mirrored, p == q direct)
        tri = plan.antisym_outputs(plan.mirror_outputs(rops))
        assert sum(1 for op in tri if op.tri is not None) == 4 + 2*4
    ref, lref = _run(rops, mode if mode != "closed" else "u", sizes, ins, src, ng)
    got, lgot = _run(tri, mode if mode != "closed" else "u", sizes, ins, src, ng)
    assert lgot.flops < lref.flops
    for nm in outs:
        assert numpy.abs(got[nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), nm


def test_singlet_reduced_program():
    """plan.singlet_reduce: for closed-shell inputs that satisfy T2aa = T2ab - T2ab(a<->b) the
    same-spin doubles residual is the antisymmetrised opposite-spin one, so the program that
    evaluates only o1.a and o2.ab reproduces all three residual blocks."""
    n, ng = 4, 2
    ints, amps, _ = util.random_u_closed(n, ng, seed=61)
    assert numpy.abs(amps[2] - (amps[3] - amps[3].transpose(0, 2, 1, 3, 4))).max() == 0.0
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    ins = {"t1.a": amps[0], "t2.aa": amps[2], "t2.ab": amps[3]}
    red = plan.mirror_reduce(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
    sing = plan.singlet_reduce(red)
    assert not any(op.out[0].startswith(("rg.aaaa", "Woooo.aa", "Wvvvv.aa")) for op in sing)
    ref, lref = _run(plan.antisym_outputs(red), "u", sizes, ins, src, ng)
    got, lgot = _run(plan.antisym_outputs(plan.mirror_outputs(sing)), "u", sizes,
                     {k: v for k, v in ins.items()}, src, ng)
    assert lgot.flops < 0.85*lref.flops
    for nm in ("o1.a", "o2.aa", "o2.ab"):
        assert numpy.abs(got[nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), nm


@pytest.mark.parametrize("singlet", [False, True])
def test_sumdiff_pairs(singlet):
    """plan.sumdiff_pairs: the W_ovvo.aaaa/.abab builds and the rg.aaaa/.abab ring contractions of
    the closed-shell program as (A1+-A2)(B1+-B2): two contractions instead of four each, same
    residual; composes with the singlet reduction and the triangle/mirror-row rewrites."""
    n, ng = 4, 2
    ints, amps, _ = util.random_u_closed(n, ng, seed=71)
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    ins = {"t1.a": amps[0], "t2.aa": amps[2], "t2.ab": amps[3]}
    red = plan.mirror_reduce(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
    base = plan.singlet_reduce(red) if singlet else red
    sd = plan.sumdiff_pairs(base)
    nq = sum(1 for op in sd if op.out[0].startswith(plan.SUMDIFF_PREFIX) and len(op.ins) == 2)
    assert nq == (2 if singlet else 4)       # with the singlet reduction rg.aaaa is gone
    ref, _ = _run(red, "u", sizes, ins, src, ng)
    got, low = _run(plan.antisym_outputs(plan.mirror_outputs(sd)), "u", sizes, ins, src, ng)
    for nm in ("o1.a", "o2.aa", "o2.ab"):
        assert numpy.abs(got[nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), nm
    # m^6 count at the benchmark size
    m = 33
    big = {("v", "a"): m, ("o", "a"): m, ("v", "b"): m, ("o", "b"): m}
    def units(ops):
        shapes = plan.slot_shapes(ops, "u", big)
        pres = [s for s in shapes if plan.is_integral_slot(s)] + [s for s in ins if s in shapes]
        lw = plan.Lowered(ops, shapes, {s: not plan.is_integral_slot(s) for s in shapes}, pres)
        return sum(2.0*d.M*d.N*d.K for d in lw.descs if d.kind == 0 and d.K >= 500)/(2.0*m**6)
    u0 = units(plan.antisym_outputs(plan.mirror_outputs(red)))
    u1 = units(plan.antisym_outputs(plan.mirror_outputs(sd)))
    assert u1 < u0 - (1.9 if singlet else 3.9)


def test_split_accumulators():
    """plan.split_accumulators: later large contractions into one slot go through scratch slots
    of their own; same residual, and at the benchmark size the m^6 contractions of the
    closed-shell singlet program need fewer launch groups."""
    n, ng = 4, 2
    ints, amps, _ = util.random_u_closed(n, ng, seed=72)
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    ins = {"t1.a": amps[0], "t2.aa": amps[2], "t2.ab": amps[3]}
    red = plan.mirror_reduce(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
    base = plan.antisym_outputs(plan.sumdiff_pairs(plan.singlet_reduce(red)))
    split = plan.split_accumulators(base)
    assert sum(1 for op in split if op.out[0].startswith(plan.SPLIT_PREFIX)) >= 2
    ref, _ = _run(red, "u", sizes, ins, src, ng)
    got, _ = _run(split, "u", sizes, ins, src, ng)
    for nm in ("o1.a", "o2.aa", "o2.ab"):
        assert numpy.abs(got[nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), nm
    m = 33
    big = {("v", "a"): m, ("o", "a"): m, ("v", "b"): m, ("o", "b"): m}

    def launches(ops):
        shapes = plan.slot_shapes(ops, "u", big)
        pres = [s for s in shapes if plan.is_integral_slot(s)] + [s for s in ins if s in shapes]
        lw = plan.Lowered(ops, shapes, {s: not plan.is_integral_slot(s) for s in shapes}, pres)
        return [len(g) for g in lw.groups if lw.descs[g[0]].kind == 0 and lw.descs[g[0]].K >= 500
                and lw.descs[g[0]].tile in (0, 2)]
    emit = plan.singlet_reduce(red, emit_aa=False)
    g0 = launches(plan.antisym_outputs(plan.sumdiff_pairs(emit)))
    g1 = launches(plan.split_accumulators(plan.antisym_outputs(plan.sumdiff_pairs(emit))))
    assert sum(g0) == sum(g1) == 10 and len(g1) < len(g0), (g0, g1)


def test_schedule_commuting_accumulations_keep_reader_order():
    """Accumulations into one slot may be reordered by the scheduler, but every op that READS a
    slot still comes after all the writes that preceded it in the program, and every overwrite
    after all earlier readers: checked on the scheduled order of the closed-shell program by
    replaying both orders symbolically (set of contributions seen by each reader)."""
    n = 24      # large enough for the m^6 terms to count as wide launches (K = n^2 >= 512)
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    red = plan.mirror_reduce(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
    ops = plan.split_accumulators(plan.antisym_outputs(plan.sumdiff_pairs(plan.singlet_reduce(red))))
    shapes = plan.slot_shapes(ops, "u", sizes)
    pres = [s for s in shapes if plan.is_integral_slot(s)] + ["t1.a", "t2.aa", "t2.ab"]
    batched = {s: not plan.is_integral_slot(s) for s in shapes}

    def replay(low):
        """slot -> frozenset of (op signature) contributions at the time of each read, per reader."""
        state, seen = {}, {}
        for d, op in zip(low.descs, low.rops):
            sig = repr(op)
            nm = low.slot_names
            reads = [nm[d.a]] + ([nm[d.b]] if d.kind != 1 and d.b >= 0 else [])
            for s_ in reads:
                seen.setdefault(sig, []).append((s_, state.get(s_, frozenset())))
            if d.beta == 0.0:
                state[nm[d.c]] = frozenset([sig])
            else:
                state[nm[d.c]] = state.get(nm[d.c], frozenset()) | frozenset([sig])
        return seen, state
    old = plan.COMMUTE_ACC
    try:
        plan.COMMUTE_ACC = False
        ref = plan.Lowered(ops, shapes, batched, pres)
        plan.COMMUTE_ACC = True
        got = plan.Lowered(ops, shapes, batched, pres)
    finally:
        plan.COMMUTE_ACC = old
    assert [repr(o) for o in ref.rops] != [repr(o) for o in got.rops]      # the order did change
    seen_ref, state_ref = replay(ref)
    seen_got, state_got = replay(got)
    assert state_ref == state_got
    assert {k: sorted(v, key=repr) for k, v in seen_ref.items()} == \
        {k: sorted(v, key=repr) for k, v in seen_got.items()}


def test_lambda_sweep_closed_shell_rewrites():
    """plan.merge_duplicates + plan.sumdiff_pairs + plan.antisym_outputs on the closed-shell
    Lambda program: mirror-duplicate contractions done once, three quartets as sum/difference
    pairs, same-spin ladder adjoints on their triangle; same Lambda map, fewer m^6 units."""
    n, ng = 4, 2
    ints, amps, lam = _closed_inputs(n, ng, 83)
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    tn = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    ln = ("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb")
    inputs = dict(zip(tn, amps))
    inputs.update(zip(ln, lam))
    keep = {k: v for k, v in inputs.items() if plan.mirror_rep(k) == k}
    inter, rest = programs.lambda_rops("u", -1.0)
    inter, rest = plan.mirror_reduce(inter), plan.mirror_reduce(rest)
    ref, _ = _run(inter + rest, "u", sizes, keep, src, ng)
    md = plan.merge_duplicates(rest)
    assert sum(1 for op in md if op.out[0].startswith(plan.DUP_PREFIX)) == 3
    sd = plan.sumdiff_pairs(md)
    assert sum(1 for op in sd if op.out[0].startswith(plan.SUMDIFF_PREFIX) and len(op.ins) == 2) == 6
    new = plan.antisym_outputs(sd)
    got, _ = _run(inter + new, "u", sizes, keep, src, ng)
    for nm in ("lo1.a", "lo2.aa", "lo2.ab"):
        assert numpy.abs(got[nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), nm
    m = 33
    big = {("v", "a"): m, ("o", "a"): m, ("v", "b"): m, ("o", "b"): m}

    def units(ops):
        shapes = plan.slot_shapes(ops, "u", big)
        written, pres = set(), []
        for op in ops:
            for s, _ in op.ins:
                if s not in written and s not in pres:
                    pres.append(s)
            written.add(op.out[0])
        lw = plan.Lowered(ops, shapes, {s: not plan.is_integral_slot(s) for s in shapes}, pres)
        return sum(2.0*d.M*d.N*d.K for d in lw.descs if d.kind == 0 and d.K >= 500)/(2.0*m**6)
    assert units(new) < units(rest) - 10.0


@pytest.mark.parametrize("prog", ["stanton", "stanton-closed", "lambda"])
@pytest.mark.parametrize("world", [2, 3])
def test_hybrid_phases(prog, world):
    """plan.hybrid_phases: the ranks that share a grid point each contract a slab of the rows of
    every large contraction; distributed parts are exchanged before they are contracted again
    and at the end.  Simulated ranks (NumPy executor) reproduce the single-rank program."""
    from plan_exec import run_hybrid
    n, ng = 5, 2
    sizes = {("v", "a"): n, ("o", "a"): n, ("v", "b"): n, ("o", "b"): n}
    tn = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    ln = ("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb")
    if prog == "stanton":
        ints, amps = util.random_u(n, n, ng, seed=91)
        inputs = dict(zip(tn, amps))
        rops = plan.antisym_outputs(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
        outs = ("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb")
    elif prog == "stanton-closed":
        ints, amps, _ = util.random_u_closed(n, ng, seed=92)
        inputs = {"t1.a": amps[0], "t2.aa": amps[2], "t2.ab": amps[3]}
        red = plan.mirror_reduce(plan.expand(programs.stanton(-1.0), programs.tensor_defs(), "u"))
        rops = plan.antisym_outputs(plan.sumdiff_pairs(plan.singlet_reduce(red)))
        outs = ("o1.a", "o2.aa", "o2.ab")
    else:
        ints, amps, lam = _closed_inputs(n, ng, 93)
        inputs = dict(zip(tn, amps))
        inputs.update(zip(ln, lam))
        inputs = {k: v for k, v in inputs.items() if plan.mirror_rep(k) == k}
        inter, rest = programs.lambda_rops("u", -1.0)
        rops = plan.mirror_reduce(inter) + plan.antisym_outputs(plan.sumdiff_pairs(
            plan.merge_duplicates(plan.mirror_reduce(rest))))
        outs = ("lo1.a", "lo2.aa", "lo2.ab")
    Fa, Fb, Ia, Ib, Iabab = ints
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    ref, _ = _run(rops, "u", sizes, inputs, src, ng)
    shapes = plan.slot_shapes(rops, "u", sizes)
    hp = plan.hybrid_phases(rops, shapes, outs, world, min_work=n**5, min_cols=8)
    nslab = sum(1 for ph in hp.phases for op in ph if op.slab)
    assert nslab >= 8 and len(hp.phases) >= 3 and any(hp.exchange)
    batched = lambda s_: not plan.is_integral_slot(s_)      # noqa: E731
    ranks = []
    for r in range(world):
        arr = {}
        for s_, shp in hp.shapes.items():
            if plan.is_integral_slot(s_):
                pre, pat = s_.split(".")
                arr[s_] = numpy.ascontiguousarray(getattr(src[pre], pat))
            elif s_ in inputs:
                arr[s_] = numpy.ascontiguousarray(inputs[s_])
            elif s_ in hp.dslots:
                arr[s_] = numpy.zeros((ng,) + tuple(shp))
        ranks.append(arr)
    run_hybrid(hp, ranks, ng, batched)
    for r in range(world):
        for nm in outs:
            assert numpy.abs(ranks[r][nm] - ref[nm]).max() < 1e-12*numpy.abs(ref[nm]).max(), (r, nm)

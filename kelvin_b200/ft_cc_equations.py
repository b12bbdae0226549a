"""Finite-temperature CC iteration maps on the GPU.

Drop-in for the on-path functions of kelvin/ft_cc_equations.py (same names and
positional signatures):

* ``ccsd_stanton`` (:96), ``uccsd_stanton`` (:130) -- one amplitude update:
  per-grid-point Stanton residual, then exp-weighted time integration;
* ``ccsd_lambda_opt`` (:385), ``uccsd_lambda_opt`` (:412),
  ``ccsd_lambda_guess`` (:502), ``uccsd_lambda_guess`` (:515);
* ``ccsd_1rdm`` (:704), ``uccsd_1rdm`` (:756), ``ccsd_2rdm`` (:725),
  ``uccsd_2rdm`` (:801).

Inputs may be NumPy arrays or torch tensors; outputs are fresh CUDA float64
tensors (inputs are never modified, as in the reference).  The per-grid-point
loop of the reference (``for y in range(ng): cqcpy._Stanton(...)``) is replaced
by one batched contraction plan over all grid points (kelvin_b200/plan.py).
"""
import os

import torch

from . import _lib, engine, plan as _plan, programs, quadrature

_F_NAMES = ("oo", "ov", "vo", "vv")


_chunk_cache = {}


def _chunk_for(p, ng, dev):
    """Largest tau chunk whose scratch fits in ~60% of the free HBM (decided once per
    plan and grid size; scratch is kept by the plan afterwards)."""
    key = (id(p), ng, dev.index)
    if key not in _chunk_cache:
        per = max(1, p.tmp_bytes_per_point())
        free, _ = torch.cuda.mem_get_info(dev)
        have = 0 if p._tmp is None else p._tmp_nb*per
        _chunk_cache[key] = int(max(1, min(ng, (0.6*(free + have))//per)))
    return _chunk_cache[key]


# ---------------------------------------------------------------------------
# plan builders
# ---------------------------------------------------------------------------
def _g_sizes(F):
    no, nv = F.ov.shape
    return {"o": int(no), "v": int(nv)}


def _u_sizes(Fa, Fb):
    noa, nva = Fa.ov.shape
    nob, nvb = Fb.ov.shape
    return {("o", "a"): int(noa), ("v", "a"): int(nva), ("o", "b"): int(nob), ("v", "b"): int(nvb)}


# tau batch from which the closed-shell program also halves the rows of its mirror-symmetric
# opposite-spin ladder terms (plan.mirror_outputs): the extra launches (diagonal pass + mirrored
# add per term) only pay once the contractions are several waves of CTAs long
MIRROR_ROWS_MIN_BATCH = int(os.environ.get("KB200_MIRROR_ROWS_MIN_BATCH", "4"))


# sum/difference form of the paired closed-shell ring contractions (plan.sumdiff_pairs): built and
# CPU-tested, off until it has had its GPU parity run
SUMDIFF = int(os.environ.get("KB200_SUMDIFF", "0"))


def stanton_plan(mode, sizes, fac=-1.0, mirror=False, mirror_rows=False, singlet=False,
                 sumdiff=None):
    """mirror (u only): the closed-shell reduction of the program (plan.mirror_reduce): only the
    alpha-leading block of every alpha <-> beta pair is evaluated; mirror_rows: additionally
    plan.mirror_outputs; singlet / sumdiff: additionally plan.singlet_reduce /
    plan.sumdiff_pairs (neither is used by the loops yet)."""
    mirror_rows = bool(mirror and mirror_rows)
    singlet = bool(mirror and singlet)
    sumdiff = bool(mirror and (SUMDIFF if sumdiff is None else sumdiff))
    key = ("stanton", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror), mirror_rows,
           singlet, sumdiff)

    def build():
        T = programs.tensor_defs()
        rops = _plan.expand(programs.stanton(fac), T, mode)
        if mode == "g":
            ins, outs = ("t1", "t2"), ("o1", "o2")
        else:
            ins = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
            outs = ("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb")
        if mirror:
            assert mode == "u"
            rops = _plan.mirror_reduce(rops)
            ins = tuple(s for s in ins if _plan.mirror_rep(s) == s)
            outs = tuple(s for s in outs if _plan.mirror_rep(s) == s)
            if singlet:
                rops = _plan.singlet_reduce(rops)
            if sumdiff:
                rops = _plan.sumdiff_pairs(rops)
            if mirror_rows:
                rops = _plan.mirror_outputs(rops)
        rops = _plan.antisym_outputs(rops)
        return engine.Plan(rops, mode, sizes, ins, outs,
                           name="stanton-" + mode + ("-closed" if mirror else ""))
    return engine.cached(key, build)


# ---------------------------------------------------------------------------
# closed-shell detection (unrestricted inputs whose alpha and beta halves coincide)
# ---------------------------------------------------------------------------
CLOSED_SHELL = int(os.environ.get("KB200_CLOSED_SHELL", "1"))
CLOSED_SHELL_TOL = 1e-12


def _same(x, y, tol=CLOSED_SHELL_TOL):
    if tuple(x.shape) != tuple(y.shape):
        return False
    if x.data_ptr() == y.data_ptr() and x.stride() == y.stride():
        return True
    d, m = _lib.max_absdiff(x, y)
    return d <= tol*m


def closed_shell_integrals(Fa, Fb, Ia, Ib, Iabab):
    """True when the dressed integrals are mirror symmetric: Fa == Fb, Ia == Ib and
    Iabab.wxyz[p,q,r,s] == Iabab.xwzy[q,p,s,r] (to CLOSED_SHELL_TOL relative).  Checked on
    the device (about forty reductions over the blocks): call once per solve."""
    if not CLOSED_SHELL:
        return False
    dev = _lib.device()
    ts = []
    for nm in _F_NAMES:
        ts += [_lib.as_dev(getattr(Fa, nm), dev), _lib.as_dev(getattr(Fb, nm), dev)]
    for nm in programs._INT2:
        ts += [_lib.as_dev(getattr(Ia, nm), dev), _lib.as_dev(getattr(Ib, nm), dev)]
    from .ov_blocks import two_e_blocks_full
    ab = {nm: _lib.as_dev(getattr(Iabab, nm), dev) for nm in two_e_blocks_full.names}
    ok = all(_same(ts[k], ts[k + 1]) for k in range(0, len(ts), 2))
    if ok:
        for nm, x in ab.items():
            img = nm[1] + nm[0] + nm[3] + nm[2]
            if img < nm:
                continue
            if not _same(x, ab[img].permute(1, 0, 3, 2)):
                ok = False
                break
    return ok


def closed_shell_amplitudes(X1a, X1b, X2aa, X2ab, X2bb):
    """True when X1a == X1b, X2aa == X2bb and X2ab[y,p,Q,r,S] == X2ab[y,Q,p,S,r]
    (amplitudes or Lambda, to CLOSED_SHELL_TOL relative).  Three device reductions and one
    synchronisation: call once per solve."""
    if not CLOSED_SHELL:
        return False
    dev = _lib.device()
    X1a, X1b, X2aa, X2ab, X2bb = [_lib.as_dev(x, dev) for x in (X1a, X1b, X2aa, X2ab, X2bb)]
    return _same(X1a, X1b) and _same(X2aa, X2bb) and _same(X2ab, X2ab.permute(0, 2, 1, 4, 3))


def _g_integral_slots(F, I, dev):
    t = {}
    for nm in _F_NAMES:
        t["F." + nm] = _lib.as_dev(getattr(F, nm), dev)
    for nm in programs._INT2:
        t["I." + nm] = _lib.as_dev(getattr(I, nm), dev)
    return t


def _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev, needed):
    t = {}
    src = {"Fa": Fa, "Fb": Fb, "Ia": Ia, "Ib": Ib, "Iabab": Iabab}
    for slot in needed:
        pre, pat = slot.split(".")
        t[slot] = _lib.as_dev(getattr(src[pre], pat), dev)
    return t


# ---------------------------------------------------------------------------
# amplitude equations
# ---------------------------------------------------------------------------
def t0_is_zero(G, amps):
    """True when the amplitudes at the first grid point (tau = 0) are exactly zero and
    stay zero under the quadrature: row 0 of G vanishes (every rule of
    kelvin/quadrature.py:215-235 integrates from tau_0 = 0), so T[0] = sum_x G[0,x] w T̄[x]
    = 0 and T̄[0] = drivers + fac*StantonTerms(T = 0) = drivers.  The pointwise solver of the
    reference uses the same fact (kelvin/cc_utils.py:205-208, "don't bother computing at
    T = inf").  One device reduction; call it once per solve, not per iteration."""
    import numpy
    G = G.cpu().numpy() if isinstance(G, torch.Tensor) else numpy.asarray(G)
    if G.shape[0] < 2 or numpy.any(G[0] != 0.0):
        return False
    for x in amps:
        if x.shape[0] < 2 or float(x[0].abs().max()) != 0.0:
            return False
    return True


def _run_rows(p, t, ins, outs, drivers, ng, dev, t0_zero):
    """Run the residual plan over the grid points; with t0_zero the first point is not
    evaluated (its amplitudes are exactly zero): T̄[0] = drivers."""
    if not t0_zero:
        p.run(t, ng, _chunk_for(p, ng, dev))
        return
    full = {nm: t[nm] for nm in tuple(ins) + tuple(outs)}
    if ng > 1:
        for nm in full:
            t[nm] = full[nm][1:]
        p.run(t, ng - 1, _chunk_for(p, ng - 1, dev))
    for nm, d in zip(outs, drivers):
        full[nm][0].copy_(d)
        full[nm][0].neg_()
        t[nm] = full[nm]


def ccsd_stanton_bar(F, I, T1old, T2old, fac=-1.0, t0_zero=False):
    """T1bar, T2bar: drivers + fac*StantonTerms at every grid point, i.e. the
    state of T1new/T2new just before the integration at
    kelvin/ft_cc_equations.py:109.  t0_zero: the caller guarantees T1old[0] = T2old[0] = 0
    (see t0_is_zero); the first grid point then costs nothing."""
    dev = _lib.device()
    T1old = _lib.as_dev(T1old, dev)
    T2old = _lib.as_dev(T2old, dev)
    ng = T1old.shape[0]
    p = stanton_plan("g", _g_sizes(F), fac)
    t = _g_integral_slots(F, I, dev)
    t["t1"], t["t2"] = T1old, T2old
    o1 = t["o1"] = torch.empty_like(T1old)
    o2 = t["o2"] = torch.empty_like(T2old)
    _run_rows(p, t, ("t1", "t2"), ("o1", "o2"), (t["F.vo"], t["I.vvoo"]), ng, dev, t0_zero)
    return o1, o2


def ccsd_stanton(F, I, T1old, T2old, D1, D2, ti, ng, G, t0_zero=False):
    """Time-dependent CCSD iteration using Stanton-Gauss intermediates
    (kelvin/ft_cc_equations.py:96-113)."""
    T1bar, T2bar = ccsd_stanton_bar(F, I, T1old, T2old, t0_zero=t0_zero)
    T1new = quadrature.int_tbar1(ng, T1bar, ti, D1, G)
    T2new = quadrature.int_tbar2(ng, T2bar, ti, D2, G)
    return T1new, T2new


def uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold, fac=-1.0,
                      t0_zero=False, closed_shell=False, beta_copies=True):
    """closed_shell: the caller guarantees mirror-symmetric integrals and amplitudes
    (closed_shell_integrals / closed_shell_amplitudes); the beta-leading blocks are then copies
    of their alpha images and only the reduced program runs (plan.mirror_reduce).
    beta_copies=False returns None in their place (the caller copies later)."""
    dev = _lib.device()
    ins = [_lib.as_dev(x, dev) for x in (T1aold, T1bold, T2aaold, T2abold, T2bbold)]
    ng = ins[0].shape[0]
    neval = ng - 1 if (t0_zero and ng > 1) else ng
    p = stanton_plan("u", _u_sizes(Fa, Fb), fac, mirror=closed_shell,
                     mirror_rows=neval >= MIRROR_ROWS_MIN_BATCH)
    t = _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                          [s for s in p.inputs if _plan.is_integral_slot(s)])
    in_names = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
    out_names = ("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb")
    drivers = [Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo]
    live = [k for k in range(5) if not closed_shell or k in (0, 2, 3)]
    outs = [None]*5
    for k in live:
        t[in_names[k]] = ins[k]
        outs[k] = t[out_names[k]] = torch.empty_like(ins[k])
    drv = [_lib.as_dev(drivers[k], dev) for k in live] if t0_zero else None
    _run_rows(p, t, [in_names[k] for k in live], [out_names[k] for k in live], drv, ng, dev,
              t0_zero)
    if closed_shell and beta_copies:
        outs[1] = outs[0].clone()
        outs[4] = outs[2].clone()
    return tuple(outs)


def uccsd_stanton(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold,
                  T2bbold, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G, t0_zero=False,
                  closed_shell=False):
    """Unrestricted CCSD iteration (kelvin/ft_cc_equations.py:130-164)."""
    b1a, b1b, b2aa, b2ab, b2bb = uccsd_stanton_bar(
        Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold, t0_zero=t0_zero,
        closed_shell=closed_shell)
    T1a = quadrature.int_tbar1(ng, b1a, ti, D1a, G)
    T1b = quadrature.int_tbar1(ng, b1b, ti, D1b, G)
    T2aa = quadrature.int_tbar2(ng, b2aa, ti, D2aa, G)
    T2ab = quadrature.int_tbar2(ng, b2ab, ti, D2ab, G)
    T2bb = quadrature.int_tbar2(ng, b2bb, ti, D2bb, G)
    return (T1a, T1b), (T2aa, T2ab, T2bb)


def _store_row(bar, ig, row):
    """bar[ig] = row for a device tensor or a NumPy array (the reference updates the caller's
    T̄ buffers in place, kelvin/ft_cc_equations.py:122-123)."""
    if isinstance(bar, torch.Tensor):
        bar[ig].copy_(row)
    else:
        bar[ig] = row.cpu().numpy()


def ccsd_stanton_single(ig, F, I, T1old, T2old, T1bar, T2bar, D1, D2, ti, ng, G):
    """Amplitude update at the single grid point ig (pointwise solver,
    kelvin/ft_cc_equations.py:116-127): T1old/T2old are the amplitudes AT that point,
    T1bar/T2bar the (ng, ...) residual buffers, whose row ig is overwritten."""
    dev = _lib.device()
    b1, b2 = ccsd_stanton_bar(F, I, _lib.as_dev(T1old, dev)[None], _lib.as_dev(T2old, dev)[None])
    _store_row(T1bar, ig, b1[0])
    _store_row(T2bar, ig, b2[0])
    T1new = quadrature.int_tbar1_single(ng, ig, T1bar, ti, D1, G)
    T2new = quadrature.int_tbar2_single(ng, ig, T2bar, ti, D2, G)
    return T1new, T2new


def uccsd_stanton_single(ig, Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab,
                         T2bb, T1bara, T1barb, T2baraa, T2barab, T2barbb,
                         D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G):
    """Unrestricted single-point update (kelvin/ft_cc_equations.py:167-192)."""
    dev = _lib.device()
    bars = uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab,
                             *[_lib.as_dev(x, dev)[None] for x in (T1a, T1b, T2aa, T2ab, T2bb)])
    bufs = (T1bara, T1barb, T2baraa, T2barab, T2barbb)
    for buf, b in zip(bufs, bars):
        _store_row(buf, ig, b[0])
    T1newa = quadrature.int_tbar1_single(ng, ig, T1bara, ti, D1a, G)
    T1newb = quadrature.int_tbar1_single(ng, ig, T1barb, ti, D1b, G)
    T2newaa = quadrature.int_tbar2_single(ng, ig, T2baraa, ti, D2aa, G)
    T2newab = quadrature.int_tbar2_single(ng, ig, T2barab, ti, D2ab, G)
    T2newbb = quadrature.int_tbar2_single(ng, ig, T2barbb, ti, D2bb, G)
    return (T1newa, T1newb), (T2newaa, T2newab, T2newbb)


# ---------------------------------------------------------------------------
# Lambda equations
# ---------------------------------------------------------------------------
_U_T = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
_U_L = ("l1.a", "l1.b", "l2.aa", "l2.ab", "l2.bb")
_U_LO = ("lo1.a", "lo1.b", "lo2.aa", "lo2.ab", "lo2.bb")


def _reps(names):
    return tuple(s for s in names if _plan.mirror_rep(s) == s)


def lambda_plan(mode, sizes, fac=-1.0, mirror=False):
    """Plan of -J(T)^T.Lbar - (F.ov + <ji||ba>t, I.oovv): intermediates of the
    forward residual followed by its mechanically derived reverse sweep."""
    key = ("lambda", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror))

    def build():
        inter, rest = programs.lambda_rops(mode, fac)
        if mode == "g":
            ins, outs = ("t1", "t2", "l1", "l2"), ("lo1", "lo2")
        else:
            ins, outs = _U_T + _U_L, _U_LO
        if mirror:
            inter, rest = _plan.mirror_reduce(inter), _plan.mirror_reduce(rest)
            ins, outs = _reps(ins), _reps(outs)
        return engine.Plan(inter + rest, mode, sizes, ins, outs,
                           name="lambda-" + mode + ("-closed" if mirror else ""))
    return engine.cached(key, build)


def lambda_split_plans(mode, sizes, fac=-1.0, mirror=False):
    """(prep, sweep): the amplitude-only forward intermediates (W_oooo, W_vvvv, W_ovvo,
    F_oo/vv/ov, tau ...) depend on T alone, which is fixed during the whole Lambda
    solve, so they are built once (prep) and every Lambda iteration runs only the
    reverse sweep (sweep) with the intermediates as inputs -- the 'cached W' cost
    model of SURVEY.md 8(d) (92 m^6 instead of 128 m^6 per grid point)."""
    key = ("lambda-split", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror))

    def build():
        inter, rest = programs.lambda_rops(mode, fac)
        tins = ("t1", "t2") if mode == "g" else _U_T
        lins = ("l1", "l2") if mode == "g" else _U_L
        louts = ("lo1", "lo2") if mode == "g" else _U_LO
        if mirror:
            inter, rest = _plan.mirror_reduce(inter), _plan.mirror_reduce(rest)
            tins, lins, louts = _reps(tins), _reps(lins), _reps(louts)
        inter_slots = []
        for op in inter:
            if op.out[0] not in inter_slots:
                inter_slots.append(op.out[0])
        prep = engine.Plan(inter, mode, sizes, tins, inter_slots, name="lambda-prep-" + mode)
        used = set()
        for op in rest:
            for slot, _ in op.ins:
                used.add(slot)
        keep = [s for s in inter_slots if s in used]
        sweep = engine.Plan(rest, mode, sizes, tuple(tins) + tuple(lins) + tuple(keep), louts,
                            name="lambda-sweep-" + mode)
        sweep.cached_slots = keep
        prep.all_slots = inter_slots
        return prep, sweep
    return engine.cached(key, build)


_lam_cache = {"key": None, "val": None, "refs": None}


def _lambda_intermediates(mode, sizes, ints_slots, tslots, ng, dev, mirror=False):
    """Forward intermediates for the current amplitudes, cached across Lambda iterations."""
    prep, sweep = lambda_split_plans(mode, sizes, mirror=mirror)
    # the cache entry keeps the key tensors alive: a live tensor's address cannot be handed to
    # another tensor, so equal (address, version) means the very same, unmodified tensors
    refs = list(tslots.values()) + list(ints_slots.values())
    key = (mode, bool(mirror), ng, tuple((v.data_ptr(), v._version) for v in refs))
    if _lam_cache["key"] == key and all(a is b for a, b in zip(_lam_cache["refs"], refs)):
        return _lam_cache["val"]
    _lam_cache["key"] = None
    _lam_cache["val"] = None
    _lam_cache["refs"] = None
    need = sum(8*ng*int(torch.tensor(prep.shapes[s]).prod()) for s in prep.all_slots)
    free, _ = torch.cuda.mem_get_info(dev)
    if need > 0.5*free:
        return None                      # too large to keep: use the fused (recompute) plan
    t = dict(ints_slots)
    t.update(tslots)
    for s in prep.all_slots:
        t[s] = torch.empty((ng,) + tuple(prep.shapes[s]), dtype=torch.float64, device=dev)
    prep.run({k: v for k, v in t.items() if k in prep.shapes}, ng)
    val = {s: t[s] for s in sweep.cached_slots}
    _lam_cache["key"] = key
    _lam_cache["val"] = val
    _lam_cache["refs"] = refs
    return val


def lambda_guess_plan(mode, sizes, beta, ls_ts_fac):
    key = ("lguess", mode, tuple(sorted(sizes.items(), key=str)), beta, ls_ts_fac)

    def build():
        rops = programs.lambda_guess_rops(mode, beta, ls_ts_fac)
        if mode == "g":
            ins, outs = ("t1",), ("lo1", "lo2")
        else:
            ins, outs = ("t1.a", "t1.b"), _U_LO
        return engine.Plan(rops, mode, sizes, ins, outs, name="lguess-" + mode)
    return engine.cached(key, build)


def _l_like(T1, T2):
    """Empty Lambda-shaped (o..v..) tensors matching amplitudes (v..o..)."""
    L1 = torch.empty((T1.shape[0], T1.shape[2], T1.shape[1]), dtype=torch.float64, device=T1.device)
    L2 = torch.empty((T2.shape[0], T2.shape[3], T2.shape[4], T2.shape[1], T2.shape[2]),
                     dtype=torch.float64, device=T2.device)
    return L1, L2


def ccsd_lambda_opt(F, I, T1old, T2old, L1old, L2old, D1, D2, ti, ng, g, G, beta):
    """Time-dependent CCSD Lambda iteration with intermediates
    (kelvin/ft_cc_equations.py:385-409)."""
    dev = _lib.device()
    T1old, T2old = _lib.as_dev(T1old, dev), _lib.as_dev(T2old, dev)
    L1int = quadrature.int_L1(ng, L1old, ti, D1, g, G)
    L2int = quadrature.int_L2(ng, L2old, ti, D2, g, G)
    sizes = _g_sizes(F)
    t = _g_integral_slots(F, I, dev)
    cached = _lambda_intermediates("g", sizes, t, {"t1": T1old, "t2": T2old}, ng, dev)
    t.update({"t1": T1old, "t2": T2old, "l1": L1int, "l2": L2int})
    t["lo1"], t["lo2"] = _l_like(T1old, T2old)
    if cached is not None:
        p = lambda_split_plans("g", sizes)[1]
        t.update(cached)
        p.run({k: v for k, v in t.items() if k in p.shapes}, ng, _chunk_for(p, ng, dev))
    else:
        p = lambda_plan("g", sizes)
        p.run(t, ng, _chunk_for(p, ng, dev))
    return t["lo1"], t["lo2"]


def uccsd_lambda_opt(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold,
                     T2bbold, L1aold, L1bold, L2aaold, L2abold, L2bbold, D1a,
                     D1b, D2aa, D2ab, D2bb, ti, ng, g, G, beta, closed_shell=False):
    """Unrestricted Lambda iteration (kelvin/ft_cc_equations.py:412-458).  closed_shell: as in
    uccsd_stanton_bar (integrals, amplitudes and Lambda all mirror symmetric)."""
    dev = _lib.device()
    Ts = [_lib.as_dev(x, dev) for x in (T1aold, T1bold, T2aaold, T2abold, T2bbold)]
    assert(Ts[0].shape[0] == ng)
    live = (0, 2, 3) if closed_shell else (0, 1, 2, 3, 4)
    Lin = (L1aold, L1bold, L2aaold, L2abold, L2bbold)
    Ds = (D1a, D1b, D2aa, D2ab, D2bb)
    Ls = {k: quadrature.int_L(ng, Lin[k], ti, Ds[k], g, G) for k in live}
    sizes = _u_sizes(Fa, Fb)
    pf = lambda_plan("u", sizes, mirror=closed_shell)
    t = _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                          [s for s in pf.inputs if _plan.is_integral_slot(s)])
    cached = _lambda_intermediates("u", sizes, t, {_U_T[k]: Ts[k] for k in live}, ng, dev,
                                   mirror=closed_shell)
    outs = [None]*5
    for k in live:
        t[_U_T[k]] = Ts[k]
        t[_U_L[k]] = Ls[k]
        outs[k] = t[_U_LO[k]] = torch.empty_like(Ls[k])
    if cached is not None:
        p = lambda_split_plans("u", sizes, mirror=closed_shell)[1]
        t.update(cached)
        p.run({k: v for k, v in t.items() if k in p.shapes}, ng, _chunk_for(p, ng, dev))
    else:
        pf.run(t, ng, _chunk_for(pf, ng, dev))
    if closed_shell:
        outs[1] = outs[0].clone()
        outs[4] = outs[2].clone()
    return tuple(outs)


def _variant_plan(method, sizes, fac=-1.0):
    key = ("residual", method, tuple(sorted(sizes.items(), key=str)), fac)

    def build():
        rops = _plan.expand(programs.residual_program(method, fac), programs.tensor_defs(), "g")
        s1 = programs.has_singles(method)
        ins = ("t1", "t2") if s1 else ("t2",)
        outs = ("o1", "o2") if s1 else ("o2",)
        return engine.Plan(rops, "g", sizes, ins, outs, name=method.lower() + "-g")
    return engine.cached(key, build)


def _variant_bar(method, F, I, T1old, T2old):
    dev = _lib.device()
    T2old = _lib.as_dev(T2old, dev)
    ng = T2old.shape[0]
    p = _variant_plan(method, _g_sizes(F))
    t = {k: v for k, v in _g_integral_slots(F, I, dev).items() if k in p.shapes}
    t["t2"] = T2old
    t["o2"] = torch.empty_like(T2old)
    if programs.has_singles(method):
        t["t1"] = _lib.as_dev(T1old, dev)
        t["o1"] = torch.empty_like(t["t1"])
    p.run(t, ng, _chunk_for(p, ng, dev))
    return t.get("o1"), t["o2"]


def lccd_simple(F, I, T2old, D2, ti, ng, G):
    """Time-dependent linearized coupled cluster doubles (LCCD) iteration
    (kelvin/ft_cc_equations.py:11-24)."""
    return quadrature.int_tbar2(ng, _variant_bar("LCCD", F, I, None, T2old)[1], ti, D2, G)


def lccsd_simple(F, I, T1old, T2old, D1, D2, ti, ng, G):
    """Time-dependent linearized coupled cluster singles and doubles (LCCSD) iteration
    (kelvin/ft_cc_equations.py:27-45)."""
    b1, b2 = _variant_bar("LCCSD", F, I, T1old, T2old)
    return quadrature.int_tbar1(ng, b1, ti, D1, G), quadrature.int_tbar2(ng, b2, ti, D2, G)


def ccd_simple(F, I, T2old, D2, ti, ng, G):
    """Time-dependent coupled cluster doubles (CCD) iteration
    (kelvin/ft_cc_equations.py:48-62)."""
    return quadrature.int_tbar2(ng, _variant_bar("CCD", F, I, None, T2old)[1], ti, D2, G)


def _variant_lambda(method, F, I, T1old, T2old, L1int, L2int, ng, beta):
    dev = _lib.device()
    sizes = _g_sizes(F)
    key = ("lambda-variant", method, tuple(sorted(sizes.items(), key=str)),
           beta if method == "LCCD" else None)
    s1 = programs.has_singles(method)

    def build():
        inter, rest = programs.lambda_rops("g", -1.0, method=method, beta=beta)
        ins = ("t1", "t2", "l1", "l2") if s1 else ("t2", "l2")
        outs = ("lo1", "lo2") if s1 else ("lo2",)
        return engine.Plan(inter + rest, "g", sizes, ins, outs, name="lambda-" + method.lower())
    p = engine.cached(key, build)
    t = {k: v for k, v in _g_integral_slots(F, I, dev).items() if k in p.shapes}
    T2old = _lib.as_dev(T2old, dev)
    if "t2" in p.shapes:
        t["t2"] = T2old
    t["l2"] = L2int
    t["lo2"] = torch.empty_like(L2int)
    if s1:
        if "t1" in p.shapes:
            t["t1"] = _lib.as_dev(T1old, dev)
        t["l1"] = L1int
        t["lo1"] = torch.empty_like(L1int)
    p.run(t, ng, _chunk_for(p, ng, dev))
    return t.get("lo1"), t["lo2"]


def lccd_lambda_simple(F, I, T2old, L2old, D2, ti, ng, g, G, beta):
    """LCCD Lambda iteration (kelvin/ft_cc_equations.py:292-310); the energy term carries the
    1/beta of the reference (:308)."""
    L2int = quadrature.int_L2(ng, L2old, ti, D2, g, G)
    return _variant_lambda("LCCD", F, I, None, T2old, None, L2int, ng, beta)[1]


def lccsd_lambda_simple(F, I, T1old, T2old, L1old, L2old, D1, D2, ti, ng, g, G, beta):
    """LCCSD Lambda iteration (kelvin/ft_cc_equations.py:313-340)."""
    L1int = quadrature.int_L1(ng, L1old, ti, D1, g, G)
    L2int = quadrature.int_L2(ng, L2old, ti, D2, g, G)
    return _variant_lambda("LCCSD", F, I, T1old, T2old, L1int, L2int, ng, beta)


def ccd_lambda_simple(F, I, T2old, L2old, D2, ti, ng, g, G, beta):
    """CCD Lambda iteration (kelvin/ft_cc_equations.py:682-701)."""
    L2int = quadrature.int_L2(ng, L2old, ti, D2, g, G)
    return _variant_lambda("CCD", F, I, None, T2old, None, L2int, ng, beta)[1]


def _rep_over_grid(x, ng, scale):
    x = _lib.as_dev(x)
    return (x*scale).expand(*((ng,) + (-1,)*x.dim())).contiguous()


def ccd_lambda_guess(I, beta, ng):
    """kelvin/ft_cc_equations.py:496-499."""
    return _rep_over_grid(I.oovv, ng, 1.0/beta)


def uccd_lambda_guess(Ia, Ib, Iabab, beta, ng):
    """kelvin/ft_cc_equations.py:529-535."""
    return tuple(_rep_over_grid(x.oovv, ng, 1.0/beta) for x in (Ia, Iabab, Ib))


def ccsd_lambda_guess(F, I, T1old, beta, ng):
    """CCSD Lambda guess (kelvin/ft_cc_equations.py:502-512)."""
    dev = _lib.device()
    T1old = _lib.as_dev(T1old, dev)
    p = lambda_guess_plan("g", _g_sizes(F), beta, 1.0/beta)
    t = _g_integral_slots(F, I, dev)
    t["t1"] = T1old
    no, nv = F.ov.shape
    t["lo1"] = torch.empty((ng, no, nv), dtype=torch.float64, device=dev)
    t["lo2"] = torch.empty((ng, no, no, nv, nv), dtype=torch.float64, device=dev)
    p.run({k: v for k, v in t.items() if k in p.shapes}, ng)
    return t["lo1"], t["lo2"]


def uccsd_lambda_guess(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, beta, ng):
    """Unrestricted Lambda guess; the <ji||ba>t term is NOT scaled by 1/beta,
    as in the reference (kelvin/ft_cc_equations.py:515-526, quirk Q3)."""
    dev = _lib.device()
    T1a, T1b = _lib.as_dev(T1aold, dev), _lib.as_dev(T1bold, dev)
    p = lambda_guess_plan("u", _u_sizes(Fa, Fb), beta, 1.0)
    t = _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                          [s for s in p.inputs if _plan.is_integral_slot(s)])
    t["t1.a"], t["t1.b"] = T1a, T1b
    noa, nva = Fa.ov.shape
    nob, nvb = Fb.ov.shape
    shp = {"lo1.a": (noa, nva), "lo1.b": (nob, nvb), "lo2.aa": (noa, noa, nva, nva),
           "lo2.ab": (noa, nob, nva, nvb), "lo2.bb": (nob, nob, nvb, nvb)}
    outs = []
    for nm in _U_LO:
        t[nm] = torch.empty((ng,) + shp[nm], dtype=torch.float64, device=dev)
        outs.append(t[nm])
    p.run(t, ng)
    return tuple(outs)


# ---------------------------------------------------------------------------
# response densities
# ---------------------------------------------------------------------------
def rdm_plan(mode, sizes):
    key = ("rdm", mode, tuple(sorted(sizes.items(), key=str)))

    def build():
        inter, rest = programs.rdm_rops(mode)
        ins = ("t1", "t2", "l1", "l2") if mode == "g" else _U_T + _U_L
        rops = inter + rest
        leaves = sorted({op.out[0] for op in rest
                         if op.out[0].endswith("~") and _plan._INT_SLOT.match(op.out[0])})
        p = engine.Plan(rops, mode, sizes, ins, leaves, name="rdm-" + mode)
        p.leaves = leaves
        return p
    return engine.cached(key, build)


def rdm_assembly_plan(mode, sizes):
    key = ("rdm-asm", mode, tuple(sorted(sizes.items(), key=str)))

    def build():
        ops, outs = programs.rdm2_assembly_rops(mode)
        shapes = _plan.slot_shapes(ops, mode, sizes)
        ins = [s for s in shapes if s.endswith("~")]
        p = engine.Plan(ops, mode, sizes, ins, [o[0] for o in outs], name="rdm-asm-" + mode,
                        shapes=shapes, batched={s: False for s in shapes})
        p.inputs = ins
        p.outs_meta = outs
        return p
    return engine.cached(key, build)


def _gsum(X, g, dev):
    """sum_y g[y] X[y]  (kelvin/ft_cc_equations.py:713,737)."""
    lib = _lib.load()
    gd = _lib.const_dev(g, dev)
    out = torch.empty(tuple(X.shape[1:]), dtype=torch.float64, device=dev)
    rc = lib.kb200_gsum(X.shape[0], out.numel(), _lib.ptr(X), _lib.ptr(gd), _lib.ptr(out),
                        _lib.stream_ptr())
    _lib.check(rc, "kb200_gsum")
    return out


_rdm_cache = {"key": None, "val": None, "refs": None}


def _rdm_leaves(mode, Ts, Ls, Ds, ti, ng, g, G):
    """Run the reverse sweep once per (T, L) and return
    (dict of g-summed integral adjoints, list of g-summed integrated Lambdas)."""
    dev = _lib.device()
    Ts = [_lib.as_dev(x, dev) for x in Ts]
    Ls = [_lib.as_dev(x, dev) for x in Ls]
    key = (mode, ng, tuple((x.data_ptr(), x._version) for x in Ts + Ls),
           tuple(float(v) for v in g), float(ti[-1]))
    refs = Ts + Ls          # kept alive by the cache entry (see _lambda_intermediates)
    if _rdm_cache["key"] == key and all(a is b for a, b in zip(_rdm_cache["refs"], refs)):
        return _rdm_cache["val"]
    Lbar = [quadrature.int_L(ng, L, ti, D, g, G) for L, D in zip(Ls, Ds)]
    t = {}
    if mode == "g":
        nv, no = Ts[0].shape[1:]
        sizes = {"o": int(no), "v": int(nv)}
        p = rdm_plan("g", sizes)
        names_t, names_l = ("t1", "t2"), ("l1", "l2")
    else:
        nva, noa = Ts[0].shape[1:]
        nvb, nob = Ts[1].shape[1:]
        sizes = {("o", "a"): int(noa), ("v", "a"): int(nva), ("o", "b"): int(nob),
                 ("v", "b"): int(nvb)}
        p = rdm_plan("u", sizes)
        names_t, names_l = _U_T, _U_L
    for nm, x in zip(names_t, Ts):
        t[nm] = x
    for nm, x in zip(names_l, Lbar):
        t[nm] = x
    for leaf in p.leaves:
        t[leaf] = torch.empty((ng,) + tuple(p.shapes[leaf]), dtype=torch.float64, device=dev)
    p.run(t, ng, _chunk_for(p, ng, dev))
    summed = {leaf: _gsum(t[leaf], g, dev) for leaf in p.leaves}
    lsum = [_gsum(x, g, dev) for x in Lbar]
    p.release()
    _rdm_cache["key"] = key
    _rdm_cache["val"] = (summed, lsum, sizes)
    _rdm_cache["refs"] = refs
    return _rdm_cache["val"]


def ccsd_1rdm(T1, T2, L1, L2, D1, D2, ti, ng, g, G):
    """pia, pba, pji, pai (kelvin/ft_cc_equations.py:704-722), evaluated as
    d(phi)/d(F blocks) by the reverse sweep of the residual plan."""
    assert(_lib.as_dev(T1).shape[0] == ng)
    summed, lsum, _ = _rdm_leaves("g", (T1, T2), (L1, L2), (D1, D2), ti, ng, g, G)
    return (lsum[0], summed["F.vv~"].t().contiguous(), summed["F.oo~"].t().contiguous(),
            summed["F.ov~"].t().contiguous())


def ccsd_2rdm(T1, T2, L1, L2, D1, D2, ti, ng, g, G):
    """(Pcdab, Pciab, Pbcai, Pijab, Pbjai, Pabij, Pjkai, Pkaij, Pklij)
    (kelvin/ft_cc_equations.py:725-753)."""
    summed, lsum, sizes = _rdm_leaves("g", (T1, T2), (L1, L2), (D1, D2), ti, ng, g, G)
    pa = rdm_assembly_plan("g", sizes)
    dev = _lib.device()
    t = {s: summed[s] for s in pa.inputs}
    for s in pa.outputs:
        t[s] = torch.empty(pa.shapes[s], dtype=torch.float64, device=dev)
    pa.run(t, 1)
    out = []
    for pname, pat in programs.RDM2_BLOCKS:
        out.append(lsum[1] if pat == "vvoo" else t["P" + pname])
    return tuple(out)


def uccsd_1rdm(T1a, T1b, T2aa, T2ab, T2bb, L1a, L1b, L2aa, L2ab, L2bb,
               D1a, D1b, D2aa, D2ab, D2bb, ti, ng, g, G):
    """((pia,pIA),(pba,pBA),(pji,pJI),(pai,pAI)) (kelvin/ft_cc_equations.py:756-798)."""
    summed, lsum, _ = _rdm_leaves("u", (T1a, T1b, T2aa, T2ab, T2bb),
                                  (L1a, L1b, L2aa, L2ab, L2bb),
                                  (D1a, D1b, D2aa, D2ab, D2bb), ti, ng, g, G)

    def tr(nm):
        return (summed["Fa.%s~" % nm].t().contiguous(), summed["Fb.%s~" % nm].t().contiguous())
    return (lsum[0], lsum[1]), tr("vv"), tr("oo"), tr("ov")


def uccsd_2rdm(T1a, T1b, T2aa, T2ab, T2bb, L1a, L1b, L2aa, L2ab, L2bb,
               D1a, D1b, D2aa, D2ab, D2bb, ti, ng, g, G):
    """The nine tuples of spin blocks in the reference's order
    (kelvin/ft_cc_equations.py:801-927)."""
    summed, lsum, sizes = _rdm_leaves("u", (T1a, T1b, T2aa, T2ab, T2bb),
                                      (L1a, L1b, L2aa, L2ab, L2bb),
                                      (D1a, D1b, D2aa, D2ab, D2bb), ti, ng, g, G)
    pa = rdm_assembly_plan("u", sizes)
    dev = _lib.device()
    t = {s: summed[s] for s in pa.inputs}
    for s in pa.outputs:
        t[s] = torch.empty(pa.shapes[s], dtype=torch.float64, device=dev)
    pa.run(t, 1)
    out = []
    for pname, pat in programs.RDM2_BLOCKS:
        if pat == "vvoo":
            out.append((lsum[2], lsum[4], lsum[3]))          # (aa, bb, ab)
        else:
            out.append(tuple(t["P%s.%s" % (pname, sp)] for sp in programs.RDM2_USPINS[pname]))
    return tuple(out)

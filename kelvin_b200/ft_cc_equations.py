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

from . import _lib, _trace, engine, plan as _plan, programs, quadrature

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
# tau batch up to which the large contractions accumulate into scratch slots of their own so that
# a whole dependency level of them shares one launch (plan.split_accumulators); beyond it the
# launches are many waves long and the extra elementwise adds do not pay
SPLIT_ACC = os.environ.get("KB200_SPLIT_ACC", "1") != "0"
SPLIT_ACC_MAX_BATCH = int(os.environ.get("KB200_SPLIT_ACC_MAX_BATCH", "2"))

# closed shell: sum/difference (singlet/triplet channel) form of the paired ring contractions
# (plan.sumdiff_pairs) and, when the amplitudes also satisfy T2aa = T2ab - T2ab(a<->b), the
# same-spin doubles residual from the opposite-spin one (plan.singlet_reduce)
SUMDIFF = int(os.environ.get("KB200_SUMDIFF", "1"))
SINGLET = int(os.environ.get("KB200_SINGLET", "1"))

_U_TIN = ("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb")
_U_TOUT = ("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb")


def _stanton_rops(mode, fac, mirror, mirror_rows, singlet, sumdiff, antisym, emit_aa=True,
                  split_acc=False):
    T = programs.tensor_defs()
    rops = _plan.expand(programs.stanton(fac), T, mode)
    if mode == "g":
        ins, outs = ("t1", "t2"), ("o1", "o2")
    else:
        ins, outs = _U_TIN, _U_TOUT
    if mirror:
        assert mode == "u"
        rops = _plan.mirror_reduce(rops)
        ins = tuple(s for s in ins if _plan.mirror_rep(s) == s)
        outs = tuple(s for s in outs if _plan.mirror_rep(s) == s)
        if singlet:
            rops = _plan.singlet_reduce(rops, emit_aa=emit_aa)
            if not emit_aa:
                outs = tuple(s for s in outs if s != "o2.aa")
        if sumdiff:
            rops = _plan.sumdiff_pairs(rops)
        if mirror_rows:
            rops = _plan.mirror_outputs(rops)
    if antisym:
        rops = _plan.antisym_outputs(rops)
    if split_acc:
        rops = _plan.split_accumulators(rops)
    return rops, ins, outs


def stanton_plan(mode, sizes, fac=-1.0, mirror=False, mirror_rows=False, singlet=False,
                 sumdiff=None, antisym=True, hybrid_world=None, emit_aa=True, split_acc=False):
    """mirror (u only): the closed-shell reduction of the program (plan.mirror_reduce): only the
    alpha-leading block of every alpha <-> beta pair is evaluated; mirror_rows: additionally
    plan.mirror_outputs; singlet / sumdiff: additionally plan.singlet_reduce /
    plan.sumdiff_pairs.  antisym=False: nothing is assumed about the permutational symmetry of
    the amplitudes (full sums over contracted pairs, full outputs, as the reference computes).
    hybrid_world=P: the same program as an engine.PhasedPlan evaluated by P ranks together.
    split_acc: plan.split_accumulators (for runs of 1-2 grid points per launch)."""
    mirror_rows = bool(mirror and mirror_rows and antisym)
    split_acc = bool(split_acc and SPLIT_ACC and not hybrid_world)
    singlet = bool(mirror and singlet and SINGLET and antisym)
    sumdiff = bool(mirror and (SUMDIFF if sumdiff is None else sumdiff))
    emit_aa = bool(emit_aa or not singlet)
    key = ("stanton", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror), mirror_rows,
           singlet, sumdiff, bool(antisym), hybrid_world, emit_aa, split_acc)

    def build():
        rops, ins, outs = _stanton_rops(mode, fac, mirror, mirror_rows, singlet, sumdiff, antisym,
                                        emit_aa, split_acc)
        name = "stanton-" + mode + ("-closed" if mirror else "")
        if hybrid_world:
            return engine.PhasedPlan(rops, mode, sizes, ins, outs, hybrid_world,
                                     name=name + "-hybrid", antisym=antisym)
        return engine.Plan(rops, mode, sizes, ins, outs, name=name, antisym=antisym)
    return engine.cached(key, build)


def antisym_ab_plan(shape):
    """aa[a,b,i,j] = ab[a,b,i,j] - ab[b,a,i,j] over all grid points: the same-spin doubles block
    of a closed-shell singlet quantity from the opposite-spin one (one fused elementwise pass)."""
    shape = tuple(int(d) for d in shape)
    key = ("antisym-ab", shape)

    def build():
        ops = [_plan.ROp(("aa", "abij"), 1.0, [("ab", "abij")]),
               _plan.ROp(("aa", "abij"), -1.0, [("ab", "baij")])]
        return engine.Plan(ops, "g", None, ["ab"], ["aa"], name="antisym-ab",
                           shapes={"ab": shape, "aa": shape}, batched={"ab": True, "aa": True})
    return engine.cached(key, build)


# ---------------------------------------------------------------------------
# symmetry detection (unrestricted inputs whose alpha and beta halves coincide;
# permutational antisymmetry of caller-supplied doubles)
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


def is_antisymmetric(X, tol=CLOSED_SHELL_TOL):
    """X[y,p,q,r,s] == -X[y,q,p,r,s] == -X[y,p,q,s,r] to `tol` relative?  The contraction plans
    sum antisymmetric contracted pairs over x < y and compute antisymmetric outputs on one
    triangle (plan.ANTISYM, plan.antisym_outputs); the amplitude equations preserve the
    antisymmetry of T2 / L2, but a caller-supplied guess need not have it -- the reference's
    full double sums (cqcpy einsums, kelvin/ft_cc_equations.py:96-113) then give something else.
    Device check, one grid point at a time; call it once per solve."""
    X = _lib.as_dev_rows(X)
    if X.dim() != 5 or X.shape[1] != X.shape[2] or X.shape[3] != X.shape[4]:
        return False
    for y in range(X.shape[0]):
        x = X[y]
        m = float(x.abs().max())
        if m == 0.0:
            continue
        if float((x + x.transpose(0, 1)).abs().max()) > tol*m \
                or float((x + x.transpose(2, 3)).abs().max()) > tol*m:
            return False
    return True


def is_singlet(X2aa, X2ab, tol=CLOSED_SHELL_TOL):
    """X2aa[y,p,q,r,s] == X2ab[y,p,q,r,s] - X2ab[y,q,p,r,s] (the same-spin block of a closed-shell
    singlet state is the antisymmetrised opposite-spin one)?  True of the MP2 guess of a
    spin-symmetric system and preserved by the update."""
    X2aa, X2ab = _lib.as_dev_rows(X2aa), _lib.as_dev_rows(X2ab)
    if tuple(X2aa.shape) != tuple(X2ab.shape):
        return False
    for y in range(X2aa.shape[0]):
        ab = X2ab[y]
        m = float(ab.abs().max())
        if float((ab - ab.transpose(0, 1) - X2aa[y]).abs().max()) > tol*max(m, 1e-300):
            return False
    return True


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
# evaluation of a per-grid-point program over the rows of the grid, sharded over the ranks
# ---------------------------------------------------------------------------
def _work_rows(work, name, ng, shapes, dev):
    """flat_rows, kept in the caller's `work` dict across calls: the iteration loops hand the
    same dict in every iteration, so the buffers -- and with them the addresses the captured
    launch graphs of the plans refer to -- stay the same."""
    shapes = [tuple(int(d) for d in shp) for shp in shapes]
    if work is None:
        return flat_rows(ng, shapes, dev)
    key = (name, ng, tuple(shapes))
    if work.get("_key_" + name) != key:
        work[name] = flat_rows(ng, shapes, dev)
        work["_key_" + name] = key
    return work[name]


def flat_rows(ng, shapes, dev):
    """One (ng, Ntot) buffer and its blocks as (ng, *shape) views (rows of the wide buffer): the
    blocks of one quantity travel in ONE collective (parallel.exchange_rows)."""
    ns = []
    for shp in shapes:
        n = 1
        for d in shp:
            n *= int(d)
        ns.append(n)
    flat = torch.empty((ng, sum(ns)), dtype=torch.float64, device=dev)
    views, off = [], 0
    for shp, n in zip(shapes, ns):
        views.append(flat[:, off:off + n].unflatten(1, tuple(int(d) for d in shp)))
        off += n
    return flat, views


class RowCache(object):
    """Per-grid-point tensors kept only for the rows this rank evaluates: `ranges` =
    [(first global row, stop)] laid out back to back in the leading axis of every tensor."""
    def __init__(self, ranges, tensors):
        self.ranges, self.tensors = list(ranges), tensors

    def local(self, a, b):
        off = 0
        for r0, r1 in self.ranges:
            if r0 <= a and b <= r1:
                return off + a - r0, off + b - r0
            off += r1 - r0
        raise Exception("RowCache: rows [%d, %d) are not held by this rank" % (a, b))

    def nrows(self):
        return sum(r1 - r0 for r0, r1 in self.ranges)


def evaluate_rows(p, hybrid, t, flat, ng, y0, dev, cache=None):
    """Fill the rows [y0, ng) of the outputs of plan `p` (column views of `flat`) on EVERY rank.
    t: slot -> tensor, batched slots with all ng rows; cache: RowCache of further batched input
    slots held for this rank's rows only; hybrid: callable returning the engine.PhasedPlan of
    the same program (or None).  Single rank: one plan run.  Sharded (parallel.active()): own
    rows locally, leftover rows together (hybrid) or by single owners, then one exchange."""
    from . import parallel

    def sliced(pl, a, b):
        tt = {}
        for s in pl.inputs + pl.outputs:
            if cache is not None and s in cache.tensors:
                la, lb = cache.local(a, b)
                tt[s] = cache.tensors[s][la:lb]
            else:
                tt[s] = t[s][a:b] if pl.batched[s] else t[s]
        return tt

    def run(a, b):
        if b > a:
            p.run(sliced(p, a, b), b - a, _chunk_for(p, b - a, dev))

    if not parallel.active():
        run(y0, ng)
        return
    sh = parallel.Shards(ng, y0)
    owner_left = False
    hp = None
    if sh.r > 0 and sh.use_hybrid() and hybrid is not None:
        try:
            hp = hybrid()
        except ValueError:
            hp = None                 # program shape the partition does not cover: single owners
    if hp is not None:
        # the shared grid points run on a high-priority side stream: their launches (small next
        # to this rank's own) and the two exchanges they contain overlap the own grid points,
        # which are enqueued first so that the device starts on the long part at once
        l0, l1 = sh.left
        cur = torch.cuda.current_stream()
        side = _side_stream(dev)
        side.wait_stream(cur)
        if sh.q > 1:
            # several own points: enqueue the long part first, the device starts on it at once
            run(*sh.own)
            _trace.mark("own")
        with torch.cuda.stream(side):
            hp.run(sliced(hp, l0, l1), l1 - l0, sh.rank, parallel.exchange_group())
            _trace.mark("shared")
        if sh.q <= 1:
            # one own point is as short as the shared one: the shared one, with its two
            # synchronisations with the other ranks, goes first (measured: 2.97 vs 3.29 ms)
            run(*sh.own)
            _trace.mark("own")
        cur.wait_stream(side)
    else:
        mine = sh.owner_row() if sh.r > 0 else None
        if sh.r > 0:
            owner_left = True
            parallel.zero_foreign_left_rows(flat, sh)
        if mine is not None and sh.q == 1 and cache is None:
            # one own row + one leftover row: two equally spaced grid points are ONE batch of 2
            # (a strided view of the row buffer), not two launches of batch 1
            a, step = sh.own[0], mine - sh.own[0]
            tt = {s_: (t[s_][a:mine + 1:step] if p.batched[s_] else t[s_])
                  for s_ in p.inputs + p.outputs}
            p.run(tt, 2, _chunk_for(p, 2, dev))
        else:
            run(*sh.own)
            if mine is not None:
                run(mine, mine + 1)
    _trace.mark("rows")
    parallel.exchange_rows(flat, sh, owner_left)
    _trace.mark("exchanged")


_side = {}


def _side_stream(dev):
    """High-priority stream for the grid points the ranks evaluate together: its many small
    launches get the next free SM instead of waiting behind the own grid points' wide launches."""
    if dev.index not in _side:
        _side[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    return _side[dev.index]


def needed_rows(ng, y0=0):
    """Row ranges of [y0, ng) this rank evaluates (alone or together with the others)."""
    from . import parallel
    if not parallel.active():
        return [(y0, ng)]
    sh = parallel.Shards(ng, y0)
    return sh.my_rows(sh.use_hybrid())


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


def ccsd_stanton_bar(F, I, T1old, T2old, fac=-1.0, t0_zero=False, antisym=None, work=None):
    """T1bar, T2bar: drivers + fac*StantonTerms at every grid point, i.e. the
    state of T1new/T2new just before the integration at
    kelvin/ft_cc_equations.py:109.  t0_zero: the caller guarantees T1old[0] = T2old[0] = 0
    (see t0_is_zero); the first grid point then costs nothing.  antisym: T2old is antisymmetric
    (None: checked here, see is_antisymmetric).  The two blocks are column ranges of one
    (ng, Ntot) buffer."""
    dev = _lib.device()
    T1old = _lib.as_dev(T1old, dev)
    T2old = _lib.as_dev(T2old, dev)
    ng = T1old.shape[0]
    if antisym is None:
        antisym = is_antisymmetric(T2old)
    sizes = _g_sizes(F)
    p = stanton_plan("g", sizes, fac, antisym=antisym)
    t = _g_integral_slots(F, I, dev)
    t["t1"], t["t2"] = T1old, T2old
    flat, (o1, o2) = _work_rows(work, "bar", ng, (T1old.shape[1:], T2old.shape[1:]), dev)
    t["o1"], t["o2"] = o1, o2
    y0 = 1 if (t0_zero and ng > 1) else 0
    from . import parallel
    hyb = (lambda: stanton_plan("g", sizes, fac, antisym=antisym,
                                hybrid_world=parallel.world_info()[1]))
    evaluate_rows(p, hyb, t, flat, ng, y0, dev)
    if y0:
        o1[0].copy_(t["F.vo"]).neg_()
        o2[0].copy_(t["I.vvoo"]).neg_()
    return o1, o2


def ccsd_stanton(F, I, T1old, T2old, D1, D2, ti, ng, G, t0_zero=False, antisym=None):
    """Time-dependent CCSD iteration using Stanton-Gauss intermediates
    (kelvin/ft_cc_equations.py:96-113)."""
    T1bar, T2bar = ccsd_stanton_bar(F, I, T1old, T2old, t0_zero=t0_zero, antisym=antisym)
    T1new = quadrature.int_tbar1(ng, T1bar, ti, D1, G)
    T2new = quadrature.int_tbar2(ng, T2bar, ti, D2, G)
    return T1new, T2new


def uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold, fac=-1.0,
                      t0_zero=False, closed_shell=False, beta_copies=True, singlet=False,
                      antisym=None, work=None):
    """closed_shell: the caller guarantees mirror-symmetric integrals and amplitudes
    (closed_shell_integrals / closed_shell_amplitudes); the beta-leading blocks are then copies
    of their alpha images and only the reduced program runs (plan.mirror_reduce);
    singlet: additionally T2aa = T2ab - T2ab(a<->b) (is_singlet).
    beta_copies=False returns None in the place of the beta blocks (the caller copies later).
    antisym: T2aa/T2bb are antisymmetric (None: checked here)."""
    dev = _lib.device()
    ins = [_lib.as_dev(x, dev) for x in (T1aold, T1bold, T2aaold, T2abold, T2bbold)]
    ng = ins[0].shape[0]
    if antisym is None:
        antisym = bool(closed_shell and singlet) or \
            (is_antisymmetric(ins[2]) and (closed_shell or is_antisymmetric(ins[4])))
    y0 = 1 if (t0_zero and ng > 1) else 0
    sizes = _u_sizes(Fa, Fb)
    from . import parallel
    nloc = max(b - a for a, b in needed_rows(ng, y0)) if ng > y0 else 0
    sing = bool(closed_shell and singlet and SINGLET and antisym)
    # singlet runs: the programs produce T̄1a and T̄2ab only; T̄2aa = T̄2ab - T̄2ab(a<->b) is formed
    # for all grid points AFTER the exchange, which then moves half the bytes
    kw = dict(mirror=closed_shell, singlet=singlet, antisym=antisym, emit_aa=not sing)
    p = stanton_plan("u", sizes, fac, mirror_rows=nloc >= MIRROR_ROWS_MIN_BATCH,
                     split_acc=0 < nloc <= SPLIT_ACC_MAX_BATCH, **kw)
    setup = None if work is None else work.get("_setup")
    if setup is not None and setup[0] is p and all(a is b for a, b in zip(setup[1], ins)):
        t = setup[2]
    else:
        t = _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                              [s for s in p.inputs if _plan.is_integral_slot(s)])
        if work is not None:
            work["_setup"] = (p, list(ins), t)
    drivers = [Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo]
    live = [k for k in range(5) if not closed_shell or k in (0, 2, 3)]
    xch = [k for k in live if not (sing and k == 2)]          # blocks the programs produce
    flat, views = _work_rows(work, "bar", ng, [ins[k].shape[1:] for k in xch], dev)
    outs = [None]*5
    for k, v in zip(xch, views):
        t[_U_TIN[k]] = ins[k]
        outs[k] = t[_U_TOUT[k]] = v
    if sing:
        t[_U_TIN[2]] = ins[2]
    hyb = (lambda: stanton_plan("u", sizes, fac, hybrid_world=parallel.world_info()[1], **kw))
    evaluate_rows(p, hyb, t, flat, ng, y0, dev)
    if sing:
        _, (aa,) = _work_rows(work, "bar-aa", ng, [ins[2].shape[1:]], dev)
        if ng > y0:
            antisym_ab_plan(ins[2].shape[1:]).run({"ab": outs[3][y0:], "aa": aa[y0:]}, ng - y0)
        outs[2] = aa
    if y0 and not (work is not None and work.get("_t0_rows") == id(flat)):
        # (row 0 is never written by the programs: with the loop's persistent buffers once is enough)
        for k in live:
            outs[k][0].copy_(_lib.as_dev(drivers[k], dev)).neg_()
        if work is not None:
            work["_t0_rows"] = id(flat)
    if closed_shell and beta_copies:
        outs[1] = outs[0]
        outs[4] = outs[2]
    return tuple(outs)


def uccsd_stanton(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold,
                  T2bbold, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G, t0_zero=False,
                  closed_shell=False, singlet=False, antisym=None):
    """Unrestricted CCSD iteration (kelvin/ft_cc_equations.py:130-164)."""
    b1a, b1b, b2aa, b2ab, b2bb = uccsd_stanton_bar(
        Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold, T2bbold, t0_zero=t0_zero,
        closed_shell=closed_shell, singlet=singlet, antisym=antisym)
    T1a = quadrature.int_tbar1(ng, b1a, ti, D1a, G)
    T2aa = quadrature.int_tbar2(ng, b2aa, ti, D2aa, G)
    T2ab = quadrature.int_tbar2(ng, b2ab, ti, D2ab, G)
    if closed_shell:
        # equal denominators (checked with the integrals): the beta blocks are copies
        T1b, T2bb = T1a.clone(), T2aa.clone()
    else:
        T1b = quadrature.int_tbar1(ng, b1b, ti, D1b, G)
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
    # the pointwise solver builds its iterates from zero by antisymmetry-preserving updates
    b1, b2 = ccsd_stanton_bar(F, I, _lib.as_dev(T1old, dev)[None], _lib.as_dev(T2old, dev)[None],
                              antisym=True)
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
                             *[_lib.as_dev(x, dev)[None] for x in (T1a, T1b, T2aa, T2ab, T2bb)],
                             antisym=True)
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


def _lambda_rops(mode, fac, mirror, antisym):
    """(forward intermediates, reverse sweep) with the closed-shell rewrites of the sweep:
    mirror-duplicate contractions done once (plan.merge_duplicates), the three quartets of ring
    contractions as sum/difference pairs (plan.sumdiff_pairs), and the same-spin ladder adjoints
    on their triangle (plan.antisym_outputs)."""
    inter, rest = programs.lambda_rops(mode, fac)
    if mirror:
        inter, rest = _plan.mirror_reduce(inter), _plan.mirror_reduce(rest)
        if SUMDIFF:
            rest = _plan.sumdiff_pairs(_plan.merge_duplicates(rest))
    if antisym:
        rest = _plan.antisym_outputs(rest)
    return inter, rest


def lambda_plan(mode, sizes, fac=-1.0, mirror=False, antisym=True, hybrid_world=None):
    """Plan of -J(T)^T.Lbar - (F.ov + <ji||ba>t, I.oovv): intermediates of the
    forward residual followed by its mechanically derived reverse sweep."""
    key = ("lambda", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror), bool(antisym),
           hybrid_world)

    def build():
        inter, rest = _lambda_rops(mode, fac, mirror, antisym)
        if mode == "g":
            ins, outs = ("t1", "t2", "l1", "l2"), ("lo1", "lo2")
        else:
            ins, outs = _U_T + _U_L, _U_LO
        if mirror:
            ins, outs = _reps(ins), _reps(outs)
        name = "lambda-" + mode + ("-closed" if mirror else "")
        if hybrid_world:
            return engine.PhasedPlan(inter + rest, mode, sizes, ins, outs, hybrid_world,
                                     name=name + "-hybrid", antisym=antisym)
        return engine.Plan(inter + rest, mode, sizes, ins, outs, name=name, antisym=antisym)
    return engine.cached(key, build)


def lambda_split_plans(mode, sizes, fac=-1.0, mirror=False, antisym=True, hybrid_world=None):
    """(prep, sweep): the amplitude-only forward intermediates (W_oooo, W_vvvv, W_ovvo,
    F_oo/vv/ov, tau ...) depend on T alone, which is fixed during the whole Lambda
    solve, so they are built once (prep) and every Lambda iteration runs only the
    reverse sweep (sweep) with the intermediates as inputs -- the 'cached W' cost
    model of SURVEY.md 8(d) (92 m^6 instead of 128 m^6 per grid point).
    hybrid_world=P: the sweep as an engine.PhasedPlan evaluated by P ranks together."""
    key = ("lambda-split", mode, tuple(sorted(sizes.items(), key=str)), fac, bool(mirror),
           bool(antisym), hybrid_world)

    def build():
        inter, rest = _lambda_rops(mode, fac, mirror, antisym)
        tins = ("t1", "t2") if mode == "g" else _U_T
        lins = ("l1", "l2") if mode == "g" else _U_L
        louts = ("lo1", "lo2") if mode == "g" else _U_LO
        if mirror:
            tins, lins, louts = _reps(tins), _reps(lins), _reps(louts)
        inter_slots = []
        for op in inter:
            if op.out[0] not in inter_slots:
                inter_slots.append(op.out[0])
        used = set()
        for op in rest:
            for slot, _ in op.ins:
                used.add(slot)
        keep = [s for s in inter_slots if s in used]
        sweep_ins = tuple(tins) + tuple(lins) + tuple(keep)
        if hybrid_world:
            sweep = engine.PhasedPlan(rest, mode, sizes, sweep_ins, louts, hybrid_world,
                                      name="lambda-sweep-" + mode + "-hybrid", antisym=antisym)
            sweep.cached_slots = keep
            return None, sweep
        prep = engine.Plan(inter, mode, sizes, tins, inter_slots, name="lambda-prep-" + mode,
                           antisym=antisym)
        sweep = engine.Plan(rest, mode, sizes, sweep_ins, louts, name="lambda-sweep-" + mode,
                            antisym=antisym)
        sweep.cached_slots = keep
        prep.all_slots = inter_slots
        return prep, sweep
    return engine.cached(key, build)


_lam_cache = {"key": None, "val": None, "refs": None}


def clear_solve_caches():
    """Drop the per-solve caches (forward intermediates of the Lambda sweep, response-density
    leaves).  They are keyed by the identity, address and version of the amplitude tensors; a
    caller that overwrites those tensors through raw pointers must call this."""
    _lam_cache.update(key=None, val=None, refs=None)
    _rdm_cache.update(key=None, val=None, refs=None)


def _lambda_intermediates(mode, sizes, ints_slots, tslots, ng, dev, mirror=False, antisym=True):
    """Forward intermediates for the current amplitudes on the rows this rank evaluates, cached
    across Lambda iterations (RowCache), or None when they do not fit."""
    prep, sweep = lambda_split_plans(mode, sizes, mirror=mirror, antisym=antisym)
    ranges = needed_rows(ng)
    # the cache entry keeps the key tensors alive: a live tensor's address cannot be handed to
    # another tensor, so equal (address, version) means the very same, unmodified tensors
    refs = list(tslots.values()) + list(ints_slots.values())
    key = (mode, bool(mirror), bool(antisym), ng, tuple(ranges),
           tuple((v.data_ptr(), v._version) for v in refs))
    if _lam_cache["key"] == key and all(a is b for a, b in zip(_lam_cache["refs"], refs)):
        return _lam_cache["val"]
    _lam_cache["key"] = None
    _lam_cache["val"] = None
    _lam_cache["refs"] = None
    nrows = sum(b - a for a, b in ranges)
    need = sum(8*nrows*int(torch.tensor(prep.shapes[s]).prod()) for s in prep.all_slots)
    free, _ = torch.cuda.mem_get_info(dev)
    if need > 0.5*free:
        return None                      # too large to keep: use the fused (recompute) plan
    bufs = {s: torch.empty((nrows,) + tuple(prep.shapes[s]), dtype=torch.float64, device=dev)
            for s in prep.all_slots}
    off = 0
    for a, b in ranges:
        t = dict(ints_slots)
        for k, v in tslots.items():
            t[k] = v[a:b]
        for s in prep.all_slots:
            t[s] = bufs[s][off:off + b - a]
        prep.run({k: v for k, v in t.items() if k in prep.shapes}, b - a)
        off += b - a
    val = RowCache(ranges, {s: bufs[s] for s in sweep.cached_slots})
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


def _work_like(work, name, x, dev):
    """Persistent buffer shaped like x in the caller's `work` dict (None: allocate afresh)."""
    if work is None:
        return None
    shp = tuple(x.shape)
    buf = work.get(name)
    if buf is None or tuple(buf.shape) != shp:
        buf = work[name] = torch.empty(shp, dtype=torch.float64, device=dev)
    return buf


def _l_shapes(T1, T2):
    """Lambda-shaped (o..v..) block shapes matching amplitudes (v..o..)."""
    return ((T1.shape[2], T1.shape[1]), (T2.shape[3], T2.shape[4], T2.shape[1], T2.shape[2]))


def _lambda_rows(mode, sizes, t, tslots, flat, ng, dev, mirror, antisym):
    """Evaluate the Lambda map on all rows: reverse sweep over cached forward intermediates when
    they fit, else the fused program."""
    from . import parallel
    ints = {k: v for k, v in t.items() if _plan.is_integral_slot(k)}
    cached = _lambda_intermediates(mode, sizes, ints, tslots, ng, dev, mirror=mirror,
                                   antisym=antisym)
    world = parallel.world_info()[1]
    if cached is not None:
        p = lambda_split_plans(mode, sizes, mirror=mirror, antisym=antisym)[1]
        hyb = (lambda: lambda_split_plans(mode, sizes, mirror=mirror, antisym=antisym,
                                          hybrid_world=world)[1])
        evaluate_rows(p, hyb, t, flat, ng, 0, dev, cache=cached)
    else:
        p = lambda_plan(mode, sizes, mirror=mirror, antisym=antisym)
        hyb = (lambda: lambda_plan(mode, sizes, mirror=mirror, antisym=antisym,
                                   hybrid_world=world))
        evaluate_rows(p, hyb, t, flat, ng, 0, dev)


def ccsd_lambda_opt(F, I, T1old, T2old, L1old, L2old, D1, D2, ti, ng, g, G, beta, antisym=None,
                    work=None):
    """Time-dependent CCSD Lambda iteration with intermediates
    (kelvin/ft_cc_equations.py:385-409).  antisym: T2old and L2old are antisymmetric (None:
    checked here)."""
    dev = _lib.device()
    T1old, T2old = _lib.as_dev(T1old, dev), _lib.as_dev(T2old, dev)
    if antisym is None:
        antisym = is_antisymmetric(T2old) and is_antisymmetric(L2old)
    L1int = quadrature.int_L(ng, L1old, ti, D1, g, G, out=_work_like(work, "l1", L1old, dev))
    L2int = quadrature.int_L(ng, L2old, ti, D2, g, G, out=_work_like(work, "l2", L2old, dev))
    sizes = _g_sizes(F)
    t = _g_integral_slots(F, I, dev)
    t.update({"t1": T1old, "t2": T2old, "l1": L1int, "l2": L2int})
    flat, (lo1, lo2) = _work_rows(work, "lo", ng, _l_shapes(T1old, T2old), dev)
    t["lo1"], t["lo2"] = lo1, lo2
    _lambda_rows("g", sizes, t, {"t1": T1old, "t2": T2old}, flat, ng, dev, False, antisym)
    return lo1, lo2


def uccsd_lambda_opt(Fa, Fb, Ia, Ib, Iabab, T1aold, T1bold, T2aaold, T2abold,
                     T2bbold, L1aold, L1bold, L2aaold, L2abold, L2bbold, D1a,
                     D1b, D2aa, D2ab, D2bb, ti, ng, g, G, beta, closed_shell=False, antisym=None,
                     work=None):
    """Unrestricted Lambda iteration (kelvin/ft_cc_equations.py:412-458).  closed_shell: as in
    uccsd_stanton_bar (integrals, amplitudes and Lambda all mirror symmetric); the beta blocks of
    the result are then the alpha tensors themselves."""
    dev = _lib.device()
    Ts = [_lib.as_dev(x, dev) for x in (T1aold, T1bold, T2aaold, T2abold, T2bbold)]
    assert(Ts[0].shape[0] == ng)
    live = (0, 2, 3) if closed_shell else (0, 1, 2, 3, 4)
    Lin = (L1aold, L1bold, L2aaold, L2abold, L2bbold)
    if antisym is None:
        antisym = all(is_antisymmetric(x) for x in
                      ([Ts[2], Lin[2]] + ([] if closed_shell else [Ts[4], Lin[4]])))
    Ds = (D1a, D1b, D2aa, D2ab, D2bb)
    Ls = {k: quadrature.int_L(ng, Lin[k], ti, Ds[k], g, G,
                              out=_work_like(work, "l%d" % k, Lin[k], dev)) for k in live}
    sizes = _u_sizes(Fa, Fb)
    pf = lambda_plan("u", sizes, mirror=closed_shell, antisym=antisym)
    t = _u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                          [s for s in pf.inputs if _plan.is_integral_slot(s)])
    shp = _l_shapes(Ts[0], Ts[2]) + _l_shapes(Ts[1], Ts[4])
    lshape = {0: shp[0], 1: shp[2], 2: shp[1], 4: shp[3],
              3: (Ts[3].shape[3], Ts[3].shape[4], Ts[3].shape[1], Ts[3].shape[2])}
    flat, views = _work_rows(work, "lo", ng, [lshape[k] for k in live], dev)
    outs = [None]*5
    for k, v in zip(live, views):
        t[_U_T[k]] = Ts[k]
        t[_U_L[k]] = Ls[k]
        outs[k] = t[_U_LO[k]] = v
    _lambda_rows("u", sizes, t, {_U_T[k]: Ts[k] for k in live}, flat, ng, dev, closed_shell,
                 antisym)
    if closed_shell:
        outs[1] = outs[0]
        outs[4] = outs[2]
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
    (dict of g-summed integral adjoints, list of g-summed integrated Lambdas).
    Sharded runs: every rank sweeps its own grid points (leftover points go to single owners)
    and the weighted sums over the grid, sum_y g_y P(y) (kelvin/ft_cc_equations.py:717-720,
    743-751), are completed by ONE all-reduce of the leaf buffer."""
    from . import parallel
    import numpy
    dev = _lib.device()
    Ts = [_lib.as_dev(x, dev) for x in Ts]
    Ls = [_lib.as_dev(x, dev) for x in Ls]
    key = (mode, ng, tuple((x.data_ptr(), x._version) for x in Ts + Ls),
           tuple(float(v) for v in g), tuple(float(v) for v in ti),
           tuple(float(v) for v in numpy.asarray(G).reshape(-1)),
           tuple(int(_lib.as_dev(d, dev).data_ptr()) for d in Ds))
    refs = Ts + Ls          # kept alive by the cache entry (see _lambda_intermediates)
    if _rdm_cache["key"] == key and all(a is b for a, b in zip(_rdm_cache["refs"], refs)):
        return _rdm_cache["val"]
    Lbar = [quadrature.int_L(ng, L, ti, D, g, G) for L, D in zip(Ls, Ds)]
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
    if parallel.active():
        ranges = parallel.Shards(ng, 0).my_rows(False)
    else:
        ranges = [(0, ng)]
    acc_flat, acc = flat_rows(1, [p.shapes[leaf] for leaf in p.leaves], dev)
    acc_flat.zero_()
    summed = {leaf: v[0] for leaf, v in zip(p.leaves, acc)}
    gh = numpy.asarray(g, dtype=numpy.float64)
    for a, b in ranges:
        t = {}
        for nm, x in zip(names_t, Ts):
            t[nm] = x[a:b]
        for nm, x in zip(names_l, Lbar):
            t[nm] = x[a:b]
        for leaf in p.leaves:
            t[leaf] = torch.empty((b - a,) + tuple(p.shapes[leaf]), dtype=torch.float64, device=dev)
        p.run(t, b - a, _chunk_for(p, b - a, dev))
        for leaf in p.leaves:
            summed[leaf].add_(_gsum(t[leaf], gh[a:b], dev))
        t = None
    parallel.sum_over_ranks(acc_flat)
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

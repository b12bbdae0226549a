"""Statement lists for the FT-CCSD residual (Stanton-Gauss-Watts-Bartlett
factorisation, JCP 94, 4334 (1991)) in the block conventions of the reference
(kelvin/lambda_stanton.py:7-34): t1[a,i], t2[a,b,i,j],
I.wxyz[p,q,r,s] = <pq||rs>, F.oo/F.vv with the orbital energies removed
(kelvin/cc_utils.py:578).

``stanton(fac)`` is the body of cqcpy.cc_equations._Stanton as the reference
uses it at kelvin/ft_cc_equations.py:101-107: the outputs o1/o2 start from the
drivers -F.vo / -I.vvoo and accumulate fac*R1 / fac*R2.
"""
from .plan import TDef, parse

_INT2 = ("vvvv", "vvvo", "vovv", "vvoo", "vovo", "oovv", "vooo", "ooov", "oooo")


def tensor_defs():
    T = {}

    def add(name, kind, role, batched, spaces, canon=None):
        T[name] = TDef(name, kind, role, batched, spaces, canon)
    for xy in ("oo", "ov", "vo", "vv"):
        add("F." + xy, "int1", "in", False, xy)
    for pat in _INT2:
        add("I." + pat, "int2", "in", False, pat)
    add("t1", "one", "in", True, "vo")
    add("t2", "amp2", "in", True, "vvoo")
    add("o1", "one", "out", True, "vo")
    add("o2", "amp2", "out", True, "vvoo")
    add("tau", "amp2", "tmp", True, "vvoo")
    add("tauh", "amp2", "tmp", True, "vvoo")
    add("Fvv", "one", "tmp", True, "vv")
    add("Foo", "one", "tmp", True, "oo")
    add("Fov", "one", "tmp", True, "ov")
    add("Xvv", "one", "tmp", True, "vv")
    add("Xoo", "one", "tmp", True, "oo")
    add("Woooo", "amp2", "tmp", True, "oooo")
    add("Wvvvv", "amp2", "tmp", True, "vvvv")
    # Storage orders of the ring-term intermediates are chosen so that every m^6 contraction
    # reads both operands as plain matrices (rows x contiguous k, or k x contiguous columns):
    # W_ovvo[m,b,e,j] is stored [e,m,b,j]: rows (m,e) of its build and the contracted pair (e,m)
    # of the ring contraction lead, the columns (b,j) are contiguous;
    add("Wovvo", "gen4", "tmp", True, "vovo", canon=(1, 2, 0, 3))
    # t2x stored [b,j,n,f]: the contracted pair (n,f) is contiguous with f innermost, the same
    # inner index as I.oovv[mnef];
    add("t2x", "gen4", "tmp", True, "voov", canon=(3, 0, 1, 2))
    # t2r[a,i,e,m] = t2[a,e,i,m]: the amplitudes as the ring contraction reads them, rows (a,i)
    # against the contiguous contracted pair (e,m) -- one permuted copy per iteration instead of
    # a strided gather in every k-tile of three m^6 contractions.
    add("t2r", "gen4", "tmp", True, "vovo", canon=(0, 2, 1, 3))
    add("Z", "gen4", "tmp", True, "vooo")
    add("rg", "gen4", "tmp", True, "vvoo")
    return T


_INTERMEDIATES = """
tau[abij] += 1 t2[abij]
tau[abij] += 1 t1[ai] t1[bj]
tau[abij] += -1 t1[bi] t1[aj]
tauh[abij] += 1 t2[abij]
tauh[abij] += 0.5 t1[ai] t1[bj]
tauh[abij] += -0.5 t1[bi] t1[aj]
Fvv[ae] += 1 F.vv[ae]
Fvv[ae] += -0.5 F.ov[me] t1[am]
Fvv[ae] += 1 I.vovv[amef] t1[fm]
Fvv[ae] += -0.5 I.oovv[mnef] tauh[afmn]
Foo[mi] += 1 F.oo[mi]
Foo[mi] += 0.5 F.ov[me] t1[ei]
Foo[mi] += 1 I.ooov[mnie] t1[en]
Foo[mi] += 0.5 I.oovv[mnef] tauh[efin]
Fov[me] += 1 F.ov[me]
Fov[me] += 1 I.oovv[mnef] t1[fn]
Woooo[mnij] += 1 I.oooo[mnij]
Woooo[mnij] += 1 I.ooov[mnie] t1[ej]
Woooo[mnij] += -1 I.ooov[mnje] t1[ei]
Woooo[mnij] += 0.25 I.oovv[mnef] tau[efij]
Wvvvv[abef] += 1 I.vvvv[abef]
Wvvvv[abef] += -1 I.vovv[amef] t1[bm]
Wvvvv[abef] += 1 I.vovv[bmef] t1[am]
Wvvvv[abef] += 0.25 I.oovv[mnef] tau[abmn]
t2x[bjnf] += 0.5 t2[fbjn]
t2x[bjnf] += 1 t1[fj] t1[bn]
t2r[aiem] += 1 t2[aeim]
Wovvo[embj] += -1 I.vovo[bmej]
Wovvo[embj] += -1 I.vovv[bmef] t1[fj]
Wovvo[embj] += 1 I.ooov[mnje] t1[bn]
Wovvo[embj] += -1 I.oovv[mnef] t2x[bjnf]
Xvv[be] += 1 Fvv[be]
Xvv[be] += -0.5 t1[bm] Fov[me]
Xoo[mj] += 1 Foo[mj]
Xoo[mj] += 0.5 t1[ej] Fov[me]
Z[bmij] += 1 I.vovo[bmej] t1[ei]
rg[abij] += 1 t2r[aiem] Wovvo[embj]
rg[abij] += 1 t1[am] Z[bmij]
"""

# residual statements; coefficient is multiplied by `fac`
_RESIDUAL = """
o1[ai] += 1 Fvv[ae] t1[ei]
o1[ai] += -1 Foo[mi] t1[am]
o1[ai] += 1 Fov[me] t2[aeim]
o1[ai] += -1 I.vovo[anfi] t1[fn]
o1[ai] += 0.5 I.vovv[amef] t2[efim]
o1[ai] += -0.5 I.ooov[mnie] t2[aemn]
o2[abij] += 1 t2[aeij] Xvv[be]
o2[abij] += -1 t2[beij] Xvv[ae]
o2[abij] += -1 t2[abim] Xoo[mj]
o2[abij] += 1 t2[abjm] Xoo[mi]
o2[abij] += 0.5 tau[abmn] Woooo[mnij]
o2[abij] += 0.5 tau[efij] Wvvvv[abef]
o2[abij] += 1 rg[abij]
o2[abij] += -1 rg[baij]
o2[abij] += -1 rg[abji]
o2[abij] += 1 rg[baji]
o2[abij] += 1 t1[ei] I.vvvo[abej]
o2[abij] += -1 t1[ej] I.vvvo[abei]
o2[abij] += 1 t1[am] I.vooo[bmij]
o2[abij] += -1 t1[bm] I.vooo[amij]
"""

_DRIVERS = """
o1[ai] += -1 F.vo[ai]
o2[abij] += -1 I.vvoo[abij]
"""


def _parse_block(text, scale=1.0):
    out = []
    for line in text.strip().splitlines():
        line = line.strip()
        if not line:
            continue
        st = parse(line)
        st.coef *= scale
        out.append(st)
    return out


def _intermediates(relayout=True):
    """relayout=False: the ring contraction reads t2 itself (no t2r copy): the form the reverse
    sweeps are derived from (their closed-shell rewrites match on the amplitude slots)."""
    text = _INTERMEDIATES
    if not relayout:
        text = text.replace("t2r[aiem] += 1 t2[aeim]\n", "")
        text = text.replace("rg[abij] += 1 t2r[aiem] Wovvo[embj]", "rg[abij] += 1 t2[aeim] Wovvo[embj]")
        assert "t2r" not in text
    return _parse_block(text)


def stanton(fac=-1.0, drivers=True, relayout=True):
    """Statements of one residual evaluation: o = drivers + fac*StantonTerms."""
    st = []
    if drivers:
        st += _parse_block(_DRIVERS)
    st += _intermediates(relayout)
    st += _parse_block(_RESIDUAL, fac)
    return st


# Terms of the residual that are linear in the amplitudes (LCCSD; LCCD = its t2-only part):
# cqcpy's _S_S + _S_D and _D_S + _D_D as the reference combines them at
# kelvin/ft_cc_equations.py:11-45.  This is the Jacobian of StantonTerms at T = 0 (checked against
# the full polynomial in tests/test_plan_cpu.py).
_LINEAR = """
o1[ai] += 1 F.vv[ae] t1[ei]
o1[ai] += -1 F.oo[mi] t1[am]
o1[ai] += -1 I.vovo[anfi] t1[fn]
o1[ai] += 1 F.ov[me] t2[aeim]
o1[ai] += 0.5 I.vovv[amef] t2[efim]
o1[ai] += -0.5 I.ooov[mnie] t2[aemn]
o2[abij] += 1 t1[ei] I.vvvo[abej]
o2[abij] += -1 t1[ej] I.vvvo[abei]
o2[abij] += 1 t1[am] I.vooo[bmij]
o2[abij] += -1 t1[bm] I.vooo[amij]
o2[abij] += 1 t2[aeij] F.vv[be]
o2[abij] += -1 t2[beij] F.vv[ae]
o2[abij] += -1 t2[abim] F.oo[mj]
o2[abij] += 1 t2[abjm] F.oo[mi]
o2[abij] += 0.5 t2[abmn] I.oooo[mnij]
o2[abij] += 0.5 t2[efij] I.vvvv[abef]
rg[abij] += -1 t2[aeim] I.vovo[bmej]
o2[abij] += 1 rg[abij]
o2[abij] += -1 rg[baij]
o2[abij] += -1 rg[abji]
o2[abij] += 1 rg[baji]
"""

METHODS = ("CCSD", "CCD", "LCCSD", "LCCD")


def has_singles(method):
    return method in ("CCSD", "LCCSD")


def _without_singles(stmts):
    """t1 := 0: drop every statement that reads t1 or writes the singles residual, then
    every intermediate nothing reads any more."""
    kept = [st for st in stmts if st.out[0] != "o1" and all(nm != "t1" for nm, _ in st.ins)]
    live = {"o2"}
    out = []
    for st in reversed(kept):
        if st.out[0] in live:
            out.append(st)
            for nm, _ in st.ins:
                live.add(nm)
    out.reverse()
    return out


def residual_program(method, fac=-1.0, drivers=True, relayout=True):
    """Statements of o = drivers + fac*R_method(t): CCSD = StantonTerms; CCD = the same with
    t1 = 0 (cqcpy _D_D + _D_DD, kelvin/ft_cc_equations.py:48-62); LCCSD / LCCD = the linear
    terms (kelvin/ft_cc_equations.py:11-45)."""
    if method not in METHODS:
        raise Exception("Unrecognized method keyword")
    if method == "CCSD":
        return stanton(fac, drivers, relayout)
    if method == "CCD":
        body = _intermediates(relayout) + _parse_block(_RESIDUAL, fac)
    else:
        body = []
        for st in _parse_block(_LINEAR):
            if st.out[0] in ("o1", "o2"):
                st.coef *= fac
            body.append(st)
    st = _parse_block(_DRIVERS) if drivers else []
    st += body
    return st if has_singles(method) else _without_singles(st)


# ---------------------------------------------------------------------------
# Lambda map = vector-Jacobian product of StantonTerms
# ---------------------------------------------------------------------------
def _subst_in(op, table):
    """Replace input slots by (new slot, letter permutation, scale)."""
    from .plan import ROp
    coef = op.coef
    ins = []
    for slot, ls in op.ins:
        if slot in table:
            new, perm, scale = table[slot]
            ins.append((new, "".join(ls[p] for p in perm)))
            coef *= scale
        else:
            ins.append((slot, ls))
    return ROp(op.out, coef, ins, op.spin)


def lambda_rops(mode, fac=-1.0, method="CCSD", beta=None):
    """Resolved ops of cqcpy.cc_equations._Lambda_opt / _uccsd_Lambda_opt as the
    reference uses them (kelvin/ft_cc_equations.py:396-407, 437-456):

        lo1[i,a]     = -F.ov - sum <ji||ba> t[b,j] + fac * d<Lbar,R(t)>/d t1[a,i]
        lo2[i,j,a,b] = -I.oovv                     + fac * P(ij)P(ab) d<Lbar,R(t)>/d t2[a,b,i,j]

    with <L,R> = L1.R1 + 1/4 L2.R2 over spin orbitals.  Inputs: t1,t2 (amplitudes),
    l1,l2 (time-integrated Lambda, o..v.. order); outputs lo1, lo2.

    method: the same construction on the CCD / LCCSD / LCCD residual (general spin orbitals
    only, as in the reference) gives ccd_lambda_simple, lccsd_lambda_simple and
    lccd_lambda_simple (kelvin/ft_cc_equations.py:682-701, 313-340, 292-310); without singles
    there is no lo1, and LCCD scales its energy term by 1/beta as the reference does (:308)."""
    from .plan import ROp, adjoint, expand, is_integral_slot
    T = tensor_defs()
    fwd = expand(residual_program(method, 1.0, drivers=False, relayout=False), T, mode)
    s1 = has_singles(method)
    if mode == "g":
        outs1, outs2 = (["o1"] if s1 else []), ["o2"]
        seed = {"o2~": ("l2", (2, 3, 0, 1), 0.25)}
        if s1:
            seed["o1~"] = ("l1", (1, 0), 1.0)
        t1s, t2s = ([("t1", "lo1")] if s1 else []), [("t2", "lo2", True)]
    else:
        assert method == "CCSD", "the unrestricted loops know CCSD only (kelvin/cc_utils.py:84)"
        outs1, outs2 = ["o1.a", "o1.b"], ["o2.aa", "o2.ab", "o2.bb"]
        seed = {"o1.a~": ("l1.a", (1, 0), 1.0), "o1.b~": ("l1.b", (1, 0), 1.0),
                "o2.aa~": ("l2.aa", (2, 3, 0, 1), 0.25), "o2.bb~": ("l2.bb", (2, 3, 0, 1), 0.25),
                "o2.ab~": ("l2.ab", (2, 3, 0, 1), 1.0)}
        t1s = [("t1.a", "lo1.a"), ("t1.b", "lo1.b")]
        t2s = [("t2.aa", "lo2.aa", True), ("t2.ab", "lo2.ab", False), ("t2.bb", "lo2.bb", True)]
    outset = set(outs1 + outs2)
    inter = [op for op in fwd if op.out[0] not in outset]        # intermediates only
    bwd = adjoint(fwd, wrt=lambda s: not is_integral_slot(s), seeds=outset)
    bwd = [_subst_in(op, seed) for op in bwd]

    # energy term (kelvin/ft_cc_equations.py:403-407) written first so that the
    # outputs are initialised by it
    Tl = dict(T)
    from .plan import TDef
    Tl["lo1"] = TDef("lo1", "one", "out", True, "ov")
    Tl["lo2"] = TDef("lo2", "amp2", "out", True, "oovv")
    if s1:
        eterm = expand(_parse_block("""
            lo1[ia] += -1 F.ov[ia]
            lo1[ia] += -1 I.oovv[jiba] t1[bj]
            lo2[ijab] += -1 I.oovv[ijab]
        """), Tl, mode)
    else:
        eterm = expand(_parse_block("lo2[ijab] += -1 I.oovv[ijab]",
                                    1.0/beta if method == "LCCD" else 1.0), Tl, mode)

    # scatter the amplitude adjoints into the outputs
    tail = []

    def sp(slot, src_letters, new_letters):
        if mode == "g":
            return None
        suf = slot.split(".")[-1]
        spins = {"a": "aa", "b": "bb", "aa": "aaaa", "ab": "abab", "bb": "bbbb"}[suf]
        return dict(zip(new_letters, spins))
    for tslot, lslot in t1s:
        tail.append(ROp((lslot, "ia"), fac, [(tslot + "~", "ai")], sp(tslot, "ai", "ai")))
    for tslot, lslot, anti in t2s:
        spn = sp(tslot, "abij", "abij")
        tail.append(ROp((lslot, "ijab"), fac, [(tslot + "~", "abij")], spn))
        if anti:
            tail.append(ROp((lslot, "ijab"), -fac, [(tslot + "~", "baij")], spn))
            tail.append(ROp((lslot, "ijab"), -fac, [(tslot + "~", "abji")], spn))
            tail.append(ROp((lslot, "ijab"), fac, [(tslot + "~", "baji")], spn))
    return prune_unused(inter, eterm + bwd + tail), eterm + bwd + tail


def prune_unused(inter, rest):
    """Dead-code elimination: keep only the forward ops whose results the reverse
    sweep (or another kept forward op) actually reads -- e.g. the ring result
    itself is never needed, only its adjoint."""
    needed = set()
    for op in rest:
        for slot, _ in op.ins:
            needed.add(slot)
    kept = []
    for op in reversed(inter):
        if op.out[0] in needed:
            kept.append(op)
            for slot, _ in op.ins:
                needed.add(slot)
    kept.reverse()
    return kept


def lambda_guess_rops(mode, beta, ls_ts_fac):
    """L1 = F.ov/beta + ls_ts_fac * sum <ji||ba> t[b,j];  L2 = I.oovv/beta
    (kelvin/ft_cc_equations.py:502-526; the g guess uses ls_ts_fac = 1/beta, the u
    guess 1 -- quirk Q3)."""
    from .plan import TDef, expand
    T = tensor_defs()
    T["lo1"] = TDef("lo1", "one", "out", True, "ov")
    T["lo2"] = TDef("lo2", "amp2", "out", True, "oovv")
    st = _parse_block("""
        lo1[ia] += 1 F.ov[ia]
        lo2[ijab] += 1 I.oovv[ijab]
    """, 1.0/beta) + _parse_block("lo1[ia] += 1 I.oovv[jiba] t1[bj]", ls_ts_fac)
    return expand(st, T, mode)


# ---------------------------------------------------------------------------
# Response densities = derivatives of the Lagrangian w.r.t. the dressed
# integral blocks (kelvin/tests/test_ft_ccsd_rdm.py:12-21)
# ---------------------------------------------------------------------------
def rdm_rops(mode):
    """Forward intermediates + reverse sweep w.r.t. every integral block of

        phi = t1.F.ov + (t2/4 + t1 t1/2).I.oovv + <Lbar, (F.vo, I.vvoo) + StantonTerms(t)>.

    Adjoint slots are named '<block>~' (g: 'I.vvvv~', 'F.oo~'; u: 'Ia.vvvv~',
    'Iabab.ovvo~', 'Fb.vv~', ...) and hold d(phi)/d(block element) with all stored
    elements treated as independent.  The driver terms <Lbar,(F.vo, I.vvoo)> are
    not swept: their derivative is Lbar itself (kelvin/ft_cc_equations.py:713,737).
    Inputs t1,t2,l1,l2 as in ``lambda_rops``."""
    from .plan import ROp, adjoint, expand, is_integral_slot
    T = tensor_defs()
    fwd = expand(stanton(1.0, drivers=False, relayout=False), T, mode)
    if mode == "g":
        outset = {"o1", "o2"}
        seed = {"o1~": ("l1", (1, 0), 1.0), "o2~": ("l2", (2, 3, 0, 1), 0.25)}
        extra = [ROp(("F.ov~", "ia"), 1.0, [("t1", "ai")]),
                 ROp(("I.oovv~", "ijab"), 0.25, [("t2", "abij")]),
                 ROp(("I.oovv~", "ijab"), 0.5, [("t1", "ai"), ("t1", "bj")])]
    else:
        outset = {"o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb"}
        seed = {"o1.a~": ("l1.a", (1, 0), 1.0), "o1.b~": ("l1.b", (1, 0), 1.0),
                "o2.aa~": ("l2.aa", (2, 3, 0, 1), 0.25), "o2.bb~": ("l2.bb", (2, 3, 0, 1), 0.25),
                "o2.ab~": ("l2.ab", (2, 3, 0, 1), 1.0)}
        sa = dict(zip("ijab", "aaaa"))
        sb = dict(zip("ijab", "bbbb"))
        sab = dict(zip("ijab", "abab"))
        extra = [ROp(("Fa.ov~", "ia"), 1.0, [("t1.a", "ai")], sa),
                 ROp(("Fb.ov~", "ia"), 1.0, [("t1.b", "ai")], sb),
                 ROp(("Ia.oovv~", "ijab"), 0.25, [("t2.aa", "abij")], sa),
                 ROp(("Ia.oovv~", "ijab"), 0.5, [("t1.a", "ai"), ("t1.a", "bj")], sa),
                 ROp(("Ib.oovv~", "ijab"), 0.25, [("t2.bb", "abij")], sb),
                 ROp(("Ib.oovv~", "ijab"), 0.5, [("t1.b", "ai"), ("t1.b", "bj")], sb),
                 ROp(("Iabab.oovv~", "ijab"), 1.0, [("t2.ab", "abij")], sab),
                 ROp(("Iabab.oovv~", "ijab"), 1.0, [("t1.a", "ai"), ("t1.b", "bj")], sab)]
    dep = set()
    for op in fwd:
        if any(is_integral_slot(s) or s in dep for s, _ in op.ins):
            dep.add(op.out[0])
    # phi is linear in the integrals, so the sweep only ever reads amplitude-only
    # intermediates (tau, tauh, t2x): nothing that depends on F/I is evaluated.
    inter = [op for op in fwd if op.out[0] not in outset and op.out[0] not in dep]
    bwd = adjoint(fwd, wrt=lambda s: is_integral_slot(s) or s in dep, seeds=outset)
    bwd = [_subst_in(op, seed) for op in bwd]
    for op in bwd:
        for slot, _ in op.ins:
            assert not is_integral_slot(slot) and slot not in dep, op
    return inter, extra + bwd


# P-block <- integral block it is conjugate to (kelvin/ft_cc_equations.py:737-751)
RDM2_BLOCKS = (("cdab", "vvvv"), ("ciab", "vvvo"), ("bcai", "vovv"), ("ijab", "vvoo"),
               ("bjai", "vovo"), ("abij", "oovv"), ("jkai", "vooo"), ("kaij", "ooov"),
               ("klij", "oooo"))
# spin patterns (over the P index order) of the unrestricted tuples
# (kelvin/ft_cc_equations.py:919-927)
RDM2_USPINS = {
    "cdab": ("aaaa", "bbbb", "abab"),
    "ciab": ("aaaa", "bbbb", "abab", "baba"),
    "bcai": ("aaaa", "bbbb", "abab", "baba"),
    "ijab": ("aaaa", "bbbb", "abab"),
    "bjai": ("aaaa", "bbbb", "abab", "abba", "baab", "baba"),
    "abij": ("aaaa", "bbbb", "abab"),
    "jkai": ("aaaa", "bbbb", "abab", "baba"),
    "kaij": ("aaaa", "bbbb", "abab", "baba"),
    "klij": ("aaaa", "bbbb", "abab"),
}


def rdm2_assembly_rops(mode, skip=("vvoo",)):
    """Unary ops building the reference's P blocks from the summed adjoints:
    P_B[r,s,p,q] = c*A(dphi/dI_B)[p,q,r,s], A = antisymmetriser over the index
    pairs of B that share a space, c = 2 per antisymmetrised pair.  In the u
    form the mixed-spin blocks are single (signed, permuted) adjoint leaves.
    Returns (ops, list of (output slot, block name, spin pattern))."""
    from .plan import ROp, TDef, resolve_u
    ops, outs = [], []

    def swap(ls, i, j):
        ls = list(ls)
        ls[i], ls[j] = ls[j], ls[i]
        return "".join(ls)
    for pname, pat in RDM2_BLOCKS:
        if pat in skip:
            continue
        a1 = pat[0] == pat[1]
        a2 = pat[2] == pat[3]
        pool = {"v": iter("abcd"), "o": iter("ijkl")}
        L = "".join(next(pool[c]) for c in pat)            # letters over the integral order
        dstl = L[2] + L[3] + L[0] + L[1]                   # P index order
        terms = [(L, 1.0)]
        if a1:
            terms += [(swap(L, 0, 1), -1.0)]
        if a2:
            terms += [(swap(ls, 2, 3), -sg) for ls, sg in list(terms)]
        if mode == "g":
            for ls, sg in terms:
                ops.append(ROp(("P" + pname, dstl), sg, [("I.%s~" % pat, ls)]))
            outs.append(("P" + pname, pname, None))
            continue
        td = TDef("I." + pat, "int2", "in", False, pat)
        for sp in RDM2_USPINS[pname]:
            ispin = sp[2] + sp[3] + sp[0] + sp[1]          # spins over the integral order
            dst = ("P%s.%s" % (pname, sp), dstl)
            spin = dict(zip(L, ispin))
            if ispin in ("aaaa", "bbbb"):
                leaf = ("Ia.%s~" if ispin == "aaaa" else "Ib.%s~") % pat
                for ls, sg in terms:
                    ops.append(ROp(dst, sg, [(leaf, ls)], spin))
            else:
                slot, ls, sg = resolve_u(td, L, list(ispin))
                ops.append(ROp(dst, sg, [(slot + "~", ls)], spin))
            outs.append((dst[0], pname, sp))
    return ops, outs

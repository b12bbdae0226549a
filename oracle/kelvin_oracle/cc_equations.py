"""Restated ``cqcpy.cc_equations`` surface used by the FT-CCSD path.

TEST INFRASTRUCTURE (see package docstring).  ``cqcpy`` is an unpinned,
un-vendored dependency of the reference (setup.cfg:8,
.github/workflows/clone_deps.sh:2); its source is not available, so the
arithmetic is restated from the published equations (Stanton, Gauss, Watts,
Bartlett, JCP 94, 4334 (1991)) in the block/index conventions visible in
kelvin/lambda_stanton.py:7-34, and pinned through the reference's golden
numbers (tests/test_oracle_golden.py).

Conventions: t1[a,i], t2[a,b,i,j]; I.xyzw[p,q,r,s] = <pq||rs>;
F.oo / F.vv have the orbital energies removed (kelvin/cc_utils.py:578).
The Lambda map and the RDM blocks are *defined* as derivatives (SURVEY.md
A.3/A.4, kelvin/tests/test_ft_lambda.py:212-285,
kelvin/tests/test_ft_ccsd_rdm.py:12-21) and evaluated with torch autograd in
float64 on the CPU.
"""
import numpy
import torch

from . import cqc


def _np_einsum(*a):
    return numpy.einsum(*a, optimize=True)


def stanton_terms(F, I, t1, t2, ein=_np_einsum, hints=False):
    """R1[a,i], R2[a,b,i,j] of SURVEY.md A.2 (no drivers, no sign).

    Called by the reference as cqcpy.cc_equations._Stanton at
    kelvin/ft_cc_equations.py:106 (which adds fac*R to T1new/T2new)."""
    kw = {"keep": True} if hints else {}    # spin-blocked back end: canonical blocks only
    tt = ein('ai,bj->abij', t1, t1)
    tt = tt - tt.transpose(1, 0, 2, 3) if not torch.is_tensor(tt) else tt - tt.permute(1, 0, 2, 3)
    tau_h = t2 + 0.5 * tt
    tau = t2 + tt

    Fvv = (F.vv - 0.5 * ein('me,am->ae', F.ov, t1)
           + ein('amef,fm->ae', I.vovv, t1)
           - 0.5 * ein('mnef,afmn->ae', I.oovv, tau_h))
    Foo = (F.oo + 0.5 * ein('me,ei->mi', F.ov, t1)
           + ein('mnie,en->mi', I.ooov, t1)
           + 0.5 * ein('mnef,efin->mi', I.oovv, tau_h))
    Fov = F.ov + ein('mnef,fn->me', I.oovv, t1)

    tmp = ein('mnie,ej->mnij', I.ooov, t1)
    Woooo = I.oooo + tmp - _swap(tmp, 2, 3) + 0.25 * ein('mnef,efij->mnij', I.oovv, tau, **kw)
    tmp = ein('amef,bm->abef', I.vovv, t1)
    Wvvvv = I.vvvv - tmp + _swap(tmp, 0, 1) + 0.25 * ein('mnef,abmn->abef', I.oovv, tau, **kw)
    # Wovvo[m,b,e,j]
    Wovvo = (-_perm(I.vovo, (1, 0, 2, 3))
             - ein('bmef,fj->mbej', I.vovv, t1)
             + ein('mnje,bn->mbej', I.ooov, t1)
             - ein('mnef,fbjn->mbej', I.oovv, 0.5 * t2 + ein('fj,bn->fbjn', t1, t1)))

    R1 = (ein('ae,ei->ai', Fvv, t1) - ein('mi,am->ai', Foo, t1)
          + ein('me,aeim->ai', Fov, t2)
          - ein('anfi,fn->ai', I.vovo, t1)
          + 0.5 * ein('amef,efim->ai', I.vovv, t2)
          - 0.5 * ein('mnie,aemn->ai', I.ooov, t2))

    Xvv = Fvv - 0.5 * ein('bm,me->be', t1, Fov)
    Xoo = Foo + 0.5 * ein('ej,me->mj', t1, Fov)
    tmp = ein('aeij,be->abij', t2, Xvv)
    R2 = tmp - _swap(tmp, 0, 1)
    tmp = ein('abim,mj->abij', t2, Xoo)
    R2 = R2 - (tmp - _swap(tmp, 2, 3))
    R2 = R2 + 0.5 * ein('abmn,mnij->abij', tau, Woooo, **kw)
    R2 = R2 + 0.5 * ein('efij,abef->abij', tau, Wvvvv, **kw)
    tmp = ein('aeim,mbej->abij', t2, Wovvo) \
        + ein('ei,am,bmej->abij', t1, t1, I.vovo)
    tmp = tmp - _swap(tmp, 0, 1)
    R2 = R2 + tmp - _swap(tmp, 2, 3)
    tmp = ein('ei,abej->abij', t1, I.vvvo)
    R2 = R2 + tmp - _swap(tmp, 2, 3)
    tmp = ein('am,bmij->abij', t1, I.vooo)
    R2 = R2 + tmp - _swap(tmp, 0, 1)
    return R1, R2


def _perm(x, p):
    return x.permute(*p) if torch.is_tensor(x) else x.transpose(*p)


def _swap(x, i, j):
    p = list(range(x.ndim))
    p[i], p[j] = p[j], p[i]
    return _perm(x, p)


# ---------------------------------------------------------------------------
# in-place cqcpy-style entry points (g form)
# ---------------------------------------------------------------------------
def _Stanton(T1, T2, F, I, T1old, T2old, fac=1.0):
    """T1 += fac*R1, T2 += fac*R2 (kelvin/ft_cc_equations.py:106)."""
    R1, R2 = stanton_terms(F, I, T1old, T2old)
    T1 += fac * R1
    T2 += fac * R2


def _LS_TS(L1, I, T1old, fac=1.0):
    """L1[i,a] += fac * sum_bj <ji||ba> t[b,j] (kelvin/ft_cc_equations.py:407)."""
    L1 += fac * numpy.einsum('jiba,bj->ia', I.oovv, T1old, optimize=True)


class _TB(object):
    pass


def _as_torch_blocks(F, I, requires_grad=False):
    Ft, It = _TB(), _TB()
    for nm in ("oo", "ov", "vo", "vv"):
        t = torch.tensor(numpy.asarray(getattr(F, nm)), dtype=torch.float64)
        t.requires_grad_(requires_grad)
        setattr(Ft, nm, t)
    for nm in cqc.two_e_blocks.names:
        t = torch.tensor(numpy.asarray(getattr(I, nm)), dtype=torch.float64)
        t.requires_grad_(requires_grad)
        setattr(It, nm, t)
    return Ft, It


def _pair(L1, L2, R1, R2):
    """<L,R> = sum L1[i,a]R1[a,i] + 1/4 sum L2[i,j,a,b]R2[a,b,i,j]."""
    return (L1 * R1.permute(1, 0)).sum() + 0.25 * (L2 * R2.permute(2, 3, 0, 1)).sum()


def lambda_terms(F, I, L1old, L2old, T1old, T2old):
    """(dL1[i,a], dL2[i,j,a,b]) of SURVEY.md A.3: the VJP of stanton_terms
    with cotangent (L1old, L2old); dL2 carries the P(ij)P(ab) projector."""
    Ft, It = _as_torch_blocks(F, I)
    t1 = torch.tensor(T1old, dtype=torch.float64, requires_grad=True)
    t2 = torch.tensor(T2old, dtype=torch.float64, requires_grad=True)
    l1 = torch.tensor(L1old, dtype=torch.float64)
    l2 = torch.tensor(L2old, dtype=torch.float64)
    R1, R2 = stanton_terms(Ft, It, t1, t2, ein=torch.einsum)
    s = _pair(l1, l2, R1, R2)
    g1, g2 = torch.autograd.grad(s, (t1, t2))
    g2 = g2 - g2.permute(1, 0, 2, 3)
    g2 = g2 - g2.permute(0, 1, 3, 2)
    return g1.permute(1, 0).numpy().copy(), g2.permute(2, 3, 0, 1).numpy().copy()


def _Lambda_opt(L1, L2, F, I, L1old, L2old, T1old, T2old, fac=1.0):
    """In-place L += fac*J(T)^T.Lold (kelvin/ft_cc_equations.py:399)."""
    d1, d2 = lambda_terms(F, I, L1old, L2old, T1old, T2old)
    L1 += fac * d1
    L2 += fac * d2


# ---------------------------------------------------------------------------
# Term classes of cqcpy.cc_equations used by the CCD / LCCSD / LCCD switches
# (kelvin/ft_cc_equations.py:11-62, 292-340, 682-701): "_X_Y..." = the part of the X (S/D)
# residual that is homogeneous of degree one in each amplitude listed after the underscore.
# They are extracted from the pinned polynomial stanton_terms (degree <= 4 in T, no
# constant term) by exact finite combinations, not restated term by term:
#   linear part   L(t) = [8(R(t) - R(-t)) - (R(2t) - R(-2t))]/12
# (odd differences cancel the degree-2/4 parts, the 8:1 weights cancel the cubic one).
# No reference golden pins these switches at finite temperature; they inherit the pin of the
# full polynomial.
# ---------------------------------------------------------------------------
def _linear_part(fn, x):
    r1, rm1 = fn(x), fn(-x)
    r2, rm2 = fn(2.0*x), fn(-2.0*x)
    return tuple((8.0*(a - b) - (c - d))/12.0 for a, b, c, d in zip(r1, rm1, r2, rm2))


def _zeros1(F):
    no, nv = F.ov.shape
    return numpy.zeros((nv, no))


def _zeros2(F):
    no, nv = F.ov.shape
    return numpy.zeros((nv, nv, no, no))


def _S_S(T1, F, I, T1old, fac=1.0):
    T1 += fac*_linear_part(lambda x: stanton_terms(F, I, x, _zeros2(F)), numpy.asarray(T1old))[0]


def _D_S(T2, F, I, T1old, fac=1.0):
    T2 += fac*_linear_part(lambda x: stanton_terms(F, I, x, _zeros2(F)), numpy.asarray(T1old))[1]


def _S_D(T1, F, I, T2old, fac=1.0):
    T1 += fac*_linear_part(lambda x: stanton_terms(F, I, _zeros1(F), x), numpy.asarray(T2old))[0]


def _D_D(T2, F, I, T2old, fac=1.0):
    T2 += fac*_linear_part(lambda x: stanton_terms(F, I, _zeros1(F), x), numpy.asarray(T2old))[1]


def _D_DD(T2, F, I, T2old, fac=1.0):
    """Quadratic part of R2(0, t2) (R2 is of degree two in t2 when t1 = 0)."""
    t2 = numpy.asarray(T2old)
    full = stanton_terms(F, I, _zeros1(F), t2)[1]
    lin = _linear_part(lambda x: stanton_terms(F, I, _zeros1(F), x), t2)[1]
    T2 += fac*(full - lin)


def _zl1(F):
    return numpy.zeros(F.ov.shape)


def _zl2(F):
    no, nv = F.ov.shape
    return numpy.zeros((no, no, nv, nv))


def _LS_LS(L1, F, I, L1old, fac=1.0):
    L1 += fac*lambda_terms(F, I, L1old, _zl2(F), _zeros1(F), _zeros2(F))[0]


def _LD_LS(L2, F, I, L1old, fac=1.0):
    L2 += fac*lambda_terms(F, I, L1old, _zl2(F), _zeros1(F), _zeros2(F))[1]


def _LS_LD(L1, F, I, L2old, fac=1.0):
    L1 += fac*lambda_terms(F, I, _zl1(F), L2old, _zeros1(F), _zeros2(F))[0]


def _LD_LD(L2, F, I, L2old, fac=1.0):
    L2 += fac*lambda_terms(F, I, _zl1(F), L2old, _zeros1(F), _zeros2(F))[1]


class _FI0(object):
    pass


def _LD_LDTD(L2, I, L2old, T2old, fac=1.0):
    """Part of the doubles Lambda map linear in T2 (no F dependence: F drops out of the
    difference)."""
    F = _FI0()
    no, nv = I.oovv.shape[0], I.oovv.shape[2]
    F.oo, F.ov = numpy.zeros((no, no)), numpy.zeros((no, nv))
    F.vo, F.vv = numpy.zeros((nv, no)), numpy.zeros((nv, nv))
    a = lambda_terms(F, I, _zl1(F), L2old, _zeros1(F), T2old)[1]
    b = lambda_terms(F, I, _zl1(F), L2old, _zeros1(F), _zeros2(F))[1]
    L2 += fac*(a - b)


# ---------------------------------------------------------------------------
# RDM blocks -- derivative of phi w.r.t. zero-valued F/I
# ---------------------------------------------------------------------------
_rdm_cache = {}


def _rdm_all(T1, T2, L1, L2):
    key = (id(T1), id(T2), id(L1), id(L2), T1.shape, float(T1.ravel()[0]), float(L2.ravel()[-1]))
    if key in _rdm_cache:
        return _rdm_cache[key]
    nv, no = T1.shape
    z = cqc.one_e_blocks(numpy.zeros((no, no)), numpy.zeros((no, nv)),
                         numpy.zeros((nv, no)), numpy.zeros((nv, nv)))
    dims = {"o": no, "v": nv}
    zi = cqc.two_e_blocks(**{p: numpy.zeros(tuple(dims[c] for c in p)) for p in cqc.two_e_blocks.names})
    Ft, It = _as_torch_blocks(z, zi, requires_grad=True)
    t1 = torch.tensor(T1, dtype=torch.float64)
    t2 = torch.tensor(T2, dtype=torch.float64)
    l1 = torch.tensor(L1, dtype=torch.float64)
    l2 = torch.tensor(L2, dtype=torch.float64)
    R1, R2 = stanton_terms(Ft, It, t1, t2, ein=torch.einsum)
    tt = torch.einsum('ai,bj->abij', t1, t1)
    phi = (torch.einsum('ai,ia->', t1, Ft.ov)
           + torch.einsum('abij,ijab->', 0.25 * t2 + 0.5 * tt, It.oovv)
           + _pair(l1, l2, Ft.vo + R1, It.vvoo + R2))
    leaves = [Ft.oo, Ft.ov, Ft.vo, Ft.vv] + [getattr(It, p) for p in cqc.two_e_blocks.names]
    grads = torch.autograd.grad(phi, leaves, allow_unused=True)
    G = {}
    for nm, gr in zip(["oo", "ov", "vo", "vv"] + list(cqc.two_e_blocks.names), grads):
        G[nm] = gr.numpy()

    def asym(x, first, second):
        if first:
            x = 0.5 * (x - x.transpose(1, 0, 2, 3))
        if second:
            x = 0.5 * (x - x.transpose(0, 1, 3, 2))
        return x
    out = {
        "ba": G["vv"].T.copy(), "ji": G["oo"].T.copy(), "ai": G["ov"].T.copy(),
        "cdab": 4.0 * asym(G["vvvv"], True, True).transpose(2, 3, 0, 1),
        "ciab": 2.0 * asym(G["vvvo"], True, False).transpose(2, 3, 0, 1),
        "bcai": 2.0 * asym(G["vovv"], False, True).transpose(2, 3, 0, 1),
        "bjai": G["vovo"].transpose(2, 3, 0, 1),
        "abij": 4.0 * asym(G["oovv"], True, True).transpose(2, 3, 0, 1),
        "jkai": 2.0 * asym(G["vooo"], False, True).transpose(2, 3, 0, 1),
        "kaij": 2.0 * asym(G["ooov"], True, False).transpose(2, 3, 0, 1),
        "klij": 4.0 * asym(G["oooo"], True, True).transpose(2, 3, 0, 1),
    }
    out = {k: numpy.ascontiguousarray(v) for k, v in out.items()}
    _rdm_cache.clear()
    _rdm_cache[key] = out
    return out


def _mk(name):
    def f(T1, T2, L1, L2):
        return _rdm_all(T1, T2, L1, L2)[name].copy()
    f.__name__ = name
    return f


ccsd_1rdm_ba_opt = _mk("ba")
ccsd_1rdm_ji_opt = _mk("ji")
ccsd_1rdm_ai_opt = _mk("ai")
ccsd_2rdm_cdab_opt = _mk("cdab")
ccsd_2rdm_ciab_opt = _mk("ciab")
ccsd_2rdm_bcai_opt = _mk("bcai")
ccsd_2rdm_bjai_opt = _mk("bjai")
ccsd_2rdm_abij_opt = _mk("abij")
ccsd_2rdm_jkai_opt = _mk("jkai")
ccsd_2rdm_kaij_opt = _mk("kaij")
ccsd_2rdm_klij_opt = _mk("klij")
ccsd_1rdm_ba, ccsd_1rdm_ji, ccsd_1rdm_ai = ccsd_1rdm_ba_opt, ccsd_1rdm_ji_opt, ccsd_1rdm_ai_opt
ccsd_2rdm_cdab, ccsd_2rdm_ciab, ccsd_2rdm_bcai = ccsd_2rdm_cdab_opt, ccsd_2rdm_ciab_opt, ccsd_2rdm_bcai_opt
ccsd_2rdm_bjai, ccsd_2rdm_abij, ccsd_2rdm_jkai = ccsd_2rdm_bjai_opt, ccsd_2rdm_abij_opt, ccsd_2rdm_jkai_opt
ccsd_2rdm_kaij, ccsd_2rdm_klij = ccsd_2rdm_kaij_opt, ccsd_2rdm_klij_opt


# ---------------------------------------------------------------------------
# unrestricted (Sz-blocked) entry points: exact embedding into the g form
# (the reference's own tests assert u == g block by block:
#  kelvin/tests/test_ft_cc_ampl.py:41-110, test_ft_lambda_equations.py:42-139)
# ---------------------------------------------------------------------------
def _embed(Fa, Fb, Ia, Ib, Iabab):
    return cqc.F_to_spin(Fa, Fb), cqc.I_to_spin(Ia, Ib, Iabab)


_emb_cache = {}


def _embed_cached(Fa, Fb, Ia, Ib, Iabab):
    key = (id(Fa), id(Fb), id(Ia), id(Ib), id(Iabab), id(Ia.vvvv), id(Iabab.vvvv))
    if key not in _emb_cache:
        _emb_cache.clear()
        _emb_cache[key] = _embed(Fa, Fb, Ia, Ib, Iabab)
    return _emb_cache[key]


def _t_to_spin(T1a, T1b, T2aa, T2ab, T2bb):
    nva, noa = T1a.shape
    nvb, nob = T1b.shape
    return (cqc.T1_to_spin(T1a, T1b, nva, noa, nvb, nob),
            cqc.T2_to_spin(T2aa, T2ab, T2bb, nva, noa, nvb, nob))


def u_stanton_terms(Fa, Fb, Ia, Ib, Iabab, T1olds, T2olds):
    T1a, T1b = T1olds
    T2aa, T2ab, T2bb = T2olds
    nva, noa = T1a.shape
    F, I = _embed_cached(Fa, Fb, Ia, Ib, Iabab)
    t1, t2 = _t_to_spin(T1a, T1b, T2aa, T2ab, T2bb)
    R1, R2 = stanton_terms(F, I, t1, t2)
    return (R1[:nva, :noa], R1[nva:, noa:],
            R2[:nva, :nva, :noa, :noa], R2[:nva, nva:, :noa, noa:], R2[nva:, nva:, noa:, noa:])


def _u_Stanton(T1a, T1b, T2aa, T2ab, T2bb, Fa, Fb, Ia, Ib, Iabab, T1olds, T2olds, fac=1.0):
    """kelvin/ft_cc_equations.py:153-155."""
    r = u_stanton_terms(Fa, Fb, Ia, Ib, Iabab, T1olds, T2olds)
    for dst, src in zip((T1a, T1b, T2aa, T2ab, T2bb), r):
        dst += fac * src


def _u_LS_TS(L1a, L1b, Ia, Ib, Iabab, T1a, T1b, fac=1.0):
    """kelvin/ft_cc_equations.py:456,524."""
    L1a += fac * (numpy.einsum('jiba,bj->ia', Ia.oovv, T1a)
                  + numpy.einsum('iJaB,BJ->ia', Iabab.oovv, T1b))
    L1b += fac * (numpy.einsum('jiba,bj->ia', Ib.oovv, T1b)
                  + numpy.einsum('jIbA,bj->IA', Iabab.oovv, T1a))


def _l_to_spin(L1a, L1b, L2aa, L2ab, L2bb):
    noa, nva = L1a.shape
    nob, nvb = L1b.shape
    return (cqc.T1_to_spin(L1a, L1b, noa, nva, nob, nvb),
            cqc.T2_to_spin(L2aa, L2ab, L2bb, noa, nva, nob, nvb))


def _uccsd_Lambda_opt(L1a, L1b, L2aa, L2ab, L2bb, Fa, Fb, Ia, Ib, Iabab,
                      L1olds, L2olds, T1olds, T2olds, fac=1.0):
    """kelvin/ft_cc_equations.py:442-445."""
    noa, nva = L1a.shape
    F, I = _embed_cached(Fa, Fb, Ia, Ib, Iabab)
    t1, t2 = _t_to_spin(T1olds[0], T1olds[1], T2olds[0], T2olds[1], T2olds[2])
    l1, l2 = _l_to_spin(L1olds[0], L1olds[1], L2olds[0], L2olds[1], L2olds[2])
    d1, d2 = lambda_terms(F, I, l1, l2, t1, t2)
    L1a += fac * d1[:noa, :nva]
    L1b += fac * d1[noa:, nva:]
    L2aa += fac * d2[:noa, :noa, :nva, :nva]
    L2ab += fac * d2[:noa, noa:, :nva, nva:]
    L2bb += fac * d2[noa:, noa:, nva:, nva:]


_U2 = {  # spin-block slices of each g 2-RDM block type, in the tuple order of
         # kelvin/ft_cc_equations.py:919-927 (0 = alpha range, 1 = beta range)
    "cdab": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1)),
    "ciab": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1), (1, 0, 1, 0)),
    "bcai": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1), (1, 0, 1, 0)),
    "bjai": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1), (0, 1, 1, 0), (1, 0, 0, 1), (1, 0, 1, 0)),
    "abij": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1)),
    "jkai": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1), (1, 0, 1, 0)),
    "kaij": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1), (1, 0, 1, 0)),
    "klij": ((0, 0, 0, 0), (1, 1, 1, 1), (0, 1, 0, 1)),
}


def _u_rdm(name, T1a, T1b, T2aa, T2ab, T2bb, L1a, L1b, L2aa, L2ab, L2bb):
    nva, noa = T1a.shape
    t1, t2 = _t_to_spin(T1a, T1b, T2aa, T2ab, T2bb)
    l1, l2 = _l_to_spin(L1a, L1b, L2aa, L2ab, L2bb)
    key = ("u", T1a.shape, T1b.shape, float(T1a.ravel()[0]), float(T2ab.ravel()[-1]),
           float(L1a.ravel()[0]), float(L2ab.ravel()[-1]), float(L2bb.ravel()[1]))
    if _rdm_cache.get("ukey") != key:
        _rdm_cache["ukey"] = key
        _rdm_cache["uval"] = dict(_rdm_all(t1, t2, l1, l2))
    P = _rdm_cache["uval"][name]
    # alpha/beta ranges of every axis: virtual letters split at nva, occupied letters at noa
    cut = [nva if c in "abcd" else noa for c in name]
    sl = [(slice(0, n), slice(n, None)) for n in cut]
    if P.ndim == 2:
        return P[sl[0][0], sl[1][0]].copy(), P[sl[0][1], sl[1][1]].copy()
    return tuple(P[sl[0][s[0]], sl[1][s[1]], sl[2][s[2]], sl[3][s[3]]].copy() for s in _U2[name])


def _mku(name):
    def f(*a):
        return _u_rdm(name, *a)
    return f


uccsd_1rdm_ba, uccsd_1rdm_ji, uccsd_1rdm_ai = _mku("ba"), _mku("ji"), _mku("ai")
uccsd_2rdm_cdab, uccsd_2rdm_ciab, uccsd_2rdm_bcai = _mku("cdab"), _mku("ciab"), _mku("bcai")
uccsd_2rdm_bjai, uccsd_2rdm_abij, uccsd_2rdm_jkai = _mku("bjai"), _mku("abij"), _mku("jkai")
uccsd_2rdm_kaij, uccsd_2rdm_klij = _mku("kaij"), _mku("klij")

"""NumPy restatement of the reference's FT-CCSD drivers (TEST INFRASTRUCTURE).

Follows kelvin/quadrature.py, kelvin/ft_cc_energy.py, kelvin/ft_cc_equations.py
and kelvin/cc_utils.py function by function (citations in each docstring).
Validated against the unmodified reference drivers run on the same restated
``cqcpy`` layer (tests/golden/make_golden.py -> tests/golden/*.npz) and the
reference's published golden numbers (tests/test_oracle_golden.py).
"""
import numpy

from . import cqc
from . import cc_equations as cqe

einsum = numpy.einsum


# --- kelvin/quadrature.py -------------------------------------------------
def simpson_G(ng, delta):
    """kelvin/quadrature.py:32-42."""
    G = numpy.zeros((ng, ng))
    G[1, 0] = G[1, 1] = 0.5*delta
    for y in range(2, ng):
        G[y] = G[y - 2]
        G[y, y - 2] += delta/3.0
        G[y, y - 1] += 4.0*delta/3.0
        G[y, y] += delta/3.0
    return G


def simpson_g(ng, delta):
    """kelvin/quadrature.py:45-60."""
    g = numpy.zeros(ng)
    if ng % 2 == 0:
        g[0] += 0.5*delta
        g[1] += 0.5*delta
        rng = range(3, ng, 2)
    else:
        rng = range(2, ng, 2)
    for y in rng:
        g[y - 2] += delta/3.0
        g[y - 1] += 4.0*delta/3.0
        g[y] += delta/3.0
    return g


def simpsons(ng, beta):
    """kelvin/quadrature.py:102-107 ('lin' rule of ft_quad, :215-217)."""
    delta = beta/(ng - 1.0)
    ti = numpy.asarray([float(i)*delta for i in range(ng)])
    return ti, simpson_g(ng, delta), simpson_G(ng, delta)


def d_simpsons(ng, beta):
    """kelvin/quadrature.py:108-113."""
    delta = beta/(ng - 1.0)
    ddelta = delta/beta
    return simpson_g(ng, ddelta), simpson_G(ng, ddelta)


def int_tbar(ng, tbar, ti, D, G):
    """kelvin/quadrature.py:292-317.  The reference allocates ``dt`` once, so
    the weight is exp(D*(ti[x]-ti[y])) for x<y and 1 for x>=y (quirk Q4)."""
    out = numpy.zeros(tbar.shape, dtype=tbar.dtype)
    for y in range(ng):
        dt = numpy.zeros(ng)
        dt[:y] = ti[:y] - ti[y]
        w = numpy.exp(dt.reshape((ng,) + (1,)*D.ndim)*D[None])
        out[y] = numpy.tensordot(G[y], w*tbar, axes=(0, 0))
    return out


def int_L(ng, Lold, ti, D, g, G):
    """kelvin/quadrature.py:320-345 (D is indexed v..o.., L is o..v..)."""
    r = D.ndim
    perm = tuple(range(r//2, r)) + tuple(range(r//2))
    Dt = D.transpose(perm)
    out = numpy.zeros(Lold.shape, dtype=Lold.dtype)
    for s in range(ng):
        dt = numpy.zeros(ng)
        dt[s:] = ti[s] - ti[s:]
        w = numpy.exp(dt.reshape((ng,) + (1,)*r)*Dt[None])
        out[s] = numpy.tensordot(g*G[:, s], w*Lold, axes=(0, 0))/g[s]
    return out


# --- kelvin/ft_cc_energy.py -----------------------------------------------
def ft_cc_energy(T1, T2, f, eri, g, beta, Qterm=True):
    """kelvin/ft_cc_energy.py:7-32."""
    t2 = 0.25*T2
    if Qterm:
        t2 = t2 + 0.5*einsum('yai,ybj->yabij', T1, T1)
    t1n = einsum('y,yai->ai', g, T1)
    t2n = einsum('y,yabij->abij', g, t2)
    return (einsum('ai,ia->', t1n, f) + einsum('abij,ijab->', t2n, eri))/beta


def ft_ucc_energy(T1a, T1b, T2aa, T2ab, T2bb, fa, fb, Ia, Ib, Iabab, g, beta, Qterm=True):
    """kelvin/ft_cc_energy.py:35-72 (including the Qterm=False T2aa quirk, :58)."""
    if Qterm:
        taa = 0.25*T2aa + 0.5*einsum('yai,ybj->yabij', T1a, T1a)
        tbb = 0.25*T2bb + 0.5*einsum('yai,ybj->yabij', T1b, T1b)
        tab = T2ab + einsum('yai,ybj->yabij', T1a, T1b)
    else:
        taa, tab, tbb = 0.25*T2aa, T2ab, 0.25*T2aa
    E = einsum('ai,ia->', einsum('y,yai->ai', g, T1a), fa)
    E += einsum('ai,ia->', einsum('y,yai->ai', g, T1b), fb)
    E += einsum('abij,ijab->', einsum('y,yabij->abij', g, taa), Ia)
    E += einsum('abij,ijab->', einsum('y,yabij->abij', g, tab), Iabab)
    E += einsum('abij,ijab->', einsum('y,yabij->abij', g, tbb), Ib)
    return E/beta


# --- kelvin/cc_utils.py: dressing -----------------------------------------
def _dressF(f, e, so, sv):
    f = f - numpy.diag(e)
    s = {"o": so, "v": sv}
    return cqc.one_e_blocks(*[einsum('pq,p,q->pq', f, s[p[0]], s[p[1]])
                              for p in ("oo", "ov", "vo", "vv")])


def _dress4(V, pat, s):
    return einsum('pqrs,p,q,r,s->pqrs', V, s[0][pat[0]], s[1][pat[1]], s[2][pat[2]], s[3][pat[3]])


def ft_integrals(sys, en, beta, mu):
    """kelvin/cc_utils.py:569-602."""
    s = {"o": numpy.sqrt(cqc.ff(beta, en, mu)), "v": numpy.sqrt(cqc.ffv(beta, en, mu))}
    F = _dressF(sys.g_fock_tot(), en, s["o"], s["v"])
    eri = sys.g_aint_tot()
    I = cqc.two_e_blocks(**{p: _dress4(eri, p, (s, s, s, s)) for p in cqc.two_e_blocks.names})
    return F, I


def uft_integrals(sys, ea, eb, beta, mu):
    """kelvin/cc_utils.py:696-777."""
    sa = {"o": numpy.sqrt(cqc.ff(beta, ea, mu)), "v": numpy.sqrt(cqc.ffv(beta, ea, mu))}
    sb = {"o": numpy.sqrt(cqc.ff(beta, eb, mu)), "v": numpy.sqrt(cqc.ffv(beta, eb, mu))}
    fa, fb = sys.u_fock_tot()
    Fa = _dressF(fa, ea, sa["o"], sa["v"])
    Fb = _dressF(fb, eb, sb["o"], sb["v"])
    eriA, eriB, eriAB = sys.u_aint_tot()
    Ia = cqc.two_e_blocks(**{p: _dress4(eriA, p, (sa,)*4) for p in cqc.two_e_blocks.names})
    Ib = cqc.two_e_blocks(**{p: _dress4(eriB, p, (sb,)*4) for p in cqc.two_e_blocks.names})
    Iabab = cqc.two_e_blocks_full(**{p: _dress4(eriAB, p, (sa, sb, sa, sb))
                                     for p in cqc.two_e_blocks_full.names})
    return Fa, Fb, Ia, Ib, Iabab


# --- kelvin/ft_cc_equations.py --------------------------------------------
def ccsd_stanton(F, I, T1old, T2old, D1, D2, ti, ng, G):
    """kelvin/ft_cc_equations.py:96-113."""
    T1 = numpy.stack([-F.vo]*ng)
    T2 = numpy.stack([-I.vvoo]*ng)
    for y in range(ng):
        cqe._Stanton(T1[y], T2[y], F, I, T1old[y], T2old[y], fac=-1.0)
    return int_tbar(ng, T1, ti, D1, G), int_tbar(ng, T2, ti, D2, G)


def uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb, ng):
    o1a = numpy.stack([-Fa.vo]*ng)
    o1b = numpy.stack([-Fb.vo]*ng)
    o2aa = numpy.stack([-Ia.vvoo]*ng)
    o2ab = numpy.stack([-Iabab.vvoo]*ng)
    o2bb = numpy.stack([-Ib.vvoo]*ng)
    for y in range(ng):
        cqe._u_Stanton(o1a[y], o1b[y], o2aa[y], o2ab[y], o2bb[y], Fa, Fb, Ia, Ib, Iabab,
                       (T1a[y], T1b[y]), (T2aa[y], T2ab[y], T2bb[y]), fac=-1.0)
    return o1a, o1b, o2aa, o2ab, o2bb


def uccsd_stanton(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb,
                  D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G):
    """kelvin/ft_cc_equations.py:130-164."""
    o = uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb, ng)
    return ((int_tbar(ng, o[0], ti, D1a, G), int_tbar(ng, o[1], ti, D1b, G)),
            (int_tbar(ng, o[2], ti, D2aa, G), int_tbar(ng, o[3], ti, D2ab, G),
             int_tbar(ng, o[4], ti, D2bb, G)))


def ccsd_lambda_opt(F, I, T1old, T2old, L1old, L2old, D1, D2, ti, ng, g, G, beta):
    """kelvin/ft_cc_equations.py:385-409."""
    L1int = int_L(ng, L1old, ti, D1, g, G)
    L2int = int_L(ng, L2old, ti, D2, g, G)
    L1 = numpy.zeros(L1old.shape)
    L2 = numpy.zeros(L2old.shape)
    for y in range(ng):
        cqe._Lambda_opt(L1[y], L2[y], F, I, L1int[y], L2int[y], T1old[y], T2old[y], fac=-1.0)
    L1 -= F.ov[None]
    L2 -= I.oovv[None]
    for y in range(ng):
        cqe._LS_TS(L1[y], I, T1old[y], fac=-1.0)
    return L1, L2


def ccsd_lambda_guess(F, I, T1old, beta, ng):
    """kelvin/ft_cc_equations.py:502-512."""
    L1 = numpy.stack([(1.0/beta)*F.ov]*ng)
    L2 = numpy.stack([(1.0/beta)*I.oovv]*ng)
    for y in range(ng):
        cqe._LS_TS(L1[y], I, T1old[y], fac=(1.0/beta))
    return L1, L2


def uccsd_lambda_opt(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb,
                     L1a, L1b, L2aa, L2ab, L2bb, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, g, G, beta):
    """kelvin/ft_cc_equations.py:412-458."""
    i1a, i1b = int_L(ng, L1a, ti, D1a, g, G), int_L(ng, L1b, ti, D1b, g, G)
    i2aa, i2ab, i2bb = (int_L(ng, L2aa, ti, D2aa, g, G), int_L(ng, L2ab, ti, D2ab, g, G),
                        int_L(ng, L2bb, ti, D2bb, g, G))
    n1a, n1b = numpy.zeros(L1a.shape), numpy.zeros(L1b.shape)
    n2aa, n2ab, n2bb = numpy.zeros(L2aa.shape), numpy.zeros(L2ab.shape), numpy.zeros(L2bb.shape)
    for y in range(ng):
        cqe._uccsd_Lambda_opt(n1a[y], n1b[y], n2aa[y], n2ab[y], n2bb[y], Fa, Fb, Ia, Ib, Iabab,
                              (i1a[y], i1b[y]), (i2aa[y], i2ab[y], i2bb[y]),
                              (T1a[y], T1b[y]), (T2aa[y], T2ab[y], T2bb[y]), fac=-1.0)
    n1a -= Fa.ov[None]
    n1b -= Fb.ov[None]
    n2aa -= Ia.oovv[None]
    n2ab -= Iabab.oovv[None]
    n2bb -= Ib.oovv[None]
    for y in range(ng):
        cqe._u_LS_TS(n1a[y], n1b[y], Ia, Ib, Iabab, T1a[y], T1b[y], fac=-1.0)
    return n1a, n1b, n2aa, n2ab, n2bb


def uccsd_lambda_guess(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, beta, ng):
    """kelvin/ft_cc_equations.py:515-526 (note: _u_LS_TS without the 1/beta
    factor, quirk Q3)."""
    L1a = numpy.stack([(1.0/beta)*Fa.ov]*ng)
    L1b = numpy.stack([(1.0/beta)*Fb.ov]*ng)
    L2aa = numpy.stack([(1.0/beta)*Ia.oovv]*ng)
    L2ab = numpy.stack([(1.0/beta)*Iabab.oovv]*ng)
    L2bb = numpy.stack([(1.0/beta)*Ib.oovv]*ng)
    for y in range(ng):
        cqe._u_LS_TS(L1a[y], L1b[y], Ia, Ib, Iabab, T1a[y], T1b[y])
    return L1a, L1b, L2aa, L2ab, L2bb


# --- kelvin/cc_utils.py: loops ---------------------------------------------
def ft_cc_iter(T1old, T2old, F, I, D1, D2, g, G, beta, ng, ti, conv, log=None):
    """kelvin/cc_utils.py:111-173.  Returns (E, T1, T2, history)."""
    norm = numpy.linalg.norm
    alpha = conv["damp"]
    i, Eold, converged, hist = 0, 888888888.888888888, False, []
    nl1 = norm(T1old) + 0.1
    nl2 = norm(T2old) + 0.1
    while i < conv["max_iter"] and not converged:
        T1, T2 = ccsd_stanton(F, I, T1old, T2old, D1, D2, ti, ng, G)
        res1 = norm(T1 - T1old)/nl1
        res2 = norm(T2 - T2old)/nl2
        T1old = alpha*T1old + (1.0 - alpha)*T1
        T2old = alpha*T2old + (1.0 - alpha)*T2
        nl1 = norm(T1old) + 0.1
        nl2 = norm(T2old) + 0.1
        E = ft_cc_energy(T1old, T2old, F.ov, I.oovv, g, beta)
        hist.append((E, res1 + res2))
        i += 1
        if abs(E - Eold) < conv["econv"] and res1 + res2 < conv["tconv"]:
            converged = True
        Eold = E
    return Eold, T1old, T2old, hist


def ft_ucc_iter(T1a, T1b, T2aa, T2ab, T2bb, Fa, Fb, Ia, Ib, Iabab,
                D1a, D1b, D2aa, D2ab, D2bb, g, G, beta, ng, ti, conv):
    """kelvin/cc_utils.py:245-317."""
    norm = numpy.linalg.norm
    alpha = conv["damp"]
    i, Eold, converged, hist = 0, 888888888.888888888, False, []
    while i < conv["max_iter"] and not converged:
        T1o, T2o = uccsd_stanton(Fa, Fb, Ia, Ib, Iabab, T1a, T1b, T2aa, T2ab, T2bb,
                                 D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G)
        nl1 = norm(T1a) + 0.1 + norm(T1b)
        nl2 = norm(T2aa) + 0.1 + norm(T2ab) + norm(T2bb)
        res1 = norm(T1o[0] - T1a)/nl1 + norm(T1o[1] - T1b)/nl1
        res2 = norm(T2o[0] - T2aa)/nl2 + norm(T2o[1] - T2ab)/nl2 + norm(T2o[2] - T2bb)/nl2
        T1a = alpha*T1a + (1.0 - alpha)*T1o[0]
        T1b = alpha*T1b + (1.0 - alpha)*T1o[1]
        T2aa = alpha*T2aa + (1.0 - alpha)*T2o[0]
        T2ab = alpha*T2ab + (1.0 - alpha)*T2o[1]
        T2bb = alpha*T2bb + (1.0 - alpha)*T2o[2]
        E = ft_ucc_energy(T1a, T1b, T2aa, T2ab, T2bb, Fa.ov, Fb.ov, Ia.oovv, Ib.oovv,
                          Iabab.oovv, g, beta)
        hist.append((E, res1 + res2))
        i += 1
        if abs(E - Eold) < conv["econv"] and res1 + res2 < conv["tconv"]:
            converged = True
        Eold = E
    return Eold, (T1a, T1b), (T2aa, T2ab, T2bb), hist


def mp2_guess_g(F, I, D1, D2, ti, ng, G):
    """kelvin/ccsd.py:676-688."""
    return (int_tbar(ng, numpy.stack([-F.vo]*ng), ti, D1, G),
            int_tbar(ng, numpy.stack([-I.vvoo]*ng), ti, D2, G))


def mp2_guess_u(Fa, Fb, Ia, Ib, Iabab, D1a, D1b, D2aa, D2ab, D2bb, ti, ng, G):
    """kelvin/ccsd.py:836-851."""
    return (int_tbar(ng, numpy.stack([-Fa.vo]*ng), ti, D1a, G),
            int_tbar(ng, numpy.stack([-Fb.vo]*ng), ti, D1b, G),
            int_tbar(ng, numpy.stack([-Ia.vvoo]*ng), ti, D2aa, G),
            int_tbar(ng, numpy.stack([-Iabab.vvoo]*ng), ti, D2ab, G),
            int_tbar(ng, numpy.stack([-Ib.vvoo]*ng), ti, D2bb, G))

"""Shared helpers for the test-suite (seeded random FT-CC inputs)."""
import numpy

from kelvin_oracle import cqc


def asym(x):
    x = x - x.swapaxes(-4, -3)
    return x - x.swapaxes(-2, -1)


def random_g(n, ng, seed=0, scale=0.3):
    """Random dressed-integral-like blocks + amplitudes, general spin orbitals."""
    rng = numpy.random.default_rng(seed)
    F = cqc.one_e_blocks(*[rng.standard_normal((n, n)) for _ in range(4)])
    blocks = {}
    for p in cqc.two_e_blocks.names:
        x = rng.standard_normal((n,)*4)
        if p[0] == p[1]:
            x = x - x.transpose(1, 0, 2, 3)
        if p[2] == p[3]:
            x = x - x.transpose(0, 1, 3, 2)
        blocks[p] = numpy.ascontiguousarray(x)
    I = cqc.two_e_blocks(**blocks)
    t1 = scale*rng.standard_normal((ng, n, n))
    t2 = numpy.ascontiguousarray(scale*asym(rng.standard_normal((ng, n, n, n, n))))
    return F, I, t1, t2


def random_g_rect(no, nv, ng, seed=0, scale=0.3):
    """As random_g with no != nv (active-space truncation, kelvin/ccsd.py:642-660)."""
    rng = numpy.random.default_rng(seed)
    dim = {"o": no, "v": nv}
    F = cqc.one_e_blocks(*[rng.standard_normal((dim[p[0]], dim[p[1]])) for p in ("oo", "ov", "vo", "vv")])
    blocks = {}
    for p in cqc.two_e_blocks.names:
        x = rng.standard_normal(tuple(dim[c] for c in p))
        if p[0] == p[1]:
            x = x - x.transpose(1, 0, 2, 3)
        if p[2] == p[3]:
            x = x - x.transpose(0, 1, 3, 2)
        blocks[p] = numpy.ascontiguousarray(x)
    I = cqc.two_e_blocks(**blocks)
    t1 = scale*rng.standard_normal((ng, nv, no))
    t2 = numpy.ascontiguousarray(scale*asym(rng.standard_normal((ng, nv, nv, no, no))))
    l1 = rng.standard_normal((ng, no, nv))
    l2 = numpy.ascontiguousarray(asym(rng.standard_normal((ng, no, no, nv, nv))))
    return F, I, t1, t2, l1, l2


def random_u(na, nb, ng, seed=1, scale=0.3):
    """Random unrestricted inputs whose 34 integral blocks are distinct but
    mutually consistent (one underlying <pq|rs> per spin case, dressed with
    random o/v factors), so that u == g-embedding holds exactly."""
    rng = numpy.random.default_rng(seed)
    rnd = rng.standard_normal
    Va, Vb, Vab = asym(rnd((na,)*4)), asym(rnd((nb,)*4)), rnd((na, nb, na, nb))
    sc = {("o", "a"): rnd(na), ("v", "a"): rnd(na), ("o", "b"): rnd(nb), ("v", "b"): rnd(nb)}

    def dress(V, pat, spins):
        f = [sc[(c, s)] for s, c in zip(spins, pat)]
        return numpy.ascontiguousarray(numpy.einsum('pqrs,p,q,r,s->pqrs', V, *f))
    Ia = cqc.two_e_blocks(**{p: dress(Va, p, "aaaa") for p in cqc.two_e_blocks.names})
    Ib = cqc.two_e_blocks(**{p: dress(Vb, p, "bbbb") for p in cqc.two_e_blocks.names})
    Iabab = cqc.two_e_blocks_full(**{p: dress(Vab, p, "abab") for p in cqc.two_e_blocks_full.names})
    fa, fb = rnd((na, na)), rnd((nb, nb))

    def dressF(f, pat, sp):
        return numpy.ascontiguousarray(
            numpy.einsum('pq,p,q->pq', f, sc[(pat[0], sp)], sc[(pat[1], sp)]))
    Fa = cqc.one_e_blocks(*[dressF(fa, p, "a") for p in ("oo", "ov", "vo", "vv")])
    Fb = cqc.one_e_blocks(*[dressF(fb, p, "b") for p in ("oo", "ov", "vo", "vv")])
    T1a, T1b = scale*rnd((ng, na, na)), scale*rnd((ng, nb, nb))
    T2aa = numpy.ascontiguousarray(scale*asym(rnd((ng, na, na, na, na))))
    T2bb = numpy.ascontiguousarray(scale*asym(rnd((ng, nb, nb, nb, nb))))
    T2ab = scale*rnd((ng, na, nb, na, nb))
    return (Fa, Fb, Ia, Ib, Iabab), (T1a, T1b, T2aa, T2ab, T2bb)


def random_u_closed(n, ng, seed=1, scale=0.3):
    """Random CLOSED-SHELL unrestricted inputs: one spatial <pq|rs> (symmetric under
    particle exchange) for both spins, Fa == Fb, Ia == Ib, Iabab mirror symmetric,
    T1a == T1b, T2aa == T2bb, T2ab[a,B,i,J] == T2ab[B,a,J,i]; plus Lambda-shaped partners."""
    rng = numpy.random.default_rng(seed)
    rnd = rng.standard_normal
    V = rnd((n,)*4)
    V = V + V.transpose(1, 0, 3, 2)
    Vs = V - V.transpose(0, 1, 3, 2)
    sc = {"o": rnd(n), "v": rnd(n)}

    def dress(X, pat):
        f = [sc[c] for c in pat]
        return numpy.ascontiguousarray(numpy.einsum('pqrs,p,q,r,s->pqrs', X, *f))
    Ia = cqc.two_e_blocks(**{p: dress(Vs, p) for p in cqc.two_e_blocks.names})
    Ib = cqc.two_e_blocks(**{p: dress(Vs, p) for p in cqc.two_e_blocks.names})
    Iabab = cqc.two_e_blocks_full(**{p: dress(V, p) for p in cqc.two_e_blocks_full.names})
    f = rnd((n, n))

    def dressF(pat):
        return numpy.ascontiguousarray(numpy.einsum('pq,p,q->pq', f, sc[pat[0]], sc[pat[1]]))
    Fa = cqc.one_e_blocks(*[dressF(p) for p in ("oo", "ov", "vo", "vv")])
    Fb = cqc.one_e_blocks(*[dressF(p) for p in ("oo", "ov", "vo", "vv")])

    def amps(sc_):
        t1 = sc_*rnd((ng, n, n))
        tab = sc_*rnd((ng, n, n, n, n))
        tab = numpy.ascontiguousarray(tab + tab.transpose(0, 2, 1, 4, 3))
        # the same-spin block of a closed-shell state is the antisymmetrised opposite-spin one
        taa = numpy.ascontiguousarray(tab - tab.transpose(0, 2, 1, 3, 4))
        return (t1, t1.copy(), taa, tab, taa.copy())
    return (Fa, Fb, Ia, Ib, Iabab), amps(scale), amps(1.0)


def random_D(n, seed=5):
    rng = numpy.random.default_rng(seed)
    e = numpy.sort(rng.uniform(0.0, 5.0, n))
    return e

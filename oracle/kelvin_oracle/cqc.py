"""Restated helper layer of the (un-vendored, unpinned) ``cqcpy`` package.

TEST INFRASTRUCTURE (see package docstring).  Semantics follow SURVEY.md
Appendix A.1; each symbol cites the reference call sites that fix its meaning.
"""
import numpy


# --- cqcpy.ft_utils -------------------------------------------------------
def ff(beta, eps, mu):
    """Fermi occupation 1/(exp(beta(eps-mu))+1).
    Used at kelvin/cc_utils.py:571,698-701; kelvin/ccsd.py:628."""
    return 1.0 / (numpy.exp(beta * (eps - mu)) + 1.0)


def ffv(beta, eps, mu):
    """Vacancy 1-f = 1/(exp(-beta(eps-mu))+1).  kelvin/cc_utils.py:572."""
    return 1.0 / (numpy.exp(-beta * (eps - mu)) + 1.0)


def GP0(beta, eps, mu):
    """Per-orbital zeroth-order grand potential.  kelvin/ccsd.py:638."""
    return -numpy.log(1.0 + numpy.exp(-beta * (eps - mu))) / beta


def uGP0(beta, ea, eb, mu):
    """kelvin/ccsd.py:738."""
    return GP0(beta, ea, mu), GP0(beta, eb, mu)


def dGP0(beta, eps, mu):
    """d(GP0)/d(beta).  kelvin/ccsd.py:178,228-229."""
    x = beta * (eps - mu)
    return numpy.log(1.0 + numpy.exp(-x)) / (beta * beta) \
        + (eps - mu) / (beta * (numpy.exp(x) + 1.0))


def HtoK(T):
    """Hartree -> Kelvin (log banner, kelvin/ccsd.py:113)."""
    return T * 315775.128914


# --- cqcpy.utils ----------------------------------------------------------
def D1(ev, eo):
    """D1[a,i] = ev[a]-eo[i]  (kelvin/ccsd.py:633;
    bench/ueg_ft_ccsd_ESN19/ulambda_19_04_17.prof:90-97)."""
    return ev[:, None] - eo[None, :]


def D2(ev, eo):
    """D2[a,b,i,j] = ev[a]+ev[b]-eo[i]-eo[j]  (kelvin/ccsd.py:634)."""
    return (ev[:, None, None, None] + ev[None, :, None, None]
            - eo[None, None, :, None] - eo[None, None, None, :])


def D2u(eva, evb, eoa, eob):
    """D2u[a,B,i,J] = eva[a]+evb[B]-eoa[i]-eob[J]  (kelvin/ccsd.py:733)."""
    return (eva[:, None, None, None] + evb[None, :, None, None]
            - eoa[None, None, :, None] - eob[None, None, None, :])


def block_diag(A, B):
    """kelvin/ueg_utils.py:71; kelvin/hubbard_system.py:346."""
    na, ma = A.shape
    nb, mb = B.shape
    out = numpy.zeros((na + nb, ma + mb), dtype=numpy.result_type(A, B))
    out[:na, :ma] = A
    out[na:, ma:] = B
    return out


# --- cqcpy.ov_blocks ------------------------------------------------------
class one_e_blocks(object):
    """Attribute bag (oo, ov, vo, vv).  kelvin/cc_utils.py:588."""
    def __init__(self, oo, ov, vo, vv):
        self.oo, self.ov, self.vo, self.vv = oo, ov, vo, vv


class two_e_blocks(object):
    """Attribute bag of 9 named 2-e blocks.  kelvin/cc_utils.py:599-601."""
    names = ("vvvv", "vvvo", "vovv", "vvoo", "vovo", "oovv", "vooo", "ooov", "oooo")

    def __init__(self, vvvv=None, vvvo=None, vovv=None, vvoo=None, vovo=None,
                 oovv=None, vooo=None, ooov=None, oooo=None):
        self.vvvv, self.vvvo, self.vovv, self.vvoo = vvvv, vvvo, vovv, vvoo
        self.vovo, self.oovv, self.vooo, self.ooov = vovo, oovv, vooo, ooov
        self.oooo = oooo


class two_e_blocks_full(object):
    """Attribute bag of all 16 o/v patterns.  kelvin/cc_utils.py:769-775."""
    names = ("vvvv", "vvvo", "vvov", "vovv", "ovvv", "vvoo", "vovo", "ovvo",
             "voov", "ovov", "oovv", "vooo", "ovoo", "oovo", "ooov", "oooo")

    def __init__(self, **kw):
        for nm in self.names:
            setattr(self, nm, kw.get(nm))


# --- cqcpy.spin_utils (alpha-then-beta packing) -----------
def T1_to_spin(Ta, Tb, nva, noa, nvb, nob):
    """kelvin/tests/test_ft_ccsd_rdm.py:529-556."""
    T = numpy.zeros((nva + nvb, noa + nob), dtype=Ta.dtype)
    T[:nva, :noa] = Ta
    T[nva:, noa:] = Tb
    return T


def T2_to_spin(Taa, Tab, Tbb, nva, noa, nvb, nob):
    """T2ab[a,B,i,J] = T2[a, nva+B, i, noa+J]; mixed images by antisymmetry."""
    T = numpy.zeros((nva + nvb, nva + nvb, noa + nob, noa + nob), dtype=Taa.dtype)
    T[:nva, :nva, :noa, :noa] = Taa
    T[nva:, nva:, noa:, noa:] = Tbb
    T[:nva, nva:, :noa, noa:] = Tab
    T[nva:, :nva, :noa, noa:] = -Tab.transpose(1, 0, 2, 3)
    T[:nva, nva:, noa:, :noa] = -Tab.transpose(0, 1, 3, 2)
    T[nva:, :nva, noa:, :noa] = Tab.transpose(1, 0, 3, 2)
    return T


def D2_to_spin(Daa, Dab, Dbb, nva, noa, nvb, nob):
    """Symmetric (not antisymmetric) packing of energy differences."""
    D = numpy.zeros((nva + nvb, nva + nvb, noa + nob, noa + nob), dtype=Daa.dtype)
    D[:nva, :nva, :noa, :noa] = Daa
    D[nva:, nva:, noa:, noa:] = Dbb
    D[:nva, nva:, :noa, noa:] = Dab
    D[nva:, :nva, :noa, noa:] = Dab.transpose(1, 0, 2, 3)
    D[:nva, nva:, noa:, :noa] = Dab.transpose(0, 1, 3, 2)
    D[nva:, :nva, noa:, :noa] = Dab.transpose(1, 0, 3, 2)
    return D


def int_to_spin(Ia, Ib, Iabab, pat):
    """Pack the u integral blocks of o/v pattern ``pat`` (e.g. 'vovv') into the
    spin-orbital block <pq||rs>.  Ia/Ib: antisymmetrised same-spin blocks
    (two_e_blocks); Iabab: two_e_blocks_full with Iabab.xyzw[p,Q,r,S]=<pQ|rS>.
    The first-pair / second-pair swaps of ``pat`` select which Iabab block
    supplies each mixed-spin image."""
    x, y, z, w = pat
    aa = getattr(Ia, pat)
    bb = getattr(Ib, pat)
    n = [aa.shape[k] for k in range(4)]
    m = [bb.shape[k] for k in range(4)]
    out = numpy.zeros(tuple(n[k] + m[k] for k in range(4)), dtype=aa.dtype)
    A = [slice(0, n[k]) for k in range(4)]
    B = [slice(n[k], n[k] + m[k]) for k in range(4)]
    out[A[0], A[1], A[2], A[3]] = aa
    out[B[0], B[1], B[2], B[3]] = bb
    # (a B a B): Iabab.xyzw[p,Q,r,S]
    out[A[0], B[1], A[2], B[3]] = getattr(Iabab, x + y + z + w)
    # (B a a B)[P,q,r,S] = -<qP|rS> = -Iabab.yxzw[q,P,r,S]
    out[B[0], A[1], A[2], B[3]] = -getattr(Iabab, y + x + z + w).transpose(1, 0, 2, 3)
    # (a B B a)[p,Q,R,s] = -<pQ|sR> = -Iabab.xywz[p,Q,s,R]
    out[A[0], B[1], B[2], A[3]] = -getattr(Iabab, x + y + w + z).transpose(0, 1, 3, 2)
    # (B a B a)[P,q,R,s] = <qP|sR> = Iabab.yxwz[q,P,s,R]
    out[B[0], A[1], B[2], A[3]] = getattr(Iabab, y + x + w + z).transpose(1, 0, 3, 2)
    return out


def F_to_spin(Fa, Fb):
    oo = block_diag(Fa.oo, Fb.oo)
    ov = block_diag(Fa.ov, Fb.ov)
    vo = block_diag(Fa.vo, Fb.vo)
    vv = block_diag(Fa.vv, Fb.vv)
    return one_e_blocks(oo, ov, vo, vv)


def I_to_spin(Ia, Ib, Iabab):
    return two_e_blocks(**{p: int_to_spin(Ia, Ib, Iabab, p) for p in two_e_blocks.names})

"""Sz-blocked evaluation of the restated Stanton terms on the CPU.

TEST INFRASTRUCTURE / timed CPU baseline ("port").  The same
``cc_equations.stanton_terms`` statement sequence is executed
on spin-blocked tensors: every einsum is expanded over the Sz-allowed spin
blocks of its operands and each block contraction is a NumPy einsum
(tensordot -> dgemm), which is how an unrestricted CPU implementation
(cqcpy._u_Stanton, called at kelvin/ft_cc_equations.py:153) spends its time.
It is validated against the exact g-embedding in ``cc_equations.u_stanton_terms``
(tests/test_oracle.py), i.e. the reference's own u == g criterion
(kelvin/tests/test_ft_cc_ampl.py:41-110).
"""
import numpy

from . import cqc
from . import cc_equations as cqe

P4 = ("aaaa", "bbbb", "abab", "baba", "abba", "baab")
CANON = ("aaaa", "bbbb", "abab")


class ST(object):
    """Spin-blocked tensor: pattern string -> ndarray.  ``anti`` tensors are
    antisymmetric in (0,1) and (2,3) and may store only the canonical blocks."""
    def __init__(self, blocks, anti=False):
        self.b = blocks
        self.anti = anti
        self.ndim = len(next(iter(blocks))) if blocks else 0

    def block(self, p):
        if p in self.b:
            return self.b[p]
        if self.anti and self.ndim == 4:
            if p == "baba":
                return self.b["abab"].transpose(1, 0, 3, 2)
            if p == "baab":
                return -self.b["abab"].transpose(1, 0, 2, 3)
            if p == "abba":
                return -self.b["abab"].transpose(0, 1, 3, 2)
        return None

    def patterns(self):
        if self.anti and self.ndim == 4:
            return P4
        return tuple(self.b.keys())

    def full(self):
        return ST({p: self.block(p) for p in self.patterns()})

    def _bin(self, o, sgn):
        if isinstance(o, ST):
            if self.anti and o.anti:
                keys = set(self.b) | set(o.b)
                return ST({k: _add(self.b.get(k), o.b.get(k), sgn) for k in keys}, anti=True)
            keys = set(self.patterns()) | set(o.patterns())
            return ST({k: _add(self.block(k), o.block(k), sgn) for k in keys})
        raise TypeError

    def __add__(self, o):
        return self._bin(o, 1.0)

    def __sub__(self, o):
        return self._bin(o, -1.0)

    def __neg__(self):
        return ST({k: -v for k, v in self.b.items()}, self.anti)

    def __mul__(self, c):
        return ST({k: c*v for k, v in self.b.items()}, self.anti)

    __rmul__ = __mul__

    def transpose(self, *perm):
        s = self.full() if self.anti else self
        return ST({"".join(k[p] for p in perm): v.transpose(*perm) for k, v in s.b.items()})


def _add(x, y, sgn):
    if x is None:
        return sgn*y
    if y is None:
        return x
    return x + sgn*y


def ein(subs, *ops, **kw):
    """Spin-summed einsum.  keep=True computes only the canonical blocks of an
    antisymmetric 4-index result."""
    keep = kw.get("keep", False)
    lhs, out = subs.split("->")
    terms = lhs.split(",")
    res = {}

    def rec(k, spin, blocks):
        if k == len(ops):
            pat = "".join(spin[l] for l in out)
            if keep and pat not in CANON:
                return
            val = numpy.einsum(subs, *blocks, optimize=True)
            res[pat] = res[pat] + val if pat in res else val
            return
        for p in ops[k].patterns():
            ok = True
            new = dict(spin)
            for l, s in zip(terms[k], p):
                if new.setdefault(l, s) != s:
                    ok = False
                    break
            if ok:
                blk = ops[k].block(p)
                if blk is not None:
                    rec(k + 1, new, blocks + [blk])
    rec(0, {}, [])
    return ST(res, anti=keep)


class _Bag(object):
    pass


def wrap_integrals(Fa, Fb, Ia, Ib, Iabab):
    F, I = _Bag(), _Bag()
    for nm in ("oo", "ov", "vo", "vv"):
        setattr(F, nm, ST({"aa": getattr(Fa, nm), "bb": getattr(Fb, nm)}))
    for pat in cqc.two_e_blocks.names:
        w, x, y, z = pat
        setattr(I, pat, ST({
            "aaaa": getattr(Ia, pat), "bbbb": getattr(Ib, pat),
            "abab": getattr(Iabab, pat),
            "baba": getattr(Iabab, x + w + z + y).transpose(1, 0, 3, 2),
            "baab": -getattr(Iabab, x + w + y + z).transpose(1, 0, 2, 3),
            "abba": -getattr(Iabab, w + x + z + y).transpose(0, 1, 3, 2)}))
    return F, I


def u_stanton_terms(Fa, Fb, Ia, Ib, Iabab, T1olds, T2olds, wrapped=None):
    """(R1a, R1b, R2aa, R2ab, R2bb) evaluated block by block."""
    F, I = wrapped if wrapped is not None else wrap_integrals(Fa, Fb, Ia, Ib, Iabab)
    t1 = ST({"aa": T1olds[0], "bb": T1olds[1]})
    t2 = ST({"aaaa": T2olds[0], "abab": T2olds[1], "bbbb": T2olds[2]}, anti=True)
    R1, R2 = cqe.stanton_terms(F, I, t1, t2, ein=ein, hints=True)
    return (R1.block("aa"), R1.block("bb"), R2.block("aaaa"), R2.block("abab"), R2.block("bbbb"))

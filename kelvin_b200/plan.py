"""Contraction-plan compiler for the per-tau-point FT-CCSD residual.

The reference evaluates the residual with ~60 ``einsum`` calls per grid point
inside the external ``cqcpy.cc_equations._Stanton`` / ``_u_Stanton`` /
``_Lambda_opt`` kernels (called from kelvin/ft_cc_equations.py:106,153,399,442).
Here the same algebra is written once as a list of statements

    ``out[idx] += coef * A[idx] * B[idx]``   /   ``out[idx] += coef * A[idx]``

over spin-orbital index letters (a-h virtual, i-p occupied) and compiled into
``kb200_op`` descriptors for the gathered DMMA GEMM kernel
(include/kelvin_b200.h): every index permutation becomes an offset table, every
statement is batched over the imaginary-time grid, nothing is transposed in
memory.  Two back ends share one statement list:

* ``mode='g'``: general spin orbitals, one dense array per tensor
  (kelvin/ft_cc_equations.py:96 ``ccsd_stanton``);
* ``mode='u'``: Sz-blocked.  Each statement is expanded over the spin cases of
  its indices; antisymmetric tensors keep only the (aa, ab, bb) blocks the
  reference stores (T2ab[a,B,i,J], Iabab[p,Q,r,S] = <pQ|rS>) and
  every other spin block is read as a signed, permuted *view* of a stored one
  (kelvin/ft_cc_equations.py:130 ``uccsd_stanton``).

The Lambda map (kelvin/ft_cc_equations.py:385,412) is the vector-Jacobian
product of the residual; ``adjoint`` derives it mechanically
from the resolved forward ops, so the g and u Lambda kernels are generated, not
hand-written.
"""
import ctypes
import os
import re
from collections import OrderedDict

import numpy

_os_environ_get = os.environ.get

VIRT = "abcdefgh"
OCC = "ijklmnop"


def space_of(letter):
    if letter in VIRT:
        return "v"
    if letter in OCC:
        return "o"
    raise ValueError("bad index letter %r" % letter)


# ---------------------------------------------------------------------------
# statements
# ---------------------------------------------------------------------------
_ref = re.compile(r"([A-Za-z_0-9.]+)\[([a-p]*)\]")


class Stmt(object):
    """out[idx] += coef * prod(ins)."""
    __slots__ = ("out", "coef", "ins")

    def __init__(self, out, coef, ins):
        self.out = out        # (name, letters)
        self.coef = float(coef)
        self.ins = ins        # list of (name, letters), len 1 or 2

    def __repr__(self):
        return "%s[%s] += %g %s" % (self.out[0], self.out[1], self.coef,
                                    " ".join("%s[%s]" % x for x in self.ins))


def parse(text):
    """'X[abij] += 0.5 A[aeim] B[mbej]' -> Stmt."""
    lhs, rhs = text.split("+=")
    m = _ref.search(lhs)
    out = (m.group(1), m.group(2))
    rhs = rhs.strip()
    first = rhs.split()[0]
    coef = float(first)
    ins = [(mm.group(1), mm.group(2)) for mm in _ref.finditer(rhs[len(first):])]
    assert 1 <= len(ins) <= 2, text
    return Stmt(out, coef, ins)


# ---------------------------------------------------------------------------
# tensor classes and spin-block resolution
# ---------------------------------------------------------------------------
class TDef(object):
    """kind: 'one'   2-index, blocks a / b
             'amp2'  4-index antisymmetric in (0,1) and (2,3): blocks aa, ab, bb
             'gen4'  4-index, no permutational symmetry: 6 Sz-allowed blocks
             'int1'  Fock block F.xy          (u: Fa.xy / Fb.xy)
             'int2'  ERI block  I.wxyz        (u: Ia / Ib / Iabab.<pattern>)
       role: 'in' (caller supplies), 'out' (caller supplies, accumulated into),
             'tmp' (plan-owned scratch); batched: has a leading tau index.
       canon: for a 'gen4' scratch tensor stored in another index order, the storage positions
             of the canonical <p q | r s> order (Sz conservation: spin(p)+spin(q) = spin(r)+spin(s))."""
    def __init__(self, name, kind, role, batched, spaces, canon=None):
        self.name, self.kind, self.role, self.batched, self.spaces = name, kind, role, batched, spaces
        self.canon = canon
        assert canon is None or kind == "gen4"


def _canon(td, xs):
    """xs (letters or spins in storage order) -> canonical order."""
    if td.canon is None or len(xs) != 4:
        return list(xs)
    return [xs[c] for c in td.canon]


SPIN4 = ("aaaa", "bbbb", "abab", "baba", "abba", "baab")


def resolve_u(td, letters, spins):
    """(slot name, letters in storage order, sign) of spin block `spins`."""
    s = "".join(spins)
    if len(letters) == 2:
        assert s in ("aa", "bb")
        if td.kind == "int1":
            return ("F%s.%s" % (s[0], td.name.split(".")[1]), letters, 1.0)
        return (td.name + "." + s[0], letters, 1.0)
    if td.kind == "gen4":
        s = "".join(_canon(td, spins))
        assert s in SPIN4, s
        return (td.name + "." + s, letters, 1.0)
    assert s in SPIN4, s
    p, q, r, t = letters
    if td.kind == "amp2":
        if s == "aaaa":
            return (td.name + ".aa", letters, 1.0)
        if s == "bbbb":
            return (td.name + ".bb", letters, 1.0)
        if s == "abab":
            return (td.name + ".ab", letters, 1.0)
        if s == "baba":
            return (td.name + ".ab", q + p + t + r, 1.0)
        if s == "baab":
            return (td.name + ".ab", q + p + r + t, -1.0)
        if s == "abba":
            return (td.name + ".ab", p + q + t + r, -1.0)
    if td.kind == "int2":
        pat = td.name.split(".")[1]
        w, x, y, z = pat
        if s == "aaaa":
            return ("Ia." + pat, letters, 1.0)
        if s == "bbbb":
            return ("Ib." + pat, letters, 1.0)
        if s == "abab":
            return ("Iabab." + pat, letters, 1.0)
        if s == "baba":
            return ("Iabab." + x + w + z + y, q + p + t + r, 1.0)
        if s == "baab":
            return ("Iabab." + x + w + y + z, q + p + r + t, -1.0)
        if s == "abba":
            return ("Iabab." + w + x + z + y, p + q + t + r, -1.0)
    raise ValueError("cannot resolve %s %s" % (td.name, s))


def canonical_out_blocks(td, nidx):
    if nidx == 2:
        return (("a", "a"), ("b", "b"))
    if td.kind == "amp2":
        return (tuple("aaaa"), tuple("abab"), tuple("bbbb"))
    if td.canon is not None:
        out = []
        for s in SPIN4:
            st = [None]*4
            for k, c in enumerate(td.canon):
                st[c] = s[k]
            out.append(tuple(st))
        return tuple(out)
    return tuple(tuple(s) for s in SPIN4)


def sz_ok(spins):
    if len(spins) == 2:
        return spins[0] == spins[1]
    n = lambda c: 1 if c == "a" else 0  # noqa: E731
    return n(spins[0]) + n(spins[1]) == n(spins[2]) + n(spins[3])


class ROp(object):
    """Resolved op on concrete slots: out[letters] += coef * prod(ins)."""
    __slots__ = ("out", "coef", "ins", "spin", "tri", "slab")

    def __init__(self, out, coef, ins, spin=None, tri=None, slab=False):
        self.out, self.coef, self.ins, self.spin = out, coef, ins, spin
        # tri: pairs of output letters (x, y); only the elements with x < y are computed and
        # written (see antisym_outputs)
        self.tri = tri
        # slab: the rows of this contraction are dealt to the ranks that share the grid point
        # (hybrid partition, see hybrid_phases); each rank writes only its own rows
        self.slab = slab

    def __repr__(self):
        return "%s[%s] += %g %s" % (self.out[0], self.out[1], self.coef,
                                    " ".join("%s[%s]" % x for x in self.ins))


def _canon_key(out, ins, spin):
    """Key identifying an expanded term up to renaming of summed letters."""
    ren = {}
    for l in out[1]:
        ren[l] = l
    nxt = [0]

    def r(l):
        if l not in ren:
            ren[l] = "#%d" % nxt[0]
            nxt[0] += 1
        return ren[l]
    parts = [(out[0], out[1])]
    for nm, ls in ins:
        parts.append((nm, tuple(r(l) for l in ls)))
    sp = tuple(sorted((ren[l], s) for l, s in spin.items()))
    return (tuple(parts), sp)


def expand(stmts, tdefs, mode):
    """Statements -> resolved ops (g: 1:1; u: spin expansion + merging)."""
    rops = []
    for st in stmts:
        tds = [tdefs[st.out[0]]] + [tdefs[nm] for nm, _ in st.ins]
        if mode == "g":
            rops.append(ROp(st.out, st.coef, list(st.ins)))
            continue
        letters = []
        for _, ls in [st.out] + st.ins:
            for l in ls:
                if l not in letters:
                    letters.append(l)
        summed = [l for l in letters if l not in st.out[1]]
        merged = OrderedDict()
        for ospin in canonical_out_blocks(tds[0], len(st.out[1])):
            base = dict(zip(st.out[1], ospin))
            for code in range(1 << len(summed)):
                spin = dict(base)
                for k, l in enumerate(summed):
                    spin[l] = "a" if not (code >> k) & 1 else "b"
                ok = True
                for td, (_, ls) in zip(tds[1:], st.ins):
                    if not sz_ok(_canon(td, [spin[l] for l in ls])):
                        ok = False
                        break
                if not ok:
                    continue
                oslot, ols, osg = resolve_u(tds[0], st.out[1], [spin[l] for l in st.out[1]])
                assert osg == 1.0 and ols == st.out[1]
                coef = st.coef
                rins = []
                for td, (_, ls) in zip(tds[1:], st.ins):
                    slot, sls, sg = resolve_u(td, ls, [spin[l] for l in ls])
                    coef *= sg
                    rins.append((slot, sls))
                # put the two operands in a canonical order for merging
                key_ins = rins
                if len(rins) == 2 and rins[0][0] == rins[1][0]:
                    key_ins = sorted(rins)
                key = _canon_key((oslot, ols), key_ins, spin)
                if key in merged:
                    merged[key].coef += coef
                else:
                    merged[key] = ROp((oslot, ols), coef, rins, dict(spin))
        for op in merged.values():
            if op.coef != 0.0:
                rops.append(op)
    return rops


# ---------------------------------------------------------------------------
# closed-shell mirror reduction (unrestricted programs on alpha == beta inputs)
# ---------------------------------------------------------------------------
_FLIP = str.maketrans("ab", "ba")


def mirror_slot(slot):
    """The stored slot that holds the same numbers once alpha and beta are interchanged
    (same index order), or None when the slot is its own mirror image (amp2 'ab' blocks,
    whose 'ba' image is a transposed view of themselves) or has no stored image by name
    (Iabab blocks: their images are other, transposed Iabab blocks)."""
    bar = "~" if slot.endswith("~") else ""
    base = slot[:-1] if bar else slot
    pre, _, suf = base.partition(".")
    if pre in ("Ia", "Ib", "Fa", "Fb"):
        return pre[0] + pre[1].translate(_FLIP) + "." + suf + bar
    if pre == "Iabab" or "@" in base:
        return None
    name, _, sp = base.rpartition(".")
    if not name or not sp or set(sp) - set("ab"):
        return None
    fl = sp.translate(_FLIP)
    if len(sp) == 2 and sp[0] != sp[1]:
        return None
    return name + "." + fl + bar


def mirror_rep(slot):
    """Representative of {slot, mirror_slot(slot)}: the alpha-leading one."""
    m = mirror_slot(slot)
    if m is None:
        return slot
    pre = slot.partition(".")[0]
    if pre in ("Ib", "Fb"):
        return m
    if pre in ("Ia", "Fa"):
        return slot
    sp = slot.rstrip("~").rpartition(".")[2]
    return slot if sp[0] == "a" else m


def mirror_reduce(rops):
    """Unrestricted program -> the ops a spin-symmetric (closed-shell) run needs.

    The spin expansion is closed under the interchange alpha <-> beta: every op writing a
    block X.s has a partner writing X.flip(s) with all spins flipped.  When every input is
    mirror symmetric (Fa == Fb, Ia == Ib, Iabab.wxyz[p,q,r,s] == Iabab.xwzy[q,p,s,r],
    T1a == T1b, T2aa == T2bb, T2ab[a,B,i,J] == T2ab[B,a,J,i]; na == nb) the partner computes
    the same numbers, so only the alpha-leading block of each pair is evaluated and every
    read of a beta-leading block is redirected to its image (same memory, same index order).
    Blocks that are their own image (o2.ab, tau.ab, W_oooo.ab, ...) keep all their ops.
    The caller copies the alpha outputs into the beta ones."""
    out = []
    for op in rops:
        if mirror_rep(op.out[0]) != op.out[0]:
            continue
        ins = [(mirror_rep(sl), ls) for sl, ls in op.ins]
        out.append(ROp(op.out, op.coef, ins, op.spin))
    return out


# ---------------------------------------------------------------------------
# closed shell: sum/difference form of paired contractions
# ---------------------------------------------------------------------------
SUMDIFF_PREFIX = "Qsd"
DUP_PREFIX = "Qdup"


def _swap_pos(ls, p, q):
    ls = list(ls)
    ls[p], ls[q] = ls[q], ls[p]
    return "".join(ls)


def _orientations(slot, ls):
    """(sign, letters) of every way of reading an operand that its permutational antisymmetry
    (antisym_pairs: known from the tensor class) makes equal up to the sign."""
    res = [(1.0, ls)]
    for p, q in antisym_pairs(slot):
        res += [(-sg, _swap_pos(l, p, q)) for sg, l in list(res)]
    return res


def _relabel(op, la, lb, k, flip=True):
    """Effective (coef, out letters) of binary op `op` once its operands are re-oriented (through
    their antisymmetry; flip=False: as written only) and its letters renamed so that they read
    [la], [lb] like the reference op; None when no orientation fits.  k: which operand of `op`
    plays the first role."""
    (sa_, lsa), (sb_, lsb) = op.ins[k], op.ins[1 - k]
    if len(lsa) != len(la) or len(lsb) != len(lb):
        return None
    for sga, va in (_orientations(sa_, lsa) if flip else [(1.0, lsa)]):
        for sgb, vb in (_orientations(sb_, lsb) if flip else [(1.0, lsb)]):
            phi, ok = {}, True
            for x, y in list(zip(va, la)) + list(zip(vb, lb)):
                if phi.setdefault(x, y) != y:
                    ok = False
                    break
            if not ok or len(set(phi.values())) != len(phi):
                continue
            if any(l not in phi for l in op.out[1]):
                continue
            return op.coef*sga*sgb, "".join(phi[l] for l in op.out[1])
    return None


def merge_duplicates(rops):
    """Closed-shell (mirror_reduce'd) reverse sweeps contain pairs of contractions with the very
    same operands whose results are added to one block under two index orders,

        X[pqrs] += c A B        X[qpsr] += c A B

    (the adjoint of a self-mirror block receives the contribution of its beta-leading image as
    a transposed copy).  The contraction is done once into a scratch block T,

        T[pqrs] (+)= c A B      ...      X[pqrs] += T[pqrs];  X[qpsr] += T[pqrs]

    and every such pair with the same (X, two index orders) shares one T, so that the sum /
    difference rewrite below still sees the members of its quartets in one output."""
    ops = list(rops)
    temps = {}            # (X, lo1, lo2) -> temp slot
    used = set()
    out = []
    pend = {}             # temp -> (X, lo1, lo2, spin, position of the last contributing op)
    n = len(ops)
    partner = {}
    for i in range(n):
        a = ops[i]
        if i in partner or len(a.ins) != 2 or a.tri is not None or len(a.out[1]) != 4 \
                or len(set(a.ins[0][1]) & set(a.ins[1][1])) < 2:
            continue                      # m^6 contractions only: a copy pass costs as much as an m^5 one
        for j in range(i + 1, n):
            b = ops[j]
            if j in partner or len(b.ins) != 2 or b.out[0] != a.out[0] or b.tri is not None:
                continue
            if b.out[1] == a.out[1] or sorted(b.out[1]) != sorted(a.out[1]):
                continue
            r = _relabel(b, a.ins[0][1], a.ins[1][1], 0, flip=False)
            if r is None or b.ins[0][0] != a.ins[0][0] or b.ins[1][0] != a.ins[1][0]:
                continue
            if abs(r[0] - a.coef) > 1e-14*abs(a.coef):
                continue
            # nothing in between may read X or write an operand
            srcs = {a.ins[0][0], a.ins[1][0]}
            if any(o.out[0] in srcs or any(sl == a.out[0] for sl, _ in o.ins)
                   for o in ops[i + 1:j]):
                continue
            partner[i] = (j, r[1])
            partner[j] = None
            break
    k = 0
    last = {}
    plan_ = []
    for i, op in enumerate(ops):
        if i in partner and partner[i] is None:
            continue                                     # the duplicate: dropped
        if i in partner:
            j, lo2 = partner[i]
            key = (op.out[0], op.out[1], lo2)
            if key not in temps:
                suf = op.out[0].partition(".")[2]
                temps[key] = "%s%d%s" % (DUP_PREFIX, k, "." + suf.rstrip("~") if suf else "")
                k += 1
            t = temps[key]
            plan_.append(ROp((t, op.out[1]), op.coef, list(op.ins), op.spin))
            last[key] = (len(plan_) - 1, max(last.get(key, (0, 0))[1], j), op.spin)
        else:
            plan_.append(op)
    # the two adds go in front of the first later reader of X (or at the end): later rewrites
    # (sumdiff_pairs) then find nothing between the members of a group that reads T
    inserts = []
    for key, (pos, _, _) in last.items():
        at = len(plan_)
        for j in range(pos + 1, len(plan_)):
            if any(sl == key[0] for sl, _ in plan_[j].ins):
                at = j
                break
        inserts.append((at, key))
    for at, key in sorted(inserts, key=lambda x: -x[0]):
        X, lo1, lo2 = key
        t = temps[key]
        spin = last[key][2]
        plan_[at:at] = [ROp((X, lo1), 1.0, [(t, lo1)], spin),
                        ROp((X, lo2), 1.0, [(t, lo1)], spin)]
    return plan_


def sumdiff_pairs(rops):
    """Closed-shell (mirror_reduce'd) programs contain quartets of contractions

        Xa += c A1 B1     Xa += c A2 B2     Xb += c' A1 B2     Xb += c' A2 B1

    with one index pattern (the W_ovvo.aaaa / W_ovvo.abab builds from Ia.oovv, Iabab.oovv and
    t2x.aaaa, t2x.abba; the ring contractions rg.aaaa / rg.abab from t2.aa, t2.ab and those two
    W blocks; their images in the Lambda sweep).  Since Xa/c + Xb/c' gets (A1+A2)(B1+B2) and
    Xa/c - Xb/c' gets (A1-A2)(B1-B2), two contractions do the work of four:

        Ap = A1+A2, Am = A1-A2, Bp = B1+B2, Bm = B1-B2        (elementwise)
        S = c Ap Bp,  D = c Am Bm                              (2 contractions)
        Xa += S/2 + D/2,  Xb += (c'/c) (S/2 - D/2)             (elementwise)

    This is the singlet/triplet channel decomposition of the closed-shell ring terms.  Members
    are matched up to a renaming of their letters and up to the permutational antisymmetry of
    same-spin operands (t2.aa[aeim] = t2.aa[eami]); Xb may carry its own index order.  The
    rewrite is placed where the last member of the quartet stood and is only done when nothing
    in between reads Xa/Xb or writes an operand."""
    ops = list(rops)
    k = 0
    while True:
        quartet = _find_quartet(ops)
        if quartet is None:
            return ops
        (i11, i22, i12, i21), (A2, B2), lob, ratio = quartet
        o11 = ops[i11]
        (A1, la), (B1, lb) = o11.ins
        Xa, Xb = o11.out, (ops[i12].out[0], lob)
        c, spin = o11.coef, o11.spin
        tag = "%s%d" % (SUMDIFF_PREFIX, k)
        k += 1
        Ap, Am, Bp, Bm, S, D = [tag + x for x in ("Ap", "Am", "Bp", "Bm", "S", "D")]
        lc = Xa[1]
        # the sums / differences are scratch: store them as the contraction wants to read them,
        # [free indices][contracted indices, one order for both operands] -- plain matrices
        kk = "".join(l for l in lb if l in la)
        sa_ = "".join(l for l in la if l in lc) + kk
        sb_ = "".join(l for l in lb if l in lc) + kk
        new = [ROp((Ap, sa_), 1.0, [(A1, la)], spin), ROp((Ap, sa_), 1.0, [(A2, la)], spin),
               ROp((Am, sa_), 1.0, [(A1, la)], spin), ROp((Am, sa_), -1.0, [(A2, la)], spin),
               ROp((Bp, sb_), 1.0, [(B1, lb)], spin), ROp((Bp, sb_), 1.0, [(B2, lb)], spin),
               ROp((Bm, sb_), 1.0, [(B1, lb)], spin), ROp((Bm, sb_), -1.0, [(B2, lb)], spin),
               ROp((S, lc), c, [(Ap, sa_), (Bp, sb_)], spin),
               ROp((D, lc), c, [(Am, sa_), (Bm, sb_)], spin),
               ROp(Xa, 0.5, [(S, lc)], spin), ROp(Xa, 0.5, [(D, lc)], spin),
               ROp(Xb, 0.5*ratio, [(S, lc)], spin), ROp(Xb, -0.5*ratio, [(D, lc)], spin)]
        q = (i11, i22, i12, i21)
        last = max(q)
        drop = set(q)
        ops = [op for j, op in enumerate(ops[:last]) if j not in drop] + new + ops[last + 1:]


SPLIT_PREFIX = "Acc"


def _is_large(op):
    """Two 4-index operands contracted over two letters into a 4-index output (an m^6 term)."""
    return (len(op.ins) == 2 and len(op.out[1]) == 4 and all(len(ls) == 4 for _, ls in op.ins)
            and len(set(op.ins[0][1]) & set(op.ins[1][1])) == 2)


def split_accumulators(rops):
    """Small tau batches (1-2 grid points per launch, the tau-sharded case): every large
    contraction after the first that accumulates into one slot gets a scratch slot of its own and
    an elementwise add.  The large contractions of one dependency level are then mutually
    independent and share ONE launch: at batch 1 the ten m^6 launches of the closed-shell
    residual go out as 5 + 5 tiles-of-81 (3 + 3 waves of 148 CTAs) instead of 5 + 3 + 2
    (3 + 2 + 2 waves).  Costs one scratch block per split term; triangular and row-slabbed
    contractions are left alone."""
    seen, out, k = {}, [], 0
    for op in rops:
        if _is_large(op) and op.tri is None and not op.slab:
            cnt = seen.get(op.out[0], 0)
            seen[op.out[0]] = cnt + 1
            if cnt >= 1:
                tmp = "%s%d" % (SPLIT_PREFIX, k)
                k += 1
                out.append(ROp((tmp, op.out[1]), op.coef, op.ins, op.spin))
                out.append(ROp(op.out, 1.0, [(tmp, op.out[1])], op.spin))
                continue
        out.append(op)
    return out


def _find_quartet(ops):
    cand = [j for j, op in enumerate(ops) if len(op.ins) == 2 and op.tri is None
            and len(op.out[1]) == 4 and not op.out[0].startswith(SUMDIFF_PREFIX)
            and len(set(op.ins[0][1]) & set(op.ins[1][1])) == 2]
    for i11 in cand:
        o11 = ops[i11]
        (A1, la), (B1, lb) = o11.ins
        if A1 == B1:
            continue
        c = o11.coef
        # every other candidate, read in the orientation of o11: (A slot, B slot) -> entries
        table = {}
        for j in cand:
            if j == i11:
                continue
            o = ops[j]
            for kk in (0, 1):
                r = _relabel(o, la, lb, kk)
                if r is not None:
                    table.setdefault((o.ins[kk][0], o.ins[1 - kk][0]), []).append((j, r[0], r[1]))
                    break
        for (A2, B2), ents in table.items():
            if A2 == A1 or B2 == B1 or A2 == B2:
                continue
            for i22, c22, lo22 in ents:
                if ops[i22].out[0] != o11.out[0] or lo22 != o11.out[1] \
                        or abs(c22 - c) > 1e-14*abs(c):
                    continue
                for i12, c12, lo12 in table.get((A1, B2), ()):
                    for i21, c21, lo21 in table.get((A2, B1), ()):
                        if ops[i12].out[0] != ops[i21].out[0] or ops[i12].out[0] == o11.out[0] \
                                or lo12 != lo21 or abs(c12 - c21) > 1e-14*abs(c12) \
                                or len({i11, i22, i12, i21}) != 4:
                            continue
                        q = (i11, i22, i12, i21)
                        lo, hi = min(q), max(q)
                        Xs = {o11.out[0], ops[i12].out[0]}
                        srcs = {A1, A2, B1, B2}
                        ok = True
                        for j in range(lo, hi + 1):
                            if j in q:
                                continue
                            o = ops[j]
                            if o.out[0] in srcs or o.out[0] in Xs \
                                    or any(sl in Xs for sl, _ in o.ins):
                                ok = False
                                break
                        if ok:
                            return q, (A2, B2), lo12, c12/c
    return None


# ---------------------------------------------------------------------------
# closed-shell singlet: the same-spin doubles residual from the opposite-spin one
# ---------------------------------------------------------------------------
def singlet_reduce(rops, o1="o1.a", o2ab="o2.ab", o2aa="o2.aa", emit_aa=True):
    """Closed-shell (mirror_reduce'd) residual programs whose amplitudes also satisfy the singlet
    relation T2aa[a,b,i,j] == T2ab[a,b,i,j] - T2ab[b,a,i,j] (true of the MP2 guess of a
    spin-symmetric system and preserved by the update): the residual obeys the same relation
    (checked against the oracle on random inputs in tests/test_plan_cpu.py), so everything that
    only feeds the same-spin doubles residual -- the same-spin ladder terms and the rg.aaaa ring
    contraction, 2.9 of the 13.9 remaining GEMM units -- is dropped and

        o2.aa[abij] = o2.ab[abij] - o2.ab[baij].

    emit_aa=False leaves the last step out: the caller derives the same-spin block from the
    opposite-spin one itself (ft_cc_equations does so AFTER the multi-GPU exchange, which then
    moves half the bytes)."""
    live = {o1, o2ab}
    kept = []
    for op in reversed(rops):
        if op.out[0] in live:
            kept.append(op)
            for sl, _ in op.ins:
                live.add(sl)
    kept.reverse()
    if not emit_aa:
        return kept
    spin = None
    for op in rops:
        if op.out[0] == o2aa:
            spin = op.spin
            ls = op.out[1]
            break
    if spin is None:
        raise ValueError("program has no %s output" % o2aa)
    sab = {l: s for l, s in zip(ls, "abab")}
    kept.append(ROp((o2aa, ls), 1.0, [(o2ab, ls)], sab))
    kept.append(ROp((o2aa, ls), -1.0, [(o2ab, ls[1] + ls[0] + ls[2:])], sab))
    return kept


# ---------------------------------------------------------------------------
# antisymmetric outputs: compute one triangle, add its four images
# ---------------------------------------------------------------------------
ANTISYM_OUT = int(_os_environ_get("KB200_ANTISYM_OUT", "1"))
TRI_PREFIX = "Ptri"


def _tri_candidate(op):
    """((x, y), (u, v)) when the product of the two operands is antisymmetric under x <-> y
    (a pair carried by the first operand) and under u <-> v (carried by the second) and these
    four letters are the whole output: the same-spin ladder terms
    tau[efij] W_vvvv[abef], tau[abmn] W_oooo[mnij] and the I.oovv.tau builds of W_oooo / W_vvvv."""
    if len(op.ins) != 2 or len(op.out[1]) != 4 or op.tri is not None:
        return None
    (na, la), (nb, lb) = op.ins
    lc = op.out[1]
    pa = [(la[p], la[q]) for p, q in antisym_pairs(na) if la[p] in lc and la[q] in lc]
    pb = [(lb[p], lb[q]) for p, q in antisym_pairs(nb) if lb[p] in lc and lb[q] in lc]
    if len(pa) != 1 or len(pb) != 1 or set(pa[0]) | set(pb[0]) != set(lc):
        return None
    if len(set(pa[0]) | set(pb[0])) != 4:
        return None
    order = lambda pr: tuple(sorted(pr, key=lc.index))      # noqa: E731
    return order(pa[0]), order(pb[0])


def antisym_outputs(rops):
    """Rewrite every contraction whose result is antisymmetric in two output index pairs

        X[abij] += c A B      as      P[abij]  = c A B   restricted to a < b, i < j
                                      X[abij] += P[abij] - P[baij] - P[abji] + P[baji]

    P is a plan-owned scratch block that is zero outside the triangle (allocated zeroed; only the
    triangle is ever written), so the four permuted adds -- one fused elementwise pass -- restore
    the full antisymmetric contribution.  The contraction runs on m(m-1)/2 rows and columns
    instead of m^2: with the x < y sum over the contracted pair (ANTISYM) a same-spin ladder
    term does 1/8 of the reference's flops.  Relies on the operands being antisymmetric to
    rounding, like ANTISYM."""
    out, k = [], 0
    for op in rops:
        t = _tri_candidate(op) if ANTISYM_OUT else None
        if t is None:
            out.append(op)
            continue
        (x, y), (u, v) = t
        ls = op.out[1]
        suf = op.out[0].partition(".")[2]
        slot = "%s%d%s" % (TRI_PREFIX, k, "." + suf if suf else "")
        k += 1
        out.append(ROp((slot, ls), op.coef, list(op.ins), op.spin, tri=((x, y), (u, v))))

        def sw(s, p, q):
            return s.translate(str.maketrans(p + q, q + p))
        for pls, sg in ((ls, 1.0), (sw(ls, x, y), -1.0), (sw(ls, u, v), -1.0),
                        (sw(sw(ls, x, y), u, v), 1.0)):
            out.append(ROp(op.out, sg, [(slot, pls)], op.spin))
    return out


# ---------------------------------------------------------------------------
# closed shell: mirror-symmetric opposite-spin outputs, computed on half of their rows
# ---------------------------------------------------------------------------
MIRROR_OUT = int(_os_environ_get("KB200_MIRROR_OUT", "1"))
_SELF_MIRROR_AMP = ("t2", "tau", "tauh", "Woooo", "Wvvvv", "o2")


def _self_mirror(slot):
    """Opposite-spin blocks that equal their own alpha <-> beta image, X[p,Q,r,S] == X[Q,p,S,r],
    in a closed-shell (mirror_reduce'd) program: amplitude-like 'ab' blocks and the Iabab blocks
    whose pattern is its own image."""
    if "~" in slot or "@" in slot:
        return False
    base, _, suf = slot.partition(".")
    if base == "Iabab":
        return suf[0] == suf[1] and suf[2] == suf[3]
    return base in _SELF_MIRROR_AMP and suf == "ab"


def _mirror_candidate(op):
    """(p, q): the leading output pair of an opposite-spin contraction X.ab[p,q,r,s] += A B whose
    value is unchanged under the joint swap p <-> q, r <-> s: every operand is its own mirror
    image and every operand's index pairs (0,1), (2,3) are {p,q}, {r,s} or the summed pair."""
    if len(op.ins) != 2 or len(op.out[1]) != 4 or op.tri is not None:
        return None
    if not _self_mirror(op.out[0]) or not all(_self_mirror(sl) and len(ls) == 4 for sl, ls in op.ins):
        return None
    lc = op.out[1]
    summed = set(l for _, ls in op.ins for l in ls) - set(lc)
    pairs = [frozenset(lc[:2]), frozenset(lc[2:]), frozenset(summed)]
    if len(summed) != 2:
        return None
    for _, ls in op.ins:
        if frozenset(ls[:2]) not in pairs or frozenset(ls[2:]) not in pairs:
            return None
    # restrict the output pair carried by the FIRST operand: it becomes the row index of the
    # contraction, so the diagonal pass (m rows x m^2 columns) runs on the big-tile kernel's
    # ragged-row path instead of a 33-column GEMM
    first = op.ins[0][1]
    for pr in (lc[:2], lc[2:]):
        if set(pr) <= set(first):
            return pr[0], pr[1]
    return None


def mirror_outputs(rops):
    """Closed-shell programs only (after mirror_reduce).  An opposite-spin ladder-type contraction
    between self-mirror blocks gives a block with X[p,q,r,s] == X[q,p,s,r]:

        X[pqrs] += c A B     becomes     P[pqrs]  = c A B   on the rows p < q   (P zero elsewhere)
                                         X[pqrs] += P[pqrs] + P[qpsr]
                                         X[pqrs] += c A B   on the rows p == q

    i.e. m(m-1)/2 + m of the m^2 rows are contracted."""
    out, k = [], 0
    for op in rops:
        t = _mirror_candidate(op) if MIRROR_OUT else None
        if t is None:
            out.append(op)
            continue
        p, q = t
        ls = op.out[1]
        slot = "%sm%d.ab" % (TRI_PREFIX, k)
        k += 1
        out.append(ROp((slot, ls), op.coef, list(op.ins), op.spin, tri=((p, q, "lt"),)))
        out.append(ROp(op.out, 1.0, [(slot, ls)], op.spin))
        out.append(ROp(op.out, 1.0, [(slot, ls[1] + ls[0] + ls[3] + ls[2])], op.spin))
        out.append(ROp(op.out, op.coef, list(op.ins), op.spin, tri=((p, q, "eq"),)))
    return out


# ---------------------------------------------------------------------------
# hybrid partition: one grid point evaluated by several ranks
# ---------------------------------------------------------------------------
HYBRID_MIN_WORK = 1 << 24      # M*N*K from which a contraction is dealt out over the ranks
DIST_SUFFIX = "^"


class HybridProgram(object):
    """Result of hybrid_phases: phases[p] = resolved ops of phase p; after phase p the
    distributed buffers exchange[p] are summed over the ranks (one all-reduce: the engine lays
    them out back to back); dslots = every distributed buffer (zero-filled before each run);
    shapes = slot shapes including the distributed companions."""
    def __init__(self, phases, exchange, dslots, shapes):
        self.phases, self.exchange, self.dslots, self.shapes = phases, exchange, dslots, shapes


def _contraction_sizes(op, shapes):
    dims = {}
    for slot, ls in [op.out] + list(op.ins):
        for l, d in zip(ls, shapes[slot]):
            dims[l] = d
    (_, la), (_, lb) = op.ins
    lc = op.out[1]
    prod = lambda ls: int(numpy.prod([dims[l] for l in ls])) if ls else 1     # noqa: E731
    M = prod([l for l in la if l in lc])
    N = prod([l for l in lb if l in lc])
    K = prod([l for l in la if l in lb])
    if op.tri:
        for ent in op.tri:
            n = dims[ent[0]]
            rel = ent[2] if len(ent) > 2 else "lt"
            keep = (n*(n - 1)//2) if rel == "lt" else n
            if ent[0] in la and ent[0] in lc:
                M = M//(n*n)*keep
            else:
                N = N//(n*n)*keep
    return M, N, K


def hybrid_phases(rops, shapes, outputs, world, min_work=None, min_cols=64):
    """Rewrite a resolved program so that `world` ranks evaluate it TOGETHER at the same grid
    points (the partition inside a grid point, for grids with fewer points than
    twice the number of GPUs).

    * Every large contraction (M*N*K >= min_work) is marked `slab`: each rank computes a slab of
      its rows.  Its result is a DISTRIBUTED contribution: the true value is the sum over ranks.
    * A slot that receives distributed contributions gets a companion buffer S^ for them (or is
      itself the distributed buffer when nothing else is written to it); everything else --
      the m^5 and elementwise terms, small next to the m^6 ones -- is evaluated by every rank
      (replicated), so replicated and distributed parts never mix in one buffer.
    * Ops that are linear in a slot with a pending distributed part and do not contract it
      (index-permuted adds, the four antisymmetric images, the triangle expansions) are applied
      to both parts: the distributed part propagates without communication.
    * A contraction needs its operands complete: the pending distributed buffers it reads are
      summed over the ranks first (all-reduce) and merged (S += S^).  The program is cut into
      phases at these points; the outputs are completed by a final exchange.

    For the FT-CCSD residual this gives two exchanges per evaluation: the W intermediates, then
    the doubles residual.  Accumulation orders change (sums over ranks), results agree with the
    single-rank program to rounding."""
    if min_work is None:
        min_work = HYBRID_MIN_WORK
    shapes = OrderedDict(shapes)
    outputs = list(outputs)

    def is_big(op):
        if len(op.ins) != 2 or not (set(op.ins[0][1]) & set(op.ins[1][1])):
            return False
        M, N, K = _contraction_sizes(op, shapes)
        return M*N*K >= min_work and min(M, N) > min_cols and max(M, N) >= 8*world

    big = [is_big(op) for op in rops]

    def sweep(rw):
        """One pass over the program.  rw: slots known to receive replicated writes (these get
        a companion buffer for their distributed part); None = dry run that finds them."""
        dry = rw is None
        hasR, dbuf, pending, read = set(), {}, set(), set()
        wphase = {}                 # buffer -> phase of its last writer
        phase_ops, merges, exch = [], [], {}
        rwrites = set(outputs)      # outputs are completed into the caller's buffers

        def emit(phase, op, target):
            phase_ops.append((phase, len(phase_ops), op))
            wphase[target] = max(wphase.get(target, 0), phase)

        def avail(buf):
            return wphase.get(buf, 0)

        def dname(slot):
            if slot not in dbuf:
                if dry or slot in rw:
                    dbuf[slot] = slot + DIST_SUFFIX
                    shapes[dbuf[slot]] = shapes[slot]
                else:
                    dbuf[slot] = slot
            return dbuf[slot]

        def complete(slot):
            """Phase from which the complete value of `slot` can be read; schedules the
            exchange of its pending distributed part."""
            if slot not in pending:
                return avail(slot)
            d = dbuf[slot]
            p = avail(d)
            exch.setdefault(p, []).append(d)
            pending.discard(slot)
            if d != slot:
                ph = max(p + 1, avail(slot))
                merges.append((ph, slot, d))
                wphase[slot] = ph
            else:
                wphase[slot] = p + 1
            hasR.add(slot)
            return wphase[slot]

        for op, b in zip(rops, big):
            target = op.out[0]
            if target in read:
                raise ValueError("hybrid partition: %s is written after it was read: %r"
                                 % (target, op))
            contracted = len(op.ins) == 2 and bool(set(op.ins[0][1]) & set(op.ins[1][1]))
            if contracted:
                ph = 0
                for sl, _ in op.ins:
                    ph = max(ph, complete(sl))
                    read.add(sl)
                if b:
                    d = dname(target)
                    pending.add(target)
                    emit(max(ph, avail(d)), ROp((d, op.out[1]), op.coef, list(op.ins), op.spin,
                                                op.tri, slab=True), d)
                else:
                    hasR.add(target)
                    rwrites.add(target)
                    emit(max(ph, avail(target)), op, target)
                continue
            # no contracted index: linear in each operand
            pend_in = [k for k, (sl, _) in enumerate(op.ins) if sl in pending]
            if len(pend_in) > 1:
                for k in pend_in[1:]:                   # a product of two incomplete tensors
                    complete(op.ins[k][0])
                pend_in = pend_in[:1]
            for sl, _ in op.ins:
                read.add(sl)
            if not pend_in:
                ph = max([avail(sl) for sl, _ in op.ins] + [avail(target)])
                hasR.add(target)
                rwrites.add(target)
                emit(ph, op, target)
                continue
            k = pend_in[0]
            src = op.ins[k][0]
            d_src = dbuf[src]
            others = [x for j, x in enumerate(op.ins) if j != k]
            if d_src != src and src in hasR:
                # replicated part of the source -> replicated part of the target
                ph = max([avail(src), avail(target)] + [avail(sl) for sl, _ in others])
                hasR.add(target)
                rwrites.add(target)
                emit(ph, ROp(op.out, op.coef, list(op.ins), op.spin, op.tri), target)
            d = dname(target)
            pending.add(target)
            ins = list(op.ins)
            ins[k] = (d_src, op.ins[k][1])
            ph = max([avail(d_src), avail(d)] + [avail(sl) for sl, _ in others])
            emit(ph, ROp((d, op.out[1]), op.coef, ins, op.spin, op.tri), d)
        for sl in outputs:
            if sl in pending:
                complete(sl)
        return rwrites, phase_ops, merges, exch, dbuf

    rw, _, _, _, _ = sweep(None)
    for k in [k for k in shapes if k.endswith(DIST_SUFFIX)]:
        del shapes[k]
    _, phase_ops, merges, exch, dbuf = sweep(rw)
    nph = 1 + max([p for p, _, _ in phase_ops] + [p for p, _, _ in merges] + [0])
    phases = [[] for _ in range(nph)]
    for ph, slot, d in merges:
        ls = "abcdefgh"[:len(shapes[slot])]
        phases[ph].append(ROp((slot, ls), 1.0, [(d, ls)], None))
    for ph, _, op in sorted(phase_ops, key=lambda x: (x[0], x[1])):
        phases[ph].append(op)
    exchange = [exch.get(p, []) for p in range(nph)]
    order = [d for p in range(nph) for d in exchange[p]]
    dslots = order + [d for d in dict.fromkeys(dbuf.values()) if d not in order]
    return HybridProgram(phases, exchange, dslots, shapes)


# ---------------------------------------------------------------------------
# reverse mode (Lambda / RDM)
# ---------------------------------------------------------------------------
def adjoint(rops, wrt, seeds, bar=lambda s: s + "~"):
    """Reverse-mode sweep over resolved ops.

    wrt:   predicate(slot name) -> True if the adjoint of that slot is needed
           (i.e. the slot depends on the differentiation variables or is one).
    seeds: slot names whose adjoints are supplied by the caller.
    Returns the list of adjoint ROps (in execution order); adjoint slots are
    named bar(slot)."""
    out = []
    for op in reversed(rops):
        cbar = (bar(op.out[0]), op.out[1])
        if not (wrt(op.out[0]) or op.out[0] in seeds):
            continue
        for k, (slot, ls) in enumerate(op.ins):
            if not wrt(slot):
                continue
            others = [x for j, x in enumerate(op.ins) if j != k]
            out.append(ROp((bar(slot), ls), op.coef, [cbar] + others, op.spin))
    return out


# ---------------------------------------------------------------------------
# lowering to kb200_op
# ---------------------------------------------------------------------------
class kb200_op(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("a", ctypes.c_int32), ("b", ctypes.c_int32),
                ("c", ctypes.c_int32),
                ("a_off", ctypes.c_int64), ("b_off", ctypes.c_int64), ("c_off", ctypes.c_int64),
                ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
                ("batch", ctypes.c_int32),
                ("bsA", ctypes.c_int64), ("bsB", ctypes.c_int64), ("bsC", ctypes.c_int64),
                ("tAm", ctypes.c_int64), ("tAk", ctypes.c_int64), ("tBk", ctypes.c_int64),
                ("tBn", ctypes.c_int64), ("tCm", ctypes.c_int64), ("tCn", ctypes.c_int64),
                ("alpha", ctypes.c_double), ("beta", ctypes.c_double),
                ("a_mode", ctypes.c_int32), ("b_mode", ctypes.c_int32),
                ("tile", ctypes.c_int32), ("splitk", ctypes.c_int32),
                ("group", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("ldA", ctypes.c_int64), ("ldB", ctypes.c_int64)]


def _strides(shape):
    st = [1] * len(shape)
    for k in range(len(shape) - 2, -1, -1):
        st[k] = st[k + 1] * shape[k + 1]
    return st


def padded_strides(shape):
    """Strides of a plan-owned 4-index block: the trailing index pair is stored with an EVEN
    pitch (1089 -> 1090 doubles at 33 orbitals), so that every row of the [p q][r s] matrix view
    starts on a 16-byte boundary -- what TMA tensor maps and 16-byte copies need; with an odd
    orbital count the natural strides are all 8 mod 16 bytes.  -> (strides, elements)."""
    if len(shape) != 4:
        return _strides(shape), int(numpy.prod(shape)) if len(shape) else 1
    P = shape[2]*shape[3]
    P += P & 1
    return [shape[1]*P, P, shape[3], 1], shape[0]*shape[1]*P


def _affine(group):
    """[(dim, stride)] of a composite index -> stride of its last letter when the composite
    value addresses memory affinely (each stride = next dim * next stride), else None."""
    if not group:
        return None
    for (d0, s0), (d1, s1) in zip(group[:-1], group[1:]):
        if s0 != d1*s1:
            return None
    return group[-1][1]


class TableBank(object):
    """Deduplicated uint32 offset tables, concatenated into one buffer."""
    def __init__(self):
        self.chunks = []
        self.pos = 0
        self.index = {}

    def get(self, dims_strides):
        key = tuple(dims_strides)
        if key in self.index:
            return self.index[key]
        off = numpy.zeros((1,), dtype=numpy.int64)
        for d, s in dims_strides:
            off = (off[:, None] + (numpy.arange(d, dtype=numpy.int64) * s)[None, :]).reshape(-1)
        assert off.max(initial=0) < 2 ** 32 - 1      # 0xffffffff marks 'no element' in the kernels
        start = self.pos
        self.chunks.append(off.astype(numpy.uint32))
        self.pos += off.size
        self.index[key] = start
        return start

    def get_lt(self, dims_strides, px, py, rel="lt"):
        """As get(), keeping only the entries whose index at position px is smaller than the
        one at position py (the x < y half of an antisymmetric contracted pair); rel="eq" keeps
        the diagonal x == y instead."""
        key = (rel, px, py) + tuple(dims_strides)
        if key in self.index:
            return self.index[key]
        dims = [d for d, _ in dims_strides]
        grids = numpy.meshgrid(*[numpy.arange(d, dtype=numpy.int64) for d in dims], indexing="ij")
        off = sum(g*s_ for g, (_, s_) in zip(grids, dims_strides))
        keep = (grids[px] < grids[py]) if rel == "lt" else (grids[px] == grids[py])
        off = off[keep].reshape(-1)          # C order of the full index space, filtered
        assert off.max(initial=0) < 2 ** 32 - 1
        start = self.pos
        self.chunks.append(off.astype(numpy.uint32))
        self.pos += off.size
        self.index[key] = start
        return start

    def buffer(self):
        if not self.chunks:
            return numpy.zeros((1,), dtype=numpy.uint32)
        return numpy.concatenate(self.chunks)


def slab_rows(M, rank, world):
    """Rows [lo, hi) of an M-row contraction dealt to `rank` of `world`: boundaries on multiples
    of 8 rows (one DMMA row group)."""
    g = (M + 7)//8
    lo = min(M, (g*rank//world)*8)
    hi = min(M, (g*(rank + 1)//world)*8)
    return lo, hi


N_SM = 148
# CTA tile used for large contractions (see kb200.cu: tile ids); overridable for experiments
import os as _os
BIG_TILE = int(_os.environ.get("KB200_BIG_TILE", "2"))
RANKK = int(_os.environ.get("KB200_RANKK", "1"))
# re-laid-out integral copies for long-K contractions, so that both operands of the
# whole-output-per-CTA kernel (tile 6) stream contiguously along k
DERIVE = int(_os.environ.get("KB200_DERIVE", "1"))
DERIVE_BIG = int(_os.environ.get("KB200_DERIVE_BIG", "1"))
# plain integral operands of m^6 contractions with an odd row pitch: contract an aligned copy
PAD_INTEGRALS = int(_os.environ.get("KB200_PAD_INTEGRALS", "1"))
# fold the tau batch into the N index of matrix-vector-like contractions with a tau-independent
# (integral) A operand: the integral block is then read once instead of once per grid point
FOLD = int(_os.environ.get("KB200_FOLD", "1"))
LONGK_MIN = 4096      # contracted length from which the long-K path (derived layouts) is used
RANKK_MIN_M = 4096    # rows from which the rank-K streaming kernel is used
# independent contractions of one kernel configuration launched together (kb200_gemm.cuh
# MAX_GROUP); 1 disables grouping and keeps the program order
MAX_GROUP = int(_os.environ.get("KB200_GROUP", "8"))
# accumulations into one slot commute in the schedule (0: they keep their program order)
COMMUTE_ACC = _os.environ.get("KB200_COMMUTE_ACC", "1") != "0"
# consecutive index-permuted sums / outer products into one output fused into one pass (kind 3)
FUSE_EW = int(_os.environ.get("KB200_FUSE", "1"))
# contracted index pairs in which both operands are antisymmetric are summed over x < y only
# (twice the half sum): the same-spin ladder terms then do half the flops of the reference's
# full double sum.  Relies on the amplitudes being antisymmetric, which the equations preserve.
ANTISYM = int(_os.environ.get("KB200_ANTISYM", "1"))
_AMP2_BASES = ("t2", "tau", "tauh", "Woooo", "Wvvvv", "l2")


def antisym_pairs(slot):
    """Index-position pairs under which the tensor in `slot` changes sign, known from its class
    (programs.tensor_defs): same-spin amplitudes / ladder intermediates and same-spin
    antisymmetrised integral blocks.  Adjoint and derived slots are not claimed."""
    if "~" in slot or "@" in slot:
        return ()
    base, _, suf = slot.partition(".")
    if base in ("I", "Ia", "Ib"):
        return tuple(pq for pq in ((0, 1), (2, 3)) if suf[pq[0]] == suf[pq[1]])
    if base in _AMP2_BASES and suf in ("", "aa", "bb"):
        return ((0, 1), (2, 3))
    return ()
EW_MAX_TERMS = 6
LONGK_TILE = int(_os.environ.get("KB200_LONGK", "1"))
SKINNY_TILE = int(_os.environ.get("KB200_SKINNY", "1"))
_TILE_BN = {0: 128, 1: 32, 2: 128, 3: 128, 4: 64, 5: 64, 6: 40, 7: 40}
_TILE_BM = {0: 128, 1: 128, 2: 128, 3: 128, 4: 128, 5: 64, 6: 40, 7: 8}


class Lowered(object):
    """A resolved program lowered to kb200_op descriptors + offset tables."""

    def __init__(self, rops, slot_shapes, batched, preset=(), antisym=True, pad=()):
        """slot_shapes: slot -> per-tau-point shape; batched: slot -> bool;
        preset: slots that already hold data when the plan starts (inputs and
        accumulate-into outputs); every other slot is overwritten (beta=0) by
        its first write.  antisym=False: the amplitudes are NOT known to be antisymmetric (a
        caller-supplied guess that failed the check): contracted pairs are summed in full, as
        the reference does."""
        self.rops = rops
        self.antisym = bool(antisym)
        # pad: slots stored with padded_strides (plan-owned scratch; derived integral copies
        # always are)
        self.pad = set(pad)
        self.slot_names = list(slot_shapes.keys())
        self.slot_index = {nm: k for k, nm in enumerate(self.slot_names)}
        self.slot_shapes = slot_shapes
        self.batched = batched
        self.bank = TableBank()
        self.derived = OrderedDict()   # derived slot -> (source slot, axis permutation)
        self.descs = []
        self._nfold = {}          # id(desc) -> N-index (dim, stride) lists on B and C (batch folding)
        self._src = {}            # id(desc) -> (resolved op, beta)
        self._slab = set()        # id(desc) of the row-slabbed contractions
        self.flops = 0.0          # per tau point, executed (2*M*N*K)
        written = set(preset)
        for op in rops:
            beta = 1.0 if op.out[0] in written else 0.0
            written.add(op.out[0])
            for slot, _ in op.ins:
                if slot not in written:
                    raise ValueError("slot %s read before it is written: %r" % (slot, op))
            d = self._lower(op, beta)
            self._src[id(d)] = (op, beta)
            if getattr(op, "slab", False):
                if d.kind != 0:
                    raise ValueError("only contractions can be row-slabbed: %r" % op)
                self._slab.add(id(d))
            self.descs.append(d)
        self.tables = self.bank.buffer()
        self.groups = [[k] for k in range(len(self.descs))]
        if MAX_GROUP > 1:
            self._schedule()
            self.tables = self.bank.buffer()

    # -- fused elementwise chains (kind 3) -------------------------------------
    def _is_ew(self, d):
        """Index-permuted copy/sum or outer product: no contracted index, output not read."""
        if not FUSE_EW:
            return False
        op, _ = self._src[id(d)]
        out = set(op.out[1])
        return all(set(ls) <= out for _, ls in op.ins) and all(sl != op.out[0] for sl, _ in op.ins)

    def _lower_fused(self, items):
        """items: [(op, beta)] writing one slot, in program order -> kind-3 descriptors."""
        cslot = items[0][0].out[0]
        shp = self.slot_shapes[cslot]
        nd = len(shp)
        stc = self.strides_of(cslot)
        numel = lambda sl: int(numpy.prod(self.slot_shapes[sl]))  # noqa: E731
        bs = lambda sl: self.slot_size(sl) if self.batched[sl] else 0      # noqa: E731
        terms = []
        votes = {}
        for op, beta in items:
            pos = {l: k for k, l in enumerate(op.out[1])}
            ins = sorted(op.ins, key=lambda x: -numel(x[0]))       # X = the larger operand
            strides = []
            for sl, ls in ins:
                st = [0]*nd
                for l, v in zip(ls, self.strides_of(sl)):
                    st[pos[l]] = v
                strides.append(st)
            xs, xl = ins[0]
            fast = pos[xl[-1]]
            big = numel(xs) > 16384
            if big and fast != nd - 1:
                votes[fast] = votes.get(fast, 0) + 1
            terms.append((op, beta, ins, strides, fast, big))
        winner = max(votes, key=lambda a: votes[a]) if votes else None
        # n runs over the last index of C, or the last two when that keeps every big operand
        # readable along its own contiguous index
        n_axes = [nd - 1]
        if nd >= 3 and winner != nd - 2 and shp[nd - 1] < 64:
            n_axes = [nd - 2, nd - 1]
        m_axes = [a for a in range(nd) if a not in n_axes]
        if winner is not None and winner in m_axes:
            m_axes.remove(winner)
            m_axes.append(winner)
        tabm = lambda st: self.bank.get([(shp[a], st[a]) for a in m_axes])    # noqa: E731
        tabn = lambda st: self.bank.get([(shp[a], st[a]) for a in n_axes])    # noqa: E731
        M = int(numpy.prod([shp[a] for a in m_axes])) if m_axes else 1
        out = []
        for k, (op, beta, ins, strides, fast, big) in enumerate(terms):
            d = kb200_op()
            d.kind = 3
            d.a = self.slot_index[ins[0][0]]
            d.b = self.slot_index[ins[1][0]] if len(ins) > 1 else -1
            d.c = self.slot_index[cslot]
            d.M, d.N, d.K, d.batch, d.splitk = M, int(numpy.prod([shp[a] for a in n_axes])), 1, 1, 1
            d.bsA = bs(ins[0][0])
            d.bsB = bs(ins[1][0]) if len(ins) > 1 else 0
            d.bsC = bs(cslot)
            d.tAm, d.tAk = tabm(strides[0]), tabn(strides[0])
            if len(ins) > 1:
                d.tBk, d.tBn = tabm(strides[1]), tabn(strides[1])
            d.tCm, d.tCn = tabm(stc), tabn(stc)
            d.alpha = op.coef
            d.beta = beta if k == 0 else 1.0
            d.a_mode = 1 if (big and winner is not None and fast == winner and winner in m_axes) else 0
            self._src[id(d)] = (op, d.beta)
            out.append(d)
        return out

    # -- launch grouping -----------------------------------------------------
    def _groupable(self, d):
        """Large-tile contractions: one launch is only a few waves of CTAs."""
        return d.kind == 0 and d.tile in (0, 2) and d.K >= 512

    def _schedule(self):
        """Reorder the ops along their dependency DAG so that independent large contractions
        of the same kernel configuration become adjacent, and record them as launch groups.
        Read-after-write, write-after-write and write-after-read orders on every slot are
        kept, so each output sees its accumulations in the program order (bit-identical
        results); scratch slots are never aliased, so no liveness changes."""
        n = len(self.descs)
        reads, writes, commut = [], [], []
        for d in self.descs:
            r = {d.a}
            if d.kind != 1:
                r.add(d.b)
            # X += ... with X not among the operands: such accumulations into one slot commute
            # (COMMUTE_ACC); they are still never in flight together (the runner orders every
            # write after write), only their order is left to the scheduler
            cm = bool(COMMUTE_ACC and d.beta == 1.0 and d.c not in r)
            if d.beta != 0.0 and not cm:
                r.add(d.c)
            reads.append(r)
            writes.append(d.c)
            commut.append(cm)
        npred = [0]*n
        succ = [[] for _ in range(n)]
        base = {}        # slot -> last op that overwrote it (or updated it non-commutatively)
        acc = {}         # slot -> commuting accumulations since then
        readers = {}     # slot -> ops that read it since then
        for i in range(n):
            pre = set()
            for s_ in reads[i]:
                if s_ in base:
                    pre.add(base[s_])
                pre.update(acc.get(s_, ()))
            w = writes[i]
            if w in base:
                pre.add(base[w])
            pre.update(readers.get(w, ()))
            if not commut[i]:
                pre.update(acc.get(w, ()))
            pre.discard(i)
            for j in pre:
                succ[j].append(i)
            npred[i] = len(pre)
            for s_ in reads[i]:
                readers.setdefault(s_, []).append(i)
            if commut[i]:
                acc.setdefault(w, []).append(i)
            else:
                base[w] = i
                acc[w] = []
                readers[w] = []
        ready = sorted(i for i in range(n) if npred[i] == 0)
        order, groups, chains = [], [], []

        def retire(i):
            for j in succ[i]:
                npred[j] -= 1
                if npred[j] == 0:
                    ready.append(j)

        while ready:
            ready.sort()
            small = [i for i in ready if not self._groupable(self.descs[i])]
            if small:
                # cheap ops first, in program order: they unlock more large contractions
                i = small[0]
                ready.remove(i)
                chain = [i]
                retire(i)
                if self._is_ew(self.descs[i]):
                    # following elementwise updates of the same output join its pass
                    while len(chain) < EW_MAX_TERMS:
                        cand = [j for j in ready if writes[j] == writes[i] and self._is_ew(self.descs[j])]
                        if not cand:
                            break
                        j = min(cand)
                        ready.remove(j)
                        chain.append(j)
                        retire(j)
                    if len(chain) > 1 or self.descs[i].kind == 0:
                        chains.append(list(range(len(order), len(order) + len(chain))))
                groups.append(list(range(len(order), len(order) + len(chain))))
                order.extend(chain)
                continue
            # (the kernel takes the operand contiguity modes per member: only the tile must agree)
            lead = self.descs[ready[0]]
            grp = [i for i in ready if self.descs[i].tile == lead.tile]
            # one slot may be accumulated into by only one member of a launch
            seen, pick = set(), []
            for i in grp:
                if writes[i] in seen or len(pick) == MAX_GROUP:
                    continue
                seen.add(writes[i])
                pick.append(i)
            groups.append(list(range(len(order), len(order) + len(pick))))
            for i in pick:
                ready.remove(i)
                order.append(i)
            for i in pick:
                retire(i)
        assert len(order) == n
        self.descs = [self.descs[i] for i in order]
        self.rops = [self.rops[i] for i in order]
        self.groups = groups
        for pos in chains:
            fused = self._lower_fused([self._src[id(self.descs[k])] for k in pos])
            for k, d in zip(pos, fused):
                self.flops -= 2.0 * self.descs[k].M * self.descs[k].N * self.descs[k].K \
                    if self.descs[k].kind == 0 else 0.0
                self.descs[k] = d

    # -- helpers ---------------------------------------------------------
    def _dims(self, op):
        dims = {}
        for slot, ls in [op.out] + list(op.ins):
            shp = self.slot_shapes[slot]
            assert len(shp) == len(ls), (slot, shp, ls)
            for l, d in zip(ls, shp):
                if dims.setdefault(l, d) != d:
                    raise ValueError("dimension clash on %s in %r" % (l, op))
        return dims

    def strides_of(self, slot):
        if slot in self.pad:
            return padded_strides(self.slot_shapes[slot])[0]
        return _strides(self.slot_shapes[slot])

    def slot_size(self, slot):
        """Elements one grid point of the slot occupies (padding included)."""
        if slot in self.pad:
            return padded_strides(self.slot_shapes[slot])[1]
        return int(numpy.prod(self.slot_shapes[slot])) if len(self.slot_shapes[slot]) else 1

    def _stride_map(self, slot, ls):
        return dict(zip(ls, self.strides_of(slot)))

    def _derive(self, slot, letters, new_letters, pad_only=False):
        """Register a re-laid-out copy of a tau-independent integral block:
        new[new_letters] = slot[letters].  Materialised once per integral tensor by the
        engine (it is constant over iterations and grid points).  pad_only: a copy in the same
        index order, only for the padded (16-byte aligned) strides derived blocks are stored with."""
        perm = tuple(letters.index(l) for l in new_letters)
        if perm == tuple(range(len(perm))) and not pad_only:
            return slot
        name = "%s@%s" % (slot, "".join(str(p) for p in perm))
        if name not in self.derived:
            self.derived[name] = (slot, perm)
            shp = self.slot_shapes[slot]
            self.slot_shapes[name] = tuple(shp[p] for p in perm)
            self.pad.add(name)
            self.batched[name] = False
            self.slot_index[name] = len(self.slot_names)
            self.slot_names.append(name)
        return name

    @staticmethod
    def _order(group, primary, secondary):
        """Order a composite index: innermost = stride-1 letter of the primary
        operand when it belongs to the group, rest by decreasing stride."""
        g = sorted(group, key=lambda l: -primary.get(l, 0))
        if not g:
            return g
        # make sure a stride-1 letter of the secondary operand is not lost
        # when the primary has none in this group
        if primary.get(g[-1], 0) != 1:
            for l in g:
                if secondary.get(l, 0) == 1:
                    g.remove(l)
                    g.append(l)
                    break
        return g

    def _lower(self, op, beta):
        d = kb200_op()
        dims = self._dims(op)
        sc = self._stride_map(*op.out)
        size = lambda grp: int(numpy.prod([dims[l] for l in grp])) if grp else 1  # noqa: E731
        tab = lambda grp, smap: self.bank.get([(dims[l], smap[l]) for l in grp])  # noqa: E731
        bs = lambda slot: self.slot_size(slot) if self.batched[slot] else 0  # noqa: E731
        d.c = self.slot_index[op.out[0]]
        d.bsC = bs(op.out[0])
        d.alpha, d.beta = op.coef, beta
        d.batch = 1
        d.splitk = 1
        if len(op.ins) == 1:
            (sa_name, la) = op.ins[0]
            sa = self._stride_map(sa_name, la)
            assert sorted(la) == sorted(op.out[1]), op
            lc = op.out[1]
            last_c = lc[-1]
            last_a = la[-1]
            ngrp = [last_c]
            if len(lc) >= 2 and lc[-2] != last_a and len(lc) > 2:
                ngrp = [lc[-2], last_c]
            mgrp = [l for l in lc if l not in ngrp]
            if last_a in mgrp:
                mgrp.remove(last_a)
                mgrp.append(last_a)
            d.kind = 1
            d.a = self.slot_index[sa_name]
            d.b = 0
            d.M, d.N, d.K = size(mgrp), size(ngrp), 1
            d.bsA = bs(sa_name)
            d.tAm, d.tAk = tab(mgrp, sa), tab(ngrp, sa)
            d.tCm, d.tCn = tab(mgrp, sc), tab(ngrp, sc)
            d.a_mode = 0 if last_a in ngrp else 1
            d.b_mode = 0
            return d
        (na, la), (nb, lb) = op.ins
        lc = op.out[1]
        M = [l for l in la if l in lc]
        N = [l for l in lb if l in lc]
        K = [l for l in la if l in lb]
        if set(M) & set(N) or set(K) & set(lc) or len(M) + len(K) != len(la) \
                or len(N) + len(K) != len(lb) or len(M) + len(N) != len(lc):
            raise ValueError("unsupported contraction pattern %r" % op)
        if size(N) > size(M):
            (na, la), (nb, lb) = (nb, lb), (na, la)
            M, N = N, M
        sa, sb = self._stride_map(na, la), self._stride_map(nb, lb)
        # long-K contraction into a tiny output (F_vv / F_oo builds, singles residual, their
        # adjoints): both operands are n^4 tensors reduced over three indices.  The amplitude
        # operand streams from HBM in its natural order; the integral operand is replaced by
        # a copy laid out [row][k in the same order], so both sides are read contiguously.
        if DERIVE and size(M) <= 64 and size(N) <= 64 and size(K) >= LONGK_MIN \
                and is_integral_slot(na) != is_integral_slot(nb):
            if is_integral_slot(na):
                K = sorted(K, key=lambda l: -sb[l])
                new = "".join(M) + "".join(K)
                na, la = self._derive(na, la, new), new
            else:
                K = sorted(K, key=lambda l: -sa[l])
                new = "".join(N) + "".join(K)
                nb, lb = self._derive(nb, lb, new), new
            sa, sb = self._stride_map(na, la), self._stride_map(nb, lb)
        # m^6 contraction with an integral operand that is neither [rows][k] nor [k][rows] with k
        # in the other operand's order (the W_ovvo builds read I.oovv[m,n,e,f] with rows (m,e)
        # against k = (n,f)): contract a copy laid out [rows][k] instead (made once per solve),
        # so that both operands are plain matrices
        if DERIVE_BIG and min(size(M), size(N)) >= 128 and size(K) >= 256 \
                and is_integral_slot(na) != is_integral_slot(nb):
            odd = lambda sl: len(self.slot_shapes[sl]) == 4 and \
                (self.slot_shapes[sl][2]*self.slot_shapes[sl][3]) % 2 == 1     # noqa: E731
            if is_integral_slot(na):
                ko = [l for l in lb if l in K]
                if not _plain_order(la, M, ko):
                    new = "".join(M) + "".join(ko)
                    na, la = self._derive(na, la, new), new
                elif PAD_INTEGRALS and odd(na):
                    na = self._derive(na, la, la, pad_only=True)
            else:
                ko = [l for l in la if l in K]
                if not _plain_order(lb, N, ko):
                    new = "".join(N) + "".join(ko)
                    nb, lb = self._derive(nb, lb, new), new
                elif PAD_INTEGRALS and odd(nb):
                    nb = self._derive(nb, lb, lb, pad_only=True)
            sa, sb = self._stride_map(na, la), self._stride_map(nb, lb)
            K = [l for l in la if l in lb]
        a_mode = 0 if la[-1] in K else 1
        b_mode = 0 if lb[-1] in K else 1
        # order of the contracted composite index: follow the operand that wants to be
        # read contiguously along k; if both do, favour the tau-batched one (it streams
        # from HBM once, whereas a tau-independent integral block stays L2-resident and
        # tolerates a strided gather)
        if a_mode == 0 and b_mode == 0:
            # (a nominal 8 grid points for a batched operand: a batched vector against an
            # integral matrix still follows the matrix)
            ea = size(M) * (8 if bs(na) != 0 else 1)
            eb = size(N) * (8 if bs(nb) != 0 else 1)
            a_first = ea >= eb
        else:
            a_first = (a_mode == 0) or (b_mode != 0)
        K = self._order(K, sa, sb) if a_first else self._order(K, sb, sa)
        # HBM-bound rank-K update (small K and N, ~n^3 rows): the streaming SIMT kernel owns
        # one C row per thread, so C (2/3 of the traffic) must be contiguous along the rows
        rankk = (RANKK and size(K) <= 64 and size(N) <= 64 and size(M) >= RANKK_MIN_M
                 and lc[-1] in M)
        if rankk:
            M = self._order(M, sc, sa)
        else:
            M = self._order(M, sa, sc) if a_mode == 1 else self._order(M, sc, sa)
        N = self._order(N, sb, sc) if b_mode == 1 else self._order(N, sc, sb)
        if not rankk and size(M) >= 128 and size(K) >= 256:
            # large contractions: a row order that makes the operand a plain matrix (rows
            # strided uniformly) beats following the output's order -- C is scattered anyway
            Ma = sorted(M, key=lambda l: -sa[l])
            if _affine([(dims[l], sa[l]) for l in Ma]) is not None:
                M = Ma
            Nb = sorted(N, key=lambda l: -sb[l])
            if size(N) >= 64 and _affine([(dims[l], sb[l]) for l in Nb]) is not None:
                N = Nb
        d.kind = 0
        d.a, d.b = self.slot_index[na], self.slot_index[nb]
        d.M, d.N, d.K = size(M), size(N), size(K)
        d.bsA, d.bsB = bs(na), bs(nb)
        d.tAm, d.tBn = tab(M, sa), tab(N, sb)
        tri_m = tri_n = None
        for ent in (op.tri or ()):
            x, y = ent[0], ent[1]
            rel = ent[2] if len(ent) > 2 else "lt"
            if dims[x] != dims[y]:
                raise ValueError("restricted output pair (%s,%s) of unequal ranges: %r" % (x, y, op))
            if x in M and y in M and tri_m is None:
                tri_m = (M.index(x), M.index(y), rel)
            elif x in N and y in N and tri_n is None:
                tri_n = (N.index(x), N.index(y), rel)
            else:
                raise ValueError("restricted output pair (%s,%s) straddles rows and columns: %r"
                                 % (x, y, op))

        def kept(grp, t):
            nx = dims[grp[t[0]]]
            return size(grp) // (nx * nx) * ((nx * (nx - 1) // 2) if t[2] == "lt" else nx)
        if tri_m is not None:
            d.tAm = self.bank.get_lt([(dims[l], sa[l]) for l in M], *tri_m)
            d.M = kept(M, tri_m)
        if tri_n is not None:
            d.tBn = self.bank.get_lt([(dims[l], sb[l]) for l in N], *tri_n)
            d.N = kept(N, tri_n)
        half = None
        if ANTISYM and self.antisym and len(K) >= 2:
            for pa, qa in antisym_pairs(na):
                x, y = la[pa], la[qa]
                if x in K and y in K and any({lb[pb], lb[qb]} == {x, y} for pb, qb in antisym_pairs(nb)):
                    half = (K.index(x), K.index(y))
                    break
        if half is not None and "@" not in na and "@" not in nb:
            d.tAk = self.bank.get_lt([(dims[l], sa[l]) for l in K], *half)
            d.tBk = self.bank.get_lt([(dims[l], sb[l]) for l in K], *half)
            nx = dims[K[half[0]]]
            d.K = size(K) // (nx * nx) * (nx * (nx - 1) // 2)
            d.alpha = 2.0 * op.coef
        else:
            d.tAk, d.tBk = tab(K, sa), tab(K, sb)
        d.tCm, d.tCn = tab(M, sc), tab(N, sc)
        if tri_m is not None:
            d.tCm = self.bank.get_lt([(dims[l], sc[l]) for l in M], *tri_m)
        if tri_n is not None:
            d.tCn = self.bank.get_lt([(dims[l], sc[l]) for l in N], *tri_n)
        d.a_mode, d.b_mode = a_mode, b_mode
        # plain operands: the composite row and k indices address memory affinely and the
        # contiguous one has unit stride -> ld = pitch of the other one (0: gathered).  The kernels
        # may then use wide copies / TMA tensor maps instead of the offset tables.
        d.ldA = d.ldB = 0
        if half is None:
            gm, gk = [(dims[l], sa[l]) for l in M], [(dims[l], sa[l]) for l in K]
            fm, fk = _affine(gm), _affine(gk)
            if tri_m is None and fm is not None and fk is not None:
                if a_mode == 0 and fk == 1:
                    d.ldA = fm
                elif a_mode == 1 and fm == 1:
                    d.ldA = fk
            gn, gk = [(dims[l], sb[l]) for l in N], [(dims[l], sb[l]) for l in K]
            fn, fk = _affine(gn), _affine(gk)
            if tri_n is None and fn is not None and fk is not None:
                if b_mode == 0 and fk == 1:
                    d.ldB = fn
                elif b_mode == 1 and fn == 1:
                    d.ldB = fk
        # (batch folding rebuilds the N tables from these lists: not for triangular columns)
        self._nfold[id(d)] = None if tri_n is not None else \
            ([(dims[l], sb[l]) for l in N], [(dims[l], sc[l]) for l in N])
        if SKINNY_TILE and d.K <= 40 and d.N <= 40 and d.M >= RANKK_MIN_M:
            d.tile = 7       # skinny streaming update (DMMA, 8 rows per warp)
            d.reserved = 1 if lc[-1] in M else 0
        elif rankk:
            d.kind = 2
            d.tile = 0
        elif LONGK_TILE and d.M <= 40 and d.N <= 40 and d.K >= 2048:
            d.tile = 6       # whole output per CTA, split over K only
        elif d.N <= 32:
            d.tile = 1 if d.M > 64 else 5
        elif d.N <= 64:
            d.tile = 5
        else:
            d.tile = BIG_TILE
        self.flops += 2.0 * d.M * d.N * d.K
        return d

    def finalize(self, nbatch, bstrides=None, part=None):
        """Set the tau batch and the split-K factors; return the ctypes array.
        bstrides: slot index -> distance (in doubles) between consecutive grid points of a
        batched slot, when it differs from the slot's size (rows of a wider buffer);
        part = (rank, world): row-slabbed contractions keep the rows of this rank only."""
        arr = (kb200_op * len(self.descs))()
        lead_size = {g[0]: len(g) for g in self.groups}
        gsize_of = {k: len(g) for g in self.groups for k in g}
        for k, d in enumerate(self.descs):
            ctypes.memmove(ctypes.byref(arr[k]), ctypes.byref(d), ctypes.sizeof(kb200_op))
            o = arr[k]
            o.group = lead_size.get(k, 0)
            if bstrides:
                if o.bsA != 0 and o.a in bstrides:
                    o.bsA = bstrides[o.a]
                if o.bsB != 0 and o.kind != 1 and o.b >= 0 and o.b in bstrides:
                    o.bsB = bstrides[o.b]
                if o.bsC != 0 and o.c in bstrides:
                    o.bsC = bstrides[o.c]
            if part is not None and id(d) in self._slab:
                r, w = part
                lo, hi = slab_rows(o.M, r, w)
                if hi <= lo:
                    raise ValueError("row slab of %d rows over %d ranks is empty" % (o.M, w))
                o.tAm += lo
                o.tCm += lo
                o.M = hi - lo
                o.ldA = 0                 # (the tensor map would have to start at row lo)
            o.batch = nbatch if (o.bsC != 0 or o.bsA != 0 or (o.kind in (0, 3) and o.bsB != 0)) else 1
            if o.batch > 1 and o.bsC == 0:
                raise ValueError("batched operands reduce into an unbatched output")
            if FOLD and o.kind == 0 and o.batch > 1 and o.bsA == 0 and o.bsB != 0 \
                    and o.N * o.batch <= 32 and o.tile != 6 and gsize_of[k] == 1 \
                    and self._nfold.get(id(d)) is not None:
                nB, nC = self._nfold[id(d)]
                o.tBn = self.bank.get([(o.batch, o.bsB)] + nB)
                o.tCn = self.bank.get([(o.batch, o.bsC)] + nC)
                o.N = o.N * o.batch
                o.ldB = 0
                o.batch, o.bsB, o.bsC = 1, 0, 0
                o.tile = 1 if o.M > 64 else 5
            if o.kind == 0:
                bm, bn = _TILE_BM[o.tile], _TILE_BN[o.tile]
                ctas = ((o.M + bm - 1) // bm) * ((o.N + bn - 1) // bn) * o.batch
                target = 4 * N_SM if o.tile == 5 else 2 * N_SM
                if gsize_of[k] > 1:
                    continue                      # launched with its group: never split
                if o.tile == 7:
                    continue
                if o.tile == 6:
                    # grid = (splitk, batch): one wave of two CTAs per SM, K chunks of >= 4 stages
                    best = max(1, min((2 * N_SM) // o.batch, o.K // 128))
                    o.splitk = best
                    continue
                if ctas < target // 2 and o.K >= 512:
                    o.splitk = int(min(max(1, -(-target // ctas)), max(1, o.K // 128)))
                elif o.tile in (0, 2) and ctas <= 2 * N_SM and o.K >= 512:
                    # few waves of one-CTA-per-SM tiles (small tau batches on a sharded run):
                    # pick the split that minimises wave quantisation, charging the
                    # deterministic reduction pass ~10 % of a wave per extra split (measured: splitting
                    # a 3-wave launch does not pay)
                    best, bests = None, 1
                    for sk in (1, 2, 3, 4):
                        cost = -(-ctas * sk // N_SM) / float(sk) + 0.10 * (sk - 1)
                        if best is None or cost < best - 1e-9:
                            best, bests = cost, sk
                    o.splitk = bests
        self.tables = self.bank.buffer()      # batch folding may have added tables
        return arr


# ---------------------------------------------------------------------------
# slot bookkeeping shared by the g and u back ends
# ---------------------------------------------------------------------------
_INT_SLOT = re.compile(r"^(I|Ia|Ib|Iabab|F|Fa|Fb)\.")


def _plain_order(ls, rows, ko):
    """Are the letters `ls` of an operand [rows...][ko] or [ko][rows...] (rows in any order, the
    contracted letters exactly in the order `ko`)?"""
    ls, ko, n = list(ls), list(ko), len(ko)
    return (ls[-n:] == ko and set(ls[:-n]) == set(rows)) or \
        (ls[:n] == ko and set(ls[n:]) == set(rows))


def is_integral_slot(slot):
    """True for the caller-supplied dressed-integral slots (not their adjoints)."""
    return (not slot.endswith("~")) and ("@" not in slot) and _INT_SLOT.match(slot) is not None


def slot_shapes(rops, mode, sizes):
    """Per-tau-point shape of every slot touched by `rops`.

    sizes: g -> {'o': no, 'v': nv};  u -> {('o','a'): noa, ('o','b'): nob, ...}."""
    shapes = OrderedDict()
    for op in rops:
        for slot, ls in [op.out] + list(op.ins):
            if mode == "g":
                shp = tuple(sizes[space_of(l)] for l in ls)
            else:
                shp = tuple(sizes[(space_of(l), op.spin[l])] for l in ls)
            if shapes.setdefault(slot, shp) != shp:
                raise ValueError("shape clash for slot %s" % slot)
    return shapes

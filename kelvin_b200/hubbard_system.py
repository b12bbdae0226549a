"""Hubbard model in a UHF-like mean-field basis (input generator).

Same constructor and on-path methods as kelvin/hubbard_system.py
(``HubbardSystem``, finite-temperature / chemical-potential branch): the basis
diagonalises the mean-field Fock matrices built from the user densities Pa, Pb
(:48-75); integrals are transformed to that basis (:571-593,658-663).
``Hubbard1D`` provides the model object the reference takes from the external
``lattice`` package.
"""
import numpy

from . import ft_utils
from .system import System


class Hubbard1D(object):
    """1-D Hubbard chain: hopping t, on-site U, boundary 'p'eriodic or 'o'pen."""
    def __init__(self, L, t, U, boundary='p'):
        self.L, self.t, self.U, self.boundary = L, t, U, boundary

    def get_tmatS(self):
        L = self.L
        T = numpy.zeros((L, L))
        idx = numpy.arange(L - 1)
        T[idx, idx + 1] = -self.t
        T[idx + 1, idx] = -self.t
        if self.boundary == 'p' and L > 2:
            T[0, L - 1] = T[L - 1, 0] = -self.t
        return T

    def get_umatS(self):
        L = self.L
        V = numpy.zeros((L, L, L, L))
        i = numpy.arange(L)
        V[i, i, i, i] = self.U
        return V

    def get_tmat(self):
        T = self.get_tmatS()
        Z = numpy.zeros_like(T)
        return numpy.block([[T, Z], [Z, T]])

    def get_umat(self):
        L = self.L
        V = numpy.zeros((2*L,)*4)
        i = numpy.arange(L)
        V[i, L + i, i, L + i] = self.U
        V[L + i, i, L + i, i] = self.U
        return V


def _transform(V, u1, u2, u3, u4):
    X = numpy.einsum('ijkl,ls->ijks', V, u4)
    X = numpy.einsum('ijks,kr->ijrs', X, u3)
    X = numpy.einsum('ijrs,jq->iqrs', X, u2)
    return numpy.einsum('iqrs,ip->pqrs', X, u1)


class HubbardSystem(System):
    def __init__(self, T, model, Pa=None, Pb=None, mu=None, na=None, nb=None,
                 ua=None, ub=None, orbtype='u'):
        if na is not None or nb is not None or mu is None:
            raise Exception("kelvin_b200.HubbardSystem supports the finite-temperature (mu) path only")
        if Pa is None or Pb is None:
            raise Exception("No reference provided")
        self.T = T
        self.model = model
        self.Pa, self.Pb = Pa, Pb
        self.orbtype = orbtype
        self.mu = mu
        self.na = self.nb = None
        self.beta = 1.0/self.T if self.T > 0.0 else 1.0e20
        V = model.get_umatS()
        Va = V - V.transpose((0, 1, 3, 2))
        Fa = model.get_tmatS().copy()
        Fb = model.get_tmatS().copy()
        Fa += numpy.einsum('pqrs,qs->pr', Va, Pa)
        Fa += numpy.einsum('pqrs,qs->pr', V, Pb)
        Fb += numpy.einsum('pqrs,qs->pr', Va, Pb)
        Fb += numpy.einsum('pqrs,pr->qs', V, Pa)
        self.Fa, self.Fb = Fa, Fb
        if ua is None:
            assert(ub is None)
            self.ea, self.ua = numpy.linalg.eigh(Fa)
            self.eb, self.ub = numpy.linalg.eigh(Fb)
        else:
            self.ua, self.ub = ua, ub
            self.ea = numpy.einsum('ij,ip,jq->pq', Fa, ua, ua).diagonal()
            self.eb = numpy.einsum('ij,ip,jq->pq', Fb, ua, ub).diagonal()
        self._cache = {}

    def has_g(self):
        return True

    def has_u(self):
        return self.orbtype != 'g'

    def has_r(self):
        return False

    def verify(self, T, mu):
        if T > 0.0:
            return T == self.T and mu == self.mu
        return T == self.T

    def const_energy(self):
        return 0.0

    def u_energies_tot(self):
        return self.ea, self.eb

    def g_energies_tot(self):
        return numpy.hstack((self.ea, self.eb))

    def u_aint_tot(self):
        if "u" not in self._cache:
            V = self.model.get_umatS()
            Va = V - V.transpose((0, 1, 3, 2))
            Vabab = _transform(V, self.ua, self.ub, self.ua, self.ub)
            Vb = _transform(Va, self.ub, self.ub, self.ub, self.ub)
            Vaa = _transform(Va, self.ua, self.ua, self.ua, self.ua)
            self._cache["u"] = (Vaa, Vb, Vabab)
        return self._cache["u"]

    def g_aint_tot(self):
        if "g" not in self._cache:
            Us = self.model.get_umatS()
            n = Us.shape[0]
            U = numpy.zeros((2*n,)*4)
            U[n:, :n, n:, :n] = Us
            U[:n, n:, n:, :n] = -Us
            U[:n, n:, :n, n:] = Us
            U[n:, :n, :n, n:] = -Us
            utot = self._utot()
            self._cache["g"] = _transform(U, utot, utot, utot, utot)
        return self._cache["g"]

    def g_int_tot(self):
        utot = self._utot()
        return _transform(self.model.get_umat(), utot, utot, utot, utot)

    def _utot(self):
        n = self.ua.shape[0]
        Z = numpy.zeros((n, n))
        return numpy.block([[self.ua, Z], [Z, self.ub]])

    def get_mp1(self):
        """kelvin/hubbard_system.py:124-137."""
        Va, Vb, Vabab = self.u_aint_tot()
        ea, eb = self.u_energies_tot()
        foa = ft_utils.ff(self.beta, ea, self.mu)
        fob = ft_utils.ff(self.beta, eb, self.mu)
        E1 = -0.5*numpy.einsum('ijij,i,j->', Va, foa, foa)
        E1 -= 0.5*numpy.einsum('ijij,i,j->', Vb, fob, fob)
        E1 -= numpy.einsum('ijij,i,j->', Vabab, foa, fob)
        Fa, Fb = self.u_fock_tot()
        E1 += numpy.einsum('ii,i->', Fa - numpy.diag(ea), foa)
        E1 += numpy.einsum('ii,i->', Fb - numpy.diag(eb), fob)
        return E1

    def u_fock_tot(self):
        """kelvin/hubbard_system.py:350-384 (finite T)."""
        Ts = self.model.get_tmatS()
        da, db = self.u_energies_tot()
        foa = ft_utils.ff(self.beta, da, self.mu)
        fob = ft_utils.ff(self.beta, db, self.mu)
        Va, Vb, Vabab = self.u_aint_tot()
        JKa = numpy.einsum('prqr,r->pq', Va, foa) + numpy.einsum('prqr,r->pq', Vabab, fob)
        JKb = numpy.einsum('prqr,r->pq', Vb, fob) + numpy.einsum('pqpr,p->qr', Vabab, foa)
        Fa = JKa + numpy.einsum('ij,ip,jq->pq', Ts, self.ua, self.ua)
        Fb = JKb + numpy.einsum('ij,ip,jq->pq', Ts, self.ub, self.ub)
        return Fa, Fb

    def g_fock_tot(self):
        """kelvin/hubbard_system.py:330-348 (finite T)."""
        T = self.model.get_tmat()
        d = self.g_energies_tot()
        fo = ft_utils.ff(self.beta, d, self.mu)
        V = self.g_aint_tot()
        JK = numpy.einsum('prqr,r->pq', V, fo)
        utot = self._utot()
        return JK + numpy.einsum('ij,ip,jq->pq', T, utot, utot)

    def u_mp1_den(self):
        """kelvin/hubbard_system.py:180-203."""
        Va, Vb, Vabab = self.u_aint_tot()
        beta = self.beta
        ea, eb = self.u_energies_tot()
        foa = ft_utils.ff(beta, ea, self.mu)
        veca = foa*ft_utils.ffv(beta, ea, self.mu)
        fob = ft_utils.ff(beta, eb, self.mu)
        vecb = fob*ft_utils.ffv(beta, eb, self.mu)
        Ts = self.model.get_tmatS()
        Ta = numpy.einsum('ij,ip,jq->pq', Ts, self.ua, self.ua)
        Tb = numpy.einsum('ij,ip,jq->pq', Ts, self.ub, self.ub)
        Da = -beta*numpy.einsum('ii,i->i', Ta - numpy.diag(ea), veca)
        Db = -beta*numpy.einsum('ii,i->i', Tb - numpy.diag(eb), vecb)
        Da += -beta*numpy.einsum('ijij,i,j->i', Va, veca, foa)
        Db += -beta*numpy.einsum('ijij,i,j->i', Vb, vecb, fob)
        Da += -beta*numpy.einsum('ijij,i,j->i', Vabab, veca, fob)
        Db += -beta*numpy.einsum('ijij,i,j->j', Vabab, foa, vecb)
        return Da, Db

    def g_mp1_den(self):
        """kelvin/hubbard_system.py:205-218."""
        V = self.g_aint_tot()
        beta = self.beta
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        vec = fo*ft_utils.ffv(beta, en, self.mu)
        utot = self._utot()
        T = numpy.einsum('ij,ip,jq->pq', self.model.get_tmat(), utot, utot)
        D = -beta*numpy.einsum('ii,i->i', T - numpy.diag(en), vec)
        D += -beta*numpy.einsum('ijij,i,j->i', V, vec, fo)
        return D

    def u_fock_d_den(self):
        """kelvin/hubbard_system.py:426-447."""
        da, db = self.u_energies_tot()
        beta = self.beta
        veca = ft_utils.ff(beta, da, self.mu)*ft_utils.ffv(beta, da, self.mu)
        vecb = ft_utils.ff(beta, db, self.mu)*ft_utils.ffv(beta, db, self.mu)
        Va, Vb, Vabab = self.u_aint_tot()
        JKaa = numpy.einsum('piqi,i->pqi', Va, veca)
        JKab = numpy.einsum('piqi,i->pqi', Vabab, vecb)
        JKbb = numpy.einsum('piqi,i->pqi', Vb, vecb)
        JKba = numpy.einsum('iris,i->rsi', Vabab, veca)
        return JKaa, JKab, JKbb, JKba

    def g_fock_d_den(self):
        """kelvin/hubbard_system.py:449-460."""
        d = self.g_energies_tot()
        beta = self.beta
        vec = ft_utils.ff(beta, d, self.mu)*ft_utils.ffv(beta, d, self.mu)
        V = self.g_aint_tot()
        return numpy.einsum('piqi,i->pqi', V, vec)

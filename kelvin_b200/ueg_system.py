"""Uniform electron gas in a plane-wave basis (input generator for the solver).

Same public interface and numerical conventions as kelvin/ueg_system.py
(``UEGSystem``) and kelvin/ueg_utils.py (``UEGBasis``) for the finite-
temperature path: basis enumeration order (ueg_utils.py:31-48), kinetic
energies, <pq|rs> with the p!=r, q!=s exclusions and momentum conservation
(:83-101), Fock matrices (ueg_system.py:366-422), MP1 (:92-135) and the
occupation-derivative helpers used by compute_ESN (:171-189,204-214,461-480,
516-528).  The reference builds the ERIs with an O(n^4) Python loop and
rebuilds them on every call; here they are vectorised and cached.
"""
import logging

import numpy

from . import ft_utils
from .system import System


def scaled_energy(x):
    return x[0]*x[0] + x[1]*x[1] + x[2]*x[2]


class UEGBasis(object):
    """Plane-wave basis (kelvin/ueg_utils.py:21-135)."""
    def __init__(self, L, cutoff, norb=None):
        self.L = L
        self.cutoff = cutoff
        kmax = numpy.sqrt(2*cutoff)
        imaxf = L*kmax/(2*numpy.pi)
        imax = int(numpy.ceil(imaxf) + 0.1)
        basis = []
        for i in range(-imax, imax):
            for j in range(-imax, imax):
                for l in range(-imax, imax):
                    k2 = 4.0*numpy.pi*numpy.pi/(L*L)*(i*i + j*j + l*l)
                    if k2 < 2.0*cutoff:
                        basis.append((i, j, l))
        basis.sort(key=scaled_energy)          # stable, like the reference
        if norb is not None:
            basis = basis[:norb]
        self.basis = basis
        ccc = 2.0*numpy.pi/L
        self.ks = [numpy.array((ccc*x[0], ccc*x[1], ccc*x[2])) for x in basis]
        self.Es = [(k[0]*k[0] + k[1]*k[1] + k[2]*k[2])/2.0 for k in self.ks]
        self._V = None

    def get_nbsf(self):
        return len(self.basis)

    def u_build_diag(self):
        return numpy.asarray(self.Es), numpy.asarray(self.Es)

    def g_build_diag(self):
        return numpy.hstack((numpy.asarray(self.Es), numpy.asarray(self.Es)))

    def build_r2e_matrix(self):
        """V[p,q,r,s] = 4pi/L^3 / |k_p-k_r|^2 if p!=r, q!=s and k_p+k_q=k_r+k_s."""
        if self._V is not None:
            return self._V
        n = self.get_nbsf()
        X = numpy.asarray(self.basis, dtype=numpy.int64).reshape(n, 3)
        K = numpy.asarray(self.ks).reshape(n, 3)
        aaa = 4.0*numpy.pi/(self.L*self.L*self.L)
        dk = K[:, None, :] - K[None, :, :]
        den = dk[..., 0]*dk[..., 0] + dk[..., 1]*dk[..., 1] + dk[..., 2]*dk[..., 2]
        with numpy.errstate(divide='ignore'):
            vpr = aaa/den                       # [p, r]; diagonal is excluded below
        # momentum transfer q = x_p - x_r must equal x_s - x_q
        dpr = X[:, None, :] - X[None, :, :]     # [p, r]
        code = (dpr[..., 0] + 64)*16384 + (dpr[..., 1] + 64)*128 + (dpr[..., 2] + 64)
        cons = code[:, None, :, None] == code.T[None, :, None, :]   # [p,q,r,s]: x_p-x_r == x_s-x_q
        eye = numpy.eye(n, dtype=bool)
        mask = cons & ~eye[:, None, :, None] & ~eye[None, :, None, :]
        V = numpy.where(mask, numpy.where(eye, 0.0, vpr)[:, None, :, None], 0.0)
        self._V = numpy.ascontiguousarray(V)
        return self._V

    def build_u2e_matrix(self, anti=True):
        V = self.build_r2e_matrix()
        Va = V - numpy.transpose(V, (0, 1, 3, 2))
        return Va, Va, V

    def build_g2e_matrix(self, anti=True):
        Vs = self.build_r2e_matrix()
        m = self.get_nbsf()
        V = numpy.zeros((2*m,)*4)
        V[:m, :m, :m, :m] = Vs
        V[m:, m:, m:, m:] = Vs
        V[:m, m:, :m, m:] = Vs
        V[m:, :m, m:, :m] = Vs
        if anti:
            return V - numpy.transpose(V, (0, 1, 3, 2))
        return V


class UEGSystem(System):
    """The uniform electron gas in a plane-wave basis set
    (constructor arguments as kelvin/ueg_system.py:32-33)."""
    def __init__(self, T, L, Emax, mu=None, na=None, nb=None, norb=None, orbtype='u',
                 madelung=None):
        self.T = T
        self.L = L
        self.basis = UEGBasis(L, Emax, norb=norb)
        if na is not None or nb is not None:
            raise Exception("kelvin_b200.UEGSystem supports the finite-temperature (mu) path only")
        assert(mu is not None)
        self.mu = mu
        beta = 1.0/self.T if self.T > 0.0 else 1.0e20
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        N = fo.sum()
        self.Na = N/2.0
        self.Nb = self.Na
        self.N = self.Na + self.Nb
        self.den = self.N/(L*L*L)
        self.rs = (3/(4.0*numpy.pi*self.den))**(1.0/3.0)
        pi2 = numpy.pi*numpy.pi
        self.Ef = 0.5*(3.0*pi2*self.den)**(2.0/3.0)
        self.Tf = self.T/self.Ef
        self.orbtype = orbtype
        self.madelung = madelung
        self._mconst = 2.837297479/(2*self.L)
        self._cache = {}

    def has_g(self):
        return True

    def has_u(self):
        return (False if self.orbtype == 'g' else True)

    def has_r(self):
        return (True if self.orbtype == 'r' else False)

    def verify(self, T, mu):
        if T > 0.0:
            return T == self.T and mu == self.mu
        return T == self.T

    def const_energy(self):
        if self.madelung == 'const':
            return -(self.Na + self.Nb)*self._mconst
        return 0.0

    # -- energies ----------------------------------------------------------
    def u_energies_tot(self):
        return self.basis.u_build_diag()

    def g_energies_tot(self):
        return self.basis.g_build_diag()

    def _beta(self):
        return 1.0/self.T

    # -- integrals ---------------------------------------------------------
    def u_aint_tot(self):
        if "u" not in self._cache:
            self._cache["u"] = self.basis.build_u2e_matrix()
        return self._cache["u"]

    def g_aint_tot(self):
        if "g" not in self._cache:
            self._cache["g"] = self.basis.build_g2e_matrix()
        return self._cache["g"]

    def g_int_tot(self):
        return self.basis.build_g2e_matrix(anti=False)

    # -- first order -------------------------------------------------------
    def get_mp1(self):
        if self.T <= 0:
            raise Exception("zero-temperature MP1 is outside the FT path")
        beta = self._beta()
        if self.has_u():
            Va, Vb, Vabab = self.u_aint_tot()
            ea, eb = self.u_energies_tot()
            foa = ft_utils.ff(beta, ea, self.mu)
            fob = ft_utils.ff(beta, eb, self.mu)
            E1 = 0.5*numpy.einsum('ijij,i,j->', Va, foa, foa)
            E1 += 0.5*numpy.einsum('ijij,i,j->', Vb, fob, fob)
            E1 += numpy.einsum('ijij,i,j->', Vabab, foa, fob)
            return E1
        V = self.g_aint_tot()
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        return 0.5*numpy.einsum('ijij,i,j->', V, fo, fo)

    def u_mp1_den(self):
        """d(MP1)/d(occupation) x f(1-f), per spin (kelvin/ueg_system.py:171-189)."""
        Va, Vb, Vabab = self.u_aint_tot()
        beta = self._beta()
        ea, eb = self.u_energies_tot()
        foa = ft_utils.ff(beta, ea, self.mu)
        veca = foa*ft_utils.ffv(beta, ea, self.mu)
        fob = ft_utils.ff(beta, eb, self.mu)
        vecb = fob*ft_utils.ffv(beta, eb, self.mu)
        Da = -beta*numpy.einsum('ijij,i,j->i', Va, veca, foa)
        Db = -beta*numpy.einsum('ijij,i,j->i', Vb, vecb, fob)
        Da -= beta*numpy.einsum('ijij,i,j->i', Vabab, veca, fob)
        Db -= beta*numpy.einsum('ijij,i,j->j', Vabab, foa, vecb)
        return Da, Db

    def g_mp1_den(self):
        """kelvin/ueg_system.py:204-214."""
        V = self.g_aint_tot()
        beta = self._beta()
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        vec = fo*ft_utils.ffv(beta, en, self.mu)
        return -beta*numpy.einsum('ijij,i,j->i', V, vec, fo)

    # -- Fock matrices -----------------------------------------------------
    def u_fock_tot(self):
        """kelvin/ueg_system.py:366-398 (finite T)."""
        da, db = self.u_energies_tot()
        beta = self._beta()
        foa = ft_utils.ff(beta, da, self.mu)
        fob = ft_utils.ff(beta, db, self.mu)
        Va, Vb, Vabab = self.u_aint_tot()
        JKa = numpy.einsum('prqr,r->pq', Va, foa) + numpy.einsum('prqr,r->pq', Vabab, fob)
        JKb = numpy.einsum('prqr,r->pq', Vb, fob) + numpy.einsum('prqr,r->pq', Vabab, foa)
        T = numpy.diag(numpy.asarray(self.basis.Es))
        return (T + JKa), (T + JKb)

    def g_fock_tot(self):
        """kelvin/ueg_system.py:400-422 (finite T)."""
        d = self.g_energies_tot()
        beta = self._beta()
        fo = ft_utils.ff(beta, d, self.mu)
        V = self.g_aint_tot()
        JK = numpy.einsum('prqr,r->pq', V, fo)
        return numpy.diag(d) + JK

    def u_fock_d_den(self):
        """kelvin/ueg_system.py:461-480."""
        da, db = self.u_energies_tot()
        beta = self._beta()
        veca = ft_utils.ff(beta, da, self.mu)*ft_utils.ffv(beta, da, self.mu)
        vecb = ft_utils.ff(beta, db, self.mu)*ft_utils.ffv(beta, db, self.mu)
        Va, Vb, Vabab = self.u_aint_tot()
        JKaa = numpy.einsum('piqi,i->pqi', Va, veca)
        JKab = numpy.einsum('piqi,i->pqi', Vabab, vecb)
        JKbb = numpy.einsum('piqi,i->pqi', Vb, vecb)
        JKba = numpy.einsum('iris,i->rsi', Vabab, veca)
        return JKaa, JKab, JKbb, JKba

    def g_fock_d_den(self):
        """kelvin/ueg_system.py:516-528."""
        d = self.g_energies_tot()
        beta = self._beta()
        vec = ft_utils.ff(beta, d, self.mu)*ft_utils.ffv(beta, d, self.mu)
        V = self.g_aint_tot()
        return numpy.einsum('piqi,i->pqi', V, vec)


class ueg_system(UEGSystem):
    def __init__(self, *a, **k):
        logging.warning("This class is deprecated, use UEGSystem instead")
        UEGSystem.__init__(self, *a, **k)

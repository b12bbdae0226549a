"""Spin-polarised uniform electron gas in a plane-wave basis (input generator).

Same public interface and numerical conventions as kelvin/pueg_system.py (``PUEGSystem``) for
the finite-temperature path: one spin species, general-spin-orbital ('g') integrals
<pq||rs> = V[p,q,r,s] - V[p,q,s,r] over the spatial plane waves (:205-208), Fock matrix
(:142-161), MP1 (:70-80) and the occupation-derivative helpers used by compute_ESN (:95-104,
178-191).  The plane-wave basis and the vectorised, cached <pq|rs> are those of
kelvin_b200/ueg_system.py.
"""
import logging

import numpy

from . import ft_utils
from .system import System
from .ueg_system import UEGBasis


class PUEGSystem(System):
    """The polarized uniform electron gas in a plane-wave basis set (constructor arguments as
    kelvin/pueg_system.py:24)."""
    def __init__(self, T, L, Emax, mu=None, n=None, norb=None):
        self.T = T
        self.L = L
        self.basis = UEGBasis(L, Emax, norb=norb)
        if n is not None:
            raise Exception("kelvin_b200.PUEGSystem supports the finite-temperature (mu) path only")
        assert(mu is not None)
        self.mu = mu
        beta = 1.0/self.T if self.T > 0.0 else 1.0e20
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        self.N = fo.sum()
        self.den = self.N/(L*L*L)
        self.rs = (3/(4.0*numpy.pi*self.den))**(1.0/3.0)
        pi2 = numpy.pi*numpy.pi
        self.Ef = 0.5*(3.0*pi2*self.den)**(2.0/3.0)
        self.Tf = self.T/self.Ef
        self.orbtype = 'g'
        self._V = None

    def has_g(self):
        return True

    def has_u(self):
        return False

    def has_r(self):
        return False

    def verify(self, T, mu):
        if T > 0.0:
            return T == self.T and mu == self.mu
        return T == self.T

    def const_energy(self):
        return 0.0

    def _occ(self):
        beta = 1.0/self.T
        en = self.g_energies_tot()
        return beta, ft_utils.ff(beta, en, self.mu), ft_utils.ffv(beta, en, self.mu)

    def g_energies_tot(self):
        return numpy.asarray(self.basis.Es)

    def g_aint_tot(self):
        if self._V is None:
            V = self.basis.build_r2e_matrix()
            self._V = numpy.ascontiguousarray(V - V.transpose((0, 1, 3, 2)))
        return self._V

    def get_mp1(self):
        if self.T <= 0:
            raise Exception("zero-temperature MP1 is outside the FT path")
        beta, fo, fv = self._occ()
        return 0.5*numpy.einsum('ijij,i,j->', self.g_aint_tot(), fo, fo)

    def g_d_mp1(self, dvec):
        beta, fo, fv = self._occ()
        return -numpy.einsum('ijij,i,j->', self.g_aint_tot(), dvec*fo*fv, fo)

    def g_mp1_den(self):
        beta, fo, fv = self._occ()
        return -beta*numpy.einsum('ijij,i,j->i', self.g_aint_tot(), fo*fv, fo)

    def g_hcore(self):
        return numpy.diag(self.g_energies_tot())

    def g_fock_tot(self):
        beta, fo, fv = self._occ()
        JK = numpy.einsum('prqr,r->pq', self.g_aint_tot(), fo)
        return numpy.diag(self.g_energies_tot()) + JK

    def g_fock_d_tot(self, dvec):
        beta, fo, fv = self._occ()
        return -numpy.einsum('prqr,r->pq', self.g_aint_tot(), dvec*fo*fv)

    def g_fock_d_den(self):
        beta, fo, fv = self._occ()
        return numpy.einsum('piqi,i->pqi', self.g_aint_tot(), fo*fv)


class pueg_system(PUEGSystem):
    def __init__(self, T, L, Emax, mu=None, n=None, norb=None):
        logging.warning("This class is deprecated, use PUEGSystem instead")
        PUEGSystem.__init__(self, T, L, Emax, mu=mu, n=n, norb=norb)

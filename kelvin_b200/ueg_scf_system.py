"""Uniform electron gas with Hartree-Fock orbital energies (input generator for the solver).

Same public interface and numerical conventions as kelvin/ueg_scf_system.py (``UEGSCFSystem``)
for the finite-temperature path.  The plane waves stay the orbitals; what changes against
``UEGSystem`` is the zeroth-order Hamiltonian: the orbital energies are the Hartree-Fock ones of
the ZERO-temperature reference determinant -- all plane waves with kinetic energy below mu (or
the first `naref`) doubly occupied (:55-79, 280-297),

    e_p = k_p^2/2 + sum_{i occ} (2 <pi|pi> - <pi|ip>),

so that the first-order energy and its occupation derivatives pick up the one-body term
(T - diag(e)) (:125-158, 160-236) and the Fock matrices are T + JK with the Fermi occupations of
the HF energies (:351-417).  The plane-wave basis and the vectorised, cached <pq|rs> are those of
kelvin_b200/ueg_system.py.
"""
import logging

import numpy

from . import ft_utils
from .ueg_system import UEGBasis, UEGSystem


class UEGSCFSystem(UEGSystem):
    """The uniform electron gas in a plane-wave basis set with HF orbital energies (constructor
    arguments as kelvin/ueg_scf_system.py:31-32)."""
    def __init__(self, T, L, Emax, mu=None, na=None, nb=None, norb=None, orbtype='u',
                 madelung=None, naref=None):
        if na is not None or nb is not None:
            raise Exception("kelvin_b200.UEGSCFSystem supports the finite-temperature (mu) path only")
        assert(mu is not None)
        self.basis = UEGBasis(L, Emax, norb=norb)
        d0 = numpy.asarray(self.basis.Es)
        n = d0.shape[0]
        if naref is not None:
            self.oidx = numpy.r_[[i for i in range(naref)]]
            self.vidx = numpy.r_[[i + naref for i in range(n - naref)]]
        else:
            occ = [p for p, d in enumerate(d0) if d < mu]
            vir = [p for p, d in enumerate(d0) if d > mu]
            self.oidx = numpy.r_[occ]
            self.vidx = numpy.r_[vir]
            self.goidx = numpy.r_[occ + [p + n for p in occ]]
            self.gvidx = numpy.r_[vir + [p + n for p in vir]]
        self._e = None
        # (the base constructor builds its own, identical, basis and derives N, rs, ... from
        # g_energies_tot() of THIS class)
        UEGSystem.__init__(self, T, L, Emax, mu=mu, norb=norb, orbtype=orbtype, madelung=madelung)

    # -- energies ----------------------------------------------------------
    def r_int_tot(self):
        return self.basis.build_r2e_matrix()

    def r_energies_tot(self):
        """kelvin/ueg_scf_system.py:280-289."""
        if self._e is None:
            e = numpy.array(self.basis.Es, dtype=float)
            if len(self.oidx) > 0:
                V = self.r_int_tot()
                o = numpy.asarray(self.oidx, dtype=int)
                p = numpy.arange(e.shape[0])
                # sum_i (2 V[p,i,p,i] - V[p,i,i,p])
                e = e + 2.0*V[p[:, None], o[None, :], p[:, None], o[None, :]].sum(axis=1) \
                    - V[p[:, None], o[None, :], o[None, :], p[:, None]].sum(axis=1)
            self._e = e
        return self._e.copy()

    def u_energies_tot(self):
        e = self.r_energies_tot()
        return e, e.copy()

    def g_energies_tot(self):
        ea, eb = self.u_energies_tot()
        return numpy.hstack((ea, eb))

    def r_hcore(self):
        return numpy.diag(numpy.asarray(self.basis.Es))

    def g_hcore(self):
        Es = numpy.asarray(self.basis.Es)
        return numpy.diag(numpy.hstack((Es, Es)))

    # -- first order -------------------------------------------------------
    def get_mp1(self):
        """kelvin/ueg_scf_system.py:125-158 (finite T)."""
        if self.T <= 0:
            raise Exception("zero-temperature MP1 is outside the FT path")
        E2 = UEGSystem.get_mp1(self)          # the two-body part, with this class's energies
        beta = self._beta()
        if self.has_u():
            ea, eb = self.u_energies_tot()
            t = numpy.asarray(self.basis.Es)
            E1 = numpy.dot(t - ea, ft_utils.ff(beta, ea, self.mu))
            E1 += numpy.dot(t - eb, ft_utils.ff(beta, eb, self.mu))
            return E2 + E1
        en = self.g_energies_tot()
        t = self.g_hcore().diagonal()
        return E2 + numpy.dot(t - en, ft_utils.ff(beta, en, self.mu))

    def u_mp1_den(self):
        """kelvin/ueg_scf_system.py:183-204."""
        Da, Db = UEGSystem.u_mp1_den(self)
        beta = self._beta()
        ea, eb = self.u_energies_tot()
        t = numpy.asarray(self.basis.Es)
        veca = ft_utils.ff(beta, ea, self.mu)*ft_utils.ffv(beta, ea, self.mu)
        vecb = ft_utils.ff(beta, eb, self.mu)*ft_utils.ffv(beta, eb, self.mu)
        return -beta*(t - ea)*veca + Da, -beta*(t - eb)*vecb + Db

    def g_mp1_den(self):
        """kelvin/ueg_scf_system.py:222-236."""
        D2 = UEGSystem.g_mp1_den(self)
        beta = self._beta()
        en = self.g_energies_tot()
        vec = ft_utils.ff(beta, en, self.mu)*ft_utils.ffv(beta, en, self.mu)
        return -beta*(self.g_hcore().diagonal() - en)*vec + D2

    def u_d_mp1(self, dveca, dvecb):
        """kelvin/ueg_scf_system.py:160-181."""
        Va, Vb, Vabab = self.u_aint_tot()
        beta = self._beta()
        ea, eb = self.u_energies_tot()
        foa, fob = ft_utils.ff(beta, ea, self.mu), ft_utils.ff(beta, eb, self.mu)
        veca = dveca*foa*ft_utils.ffv(beta, ea, self.mu)
        vecb = dvecb*fob*ft_utils.ffv(beta, eb, self.mu)
        t = numpy.asarray(self.basis.Es)
        D = -numpy.dot(t - ea, veca) - numpy.dot(t - eb, vecb)
        D += -numpy.einsum('ijij,i,j->', Va, veca, foa)
        D += -numpy.einsum('ijij,i,j->', Vb, vecb, fob)
        D += -numpy.einsum('ijij,i,j->', Vabab, veca, fob)
        D += -numpy.einsum('ijij,i,j->', Vabab, foa, vecb)
        return D

    def g_d_mp1(self, dvec):
        """kelvin/ueg_scf_system.py:206-220."""
        V = self.g_aint_tot()
        beta = self._beta()
        en = self.g_energies_tot()
        fo = ft_utils.ff(beta, en, self.mu)
        vec = dvec*fo*ft_utils.ffv(beta, en, self.mu)
        E1_2 = -numpy.einsum('ijij,i,j->', V, vec, fo)
        E1_1 = -numpy.dot(self.g_hcore().diagonal() - en, vec)
        return E1_2 + E1_1

    # -- Fock matrices -----------------------------------------------------
    def u_fock_tot(self):
        """kelvin/ueg_scf_system.py:372-396 (finite T): T + JK, occupations of the HF energies."""
        Fa, Fb = UEGSystem.u_fock_tot(self)
        return Fa, Fb

    def g_fock_tot(self):
        """kelvin/ueg_scf_system.py:398-417 (finite T)."""
        d = self.g_energies_tot()
        fo = ft_utils.ff(self._beta(), d, self.mu)
        JK = numpy.einsum('prqr,r->pq', self.g_aint_tot(), fo)
        return self.g_hcore() + JK

    def u_fock_d_tot(self, dveca, dvecb):
        """kelvin/ueg_scf_system.py:419-442."""
        da, db = self.u_energies_tot()
        beta = self._beta()
        veca = dveca*ft_utils.ff(beta, da, self.mu)*ft_utils.ffv(beta, da, self.mu)
        vecb = dvecb*ft_utils.ff(beta, db, self.mu)*ft_utils.ffv(beta, db, self.mu)
        Va, Vb, Vabab = self.u_aint_tot()
        JKa = numpy.einsum('prqr,r->pq', Va, veca) + numpy.einsum('prqr,r->pq', Vabab, vecb)
        JKb = numpy.einsum('prqr,r->pq', Vb, vecb) + numpy.einsum('prps,p->rs', Vabab, veca)
        return -JKa, -JKb

    def g_fock_d_tot(self, dvec):
        """kelvin/ueg_scf_system.py:465-479."""
        d = self.g_energies_tot()
        beta = self._beta()
        vec = dvec*ft_utils.ff(beta, d, self.mu)*ft_utils.ffv(beta, d, self.mu)
        return -numpy.einsum('prqr,r->pq', self.g_aint_tot(), vec)


class ueg_scf_system(UEGSCFSystem):
    def __init__(self, *a, **k):
        logging.warning("This class is deprecated, use UEGSCFSystem instead")
        UEGSCFSystem.__init__(self, *a, **k)

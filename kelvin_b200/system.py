"""Abstract physical-system interface consumed by the ``ccsd`` driver.

Mirrors the on-path part of kelvin/system.py:1-166: a system supplies orbital
energies, the Fock matrix and antisymmetrised ERIs in the full (unpartitioned)
basis; at finite temperature every orbital is both "occupied" and "virtual".
"""


class System(object):
    def _missing(self, what):
        raise Exception("Base class function {}()".format(what))

    def verify(self, T, mu):
        self._missing("verify")

    def has_g(self):
        self._missing("has_g")

    def has_u(self):
        self._missing("has_u")

    def has_r(self):
        self._missing("has_r")

    def const_energy(self):
        self._missing("const_energy")

    def get_mp1(self):
        self._missing("get_mp1")

    def u_energies_tot(self):
        self._missing("u_energies_tot")

    def g_energies_tot(self):
        self._missing("g_energies_tot")

    def u_fock_tot(self):
        self._missing("u_fock_tot")

    def g_fock_tot(self):
        self._missing("g_fock_tot")

    def u_aint_tot(self):
        self._missing("u_aint_tot")

    def g_aint_tot(self):
        self._missing("g_aint_tot")

    def g_int_tot(self):
        self._missing("g_int_tot")

    def u_mp1_den(self):
        self._missing("u_mp1_den")

    def g_mp1_den(self):
        self._missing("g_mp1_den")

    def u_fock_d_den(self):
        self._missing("u_fock_d_den")

    def g_fock_d_den(self):
        self._missing("g_fock_d_den")

"""Model Hamiltonians needed to feed the oracle (TEST INFRASTRUCTURE).

``Hubbard1D`` restates the un-vendored ``lattice.hubbard.Hubbard1D``
(used at kelvin/hubbard_system.py:59,332,351,572,580,599);
validated by reproducing examples/hubbard1d.out to 16 digits through the
unmodified reference HubbardSystem (tests/golden/make_golden.py).
"""
import numpy


class Hubbard1D(object):
    def __init__(self, L, t, U, boundary='p'):
        self.L = L
        self.t = t
        self.U = U
        self.boundary = boundary

    def get_tmatS(self):
        """L x L hopping matrix, -t on nearest neighbours (wrap if periodic)."""
        L = self.L
        T = numpy.zeros((L, L))
        for i in range(L - 1):
            T[i, i + 1] = -self.t
            T[i + 1, i] = -self.t
        if self.boundary == 'p' and L > 2:
            T[0, L - 1] = -self.t
            T[L - 1, 0] = -self.t
        return T

    def get_umatS(self):
        """Spatial on-site repulsion U[i,i,i,i] = U."""
        L = self.L
        V = numpy.zeros((L, L, L, L))
        for i in range(L):
            V[i, i, i, i] = self.U
        return V

    def get_tmat(self):
        T = self.get_tmatS()
        L = self.L
        out = numpy.zeros((2 * L, 2 * L))
        out[:L, :L] = T
        out[L:, L:] = T
        return out

    def get_umat(self):
        """Spin-orbital (not antisymmetrised) <pq|rs>, alpha block then beta."""
        L = self.L
        V = numpy.zeros((2 * L, 2 * L, 2 * L, 2 * L))
        for i in range(L):
            V[i, L + i, i, L + i] = self.U
            V[L + i, i, L + i, i] = self.U
        return V

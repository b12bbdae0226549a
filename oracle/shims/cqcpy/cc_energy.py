"""Imported by kelvin/ccsd.py:5 and kelvin/zt_mp.py:3-4; only the
zero-temperature path (out of scope) calls into it."""


def _zero_T_only(*a, **k):
    raise NotImplementedError("zero-temperature cqcpy.cc_energy is outside the FT-CCSD path")


cc_energy_d = cc_energy_s1 = cc_energy = ucc_energy = _zero_T_only

"""Shim for lattice.hubbard.Hubbard1D (kelvin/hubbard_system.py:59,332,351,572,580,599;
semantics in SURVEY.md A.6, validated against examples/hubbard1d.out)."""
from kelvin_oracle.systems import Hubbard1D  # noqa: F401

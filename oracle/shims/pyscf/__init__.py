"""Shim: only pyscf.lib.einsum is used on the FT-CCSD path (kelvin/ccsd.py:4,16)."""
from . import lib  # noqa: F401

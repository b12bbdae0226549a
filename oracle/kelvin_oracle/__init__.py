"""CPU oracle for the FT-CCSD hot path of awhite862/kelvin.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline.  The product (``kelvin_b200``) never
imports this package and has no CPU fallback.

What it is: a NumPy (FP64) restatement of

* the un-vendored ``cqcpy`` arithmetic the reference calls
  (``cqcpy.cc_equations._Stanton`` & friends), and
* the reference's own drivers for the path
  (``kelvin/quadrature.py``, ``kelvin/ft_cc_energy.py``,
  ``kelvin/ft_cc_equations.py``, ``kelvin/cc_utils.py``, ``kelvin/ccsd.py``,
  ``kelvin/ueg_system.py``, ``kelvin/hubbard_system.py``).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` replays the golden
numbers the reference itself publishes (``kelvin/tests/test_ft_ccsd.py:23``,
``examples/*.out``, ``bench/*/*.out``) through this oracle, and
``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` exists) drives the *unmodified* reference drivers on top
of the restated ``cqcpy`` layer (``oracle/shims``) to produce the committed
fixtures the standalone oracle and the CUDA path are compared with.
"""

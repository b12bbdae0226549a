"""kelvin_b200 -- B200-native finite-temperature CCSD (amplitudes + Lambda).

A from-scratch, sm_100a implementation of the hot path behind
``kelvin.ccsd(sys, T=..., mu=..., ngrid=...).run()`` / ``.compute_ESN()`` of
awhite862/kelvin, importable under the reference's module names:

    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem
    from kelvin_b200 import ft_cc_equations, quadrature, ft_cc_energy, cc_utils

Python + PyTorch hold device memory and streams; every arithmetic step runs in
hand-written CUDA kernels behind the C ABI in include/kelvin_b200.h
(libkb200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"

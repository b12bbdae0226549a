"""Sanity + timing of the larger BASELINE configs (one or two iterations each)."""
import logging
import sys
import os
import time
import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
logging.basicConfig(level=logging.INFO, format="%(message)s", stream=sys.stdout)
from kelvin_b200.ccsd import ccsd  # noqa: E402
from kelvin_b200.ueg_system import UEGSystem  # noqa: E402
from kelvin_b200.hubbard_system import HubbardSystem, Hubbard1D  # noqa: E402


def hubbard(L, T, ng, iters):
    hub = Hubbard1D(L, 1.0, 1.0, boundary='p')
    Oa, Ob = numpy.zeros(L), numpy.zeros(L)
    Oa[0::2] = 1.0
    Ob[1::2] = 1.0
    s = HubbardSystem(T, hub, numpy.einsum('i,j->ij', Oa, Oa), numpy.einsum('i,j->ij', Ob, Ob), mu=0.0)
    t0 = time.time()
    cc = ccsd(s, T=T, mu=0.0, iprint=1, max_iter=iters, econv=1e-11, ngrid=ng)
    out = cc.run()
    torch.cuda.synchronize()
    print("hubbard L=%d ng=%d T=%g: %s  wall %.2f s, peak mem %.1f GB" %
          (L, ng, T, out, time.time() - t0, torch.cuda.max_memory_allocated()/1e9))
    return cc


def ueg(norb, ng, iters, emax=30.0):
    t0 = time.time()
    s = UEGSystem(0.5, 1.942, emax, mu=7.0, norb=norb, orbtype='u')
    print("UEG-%d system built in %.1f s" % (norb, time.time() - t0))
    t0 = time.time()
    cc = ccsd(s, T=0.5, mu=7.0, iprint=1, max_iter=iters, damp=0.0, ngrid=ng)
    out = cc.run()
    torch.cuda.synchronize()
    print("UEG-%d ng=%d: %s  wall %.2f s, peak mem %.1f GB" %
          (norb, ng, out, time.time() - t0, torch.cuda.max_memory_allocated()/1e9))
    return cc


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "hubbard32":
        hubbard(32, 1.0, 40, 3)
    elif which == "ueg57":
        ueg(57, 16, 2)
    elif which == "ueg81":
        ueg(81, 24, 1, emax=35.0)
    elif which == "esn33_full":
        cc = ueg(33, 10, 50)
        t0 = time.time()
        cc.compute_ESN()
        torch.cuda.synchronize()
        print("ESN33 E S N:", repr(cc.E), repr(cc.S), repr(cc.N), " compute_ESN wall %.2f s" % (time.time() - t0))

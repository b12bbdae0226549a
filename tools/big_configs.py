"""The larger BASELINE configs through the public API: full amplitude solve (iterations, seconds
per iteration, time to convergence, peak HBM), optionally Lambda + E/S/N, and a spot check of the
per-grid-point residual at ONE grid point against the oracle's Sz-blocked CPU port (SURVEY 7,
"Oracle cost": the dense oracle cannot run these sizes in full).

  python tools/big_configs.py hubbard32 [T] [max_iter] [spot] [esn]
  python tools/big_configs.py ueg57 | ueg81 | esn33 | esn19 [max_iter] [spot] [esn]
  python -m torch.distributed.run --nproc-per-node 8 tools/big_configs.py ueg81 2

Runs under whatever torch.distributed world it is launched in (rank 0 prints).
"""
import logging
import os
import sys
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    which = sys.argv[1]
    rest = sys.argv[2:]
    flags = {a for a in rest if a in ("spot", "esn")}
    nums = [a for a in rest if a not in flags]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    logging.basicConfig(level=logging.INFO if rank == 0 else logging.WARNING, format="%(message)s",
                        stream=sys.stdout)
    from kelvin_b200 import ft_cc_equations as fe
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.hubbard_system import HubbardSystem, Hubbard1D
    from kelvin_b200.ueg_system import UEGSystem

    def say(*a):
        if rank == 0:
            print(*a, flush=True)
    t0 = time.time()
    if which.startswith("hubbard"):
        L = int(which[7:])
        T = float(nums[0]) if nums else 1.0
        iters = int(nums[1]) if len(nums) > 1 else 80
        hub = Hubbard1D(L, 1.0, 1.0, boundary='p')
        Oa, Ob = numpy.zeros(L), numpy.zeros(L)
        Oa[0::2] = 1.0
        Ob[1::2] = 1.0
        s = HubbardSystem(T, hub, numpy.einsum('i,j->ij', Oa, Oa), numpy.einsum('i,j->ij', Ob, Ob), mu=0.0)
        mu, ng = 0.0, 40
        kw = dict(max_iter=iters, econv=1e-11)
        m = L
    else:
        norb, ng, emax = {"ueg57": (57, 16, 30.0), "ueg81": (81, 24, 35.0), "esn33": (33, 10, 30.0),
                          "esn19": (19, 10, 30.0)}[which]
        iters = int(nums[0]) if nums else 50
        T, mu = 0.5, 7.0
        s = UEGSystem(T, 1.942, emax, mu=mu, norb=norb, orbtype='u')
        kw = dict(max_iter=iters, damp=0.0)
        m = norb
    say("%s: system built in %.1f s (world %d)" % (which, time.time() - t0, world))
    cc = ccsd(s, T=T, mu=mu, iprint=1, ngrid=ng, **kw)
    torch.cuda.reset_peak_memory_stats()
    t0 = time.time()
    out = cc.run()
    torch.cuda.synchronize()
    wall = time.time() - t0
    say("%s ng=%d T=%g: Omega = %.10f  Omega_cc = %.10f   wall %.2f s (first call: includes plan "
        "compilation), peak HBM %.1f GB" % (which, ng, T, out[0], out[1], wall,
                                              torch.cuda.max_memory_allocated()/1e9))
    # a second solve with everything compiled: the timing
    cc2 = ccsd(s, T=T, mu=mu, iprint=0, ngrid=ng, **kw)

    class Count(logging.Handler):
        n = 0

        def emit(self, r):
            f = r.getMessage().split()
            if len(f) == 3 and f[0].isdigit():
                Count.n += 1
    h = Count()
    logging.getLogger().addHandler(h)
    t0 = time.time()
    out2 = cc2.run()
    torch.cuda.synchronize()
    wall2 = time.time() - t0
    logging.getLogger().removeHandler(h)
    nit = max(1, Count.n)
    fl = ng*(64.0*m**6 + 120.0*m**5)
    say("%s: %d iterations in %.3f s = %.4f s/iteration (algorithmic %.2f TFLOP/iteration -> %.1f "
        "TFLOP/s of the reference's work), Omega_cc %.12f" % (which, nit, wall2, wall2/nit, fl/1e12,
                                                             fl/1e12/(wall2/nit), out2[1]))
    say("flags of the solve:", cc2._flags)
    if "spot" in flags and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from kelvin_oracle import spin_blocked as sb
        y = ng//2
        ea, eb, Ds, ints = cc2._u_setup()
        amps = list(cc2.T1) + list(cc2.T2)
        bars = fe.uccsd_stanton_bar(*ints, *[a[y:y + 1] for a in amps])
        cpu = lambda t: t.cpu().numpy()     # noqa: E731

        class Bag(object):
            pass

        def host(x):
            b = Bag()
            for k, v in x.__dict__.items():
                setattr(b, k, cpu(v) if isinstance(v, torch.Tensor) else v)
            return b
        hints = [host(x) for x in ints]
        t0 = time.time()
        rs = sb.u_stanton_terms(*hints, (cpu(amps[0][y]), cpu(amps[1][y])),
                                (cpu(amps[2][y]), cpu(amps[3][y]), cpu(amps[4][y])))
        tcpu = time.time() - t0
        drv = (hints[0].vo, hints[1].vo, hints[2].vvoo, hints[4].vvoo, hints[3].vvoo)
        err = 0.0
        for k in range(5):
            ref = -drv[k] - rs[k]
            err = max(err, float(numpy.abs(cpu(bars[k][0]) - ref).max()/numpy.abs(ref).max()))
        say("%s spot check at grid point %d: max rel. deviation of the residual blocks from the "
            "oracle's Sz-blocked CPU port %.2e (CPU: %.1f s for this one point)" % (which, y, err, tcpu))
    if "esn" in flags:
        t0 = time.time()
        cc2.compute_ESN()
        torch.cuda.synchronize()
        say("%s E S N: %r %r %r   compute_ESN wall %.2f s" % (which, cc2.E, cc2.S, cc2.N,
                                                              time.time() - t0))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

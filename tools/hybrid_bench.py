"""Where a sharded amplitude iteration goes (run under torchrun): device and host time of the
own grid points (plan A), the shared grid point (hybrid plan B), both together, the row
exchange and the fused update.

  python -m torch.distributed.run --nproc-per-node N tools/hybrid_bench.py [norb] [ngrid]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import _lib, cc_utils, ft_cc_equations as fe, ft_utils, parallel, quadrature  # noqa: E402
from kelvin_b200 import plan as _plan  # noqa: E402
from kelvin_b200.ueg_system import UEGSystem  # noqa: E402


def main():
    norb = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    dev = _lib.device()
    T_, MU_, L_ = 0.5, 7.0, 1.942
    beta = 1.0/T_
    sysm = UEGSystem(T_, L_, 30.0, mu=MU_, norb=norb, orbtype='u')
    ea, eb = sysm.u_energies_tot()
    ti, g, G = quadrature.ft_quad(ng, beta, 'lin')
    ints = cc_utils.uft_integrals(sysm, ea, eb, beta, MU_)
    Fa, Fb, Ia, Ib, Iabab = ints
    Ds = (ft_utils.D1(ea, ea), ft_utils.D1(eb, eb), ft_utils.D2(ea, ea),
          ft_utils.D2u(ea, eb, ea, eb), ft_utils.D2(eb, eb))

    def rep(x):
        return (-x).expand(*((ng,) + (-1,)*x.dim())).contiguous()
    guess = [quadrature.int_tbar(ng, rep(x), ti, d, G) for x, d in
             zip((Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo), Ds)]
    st = cc_utils.UccStep(guess, *ints, Ds, g, G, beta, ng, ti)
    for _ in range(3):
        st.step(0.0)
    sh = parallel.Shards(ng, 1)
    sizes = fe._u_sizes(Fa, Fb)
    kw = dict(mirror=True, singlet=True, antisym=True, emit_aa=False)
    nown = sh.own[1] - sh.own[0]
    pA = fe.stanton_plan("u", sizes, -1.0, mirror_rows=nown >= fe.MIRROR_ROWS_MIN_BATCH, **kw)
    pB = fe.stanton_plan("u", sizes, -1.0, hybrid_world=world, **kw) if (world > 1 and sh.r) else None
    t = fe._u_integral_slots(*ints, dev, [s for s in pA.inputs if _plan.is_integral_slot(s)])
    live = (0, 3)
    flat, views = fe.flat_rows(ng, [st.old[k].shape[1:] for k in live], dev)
    t[fe._U_TIN[2]] = st.old[2]
    for k, v in zip(live, views):
        t[fe._U_TIN[k]] = st.old[k]
        t[fe._U_TOUT[k]] = v

    def sl(pl, a, b):
        return {s: (t[s][a:b] if pl.batched[s] else t[s]) for s in pl.inputs + pl.outputs}

    def runA():
        pA.run(sl(pA, *sh.own), nown)

    def runB():
        pB.run(sl(pB, *sh.left), sh.r, sh.rank, parallel.exchange_group())

    side = fe._side_stream(dev)

    def both(order):
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        if order == "BA":
            with torch.cuda.stream(side):
                runB()
            runA()
        else:
            runA()
            with torch.cuda.stream(side):
                runB()
        cur.wait_stream(side)

    def measure(name, fn, n=10):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = 0.0
        tot = 0.0
        for _ in range(n):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            h0 = time.time()
            fn()
            host += time.time() - h0
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        print("rank %d  %-28s device %7.3f ms   host enqueue %7.3f ms" % (rank, name, tot/n, host/n*1e3),
              flush=True)

    measure("plan A (own %d rows)" % nown, runA)
    if pB is not None:
        measure("plan B (hybrid, %d rows)" % sh.r, runB)
        measure("A then B, two streams", lambda: both("AB"))
        measure("B then A, two streams", lambda: both("BA"))
        old = _lib.load().kb200_set_plan_streams(1)
        measure("plan B, one stream", runB)
        _lib.load().kb200_set_plan_streams(old)
    measure("exchange_rows", lambda: parallel.exchange_rows(flat, sh, False))
    measure("whole step", lambda: st.step(0.0))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

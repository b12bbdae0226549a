"""Device time of one residual-plan run (all launches, as the solver issues them) for several
tau-batch sizes: the small batches are what a tau-sharded rank sees (1-3 grid points per GPU).

  python tools/batch_bench.py [norb] [closed|general] [ng ...]
"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import _lib, cc_utils, ft_cc_equations, plan as _plan  # noqa: E402
from kelvin_b200.ueg_system import UEGSystem  # noqa: E402


def main():
    norb = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    closed = not (len(sys.argv) > 2 and sys.argv[2] == "general")
    ngs = [int(x) for x in sys.argv[3:]] or [9, 3, 2, 1]
    T_, MU_, L_ = 0.5, 7.0, 1.942
    dev = _lib.device()
    beta = 1.0/T_
    sysm = UEGSystem(T_, L_, 30.0, mu=MU_, norb=norb, orbtype='u')
    ea, eb = sysm.u_energies_tot()
    Fa, Fb, Ia, Ib, Iabab = cc_utils.uft_integrals(sysm, ea, eb, beta, MU_)
    sizes = ft_cc_equations._u_sizes(Fa, Fb)
    rows = closed and os.environ.get("KB200_MIRROR_ROWS", "auto")
    plans = {}

    def plan_for(ng):
        mr = bool(closed) and (ng >= ft_cc_equations.MIRROR_ROWS_MIN_BATCH if rows == "auto"
                               else rows == "1")
        sa = ng <= ft_cc_equations.SPLIT_ACC_MAX_BATCH
        if (mr, sa) not in plans:
            plans[mr, sa] = ft_cc_equations.stanton_plan(
                "u", sizes, -1.0, mirror=closed, mirror_rows=mr, singlet=closed, emit_aa=not closed,
                split_acc=sa)
        return plans[mr, sa]
    p = plan_for(max(ngs))
    ints = ft_cc_equations._u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                                             [s for s in p.inputs if _plan.is_integral_slot(s)])
    print("streams=%s group=%s closed=%s" % (os.environ.get("KB200_STREAMS", "3"),
                                              os.environ.get("KB200_GROUP", "8"), closed))
    for ng in ngs:
        p = plan_for(ng)
        torch.manual_seed(0)
        t = dict(ints)
        for s in p.inputs + p.outputs:
            if s not in t:
                t[s] = 0.01*torch.randn((ng,) + tuple(p.shapes[s]), dtype=torch.float64, device=dev)
        for _ in range(3):
            p.run(t, ng)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            p.run(t, ng)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/n
        chk = float(sum(t[s].double().abs().sum() for s in p.outputs))
        print("ng=%2d  %8.3f ms/run  %6.3f ms/point  checksum %.15e" % (ng, ms, ms/ng, chk))
        if os.environ.get("KB200_BATCH_TIMELINE"):
            timeline(lambda: p.run(t, ng), ng)


def timeline(fn, ng):
    """Kernel timeline of one run (torch.profiler / CUPTI): start, duration, stream of every
    kernel, plus how many SM-filling GEMM kernels overlap -- gpurun_out/timeline_ng<ng>.txt."""
    from torch.profiler import profile, ProfilerActivity
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    if not ev:
        print("no device events")
        return
    t0 = ev[0].time_range.start
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    end = max(e.time_range.end for e in ev)
    busy = 0.0
    cur_end = t0
    for e in ev:
        a, b = e.time_range.start, e.time_range.end
        if b > cur_end:
            busy += b - max(a, cur_end)
            cur_end = b
    with open(os.path.join(ROOT, "gpurun_out", "timeline_ng%d.txt" % ng), "w") as f:
        f.write("# %d kernels, span %.1f us, some kernel running %.1f us\n" % (len(ev), end - t0, busy))
        for e in ev:
            f.write("%9.1f %8.1f  s%-3s %s\n" % (e.time_range.start - t0, e.time_range.end - e.time_range.start,
                                              getattr(e, "device_resource_id", "?"), e.name[:90]))
    print("timeline: %d kernels, span %.1f us, union of kernel intervals %.1f us" % (len(ev), end - t0, busy))


if __name__ == "__main__":
    main()

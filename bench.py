#!/usr/bin/env python
"""Benchmark of the FT-CCSD amplitude iteration (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo, N GPUs of one node
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

A "step" is ONE FT-CCSD amplitude iteration of the unrestricted UEG benchmark
(bench/ueg_ft_ccsd_ESN33.py: 33 plane waves, T=0.5, mu=7, L=1.942, ngrid=10):
per-grid-point Stanton residual for all grid points, exp-weighted time
integration, damping + residual norms, energy.  For N > 1 the imaginary-time
grid is sharded over ranks (strong scaling: the job is fixed).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (norb, ngrid, Emax)
    "ueg_ft_ccsd_ESN19": (19, 10, 30.0),
    "ueg_ft_ccsd_ESN33": (33, 10, 30.0),
    "ueg57_ng16": (57, 16, 30.0),
}
T_, MU_, L_ = 0.5, 7.0, 1.942


def _traffic(workload, world):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/r1_traffic.json), or None when no capture matches this configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["gemm_tab_kernel"]
        if t["workload"] == workload and t["n_gpus"] == world:
            return t["bytes_per_launch"]
    except Exception:
        pass
    return None


def algorithmic_flops(m, ng):
    """F_T of SURVEY.md 8(d): ng*(64 m^6 + 120 m^5), unrestricted formulation."""
    return ng*(64.0*m**6 + 120.0*m**5)


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                              "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(self.rows), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
def cpu_port_step(m, ng, sample_points, threads):
    """Time the oracle's Sz-blocked CPU restatement of one residual evaluation
    on `sample_points` grid points; returns seconds per full iteration
    (scaled by ng/sample_points) and the sample description."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy
    from kelvin_oracle import spin_blocked as sb, driver as odrv
    import util
    ints, amps = util.random_u(m, m, sample_points, seed=0, scale=0.05)
    w = sb.wrap_integrals(*ints)
    t0 = time.time()
    outs = []
    for y in range(sample_points):
        outs.append(sb.u_stanton_terms(*ints, (amps[0][y], amps[1][y]),
                                       (amps[2][y], amps[3][y], amps[4][y]), wrapped=w))
    t_res = time.time() - t0
    # integration of the sampled points (exp-weighted quadrature, kelvin/quadrature.py:292-317)
    e = numpy.linspace(0.1, 5.0, m)
    D2 = e[:, None, None, None] + e[None, :, None, None] - e[None, None, :, None] - e[None, None, None, :]
    ti, g, G = odrv.simpsons(ng, 1.0/T_)
    tb = numpy.broadcast_to(outs[0][3][None], (ng,) + outs[0][3].shape)
    t0 = time.time()
    odrv.int_tbar(ng, tb, ti, D2, G)
    t_int = (time.time() - t0)*3.0          # three T2 blocks
    return t_res*ng/sample_points + t_int, t_res, t_int


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle port)
    on all host threads; one step = a bounded sample scaled to one iteration."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    norb, ng, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample = max(1, min(ng, args.cpu_points))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_port_step(norb, ng, 1, threads)
    ts = []
    for _ in range(max(1, min(args.steps, 3))):
        s_iter, _, _ = cpu_port_step(norb, ng, sample, threads)
        ts.append(s_iter)
    val = sum(ts)/len(ts)
    line = {
        "impl": "reference", "metric": "ft_ccsd_seconds_per_amplitude_iteration", "value": val,
        "unit": "s", "n_gpus": args.gpus, "steps": len(ts), "warmup": min(args.warmup, 1),
        "ms_per_step": val*1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "norb": norb, "ngrid": ng, "formulation": "u"},
        "cpu_baseline": {"value": val, "unit": "s", "cores": threads, "kind": "port",
                         "sample": "oracle Sz-blocked NumPy/BLAS residual on %d of %d grid points "
                                   "(random amplitudes of the benchmark's shape) + NumPy "
                                   "integration, scaled to one iteration" % (sample, ng)},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kb200", choices=["kb200", "reference"])
    ap.add_argument("--workload", default="ueg_ft_ccsd_ESN33", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-points", type=int, default=6,
                    help="grid points sampled by the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from kelvin_b200 import _lib, cc_utils, ft_utils, parallel, quadrature, ft_cc_equations
    from kelvin_b200 import plan as _plan
    from kelvin_b200.ueg_system import UEGSystem
    lib = _lib.load()
    dev = _lib.device()

    norb, ng, emax = WORKLOADS[args.workload]
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    beta = 1.0/T_
    sysm = UEGSystem(T_, L_, emax, mu=MU_, norb=norb, orbtype='u')
    ea, eb = sysm.u_energies_tot()
    ti, g, G = quadrature.ft_quad(ng, beta, 'lin')
    Fa, Fb, Ia, Ib, Iabab = cc_utils.uft_integrals(sysm, ea, eb, beta, MU_)
    Ds = (ft_utils.D1(ea, ea), ft_utils.D1(eb, eb), ft_utils.D2(ea, ea),
          ft_utils.D2u(ea, eb, ea, eb), ft_utils.D2(eb, eb))
    solver = parallel.TauShardedUCCSD(Fa, Fb, Ia, Ib, Iabab, Ds, g, G, beta, ng, ti)

    def mp2_guess():
        def rep(x):
            return (-x).expand(*((ng,) + (-1,)*x.dim())).contiguous()
        return [quadrature.int_tbar(ng, rep(x), ti, d, G) for x, d in
                zip((Fa.vo, Fb.vo, Ia.vvoo, Iabab.vvoo, Ib.vvoo), Ds)]
    guess = mp2_guess()
    solver.set_amplitudes(*guess)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------
    for _ in range(W):
        E, res = solver.step(0.0)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    lib.kb200_launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        E, res = solver.step(0.0)
    ev1.record()
    barrier()
    launches = int(lib.kb200_launch_count())
    t_dev = ev0.elapsed_time(ev1)*1e-3
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = float(tt.item())/K

    # ---- auxiliary: the same step without the closed-shell reduction (every beta block
    # evaluated, all 32 block GEMMs per grid point), for a like-for-like flop count
    t_general = None
    closed = bool(solver.closed_shell)
    fc = torch.tensor([1.0 if closed else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fc, op=dist.ReduceOp.MAX)      # ranks without grid points report False
    if fc.item() > 0:
        solver.closed_shell = False
        for _ in range(2):
            solver.step(0.0)
        barrier()
        ev0.record()
        for _ in range(K):
            solver.step(0.0)
        ev1.record()
        barrier()
        tg = torch.tensor([ev0.elapsed_time(ev1)*1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        t_general = float(tg.item())/K
        solver.closed_shell = closed
        solver.step(0.0)

    # ---- optional per-phase breakdown of the sharded step (diagnostics, outside the timed region)
    phases = None
    if os.environ.get("KB200_PHASES"):
        solver.phase_ms = {}
        for _ in range(3):
            solver.step(0.0)
        phases = {k: v/3.0 for k, v in solver.phase_ms.items()}
        solver.phase_ms = None
        print("rank %d phases (ms/step): %s" % (rank, json.dumps(phases)), flush=True)

    # ---- end-to-end: host amplitudes in pinned memory -> device -> iterate -> E,res to host
    host = [x.cpu().pin_memory() for x in solver.old]
    h2d = sum(x.numel()*8 for x in host)

    def e2e_step():
        # the host copy is the solver's own state: whether T[0] vanishes is already known
        solver.set_local_amplitudes([x.to(dev, non_blocking=True) for x in host],
                                    t0_zero=solver.t0_zero, closed_shell=solver.closed_shell)
        return solver.step(0.0)
    for _ in range(2):
        e2e_step()
    barrier()
    ev0.record()
    for _ in range(K):
        Ee, rese = e2e_step()
    ev1.record()
    barrier()
    tt[0] = ev0.elapsed_time(ev1)*1e-3
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_e2e = float(tt.item())/K

    # ---- live roofline of the dominant kernel (DMMA contraction GEMM) --------
    roof = None
    if rank == 0:
        # the plan the step runs (closed-shell reduction when the inputs allow it), on the grid
        # points the step evaluates (without tau_0 when the shortcut applies)
        skip = 1 if (solver.t0_zero and solver.nloc > 1) else 0
        nloc = solver.nloc - skip
        mrows = closed and nloc >= ft_cc_equations.MIRROR_ROWS_MIN_BATCH
        p = ft_cc_equations.stanton_plan("u", ft_cc_equations._u_sizes(Fa, Fb), -1.0,
                                         mirror=closed, mirror_rows=mrows)
        t = ft_cc_equations._u_integral_slots(
            Fa, Fb, Ia, Ib, Iabab, dev, [s for s in p.inputs if _plan.is_integral_slot(s)])
        for nm, x in zip(("t1.a", "t1.b", "t2.aa", "t2.ab", "t2.bb"), solver.old):
            if nm in p.shapes:
                t[nm] = x[skip:]
        for nm, x in zip(("o1.a", "o1.b", "o2.aa", "o2.ab", "o2.bb"), solver.old):
            if nm in p.shapes:
                t[nm] = torch.empty_like(x[skip:])
        tim = []
        for _ in range(3):
            tim = []
            p.run(t, nloc, timings=tim)
        gemm = [(fl, dt) for kind, fl, dt, meta in tim if kind == 0]
        # the dominant kernel: every launch of the big-tile contraction kernel
        # (gemm_tab_kernel<4,4,32,32,...>: the m^6 block GEMMs -- full, half-K, half-row and
        # triangular members -- and the diagonal passes grouped with them); a launch group's time
        # is split over its members in proportion to their flops, so sums over members = sums
        # over launches
        big = [(fl, dt, meta) for kind, fl, dt, meta in tim if kind == 0 and meta[4] in (0, 2, 3)]
        fl_big = sum(x[0] for x in big)
        dt_big = sum(x[1] for x in big)
        n_launch = max(1, sum(1 for x in big if x[2][8] >= 1))      # group leaders
        n_m6 = sum(1 for x in big if x[0] >= 0.2*2.0*nloc*norb**6)
        # measured FP64 tensor peak: cuBLAS DGEMM 8192^3, best of 5 (same box, same run)
        n = 8192
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        best = 1e9
        for _ in range(6):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            torch.matmul(A, B)
            a1.record()
            torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(a1)*1e-3)
        del A, B
        peak = 2.0*n**3/best/1e12
        ach = fl_big/dt_big/1e12 if dt_big > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "kb200::gemm_tab_kernel (FP64 DMMA, 128x128x16 CTA tile, gathered operands)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach/peak,
                "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "launches_per_step": n_launch, "contractions_per_step": len(big),
                "m6_block_gemms_per_step": n_m6,
                "flops_per_launch": fl_big/n_launch, "avg_launch_s": dt_big/n_launch,
                "gemm_share_of_plan": sum(d for _, d in gemm)/max(1e-12, sum(x[2] for x in tim)),
                "plan_s": sum(x[2] for x in tim),
                "traffic": _traffic(args.workload, world)}

    # ---- auxiliary: one Lambda iteration (reverse sweep, intermediates cached), N = 1 only
    lam = None
    if world == 1:
        Ts = solver.old
        Ls = ft_cc_equations.uccsd_lambda_guess(Fa, Fb, Ia, Ib, Iabab, Ts[0], Ts[1], beta, ng)

        def lam_step(Lc):
            return ft_cc_equations.uccsd_lambda_opt(Fa, Fb, Ia, Ib, Iabab, *Ts, *Lc, *Ds, ti, ng, g,
                                                    G, beta, closed_shell=closed)
        for _ in range(2):
            Ls = lam_step(Ls)
        torch.cuda.synchronize()
        ev0.record()
        nl = 3
        for _ in range(nl):
            Ls = lam_step(Ls)
        ev1.record()
        torch.cuda.synchronize()
        lam = ev0.elapsed_time(ev1)*1e-3/nl
        del Ls

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        s_iter, t_res, t_int = cpu_port_step(norb, ng, args.cpu_points, threads)
        cpu = {"value": s_iter, "unit": "s", "cores": threads, "kind": "port",
               "sample": "oracle Sz-blocked NumPy/BLAS residual on %d of %d grid points (%.1f s) + "
                         "NumPy integration (%.1f s), scaled to one iteration" %
                         (args.cpu_points, ng, t_res, t_int)}

    # rank 0 owns tau_0 (contiguous blocks from y = 0)
    t0 = bool(solver.t0_zero) if rank == 0 else None
    if rank == 0:
        fl = algorithmic_flops(norb, ng)
        # tau_0 shortcut: T[0] = 0 identically (row 0 of G vanishes), so T̄[0] = drivers and that
        # grid point is not evaluated (SURVEY 8d allows it; the reference's own pointwise solver
        # does the same, kelvin/cc_utils.py:205-208).  TFLOP/s is quoted on the EXECUTED flops.
        npts = ng - 1 if t0 else ng
        fl_exec = algorithmic_flops(norb, npts)
        if closed:
            # executed 2*M*N*K of the reduced program (all contraction classes)
            pc = ft_cc_equations.stanton_plan(
                "u", ft_cc_equations._u_sizes(Fa, Fb), -1.0, mirror=True,
                mirror_rows=npts >= ft_cc_equations.MIRROR_ROWS_MIN_BATCH)
            fl_exec = float(pc.flops_per_point)*npts
        line = {
            "metric": "ft_ccsd_seconds_per_amplitude_iteration", "value": t_step, "unit": "s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_step*1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "system": "UEG T=0.5 mu=7 L=1.942",
                       "norb": norb, "ngrid": ng, "formulation": "u", "damp": 0.0,
                       "parallelism": "tau%d" % world,
                       "cache": "working set (amplitudes+integrals+intermediates) >> 126 MB L2",
                       "algorithmic_tflop_per_step": fl/1e12,
                       "tau0_shortcut": t0, "tau_points_evaluated": ng - 1 if t0 else ng,
                       "closed_shell_reduction": closed,
                       "general_path_s_per_iteration": t_general,
                       "executed_tflop_per_step": fl_exec/1e12,
                       "fp64_tflops_whole_step": fl_exec/t_step/1e12,
                       "published_cpu_s_per_iter_unknown_hw": 321.4 if norb == 33 else None,
                       "lambda_s_per_iteration": lam,
                       "lambda_algorithmic_tflop": ng*92.0*norb**6/1e12},
            "e2e": {"value": t_e2e, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 160},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "last_energy": E, "last_residual": res,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

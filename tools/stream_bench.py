"""Achieved HBM GB/s of the streaming kernels on ESN33-sized tensors (algorithmic bytes / CUDA-event
time of a graph replay of the call: no host time between the launches)."""
import ctypes
import json
import os
import sys
import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from kelvin_b200 import _lib, cc_utils, ft_cc_energy, quadrature  # noqa: E402
from gemm_bench import time_fn  # noqa: E402


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = _lib.device()
    peak = 6542.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    N = m**4
    e = numpy.sort(numpy.random.default_rng(0).uniform(0, 5, m))
    D2 = torch.as_tensor(e[:, None, None, None] + e[None, :, None, None] - e[None, None, :, None]
                         - e[None, None, None, :]).to(dev).contiguous()
    ti, g, G = quadrature.ft_quad(ng, 2.0, 'lin')
    X = 0.1*torch.randn((ng, m, m, m, m), dtype=torch.float64, device=dev)
    Y = 0.1*torch.randn((ng, m, m, m, m), dtype=torch.float64, device=dev)
    I4 = torch.randn((m, m, m, m), dtype=torch.float64, device=dev)
    s = torch.rand(m, dtype=torch.float64, device=dev)
    T1 = 0.1*torch.randn((ng, m, m), dtype=torch.float64, device=dev)
    rows = []

    def graph_time(fn, reps=5):
        """Device time of one call: `reps` calls captured into one CUDA graph, replayed (the
        eager call is host-bound for these 20-200 us kernels: Python + ctypes take ~100 us)."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(reps):
                    fn()
        except Exception:
            torch.cuda.synchronize()
            return time_fn(fn, iters=7, warm=3)[0]
        return time_fn(gr.replay, iters=7, warm=2)[0]/reps

    def rec(name, nbytes, fn, note=""):
        t = graph_time(fn)
        gbs = nbytes/t/1e9
        rows.append((name, nbytes/1e6, t*1e6, gbs, gbs/peak, note))
        print("%-34s %8.1f MB %8.1f us %8.1f GB/s  %.2f of %.0f  %s" % (name, nbytes/1e6, t*1e6, gbs, gbs/peak, peak, note),
              flush=True)
    for mode in (1, 0):
        rec("int_tbar mode %d (T2 block)" % mode, (16*ng + 8)*N,
            lambda: quadrature.int_tbar(ng, X, ti, D2, G, mode=mode), "ng-1 exps" if mode else "ng^2/2 exps")
    rec("int_L mode 1 (L2 block)", (16*ng + 8)*N, lambda: quadrature.int_L(ng, X, ti, D2, g, G, mode=1))
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    Iab0 = ft_cc_energy.oovv_to_abij(I4)
    rec("int_tbar_update (fused, T2 block)", (24*ng + 16)*N,
        lambda: quadrature.int_tbar_update(ng, X, ti, D2, G, Y, 0.3, stats, g=g, W=Iab0, T1x=T1, T1y=T1,
                                           c2=0.25, c11=0.5),
        "integrate + norms + damp + energy: read tbar, read+write T, D, <ij||ab>")
    st = cc_utils._Stats(1, dev)
    rec("damp_norms (T2 block)", 24*ng*N, lambda: st.damp(0, X, Y, 0.3), "read old,new; write old")
    Iab = ft_cc_energy.oovv_to_abij(I4)
    rec("energy_pair (T2 block, Qterm)", (8*ng + 8)*N,
        lambda: ft_cc_energy.energy_terms([], [(X, T1, T1, Iab, 0.25, 0.5)], g, dev))
    rec("dress4", 16*N, lambda: cc_utils._dress4(I4, s, s, s, s))
    rec("permute oovv->abij", 16*N, lambda: ft_cc_energy.oovv_to_abij(I4))
    from kelvin_b200 import ft_cc_equations
    rec("gsum (sum_y g_y X_y)", 8*(ng + 1)*N, lambda: ft_cc_equations._gsum(X, g, dev))
    rec("dot_keep 'yijab,yabij->y'", 16*ng*N, lambda: _lib.dot_keep(X, "yijab", Y, "yabij", "y"))
    rec("dot_keep 'cdab,abcd->c'", 16*N, lambda: _lib.dot_keep(I4, "cdab", Iab, "abcd", "c"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_stream_bench.txt"), "w") as f:
        f.write("# streaming kernels, m=%d ng=%d, algorithmic bytes / CUDA-event time, peak = %.0f GB/s (MEASURED_PEAKS.json)\n" % (m, ng, peak))
        for r in rows:
            f.write("%-34s %8.1f MB %8.1f us %8.1f GB/s  frac %.2f  %s\n" % r)


if __name__ == "__main__":
    main()

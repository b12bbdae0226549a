"""Micro-benchmark: cuBLAS DGEMM peak vs the kb200 gathered DMMA GEMM (run on the GPU box)."""
import json
import sys
import os
import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import engine, plan  # noqa: E402


def time_fn(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)*1e-3)
    return min(ts), sorted(ts)[len(ts)//2]


def main():
    out = {}
    for n in (() if os.environ.get("KB200_SKIP_CUBLAS") else (1089, 2048, 4096, 8192)):
        A = torch.randn(n, n, dtype=torch.float64, device="cuda")
        B = torch.randn(n, n, dtype=torch.float64, device="cuda")
        t, tm = time_fn(lambda: torch.matmul(A, B))
        out["cublas_dgemm_%d" % n] = 2.0*n**3/t/1e12
        print("cuBLAS dgemm %5d: %.2f TF (best) %.2f TF (median)" % (n, 2.0*n**3/t/1e12, 2.0*n**3/tm/1e12), flush=True)
    nb = int(os.environ.get("KB200_NB", "10"))
    A = torch.randn(nb, 1089, 1089, dtype=torch.float64, device="cuda")
    B = torch.randn(nb, 1089, 1089, dtype=torch.float64, device="cuda")
    t, tm = time_fn(lambda: torch.bmm(A, B))
    out["cublas_bmm_10x1089"] = 2.0*nb*1089**3/t/1e12
    print("cuBLAS bmm 10x1089: %.2f TF" % out["cublas_bmm_10x1089"], flush=True)
    # kb200: ladder-like [ab][ef] x [ef][ij], m = 33, batch 10, all four operand modes
    m = 33
    dims = dict(a=m, b=m, e=m, f=m, i=m, j=m)
    tiles = [int(x) for x in os.environ.get("KB200_TILES", "0").split(",")]
    cases = [(t_, la, lb, "tile%d %s" % (t_, tag)) for t_ in tiles for la, lb, tag in
             (("abef", "efij", "A:kc B:nc"), ("efab", "ijef", "A:mc B:kc"),
              ("abef", "ijef", "A:kc B:kc"), ("efab", "efij", "A:mc B:nc"),
              ("aeim", "mbej", "ring-like"))]
    for tile_, la, lb, tag in cases:
        plan.BIG_TILE = tile_
        lc = "abij"
        Ash = tuple(dims.get(l, m) for l in la)
        Bsh = tuple(dims.get(l, m) for l in lb)
        ops = [plan.ROp(("C", lc), 1.0, [("A", la), ("B", lb)])]
        shapes = {"C": (m,)*4, "A": Ash, "B": Bsh}
        p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], shapes=shapes,
                        batched={"C": True, "A": True, "B": False})
        tns = {"A": torch.randn((nb,) + Ash, dtype=torch.float64, device="cuda"),
               "B": torch.randn(Bsh, dtype=torch.float64, device="cuda"),
               "C": torch.empty((nb,) + (m,)*4, dtype=torch.float64, device="cuda")}
        t, tm = time_fn(lambda: p.run(tns, nb))
        fl = 2.0*nb*m**6
        out["kb200_%s" % tag] = fl/t/1e12
        print("kb200 gemm m=33 xNB %-10s: %.2f TF (best) %.2f TF (median)  %.1f us" %
              (tag, fl/t/1e12, fl/tm/1e12, t*1e6), flush=True)
    # large single problem
    for m2 in (57,):
        dims = dict(a=m2, b=m2, e=m2, f=m2, i=m2, j=m2)
        ops = [plan.ROp(("C", "abij"), 1.0, [("A", "abef"), ("B", "efij")])]
        shapes = {"C": (m2,)*4, "A": (m2,)*4, "B": (m2,)*4}
        p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], shapes=shapes,
                        batched={"C": True, "A": True, "B": False})
        tns = {"A": torch.randn((2,) + (m2,)*4, dtype=torch.float64, device="cuda"),
               "B": torch.randn((m2,)*4, dtype=torch.float64, device="cuda"),
               "C": torch.empty((2,) + (m2,)*4, dtype=torch.float64, device="cuda")}
        t, tm = time_fn(lambda: p.run(tns, 2), iters=3, warm=1)
        fl = 2.0*2*m2**6
        out["kb200_m%d" % m2] = fl/t/1e12
        print("kb200 gemm m=%d x2: %.2f TF" % (m2, fl/t/1e12), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

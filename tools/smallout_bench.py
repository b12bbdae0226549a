"""Micro-benchmark of the small-output long-K contraction class (F_vv-type builds)."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import engine, plan  # noqa: E402
from gemm_bench import time_fn  # noqa: E402


def main():
    m, nb = 33, 10
    dims = dict(a=m, e=m, m=m, n=m, f=m)
    for la, lb, tag in (("mnef", "afmn", "Fvv-like"), ("efmn", "afmn", "both K-contig")):
        lc = "ae"
        Ash = tuple(dims[l] for l in la)
        Bsh = tuple(dims[l] for l in lb)
        ops = [plan.ROp(("C", lc), 1.0, [("A", la), ("B", lb)])]
        shapes = {"C": (m, m), "A": Ash, "B": Bsh}
        p = engine.Plan(ops, "g", None, ["A", "B"], ["C"], shapes=shapes,
                        batched={"C": True, "A": False, "B": True})
        tns = {"A": torch.randn(Ash, dtype=torch.float64, device="cuda"),
               "B": torch.randn((nb,) + Bsh, dtype=torch.float64, device="cuda"),
               "C": torch.empty((nb, m, m), dtype=torch.float64, device="cuda")}
        for tile in (5, 1):
            for sk in (8, 15, 30, 60, 120, 240):
                arr, ws = p._ops_for(nb)
                arr[0].tile = tile
                arr[0].splitk = sk
                p._ops[nb] = (arr, 10*240*m*m*8*2)
                t, tm = time_fn(lambda: p.run(tns, nb))
                print("%-14s tile=%d splitk=%3d: %7.1f us" % (tag, tile, sk, t*1e6), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    main()

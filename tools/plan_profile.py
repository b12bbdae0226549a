"""Per-op device-time breakdown of a contraction plan on the benchmark workload."""
import os
import sys
import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import _lib, cc_utils, ft_cc_equations, ft_utils, plan as _plan, quadrature  # noqa: E402
from kelvin_b200.ueg_system import UEGSystem  # noqa: E402


def main():
    norb = int(sys.argv[1]) if len(sys.argv) > 1 else 33
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    which = sys.argv[3] if len(sys.argv) > 3 else "stanton"
    closed = len(sys.argv) > 4 and sys.argv[4] == "closed"     # closed-shell (mirror) reduction
    T_, MU_, L_ = 0.5, 7.0, 1.942
    dev = _lib.device()
    beta = 1.0/T_
    sysm = UEGSystem(T_, L_, 30.0, mu=MU_, norb=norb, orbtype='u')
    ea, eb = sysm.u_energies_tot()
    Fa, Fb, Ia, Ib, Iabab = cc_utils.uft_integrals(sysm, ea, eb, beta, MU_)
    sizes = ft_cc_equations._u_sizes(Fa, Fb)
    if which == "stanton":
        p = ft_cc_equations.stanton_plan(
            "u", sizes, -1.0, mirror=closed, singlet=closed, emit_aa=not closed,
            mirror_rows=closed and ng >= ft_cc_equations.MIRROR_ROWS_MIN_BATCH)
    elif which == "lambda-sweep":
        p = ft_cc_equations.lambda_split_plans("u", sizes, -1.0, mirror=closed)[1]
    else:
        p = ft_cc_equations.lambda_plan("u", sizes, -1.0, mirror=closed)
    torch.manual_seed(0)
    t = ft_cc_equations._u_integral_slots(Fa, Fb, Ia, Ib, Iabab, dev,
                                          [s for s in p.inputs if _plan.is_integral_slot(s)])
    m = norb
    shp1, shp2 = (ng, m, m), (ng, m, m, m, m)
    for s in p.inputs + p.outputs:
        if s in t:
            continue
        shape = (ng,) + tuple(p.shapes[s])
        t[s] = 0.01*torch.randn(shape, dtype=torch.float64, device=dev)
    for _ in range(2):
        tim = []
        p.run(t, ng, timings=tim)
    rops = p.low.rops
    rows = []
    for (kind, fl, dt, meta), op in zip(tim, rops):
        rows.append((dt, kind, fl, meta, repr(op)))
    tot = sum(r[0] for r in rows)
    print("total %.3f ms over %d ops" % (tot*1e3, len(rows)))
    cls = {}
    for dt, kind, fl, meta, txt in rows:
        if kind == 1:
            key = "permute"
        elif kind == 3:
            key = "fused elementwise (kind 3)"
        elif kind == 0 and meta[4] == 7:
            key = "skinny streaming (tile 7)"
        elif kind == 0 and meta[4] == 6:
            key = "long-K tiny output (tile 6)"
        elif kind == 2:
            key = "rank-k update (kind 2)"
        elif fl >= 0.4*2*ng*m**6:
            key = "gemm m^6"
        elif meta[5] > 1:
            key = "gemm split-K"
        elif meta[4] == 1:
            key = "gemm tile1 (N<=48)"
        else:
            key = "gemm other"
        c = cls.setdefault(key, [0, 0.0, 0.0])
        c[0] += 1
        c[1] += dt
        c[2] += fl
    for k, (n, dt, fl) in sorted(cls.items(), key=lambda x: -x[1][1]):
        print("%-22s n=%3d  %8.3f ms  %5.1f%%  %7.2f TF" % (k, n, dt*1e3, 100*dt/tot, fl/max(dt, 1e-12)/1e12))
    print()
    top = int(sys.argv[5]) if len(sys.argv) > 5 else (0 if closed else 70)
    for dt, kind, fl, meta, txt in sorted(rows, key=lambda r: -r[0])[:top]:
        print("%7.1f us k%d M=%6d N=%5d K=%6d b=%2d tile=%d sk=%2d am=%d bm=%d %6.2f TF  %s" %
              (dt*1e6, kind, meta[0], meta[1], meta[2], meta[3], meta[4], meta[5], meta[6], meta[7],
               fl/max(dt, 1e-12)/1e12, txt[:70]))


if __name__ == "__main__":
    main()

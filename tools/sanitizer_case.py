"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): one closed-shell and one
general FT-UCCSD amplitude iteration + one Lambda iteration at 19 plane waves on a short grid
(every kernel class of the path: big-tile / skinny / long-K / matrix-vector contractions,
fused elementwise, permutes, fused integrate-update, int_L, damping), concurrent launches on.

  compute-sanitizer --tool racecheck python tools/sanitizer_case.py [norb] [ngrid]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kelvin_b200 import cc_utils, ft_cc_equations as fe, ft_utils, quadrature  # noqa: E402
from kelvin_b200.ueg_system import UEGSystem  # noqa: E402


def main():
    norb = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    ng = int(sys.argv[2]) if len(sys.argv) > 2 else 3
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
    print("closed-shell step:", st.flags(), st.step(0.1))
    st.set_flags({"t0": st.t0, "closed_shell": False, "singlet": False, "antisym": True})
    print("general step:", st.step(0.1))
    Ls = fe.uccsd_lambda_guess(*ints, st.old[0], st.old[1], beta, ng)
    for cs in (True, False):
        out = fe.uccsd_lambda_opt(*ints, *st.old, *Ls, *Ds, ti, ng, g, G, beta, closed_shell=cs,
                                  antisym=True)
        torch.cuda.synchronize()
        print("lambda map (closed=%s): |lo2ab| = %.12e" % (cs, float(out[3].norm())))


if __name__ == "__main__":
    main()

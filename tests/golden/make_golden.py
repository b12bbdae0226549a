"""Generate tests/golden/*.npz by running the UNMODIFIED reference drivers.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's kelvin/*.py (ccsd, cc_utils, ft_cc_equations, quadrature,
ft_cc_energy, ueg_system, hubbard_system) are imported as they are; their three
absent dependencies resolve to the shims in oracle/shims (pyscf.lib.einsum ->
numpy.einsum, cqcpy -> oracle/kelvin_oracle restatement, lattice -> Hubbard1D).
The fixtures therefore pin (i) every host-side convention of the reference
(grids, dressing, loops, RDM assembly, E/S/N) exactly and (ii) the restated
cqcpy arithmetic to the extent the reference's published numbers do
(tests/golden/published.py).
"""
import logging
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "oracle"), "/root/reference"]

from kelvin import quadrature  # noqa: E402
from kelvin.ccsd import ccsd  # noqa: E402
from kelvin.hubbard_system import HubbardSystem  # noqa: E402
from kelvin.ueg_system import UEGSystem  # noqa: E402
from lattice.hubbard import Hubbard1D  # noqa: E402

logging.basicConfig(level=logging.WARNING)


def quad_fixture():
    out = {}
    beta = 2.5
    for quad in ('lin', 'ln', 'sin', 'exp', 'quad', 'cub', 'quar', 'mid', 'L'):
        for ng in (7, 8):
            ti, g, G = quadrature.ft_quad(ng, beta, quad)
            out["%s_%d_ti" % (quad, ng)] = ti
            out["%s_%d_g" % (quad, ng)] = g
            out["%s_%d_G" % (quad, ng)] = G
            if quad != 'mid':
                gd, Gd = quadrature.d_ft_quad(ng, beta, quad)
                out["%s_%d_gd" % (quad, ng)] = gd
                out["%s_%d_Gd" % (quad, ng)] = Gd
    numpy.savez_compressed(os.path.join(HERE, "quadrature.npz"), **out)


def system_fixture():
    out = {}
    for orb in ("u", "g"):
        s = UEGSystem(0.5, 1.942, 30.0, mu=7.0, norb=7, orbtype=orb)
        out["ueg_%s_N" % orb] = s.N
        out["ueg_%s_mp1" % orb] = s.get_mp1()
        if orb == "u":
            out["ueg_u_ea"] = s.u_energies_tot()[0]
            out["ueg_u_eriab"] = s.u_aint_tot()[2]
            fa, fb = s.u_fock_tot()
            out["ueg_u_fa"] = fa
            da, db = s.u_mp1_den()
            out["ueg_u_mp1den"] = da
            out["ueg_u_fdd"] = numpy.stack(s.u_fock_d_den())
        else:
            out["ueg_g_en"] = s.g_energies_tot()
            out["ueg_g_eri"] = s.g_aint_tot()
            out["ueg_g_f"] = s.g_fock_tot()
            out["ueg_g_mp1den"] = s.g_mp1_den()
    hub, Pa, Pb = hubbard_inputs(4, 2.0)
    for orb in ("u", "g"):
        s = HubbardSystem(1.0, hub, Pa, Pb, mu=0.3, orbtype=orb)
        out["hub_%s_mp1" % orb] = s.get_mp1()
        if orb == "u":
            out["hub_u_ea"], out["hub_u_eb"] = s.u_energies_tot()
            va, vb, vab = s.u_aint_tot()
            out["hub_u_va"], out["hub_u_vb"], out["hub_u_vab"] = va, vb, vab
            out["hub_u_fa"], out["hub_u_fb"] = s.u_fock_tot()
            out["hub_u_mp1den"] = numpy.stack(s.u_mp1_den())
        else:
            out["hub_g_eri"] = s.g_aint_tot()
            out["hub_g_f"] = s.g_fock_tot()
    numpy.savez_compressed(os.path.join(HERE, "systems.npz"), **out)


def hubbard_inputs(L, U):
    hub = Hubbard1D(L, 1.0, U, boundary='p')
    Oa = numpy.zeros(L)
    Ob = numpy.zeros(L)
    Oa[0::2] = 1.0
    Ob[1::2] = 1.0
    return hub, numpy.einsum('i,j->ij', Oa, Oa), numpy.einsum('i,j->ij', Ob, Ob)


def run_fixture(name, sysm, **kw):
    cc = ccsd(sysm, **kw)
    Etot, Ecc = cc.run()
    cc.compute_ESN()
    out = dict(Etot=Etot, Ecc=Ecc, E=cc.E, S=cc.S, N=cc.N, E0=cc.E0, E1=cc.E1, Ecc_=cc.Ecc,
               N0=cc.N0, N1=cc.N1, Ncc=cc.Ncc, S0=cc.S0, S1=cc.S1, Scc=cc.Scc)
    try:
        full1 = cc.full_1rdm()
    except ValueError:      # g branch with truncated occupied set: kelvin/ccsd.py:2001 mis-shapes
        full1 = None
    rel1 = cc.full_1rdm(relax=True)       # kelvin/ccsd.py:1961-2006 (runs _{g,u}_ft_rorb)
    full2 = cc.full_2rdm()                # kelvin/ccsd.py:2008-2053
    if sysm.has_u():
        for k in (0, 1):
            out["full1rdm%d" % k] = full1[k]
            out["rel1rdm%d" % k] = rel1[k]
            out["rorbo%d" % k] = cc.rorbo[k]
            out["rorbv%d" % k] = cc.rorbv[k]
        for k in (0, 1, 2):
            out["full2rdm%d" % k] = full2[k]
    else:
        out.update(rel1rdm=rel1, rorbo=cc.rorbo, rorbv=cc.rorbv, full2rdm=full2)
        if full1 is not None:
            out["full1rdm"] = full1
    if sysm.has_u():
        for k, nm in enumerate(("T1a", "T1b")):
            out[nm] = cc.T1[k]
        for k, nm in enumerate(("T2aa", "T2ab", "T2bb")):
            out[nm] = cc.T2[k]
        for k, nm in enumerate(("L1a", "L1b")):
            out[nm] = cc.L1[k]
        for k, nm in enumerate(("L2aa", "L2ab", "L2bb")):
            out[nm] = cc.L2[k]
        for k in (0, 1):
            out["n1rdm%d" % k] = cc.n1rdm[k]
            out["rono%d" % k] = cc.rono[k]
            out["ronv%d" % k] = cc.ronv[k]
            out["ron1%d" % k] = cc.ron1[k]
            for nm in ("dia", "dba", "dji", "dai"):
                out["%s%d" % (nm, k)] = getattr(cc, nm)[k]
        for k in (0, 1, 2):
            out["n2rdm%d" % k] = cc.n2rdm[k]
        for b, tup in enumerate(cc.P2):
            for k, P in enumerate(tup):
                out["P2_%d_%d" % (b, k)] = P
    else:
        out.update(T1=cc.T1, T2=cc.T2, L1=cc.L1, L2=cc.L2, n1rdm=cc.n1rdm, n2rdm=cc.n2rdm,
                   rono=cc.rono, ronv=cc.ronv, ron1=cc.ron1, dia=cc.dia, dba=cc.dba,
                   dji=cc.dji, dai=cc.dai)
        for b, P in enumerate(cc.P2):
            out["P2_%d" % b] = P
    numpy.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, Etot, Ecc, cc.E, cc.S, cc.N)


def main():
    quad_fixture()
    system_fixture()
    T, mu = 0.1, 0.1
    for orb in ("u", "g"):
        ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype=orb)
        run_fixture("ueg7_" + orb, ueg, T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6,
                    econv=1e-11, tconv=1e-9)
    hub, Pa, Pb = hubbard_inputs(4, 2.0)
    sysm = HubbardSystem(1.0, hub, Pa, Pb, mu=0.3, orbtype='u')
    run_fixture("hubbard4_u", sysm, T=1.0, mu=0.3, iprint=0, max_iter=80, ngrid=8, quad='quad',
                econv=1e-11, tconv=1e-9)
    active_fixtures()
    variant_fixtures()
    pueg_fixture()


def active_fixtures():
    """Occupation-threshold truncation (athresh > 0): rectangular no != nv blocks.
    UEG-7 g at T=0.05, mu=0.3: f_o = (0.9975, 0.018 x6), f_v = (0.0025, 0.982 x6) per spin;
    athresh=0.01 drops the two nearly-empty virtual spin orbitals (nocc=14, nvir=12).
    Hubbard-4 u at T=0.5, mu=0.3: f_o = (0.956, 0.646, 0.032, 0.0028), athresh=0.05 keeps
    nocc=2, nvir=3 per spin."""
    T, mu = 0.05, 0.3
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='g')
    run_fixture("ueg7_g_active", ueg, T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6,
                econv=1e-11, tconv=1e-9, athresh=0.01)
    hub, Pa, Pb = hubbard_inputs(4, 2.0)
    sysm = HubbardSystem(0.5, hub, Pa, Pb, mu=0.3, orbtype='u')
    run_fixture("hubbard4_u_active", sysm, T=0.5, mu=0.3, iprint=0, max_iter=150, damp=0.2, ngrid=8,
                econv=1e-11, tconv=1e-9, athresh=0.05)


def variant_fixtures():
    """Sibling solvers on the same kernels: the pointwise-extrapolated solver
    (rt_iter='point', kelvin/cc_utils.py:176-242,320-411; parameters of
    kelvin/tests/test_ft_ccsd.py:166-201 on a 6-point grid) and FT-CCD (singles=False,
    general spin orbitals) with its Lambda solve."""
    T, mu = 0.1, 0.1
    out = {}
    for orb in ("u", "g"):
        ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype=orb)
        cc = ccsd(ueg, T=T, mu=mu, iprint=0, max_iter=50, damp=0.2, tconv=1e-8, ngrid=6,
                  rt_iter="point")
        Etot, Ecc = cc.run()
        out["point_%s_Etot" % orb], out["point_%s_Ecc" % orb] = Etot, Ecc
        if orb == "u":
            for k, nm in enumerate(("T1a", "T1b")):
                out["point_u_" + nm] = cc.T1[k]
            for k, nm in enumerate(("T2aa", "T2ab", "T2bb")):
                out["point_u_" + nm] = cc.T2[k]
        else:
            out["point_g_T1"], out["point_g_T2"] = cc.T1, cc.T2
        print("point", orb, Etot, Ecc)
    ueg = UEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='g')
    cc = ccsd(ueg, T=T, mu=mu, iprint=0, max_iter=80, damp=0.2, ngrid=6, econv=1e-11, tconv=1e-9,
              singles=False)
    Etot, Ecc = cc.run()
    cc._ft_ccsd_lambda()
    out.update(ccd_Etot=Etot, ccd_Ecc=Ecc, ccd_T1=cc.T1, ccd_T2=cc.T2, ccd_L2=cc.L2)
    print("ccd", Etot, Ecc)
    numpy.savez_compressed(os.path.join(HERE, "ueg7_variants.npz"), **out)


def pueg_fixture():
    """Spin-polarised UEG (kelvin/pueg_system.py) with the parameters of
    kelvin/tests/test_ft_ccsd.py:157-170 (pinned there: Omega_cc = -0.001403909274)."""
    from kelvin.pueg_system import PUEGSystem
    T, mu = 0.1, 0.1
    s = PUEGSystem(T, 2*numpy.pi/numpy.sqrt(1.0), 1.2, mu=mu, norb=7)
    out = dict(N=s.N, mp1=s.get_mp1(), en=s.g_energies_tot(), eri=s.g_aint_tot(), f=s.g_fock_tot(),
               mp1den=s.g_mp1_den(), fdd=s.g_fock_d_den(), fdt=s.g_fock_d_tot(numpy.arange(7.0)),
               dmp1=s.g_d_mp1(numpy.arange(7.0)))
    cc = ccsd(s, T=T, mu=mu, iprint=0, max_iter=50, damp=0.2, ngrid=10)
    Etot, Ecc = cc.run()
    cc.compute_ESN()
    out.update(Etot=Etot, Ecc=Ecc, E=cc.E, S=cc.S, N_=cc.N, T1=cc.T1, T2=cc.T2)
    numpy.savez_compressed(os.path.join(HERE, "pueg7.npz"), **out)
    print("pueg", Etot, Ecc, cc.E, cc.S, cc.N)


def uegscf_fixture():
    """UEG with HF orbital energies (kelvin/ueg_scf_system.py), parameters of
    kelvin/tests/test_ft_deriv.py:277-292 (u path) and the same system through the g path."""
    from kelvin.ueg_scf_system import UEGSCFSystem
    T, mu = 0.1, 0.1
    L = 2*numpy.pi/numpy.sqrt(1.0)
    out = {}
    for orb in ("u", "g"):
        s = UEGSCFSystem(T, L, 1.2, mu=mu, norb=7, orbtype=orb)
        out["N_" + orb] = s.N
        out["mp1_" + orb] = s.get_mp1()
        if orb == "u":
            out["ea"] = s.u_energies_tot()[0]
            out["fa"], out["fb"] = s.u_fock_tot()
            out["mp1den_u"] = numpy.stack(s.u_mp1_den())
            out["fdd_u"] = numpy.stack(s.u_fock_d_den())
            out["fdt_u"] = numpy.stack(s.u_fock_d_tot(numpy.arange(7.0), numpy.arange(7.0) + 1.0))
            out["dmp1_u"] = s.u_d_mp1(numpy.arange(7.0), numpy.arange(7.0) + 1.0)
        else:
            out["en"] = s.g_energies_tot()
            out["f"] = s.g_fock_tot()
            out["mp1den_g"] = s.g_mp1_den()
            out["fdd_g"] = s.g_fock_d_den()
            out["fdt_g"] = s.g_fock_d_tot(numpy.arange(14.0))
            out["dmp1_g"] = s.g_d_mp1(numpy.arange(14.0))
        cc = ccsd(s, T=T, mu=mu, iprint=0, max_iter=50, damp=0.2, ngrid=8)
        Etot, Ecc = cc.run()
        cc.compute_ESN()
        out.update({"Etot_" + orb: Etot, "Ecc_" + orb: Ecc, "E_" + orb: cc.E, "S_" + orb: cc.S,
                    "Ncc_" + orb: cc.N})
        if orb == "u":
            out["T2ab"] = cc.T2[1]
        else:
            out["T2"] = cc.T2
        print("uegscf", orb, repr(Etot), repr(Ecc), repr(cc.E), repr(cc.S), repr(cc.N))
    numpy.savez_compressed(os.path.join(HERE, "uegscf7.npz"), **out)


def esn19_tight_fixture():
    """BASELINE config 0 (bench/ueg_ft_ccsd_ESN19.py:5-22) converged TIGHTLY (econv 1e-12,
    tconv 1e-10) by the unmodified reference drivers: the scalars north_star asks to reproduce
    to 1e-10 Hartree (the published log stops at tconv = 1e-5)."""
    T, mu = 0.5, 7.0
    ueg = UEGSystem(T, 1.942, 30.0, mu=mu, norb=19, orbtype='u')
    cc = ccsd(ueg, T=T, mu=mu, iprint=0, max_iter=80, damp=0.0, ngrid=10, econv=1e-12, tconv=1e-10)
    Etot, Ecc = cc.run()
    cc.compute_ESN()
    out = dict(Etot=Etot, Ecc=Ecc, E=cc.E, S=cc.S, N=cc.N, E0=cc.E0, E1=cc.E1, Ecc_=cc.Ecc,
               N0=cc.N0, N1=cc.N1, Ncc=cc.Ncc, S0=cc.S0, S1=cc.S1, Scc=cc.Scc)
    numpy.savez_compressed(os.path.join(HERE, "esn19_tight.npz"), **out)
    print("esn19_tight", repr(Etot), repr(Ecc), repr(cc.E), repr(cc.S), repr(cc.N))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "esn19_tight":
        logging.getLogger().setLevel(logging.INFO)
        esn19_tight_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "pueg":
        pueg_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "uegscf":
        uegscf_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "active":
        active_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "variants":
        variant_fixtures()
    else:
        main()

"""CPU: host-side logic of the product (grids, systems, C-ABI surface, tau sharding)."""
import ctypes
import os
import re

import numpy
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_quadrature_matches_reference_fixtures():
    from kelvin_b200 import quadrature
    ref = numpy.load(os.path.join(HERE, "golden", "quadrature.npz"))
    beta = 2.5
    for quad in ('lin', 'ln', 'sin', 'exp', 'quad', 'cub', 'quar', 'mid', 'L'):
        for ng in (7, 8):
            ti, g, G = quadrature.ft_quad(ng, beta, quad)
            assert numpy.abs(ti - ref["%s_%d_ti" % (quad, ng)]).max() < 1e-15
            assert numpy.abs(g - ref["%s_%d_g" % (quad, ng)]).max() < 1e-15
            assert numpy.abs(G - ref["%s_%d_G" % (quad, ng)]).max() < 1e-15
            if quad == 'mid':
                with pytest.raises(Exception):
                    quadrature.d_ft_quad(ng, beta, quad)
                continue
            gd, Gd = quadrature.d_ft_quad(ng, beta, quad)
            assert numpy.abs(gd - ref["%s_%d_gd" % (quad, ng)]).max() < 1e-15
            assert numpy.abs(Gd - ref["%s_%d_Gd" % (quad, ng)]).max() < 1e-15
    with pytest.raises(Exception):
        quadrature.ft_quad(5, 1.0, "nope")


def test_quadrature_beta_derivative_by_fd():
    """d(g)/d(beta) by central differences (kelvin/tests/test_quadrature.py:98-110),
    here also for G (the reference's G check is vacuous)."""
    from kelvin_b200 import quadrature
    beta, d = 1.3, 1e-6
    for quad in ('lin', 'ln', 'sin', 'exp', 'quad', 'cub', 'quar'):
        _, gp, Gp = quadrature.ft_quad(9, beta + d, quad)
        _, gm, Gm = quadrature.ft_quad(9, beta - d, quad)
        gd, Gd = quadrature.d_ft_quad(9, beta, quad)
        assert numpy.abs((gp - gm)/(2*d) - gd).max() < 1e-8
        assert numpy.abs((Gp - Gm)/(2*d) - Gd).max() < 1e-8


def test_systems_match_reference_fixtures():
    from kelvin_b200.ueg_system import UEGSystem
    from kelvin_b200.hubbard_system import HubbardSystem, Hubbard1D
    ref = numpy.load(os.path.join(HERE, "golden", "systems.npz"))
    su = UEGSystem(0.5, 1.942, 30.0, mu=7.0, norb=7, orbtype="u")
    assert su.N == float(ref["ueg_u_N"])
    assert abs(su.get_mp1() - float(ref["ueg_u_mp1"])) < 1e-13
    assert numpy.abs(su.u_energies_tot()[0] - ref["ueg_u_ea"]).max() == 0.0
    assert numpy.abs(su.u_aint_tot()[2] - ref["ueg_u_eriab"]).max() == 0.0
    assert numpy.abs(su.u_fock_tot()[0] - ref["ueg_u_fa"]).max() < 1e-14
    assert numpy.abs(su.u_mp1_den()[0] - ref["ueg_u_mp1den"]).max() < 1e-14
    assert numpy.abs(numpy.stack(su.u_fock_d_den()) - ref["ueg_u_fdd"]).max() < 1e-14
    sg = UEGSystem(0.5, 1.942, 30.0, mu=7.0, norb=7, orbtype="g")
    assert numpy.abs(sg.g_aint_tot() - ref["ueg_g_eri"]).max() == 0.0
    assert numpy.abs(sg.g_fock_tot() - ref["ueg_g_f"]).max() < 1e-14
    assert numpy.abs(sg.g_mp1_den() - ref["ueg_g_mp1den"]).max() < 1e-14
    L = 4
    hub = Hubbard1D(L, 1.0, 2.0, boundary='p')
    Oa, Ob = numpy.zeros(L), numpy.zeros(L)
    Oa[0::2] = 1.0
    Ob[1::2] = 1.0
    Pa, Pb = numpy.einsum('i,j->ij', Oa, Oa), numpy.einsum('i,j->ij', Ob, Ob)
    hu = HubbardSystem(1.0, hub, Pa, Pb, mu=0.3, orbtype='u')
    assert abs(hu.get_mp1() - float(ref["hub_u_mp1"])) < 1e-13
    for got, key in zip(hu.u_aint_tot(), ("hub_u_va", "hub_u_vb", "hub_u_vab")):
        assert numpy.abs(got - ref[key]).max() < 1e-14
    for got, key in zip(hu.u_fock_tot(), ("hub_u_fa", "hub_u_fb")):
        assert numpy.abs(got - ref[key]).max() < 1e-14
    assert numpy.abs(numpy.stack(hu.u_mp1_den()) - ref["hub_u_mp1den"]).max() < 1e-14
    hg = HubbardSystem(1.0, hub, Pa, Pb, mu=0.3, orbtype='g')
    assert numpy.abs(hg.g_aint_tot() - ref["hub_g_eri"]).max() < 1e-14
    assert numpy.abs(hg.g_fock_tot() - ref["hub_g_f"]).max() < 1e-14
    assert not su.verify(0.4, 7.0) and su.verify(0.5, 7.0)


def test_capi_exports_every_declared_symbol(built):
    """The C-ABI library loads and exports every function include/kelvin_b200.h declares
    (no compute call is made: there is no GPU here)."""
    from kelvin_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kelvin_b200.h")).read()
    declared = set(re.findall(r"\b(kb200_[A-Za-z0-9_]+)\s*\(", hdr))
    declared.discard("kb200_op")
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for nm in sorted(declared):
        assert hasattr(lib, nm), nm
    assert set(_lib.EXPORTS) <= declared
    assert _lib.load().kb200_version() >= 100
    # the descriptor struct has the layout the header promises
    from kelvin_b200.plan import kb200_op
    assert ctypes.sizeof(kb200_op) == 4*4 + 3*8 + 4*4 + 3*8 + 6*8 + 2*8 + 6*4 + 2*8


def test_product_fails_loudly_without_cuda(built):
    """No CPU fallback: compute entry points raise when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from kelvin_b200 import quadrature, _lib
    ti, g, G = quadrature.ft_quad(4, 1.0, 'lin')
    with pytest.raises(_lib.KB200Error):
        quadrature.int_tbar(4, numpy.zeros((4, 2, 2)), ti, numpy.zeros((2, 2)), G)


def test_shards_partition():
    """parallel.Shards: q whole rows per rank, r leftover rows; every row of [y0, ng) is either
    in exactly one rank's own block or a leftover row; owner mode gives each of the first r
    ranks one leftover row."""
    from kelvin_b200.parallel import Shards
    for ng in (1, 2, 7, 10, 16, 24, 40):
        for y0 in (0, 1):
            for world in (1, 2, 3, 4, 8):
                seen = []
                owners = []
                for rank in range(world):
                    sh = Shards(ng, y0, rank, world)
                    assert sh.q*world + sh.r == max(0, ng - y0)
                    seen += list(range(*sh.own))
                    if sh.owner_row() is not None:
                        owners.append(sh.owner_row())
                    rows = sh.my_rows(False)
                    assert sum(b - a for a, b in rows) == sh.q + (1 if rank < sh.r else 0)
                    if sh.r:
                        assert sh.my_rows(True)[-1] == sh.left
                sh = Shards(ng, y0, 0, world)
                assert sorted(seen) == list(range(*sh.whole))
                assert owners == list(range(*sh.left))
                assert sh.whole[0] == min(y0, ng) or ng <= y0
                assert sh.left[1] == ng
    # ESN33 on 8 GPUs: 9 evaluated points -> one whole row each + one row shared by all
    sh = Shards(10, 1, 3, 8)
    assert (sh.q, sh.r, sh.own, sh.left) == (1, 1, (4, 5), (9, 10)) and sh.use_hybrid()
    # UEG-57 (ng 16, tau_0 skipped) on 8 GPUs: 7 leftover rows go to single owners
    assert not Shards(16, 1, 0, 8).use_hybrid()


def _gloo_worker(rank, world, port, ng, y0, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kelvin_b200 import parallel
    assert parallel.active() and parallel.world_info() == (rank, world)
    sh = parallel.Shards(ng, y0)
    full = torch.arange(ng*6, dtype=torch.float64).reshape(ng, 6) + 1.0
    ok = True
    for owner_left in (True, False):
        flat = torch.full((ng, 6), -7.0, dtype=torch.float64)
        flat[:y0] = full[:y0]                      # rows nobody evaluates (tau_0: the drivers)
        flat[sh.own[0]:sh.own[1]] = full[sh.own[0]:sh.own[1]]
        if owner_left:
            parallel.zero_foreign_left_rows(flat, sh)
            if sh.owner_row() is not None:
                flat[sh.owner_row()] = full[sh.owner_row()]
        else:
            # hybrid mode: the leftover rows were completed inside the evaluation on every rank
            flat[sh.left[0]:sh.left[1]] = full[sh.left[0]:sh.left[1]]
        parallel.exchange_rows(flat, sh, owner_left)
        ok = ok and bool(torch.equal(flat, full))
    acc = torch.full((3,), float(rank + 1), dtype=torch.float64)
    parallel.sum_over_ranks(acc)
    q.put((rank, ok, float(acc[0].item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("ng,y0", [(5, 0), (8, 1), (10, 1), (2, 0)])
def test_row_exchange_gloo_world2(ng, y0):
    """The N>1 exchange step (in-place all-gather of the whole rows + sum of the owner-mode
    leftover rows) and the response-density sum on 2 CPU ranks over gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7*ng + y0) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ng, y0, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, tot in res:
        assert ok and tot == 3.0


@pytest.mark.parametrize("ng,y0", [(10, 1), (7, 0)])
def test_row_exchange_gloo_world4(ng, y0):
    """The same on 4 CPU ranks: ESN33's grid on 4 ranks (2 whole rows each + 1 leftover) and a grid
    with one whole row per rank and three leftover rows."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + 7*ng + y0) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 4, port, ng, y0, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1, 2, 3]
    for rank, ok, tot in res:
        assert ok and tot == 10.0


def test_t0_is_zero_predicate():
    """The tau_0 shortcut is only taken when row 0 of G vanishes and every amplitude block is
    exactly zero at the first grid point."""
    import torch
    from kelvin_b200 import ft_cc_equations, quadrature
    for quad in ("lin", "ln", "sin", "quad", "cub"):
        ti, g, G = quadrature.ft_quad(6, 2.0, quad)
        assert numpy.all(G[0] == 0.0), quad
    ti, g, G = quadrature.ft_quad(6, 2.0, "lin")
    a = torch.zeros(6, 3, 3, dtype=torch.float64)
    b = torch.randn(6, 3, 3, 3, 3, dtype=torch.float64)
    assert not ft_cc_equations.t0_is_zero(G, (a, b))
    b[0] = 0.0
    assert ft_cc_equations.t0_is_zero(G, (a, b))
    G2 = G.copy()
    G2[0, 0] = 1e-3
    assert not ft_cc_equations.t0_is_zero(G2, (a, b))
    assert not ft_cc_equations.t0_is_zero(G[:1, :1], (a[:1], b[:1]))


def test_pueg_system_matches_reference_fixture():
    """kelvin_b200.pueg_system.PUEGSystem against the arrays of the unmodified
    kelvin/pueg_system.py (tests/golden/make_golden.py pueg), and -- through the CPU oracle
    loop -- against the grand potential the reference pins for it
    (kelvin/tests/test_ft_ccsd.py:27,157-170: -0.001403909274 to 1e-8)."""
    from kelvin_b200.pueg_system import PUEGSystem
    from kelvin_oracle import cqc, driver as odrv
    ref = numpy.load(os.path.join(HERE, "golden", "pueg7.npz"))
    T, mu = 0.1, 0.1
    s = PUEGSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7)
    assert s.has_g() and not s.has_u() and s.verify(T, mu) and not s.verify(T, 0.2)
    assert s.N == float(ref["N"])
    assert abs(s.get_mp1() - float(ref["mp1"])) < 1e-14
    assert numpy.abs(s.g_energies_tot() - ref["en"]).max() == 0.0
    assert numpy.abs(s.g_aint_tot() - ref["eri"]).max() == 0.0
    assert numpy.abs(s.g_fock_tot() - ref["f"]).max() < 1e-14
    assert numpy.abs(s.g_mp1_den() - ref["mp1den"]).max() < 1e-14
    assert numpy.abs(s.g_fock_d_den() - ref["fdd"]).max() < 1e-14
    assert numpy.abs(s.g_fock_d_tot(numpy.arange(7.0)) - ref["fdt"]).max() < 1e-14
    assert abs(s.g_d_mp1(numpy.arange(7.0)) - float(ref["dmp1"])) < 1e-14
    # the solve (CPU oracle loop, parameters of the reference test)
    beta, ng = 1.0/T, 10
    en = s.g_energies_tot()
    ti, g, G = odrv.simpsons(ng, beta)
    F, I = odrv.ft_integrals(s, en, beta, mu)
    D1, D2 = cqc.D1(en, en), cqc.D2(en, en)
    T1, T2 = odrv.mp2_guess_g(F, I, D1, D2, ti, ng, G)
    conv = {"econv": 1e-8, "tconv": 1e-5, "max_iter": 50, "damp": 0.2}
    Ecc, T1, T2, hist = odrv.ft_cc_iter(T1, T2, F, I, D1, D2, g, G, beta, ng, ti, conv)
    assert abs(Ecc - float(ref["Ecc"])) < 1e-12
    assert abs(Ecc - (-0.001403909274)) < 1e-8
    assert numpy.abs(T2 - ref["T2"]).max() < 1e-10


def test_uegscf_system_matches_reference_fixture():
    """kelvin_b200.ueg_scf_system.UEGSCFSystem (HF orbital energies of the zero-temperature
    reference determinant) against the arrays of the unmodified kelvin/ueg_scf_system.py
    (tests/golden/make_golden.py uegscf; parameters of kelvin/tests/test_ft_deriv.py:277-292), and
    -- through the CPU oracle loops -- against the reference drivers' grand potential for it on
    the g and the u path."""
    from kelvin_b200.ueg_scf_system import UEGSCFSystem
    from kelvin_oracle import cqc, driver as odrv
    ref = numpy.load(os.path.join(HERE, "golden", "uegscf7.npz"))
    T, mu, ng = 0.1, 0.1, 8
    beta = 1.0/T
    su = UEGSCFSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7)
    sg = UEGSCFSystem(T, 2*numpy.pi, 1.2, mu=mu, norb=7, orbtype='g')
    assert su.has_u() and not sg.has_u() and su.verify(T, mu) and not su.verify(T, 0.2)
    assert list(su.oidx) == [0] and list(su.goidx) == [0, 7]
    assert abs(su.N - float(ref["N_u"])) < 1e-14 and abs(sg.N - float(ref["N_g"])) < 1e-14
    assert abs(su.get_mp1() - float(ref["mp1_u"])) < 1e-14
    assert abs(sg.get_mp1() - float(ref["mp1_g"])) < 1e-14
    assert numpy.abs(su.u_energies_tot()[0] - ref["ea"]).max() < 1e-15
    assert numpy.abs(sg.g_energies_tot() - ref["en"]).max() < 1e-15
    fa, fb = su.u_fock_tot()
    assert numpy.abs(fa - ref["fa"]).max() < 1e-14 and numpy.abs(fb - ref["fb"]).max() < 1e-14
    assert numpy.abs(sg.g_fock_tot() - ref["f"]).max() < 1e-14
    assert numpy.abs(numpy.stack(su.u_mp1_den()) - ref["mp1den_u"]).max() < 1e-14
    assert numpy.abs(sg.g_mp1_den() - ref["mp1den_g"]).max() < 1e-14
    assert numpy.abs(numpy.stack(su.u_fock_d_den()) - ref["fdd_u"]).max() < 1e-14
    assert numpy.abs(sg.g_fock_d_den() - ref["fdd_g"]).max() < 1e-14
    v = numpy.arange(7.0)
    assert numpy.abs(numpy.stack(su.u_fock_d_tot(v, v + 1.0)) - ref["fdt_u"]).max() < 1e-14
    assert numpy.abs(sg.g_fock_d_tot(numpy.arange(14.0)) - ref["fdt_g"]).max() < 1e-14
    assert abs(su.u_d_mp1(v, v + 1.0) - float(ref["dmp1_u"])) < 1e-14
    assert abs(sg.g_d_mp1(numpy.arange(14.0)) - float(ref["dmp1_g"])) < 1e-14
    conv = {"econv": 1e-8, "tconv": 1e-5, "max_iter": 50, "damp": 0.2}
    ti, g, G = odrv.simpsons(ng, beta)
    # g path
    en = sg.g_energies_tot()
    F, I = odrv.ft_integrals(sg, en, beta, mu)
    D1, D2 = cqc.D1(en, en), cqc.D2(en, en)
    T1, T2 = odrv.mp2_guess_g(F, I, D1, D2, ti, ng, G)
    Ecc, T1, T2, hist = odrv.ft_cc_iter(T1, T2, F, I, D1, D2, g, G, beta, ng, ti, conv)
    assert abs(Ecc - float(ref["Ecc_g"])) < 1e-12
    assert numpy.abs(T2 - ref["T2"]).max() < 1e-10
    # u path
    ea, eb = su.u_energies_tot()
    Fa, Fb, Ia, Ib, Iabab = odrv.uft_integrals(su, ea, eb, beta, mu)
    Ds = (cqc.D1(ea, ea), cqc.D1(eb, eb), cqc.D2(ea, ea), cqc.D2u(ea, eb, ea, eb), cqc.D2(eb, eb))
    amps = odrv.mp2_guess_u(Fa, Fb, Ia, Ib, Iabab, *Ds, ti, ng, G)
    Ecc, T1s, T2s, hist = odrv.ft_ucc_iter(*amps, Fa, Fb, Ia, Ib, Iabab, *Ds, g, G, beta, ng, ti, conv)
    assert abs(Ecc - float(ref["Ecc_u"])) < 1e-12
    assert numpy.abs(T2s[1] - ref["T2ab"]).max() < 1e-10

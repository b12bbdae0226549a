"""Imaginary-time sharding of the FT-CCSD iteration over the GPUs of one node.

Given the previous amplitudes, the residual at each grid point is independent
(kelvin/ft_cc_equations.py:145-155); only the quadrature couples grid points
(kelvin/quadrature.py:292-317).  Rank r owns a contiguous block of grid points:

  1. residual T̄[y] for its own y (batched contraction plan, no communication);
  2. ONE exchange: NCCL all-gather of T̄ over NVLink/NVSwitch (ng*N*8 bytes total);
  3. exp-weighted integration of its own rows y from the gathered T̄
     (kb200_int_tbar_rows), damping and partial norms / energy on its own rows;
  4. all-reduce of the <= 20 scalars the convergence test needs.

One process per GPU (torch.distributed, backend nccl; gloo on CPU for the
host-logic tests).  With world_size == 1 this is exactly the loop body of
kelvin/cc_utils.py:274-299.
"""
import numpy
import torch

from . import _lib, ft_cc_energy, ft_cc_equations, quadrature

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def shard_bounds(ng, world):
    """Contiguous blocks of ceil(ng/world) grid points; trailing ranks may be empty.
    (Keeps the all-gathered buffer contiguous in y.)"""
    chunk = (ng + world - 1)//world
    return [(min(ng, r*chunk), min(ng, (r + 1)*chunk)) for r in range(world)], chunk


def world_info(group=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class TauShardedUCCSD(object):
    """State + one-iteration step of the unrestricted FT-CCSD fixed-point loop
    with the tau grid partitioned over ranks."""

    def __init__(self, Fa, Fb, Ia, Ib, Iabab, Ds, g, G, beta, ng, ti, group=None):
        self.dev = _lib.device()
        self.ints = (Fa, Fb, Ia, Ib, Iabab)
        self.Ds = [_lib.as_dev(d, self.dev) for d in Ds]          # D1a, D1b, D2aa, D2ab, D2bb
        self.g = numpy.asarray(g, dtype=numpy.float64)
        self.G = numpy.asarray(G, dtype=numpy.float64)
        self.ti = numpy.asarray(ti, dtype=numpy.float64)
        self.beta = beta
        self.ng = ng
        self.group = group
        self.rank, self.world = world_info(group)
        bounds, self.chunk = shard_bounds(ng, self.world)
        self.y0, self.y1 = bounds[self.rank]
        self.nloc = self.y1 - self.y0
        self.abij = (ft_cc_energy.oovv_to_abij(Ia.oovv), ft_cc_energy.oovv_to_abij(Iabab.oovv),
                     ft_cc_energy.oovv_to_abij(Ib.oovv))
        self.faT = _lib.as_dev(Fa.ov, self.dev).t().contiguous()
        self.fbT = _lib.as_dev(Fb.ov, self.dev).t().contiguous()
        self.stats = torch.zeros(20, dtype=torch.float64, device=self.dev)
        self.old = None
        self.t0_zero = False      # T[0] == 0 on the rank that owns tau_0: that point is skipped
        self.closed_shell = False  # alpha == beta inputs: the reduced (mirror) program runs
        self._gather = None
        self.phase_ms = None      # set to {} to collect per-phase device times (diagnostics)

    # -- amplitudes ---------------------------------------------------------
    def set_amplitudes(self, T1a, T1b, T2aa, T2ab, T2bb):
        """Full (ng, ...) amplitudes -> local shard (copied)."""
        self.old = [_lib.as_dev(x, self.dev)[self.y0:self.y1].clone().contiguous()
                    for x in (T1a, T1b, T2aa, T2ab, T2bb)]
        self._check_t0(None)
        self._check_closed_shell()

    def set_local_amplitudes(self, local, t0_zero=None, closed_shell=None):
        """t0_zero / closed_shell: None = look at the data (device reductions + sync);
        True/False = the caller's knowledge of whether the amplitudes at tau_0 vanish (only
        read on the rank that owns tau_0) / are mirror symmetric (must then be the same on
        every rank)."""
        self.old = [_lib.as_dev(x, self.dev).contiguous() for x in local]
        self._check_t0(t0_zero)
        if closed_shell is None:
            self._check_closed_shell()
        else:
            self.closed_shell = bool(closed_shell)

    def _check_closed_shell(self):
        """Closed shell (alpha == beta integrals, denominators and local amplitudes): the update
        preserves it.  The ranks must agree (the step gathers three blocks instead of five), so
        the local verdicts are combined with a MIN all-reduce; a rank without grid points
        abstains."""
        from . import cc_utils
        ok = True if self.nloc == 0 else bool(cc_utils._closed_shell(*self.ints, self.Ds, self.old))
        if self.world > 1:
            f = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=self.dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=self.group)
            ok = bool(f.item() > 0.5)
        self.closed_shell = ok

    def _check_t0(self, known):
        if self.y0 != 0 or self.nloc < 1 or self.ng < 2 or numpy.any(self.G[0] != 0.0):
            self.t0_zero = False
        elif known is not None:
            self.t0_zero = bool(known)
        else:
            self.t0_zero = all(float(x[0].abs().max()) == 0.0 for x in self.old)

    def full_amplitudes(self):
        """All-gather the local shards into full (ng, ...) tensors."""
        return [self._allgather_rows(x)[:self.ng].clone() for x in self.old]

    def _allgather_rows(self, local):
        shp = tuple(local.shape[1:])
        if self.world == 1:
            return local
        if local.shape[0] == self.chunk:
            pad = local                        # full shard: sent in place
        else:
            pad = torch.zeros((self.chunk,) + shp, dtype=torch.float64, device=local.device)
            pad[:local.shape[0]] = local
        out = torch.empty((self.world*self.chunk,) + shp, dtype=torch.float64, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        return out

    # -- one iteration --------------------------------------------------------
    def step(self, alpha):
        """One damped fixed-point iteration; returns (E, res1 + res2) as the
        reference logs them (kelvin/cc_utils.py:274-305)."""
        lib = _lib.load()
        ng, nloc = self.ng, self.nloc
        Fa, Fb, Ia, Ib, Iabab = self.ints
        marks = []

        def mark(name):
            if self.phase_ms is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))
        mark("start")
        # closed shell: the beta blocks (T1b, T2bb) are copies of the alpha ones -- they are
        # neither gathered nor integrated nor damped, only copied at the end
        live = (0, 2, 3) if self.closed_shell else (0, 1, 2, 3, 4)
        if nloc > 0:
            bars = ft_cc_equations.uccsd_stanton_bar(Fa, Fb, Ia, Ib, Iabab, *self.old,
                                                     t0_zero=self.t0_zero,
                                                     closed_shell=self.closed_shell,
                                                     beta_copies=False)
        else:
            bars = [torch.zeros((0,) + tuple(d.shape), dtype=torch.float64, device=self.dev)
                    for d in self.Ds]
        self.stats.zero_()
        mark("residual")
        new = {}
        fulls = {k: self._allgather_rows(bars[k]) for k in live}
        mark("all-gather")
        for k in live:
            if nloc > 0:
                new[k] = quadrature.int_tbar(ng, fulls[k][:ng], self.ti, self.Ds[k], self.G,
                                             rows=(self.y0, self.y1))
        fulls = None
        mark("integrate")
        if nloc > 0:
            scratch = _lib.reduce_scratch(self.dev)
            for k in live:
                rc = lib.kb200_damp_norms(self.old[k].numel(), _lib.ptr(self.old[k]),
                                          _lib.ptr(new[k]), alpha,
                                          self.stats.data_ptr() + 24*k, _lib.ptr(scratch),
                                          _lib.stream_ptr())
                _lib.check(rc, "kb200_damp_norms")
            if self.closed_shell:
                self.old[1].copy_(self.old[0])
                self.old[4].copy_(self.old[2])
                self.stats[3:6] = self.stats[0:3]
                self.stats[12:15] = self.stats[6:9]
            T1a, T1b, T2aa, T2ab, T2bb = self.old
            parts = ft_cc_energy.energy_terms(
                [(T1a, self.faT), (T1b, self.fbT)],
                [(T2aa, T1a, T1a, self.abij[0], 0.25, 0.5),
                 (T2ab, T1a, T1b, self.abij[1], 1.0, 1.0),
                 (T2bb, T1b, T1b, self.abij[2], 0.25, 0.5)],
                self.g[self.y0:self.y1], self.dev)
            self.stats[15:20] = parts
        mark("damp+energy")
        if self.world > 1:
            dist.all_reduce(self.stats, group=self.group)
        s = self.stats.cpu().numpy()
        mark("all-reduce")
        if self.phase_ms is not None:
            torch.cuda.synchronize()
            for (_, e0), (nm, e1) in zip(marks[:-1], marks[1:]):
                self.phase_ms[nm] = self.phase_ms.get(nm, 0.0) + e0.elapsed_time(e1)
        n = numpy.sqrt(s[:15].reshape(5, 3))
        nl1 = n[0, 1] + 0.1 + n[1, 1]
        nl2 = n[2, 1] + 0.1 + n[3, 1] + n[4, 1]
        res1 = n[0, 0]/nl1 + n[1, 0]/nl1
        res2 = n[2, 0]/nl2 + n[3, 0]/nl2 + n[4, 0]/nl2
        E = float(s[15:20].sum())/self.beta
        return E, float(res1 + res2)

"""Imaginary-time sharding of the FT-CCSD path over the GPUs of one node.

Given the amplitudes of the previous iteration, the residual at each grid point is independent
(kelvin/ft_cc_equations.py:105-107,145-155); the same holds for the Lambda map (:399,442) and
the response densities (:717-720,743-751).  Only the quadrature couples grid points
(kelvin/quadrature.py:292-345), and it is elementwise in the orbital indices.

One process per GPU (torch.distributed, backend nccl; gloo for the CPU tests of the host logic).
State is REPLICATED: every rank holds the full (ng, ...) amplitudes, so the drivers, the loops
and everything cheap (integration, damping, norms, energy: streaming passes) run unchanged and
identically on every rank, and `ccsd.run()` / `compute_ESN()` return the same numbers on every
rank.  What is partitioned is the expensive part, the per-grid-point contraction programs:

  * the rows [y0, ng) to evaluate are dealt out as q = n // P WHOLE rows per rank;
  * the r = n % P LEFTOVER rows (ESN33: 9 evaluated points on 8 GPUs -> q = 1, r = 1) are either
    evaluated by all ranks together, each contracting a slab of the rows of every m^6
    contraction (plan.hybrid_phases / engine.PhasedPlan; two all-reduces inside the evaluation),
    or -- when that does not pay (many leftover rows) -- given one each to the first r ranks;
  * ONE exchange completes the row buffer on every rank: an in-place NCCL all-gather of the
    whole rows (+ an all-reduce of the owner-mode leftover rows).  All blocks of a quantity
    (T1a, T2aa, T2ab, ...) are column ranges of one (ng, Ntot) buffer, so this is one collective.

With world size 1 nothing here is used.
"""
import os

import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None

_cfg = {"group": None,
        "enabled": os.environ.get("KB200_SHARD", "1") != "0",
        "hybrid": os.environ.get("KB200_HYBRID", "1") != "0"}


def configure(group=None, enabled=None, hybrid=None):
    """group: process group to shard over (default: the world); enabled=False: every rank
    evaluates every grid point (replicas); hybrid=False: leftover rows go to single owners."""
    _cfg["group"] = group
    _cfg.pop("xgroup", None)
    if enabled is not None:
        _cfg["enabled"] = bool(enabled)
    if hybrid is not None:
        _cfg["hybrid"] = bool(hybrid)


def group():
    return _cfg["group"]


def exchange_group():
    """Process group for the exchanges INSIDE an evaluation (hybrid partition): the same ranks,
    but NCCL kernels on a high-priority stream, so that they are not queued behind the wide
    contraction launches of the ranks' own grid points.  Created at first use (a collective call:
    every rank gets here in the same evaluation)."""
    if "xgroup" not in _cfg:
        grp = _cfg["group"]
        xg = grp
        try:
            if dist.get_backend(grp) == "nccl":
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                ranks = dist.get_process_group_ranks(grp if grp is not None else dist.group.WORLD)
                xg = dist.new_group(ranks=ranks, backend="nccl", pg_options=opts)
        except Exception:
            xg = grp
        _cfg["xgroup"] = xg
    return _cfg["xgroup"]


def world_info(grp=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        grp = grp if grp is not None else _cfg["group"]
        return dist.get_rank(grp), dist.get_world_size(grp)
    return 0, 1


def active():
    return _cfg["enabled"] and world_info()[1] > 1


def hybrid_enabled():
    return _cfg["hybrid"]


def shard_bounds(ng, world):
    """Contiguous blocks of ceil(ng/world) grid points; trailing ranks may be empty."""
    chunk = (ng + world - 1)//world
    return [(min(ng, r*chunk), min(ng, (r + 1)*chunk)) for r in range(world)], chunk


class Shards(object):
    """Partition of the grid rows [y0, ng) over the ranks: `own` = this rank's whole rows,
    `whole` = all whole rows (rank k: y0 + k q ... y0 + (k+1) q), `left` = the r leftover rows."""

    def __init__(self, ng, y0=0, rank=None, world=None):
        if rank is None or world is None:
            rank, world = world_info()
        self.ng, self.y0, self.rank, self.world = int(ng), int(y0), int(rank), int(world)
        n = max(0, self.ng - self.y0)
        self.q, self.r = divmod(n, self.world)
        self.own = (self.y0 + self.rank*self.q, self.y0 + (self.rank + 1)*self.q)
        self.whole = (self.y0, self.y0 + self.world*self.q)
        self.left = (self.whole[1], self.ng)

    def use_hybrid(self):
        """Deal the leftover rows out by contraction rows (all ranks together) rather than one
        row per rank?  Owner mode costs the busiest rank q + 1 rows, the hybrid q + r/P plus the
        replicated m^5 terms and two exchanges per leftover row (about a tenth of a row) -- and a
        chain of ~100 small launches, which for the smallest benchmark is as long as a row itself
        (ESN33 on 8 B200: own row 1.49 ms, own + shared 3.0 ms, an owner's two rows as one batch
        2.67 ms + the wait of the other ranks: a tie, measured 3.45 ms with the shared row)."""
        return hybrid_enabled() and self.r > 0 and self.r*(1.0/self.world + 0.1) < 1.0

    def owner_row(self):
        """Owner mode: the leftover row this rank evaluates alone, or None."""
        if self.rank < self.r:
            return self.left[0] + self.rank
        return None

    def my_rows(self, hybrid):
        """Row ranges whose data this rank needs: [(start, stop)], own rows first."""
        out = []
        if self.own[1] > self.own[0]:
            out.append(self.own)
        if self.r > 0:
            if hybrid:
                out.append(self.left)
            elif self.owner_row() is not None:
                out.append((self.owner_row(), self.owner_row() + 1))
        return out


def exchange_rows(flat, sh, owner_left):
    """Complete the (ng, Ntot) row buffer on every rank: all-gather of the whole rows in place;
    owner_left: the leftover rows were evaluated by single owners (everybody else holds zeros
    there): summed over the ranks (x + 0 = x: exact)."""
    grp = _cfg["group"]
    if sh.world == 1:
        return
    if not flat.is_contiguous():
        raise Exception("exchange_rows: the row buffer must be contiguous")
    if sh.q > 0:
        out = flat[sh.whole[0]:sh.whole[1]]
        inp = flat[sh.own[0]:sh.own[1]]
        if flat.device.type == "cpu":
            inp = inp.clone()                  # gloo: no aliasing of input and output
        dist.all_gather_into_tensor(out, inp, group=grp)
    if owner_left and sh.r > 0:
        dist.all_reduce(flat[sh.left[0]:sh.left[1]], group=grp)


def zero_foreign_left_rows(flat, sh):
    """Owner mode: clear the leftover rows this rank does not own (before exchange_rows)."""
    mine = sh.owner_row()
    for y in range(sh.left[0], sh.left[1]):
        if y != mine:
            flat[y].zero_()


def sum_over_ranks(t):
    """In-place sum of a tensor over the ranks (response-density accumulation,
    kelvin/ft_cc_equations.py:717-720,743-751)."""
    if active():
        dist.all_reduce(t, group=_cfg["group"])
    return t

"""Execution engine for compiled contraction plans.

A ``Plan`` owns the lowered ``kb200_op`` list, the device copy of the offset
tables and the scratch tensors (intermediates W_oooo, W_vvvv, W_ovvo, ... of
kelvin/lambda_stanton.py:15-34), batched over a chunk of imaginary-time grid
points.  One ``run`` = one C call (kb200_plan_run) that enqueues every kernel
of the residual for that chunk on the current CUDA stream.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib, plan as _plan


GRAPHS = int(os.environ.get("KB200_GRAPH", "1"))
PAD_SCRATCH = int(os.environ.get("KB200_PAD", "1"))


class Plan(object):
    def __init__(self, rops, mode, sizes, inputs, outputs, preset_outputs=(), name="plan",
                 shapes=None, batched=None, antisym=True):
        """rops: resolved ops; inputs/outputs: slot names supplied by the caller
        (outputs are overwritten by their first write unless listed in
        preset_outputs); all other slots are plan-owned scratch."""
        self.name = name
        self.mode = mode
        self.shapes = shapes if shapes is not None else _plan.slot_shapes(rops, mode, sizes)
        self.batched = batched if batched is not None else \
            {s: not _plan.is_integral_slot(s) for s in self.shapes}
        self.inputs = [s for s in self.shapes if s in set(inputs) or _plan.is_integral_slot(s)]
        self.outputs = [s for s in outputs if s in self.shapes]
        preset = list(self.inputs) + list(preset_outputs)
        # plan-owned 4-index scratch is stored with an even pitch of its trailing index pair
        # (plan.padded_strides): rows of every matrix view start on 16-byte boundaries
        pad = [s for s in self.shapes if s not in self.inputs and s not in self.outputs
               and len(self.shapes[s]) == 4] if PAD_SCRATCH else []
        self.low = _plan.Lowered(rops, self.shapes, self.batched, preset, antisym=antisym, pad=pad)
        self.derived = self.low.derived
        self.tmp_slots = [s for s in self.shapes if s not in self.inputs and s not in self.outputs
                          and s not in self.derived]
        self._derived_bufs = {}
        self._dev_tables = None
        self._tmp = None
        self._tmp_nb = 0
        self._tmp_dev = None
        self._ops = {}
        self._ws = None
        self._graphs = {}         # launch signature -> [times seen, CUDA graph or None]
        self.flops_per_point = self.low.flops

    # -- device state ------------------------------------------------------
    def _tables(self, dev):
        if self._dev_tables is None or self._dev_tables.device != dev \
                or self._dev_tables.numel() != self.low.tables.size:
            # uint32 payload carried in an int32 tensor (bit pattern preserved)
            self._dev_tables = torch.from_numpy(self.low.tables.view("int32").copy()).to(dev)
        return self._dev_tables

    def tmp_bytes_per_point(self):
        return sum(8*self.low.slot_size(s) for s in self.tmp_slots)

    def _ensure_tmp(self, nb, dev):
        if self._tmp is not None and self._tmp_nb >= nb and self._tmp_dev == dev:
            return
        self._graphs = {}          # captured launches hold the old scratch addresses
        # triangle blocks (plan.antisym_outputs) must be zero outside the part the plan writes
        self._tmp = {s: (torch.zeros if s.startswith(_plan.TRI_PREFIX) else torch.empty)(
            (nb, self.low.slot_size(s)), dtype=torch.float64, device=dev)
            for s in self.tmp_slots}
        self._tmp_nb = nb
        self._tmp_dev = dev

    def release(self):
        self._tmp = None
        self._tmp_nb = 0
        self._ws = None
        self._graphs = {}

    def _ops_for(self, nb, bstr=(), part=None):
        key = (nb, bstr, part)
        if key not in self._ops:
            arr = self.low.finalize(nb, dict(bstr) if bstr else None, part)
            ws = _lib.load().kb200_plan_workspace_bytes(arr, len(arr))
            self._ops[key] = (arr, int(ws))
        return self._ops[key]

    # -- run -----------------------------------------------------------------
    def run(self, tensors, ng, chunk=None, timings=None, part=None, graph=None):
        """tensors: slot -> CUDA float64 tensor (batched slots: leading axis ng; the grid points
        of a batched slot may be rows of a wider buffer, i.e. any leading stride).
        Scratch is allocated for `chunk` grid points at a time (default: all).
        part = (rank, world): row-slabbed contractions (plan.hybrid_phases) do this rank's rows.
        graph: replay the plan's ~100 launches as ONE CUDA graph once the same call (same
        tensors at the same addresses: the iteration loops keep their buffers) has been seen
        twice; the first call runs eagerly, the second captures.  Default: KB200_GRAPH (on)."""
        lib = _lib.load()
        dev = _lib.device()
        nb_max = ng if chunk is None else max(1, min(int(chunk), ng))
        if graph is None:
            graph = GRAPHS
        if graph and timings is None and nb_max >= ng and ng > 0:
            sig = (ng, part, torch.cuda.current_stream().cuda_stream,
                   tuple((s, tensors[s].data_ptr(), tensors[s].stride(0) if tensors[s].dim() else 0)
                         for s in self.inputs + self.outputs))
            ent = self._graphs.get(sig)
            if ent is not None and ent[1] is not None:
                ent[1].replay()
                lib.kb200_launch_count_add(ent[2])       # the kernels the graph holds
                return
            if ent is None:
                if len(self._graphs) >= 6:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[sig] = [1, None, 0]
            elif not torch.cuda.is_current_stream_capturing():
                ent[0] += 1
                g = torch.cuda.CUDAGraph()
                n0 = lib.kb200_launch_count()
                # thread-local capture mode: the NCCL watchdog thread of a sharded run keeps
                # querying its events and must not invalidate the capture
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.run(tensors, ng, chunk, None, part, graph=False)
                ent[2] = int(lib.kb200_launch_count() - n0)
                lib.kb200_launch_count_add(-ent[2])      # captured, not launched
                ent[1] = g
                g.replay()
                lib.kb200_launch_count_add(ent[2])
                return
        self._ensure_tmp(nb_max, dev)
        names = self.low.slot_names
        bstr = []
        for s in self.inputs + self.outputs:
            t = tensors[s]
            want = ((ng,) if self.batched[s] else ()) + tuple(self.shapes[s])
            inner = t[0] if (self.batched[s] and t.dim() > 0 and t.shape[0] > 0) else t
            if tuple(t.shape) != want or t.dtype != torch.float64 or not t.is_cuda \
                    or not inner.is_contiguous():
                raise Exception("plan %s: slot %s expects contiguous cuda float64 %s, got %s %s"
                                % (self.name, s, want, tuple(t.shape), t.dtype))
            if self.batched[s] and ng > 1 and t.stride(0) != inner.numel():
                bstr.append((self.low.slot_index[s], int(t.stride(0))))
        bstr = tuple(sorted(bstr))
        for name, (src, perm) in self.derived.items():
            st = tensors[src]
            key = (st.data_ptr(), st._version)
            have = self._derived_bufs.get(name)
            # the entry keeps its source tensor alive, so an equal (address, version) can only
            # be that same tensor (a freed tensor's address could be reused by another one)
            if have is None or have[0] != key or have[2] is not st:
                # (once per integral tensor: a torch copy into the padded layout)
                buf = torch.zeros(self.low.slot_size(name), dtype=torch.float64, device=dev)
                buf.as_strided(tuple(self.shapes[name]), tuple(self.low.strides_of(name))).copy_(
                    st.permute(*perm))
                self._derived_bufs[name] = (key, buf, st)
        y0 = 0
        while y0 < ng:
            nb = min(nb_max, ng - y0)
            ops, wsb = self._ops_for(nb, bstr, part)
            tables = self._tables(dev)        # after finalize: it may add (batch-folded) tables
            if wsb > 0 and (self._ws is None or self._ws.numel()*8 < wsb or self._ws.device != dev):
                self._ws = torch.empty((wsb + 7)//8, dtype=torch.float64, device=dev)
            ptrs = (ctypes.c_void_p*len(names))()
            for k, s in enumerate(names):
                if s in self._tmp:
                    ptrs[k] = self._tmp[s].data_ptr()
                elif s in self.derived:
                    ptrs[k] = self._derived_bufs[s][1].data_ptr()
                else:
                    t = tensors[s]
                    off = y0*t.stride(0)*8 if self.batched[s] else 0
                    ptrs[k] = t.data_ptr() + off
            if timings is None:
                rc = lib.kb200_plan_run(ops, len(ops), _lib.ptr(tables), ptrs, len(names),
                                        _lib.ptr(self._ws) if wsb > 0 else None, wsb,
                                        _lib.stream_ptr())
            else:
                ms = (ctypes.c_float*len(ops))()
                rc = lib.kb200_plan_run_timed(ops, len(ops), _lib.ptr(tables), ptrs, len(names),
                                              _lib.ptr(self._ws) if wsb > 0 else None, wsb,
                                              _lib.stream_ptr(), ms)
                # a launch group reports its time on the leader: share it among the members
                k = 0
                while k < len(ops):
                    gs = int(ops[k].group)
                    if ops[k].kind in (0, 3) and gs > 1:
                        # in proportion to the members' work (flops; equal for elementwise terms)
                        w = [float(ops[j].M)*ops[j].N*ops[j].K for j in range(k, k + gs)]
                        tot = float(ms[k])
                        for j in range(k, k + gs):
                            ms[j] = tot*w[j - k]/sum(w)
                        k += gs
                    else:
                        k += 1
                for k in range(len(ops)):
                    o = ops[k]
                    timings.append((int(o.kind), 2.0*o.M*o.N*o.K*o.batch if o.kind in (0, 2) else 0.0,
                                    float(ms[k])*1e-3, (int(o.M), int(o.N), int(o.K), int(o.batch),
                                                        int(o.tile), int(o.splitk),
                                                        int(o.a_mode), int(o.b_mode),
                                                        int(o.group) if o.kind in (0, 3) else 1)))
            _lib.check(rc, "kb200_plan_run(%s)" % self.name)
            y0 += nb

    def n_ops(self):
        return len(self.low.descs)


_perm_plans = {}


def permute_copy(src, perm):
    """dst = src.permute(perm), materialised with the kb200 permuted-axpby kernel."""
    letters = "abcdefgh"[:src.dim()]
    key = (tuple(src.shape), tuple(perm))
    if key not in _perm_plans:
        dst_letters = "".join(letters[p] for p in perm)
        op = _plan.ROp(("dst", dst_letters), 1.0, [("src", letters)])
        shapes = {"src": tuple(src.shape), "dst": tuple(src.shape[p] for p in perm)}
        _perm_plans[key] = Plan([op], "g", None, ["src"], ["dst"], name="permute",
                                shapes=shapes, batched={"src": False, "dst": False})
    p = _perm_plans[key]
    dst = torch.empty(p.shapes["dst"], dtype=torch.float64, device=src.device)
    p.run({"src": src.contiguous(), "dst": dst}, 1)
    return dst


_cache = {}


def cached(key, builder):
    if key not in _cache:
        _cache[key] = builder()
    return _cache[key]


def clear_cache():
    for p in _cache.values():
        if hasattr(p, "release"):
            p.release()
    _cache.clear()


class PhasedPlan(object):
    """A program evaluated by all ranks TOGETHER at the same grid points (plan.hybrid_phases):
    one Plan per phase, the buffers that cross phases owned here, the distributed buffers laid
    out back to back in one pool so that each exchange is ONE all-reduce over NVLink."""

    def __init__(self, rops, mode, sizes, inputs, outputs, world, name="hybrid", antisym=True,
                 min_work=None):
        self.name, self.world = name, int(world)
        shapes = _plan.slot_shapes(rops, mode, sizes)
        self.outputs = [s for s in outputs if s in shapes]
        self.hp = _plan.hybrid_phases(rops, shapes, self.outputs, self.world, min_work)
        self.shapes = self.hp.shapes
        self.batched = {s: not _plan.is_integral_slot(s) for s in self.shapes}
        ext = set(s for s in self.shapes if s in set(inputs) or _plan.is_integral_slot(s))
        self.inputs = [s for s in self.shapes if s in ext]
        dset = set(self.hp.dslots)
        phase_slots = []
        for ops in self.hp.phases:
            sl = []
            for op in ops:
                for s, _ in [op.out] + list(op.ins):
                    if s not in sl:
                        sl.append(s)
            phase_slots.append(sl)
        count = {}
        for sl in phase_slots:
            for s in sl:
                count[s] = count.get(s, 0) + 1
        # buffers owned here: everything distributed, and every scratch slot seen by two phases
        self.shared = [s for s in self.shapes
                       if s not in ext and s not in self.outputs and (s in dset or count.get(s, 0) > 1)]
        self.plans = []
        have = set(ext) | dset
        for p, (ops, sl) in enumerate(zip(self.hp.phases, phase_slots)):
            if not ops:
                self.plans.append(None)
                continue
            ins = [s for s in sl if s in have]
            outs = [s for s in sl if s not in have and (s in self.outputs or s in self.shared)]
            sh = OrderedDict((s, self.shapes[s]) for s in sl)
            self.plans.append(Plan(ops, mode, None, ins, outs, name="%s-phase%d" % (name, p),
                                   shapes=sh, batched={s: self.batched[s] for s in sl},
                                   antisym=antisym))
            have.update(sl)
        # pool layout: distributed buffers in exchange order
        self._pool_off = {}
        off = 0
        for s in self.hp.dslots:
            n = 1
            for d in self.shapes[s]:
                n *= d
            self._pool_off[s] = (off, n)
            off += n
        self._pool_per_point = off
        self._exch = []
        for ex in self.hp.exchange:
            if ex:
                lo = self._pool_off[ex[0]][0]
                hi = self._pool_off[ex[-1]][0] + self._pool_off[ex[-1]][1]
                self._exch.append((lo, hi))
            else:
                self._exch.append(None)
        self._bufs = None
        self._nb = 0
        self.flops_per_point = sum(p.flops_per_point for p in self.plans if p is not None)

    def _ensure(self, nb, dev):
        if self._bufs is not None and self._nb == nb and self._pool.device == dev:
            return
        # [buffer][grid point]: each exchange range is contiguous over all nb points
        self._pool = torch.zeros(nb*self._pool_per_point, dtype=torch.float64, device=dev)
        self._bufs = {}
        for s, (off, n) in self._pool_off.items():
            self._bufs[s] = self._pool[nb*off:nb*(off + n)].view((nb,) + tuple(self.shapes[s]))
        for s in self.shared:
            if s not in self._bufs:
                alloc = torch.zeros if s.startswith(_plan.TRI_PREFIX) else torch.empty
                self._bufs[s] = alloc((nb,) + tuple(self.shapes[s]), dtype=torch.float64, device=dev)
        self._nb = nb

    def n_ops(self):
        return sum(p.n_ops() for p in self.plans if p is not None)

    def release(self):
        self._bufs = None
        self._pool = None
        self._nb = 0
        for p in self.plans:
            if p is not None:
                p.release()

    def begin(self, tensors, nb, rank):
        """Start an evaluation: tensors = caller-supplied slots (inputs, integrals, outputs) for
        nb grid points; clears the distributed buffers."""
        dev = _lib.device()
        self._ensure(nb, dev)
        self._pool.zero_()
        self._cur = (tensors, nb, (int(rank), self.world))

    def run_phase(self, p):
        tensors, nb, part = self._cur
        pl = self.plans[p]
        if pl is not None:
            t = {}
            for s in pl.inputs + pl.outputs:
                t[s] = self._bufs[s] if s in self._bufs else tensors[s]
            pl.run(t, nb, part=part)

    def exchange_buffer(self, p):
        """The flat buffer that has to be summed over the ranks after phase p, or None."""
        rng = self._exch[p]
        if rng is None:
            return None
        nb = self._cur[1]
        return self._pool[nb*rng[0]:nb*rng[1]]

    def run(self, tensors, nb, rank, group=None):
        """All phases, with the exchanges as NCCL all-reduces on `group`."""
        self.begin(tensors, nb, rank)
        for p in range(len(self.plans)):
            self.run_phase(p)
            buf = self.exchange_buffer(p)
            if buf is not None and self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(buf, group=group)
        self._cur = None

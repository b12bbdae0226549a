"""Execution engine for compiled contraction plans.

A ``Plan`` owns the lowered ``kb200_op`` list, the device copy of the offset
tables and the scratch tensors (intermediates W_oooo, W_vvvv, W_ovvo, ... of
kelvin/lambda_stanton.py:15-34), batched over a chunk of imaginary-time grid
points.  One ``run`` = one C call (kb200_plan_run) that enqueues every kernel
of the residual for that chunk on the current CUDA stream.
"""
import ctypes

import torch

from . import _lib, plan as _plan


class Plan(object):
    def __init__(self, rops, mode, sizes, inputs, outputs, preset_outputs=(), name="plan",
                 shapes=None, batched=None):
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
        self.low = _plan.Lowered(rops, self.shapes, self.batched, preset)
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
        self.flops_per_point = self.low.flops

    # -- device state ------------------------------------------------------
    def _tables(self, dev):
        if self._dev_tables is None or self._dev_tables.device != dev \
                or self._dev_tables.numel() != self.low.tables.size:
            # uint32 payload carried in an int32 tensor (bit pattern preserved)
            self._dev_tables = torch.from_numpy(self.low.tables.view("int32").copy()).to(dev)
        return self._dev_tables

    def tmp_bytes_per_point(self):
        n = 0
        for s in self.tmp_slots:
            m = 1
            for d in self.shapes[s]:
                m *= d
            n += 8*m
        return n

    def _ensure_tmp(self, nb, dev):
        if self._tmp is not None and self._tmp_nb >= nb and self._tmp_dev == dev:
            return
        # triangle blocks (plan.antisym_outputs) must be zero outside the part the plan writes
        self._tmp = {s: (torch.zeros if s.startswith(_plan.TRI_PREFIX) else torch.empty)(
            (nb,) + tuple(self.shapes[s]), dtype=torch.float64, device=dev)
            for s in self.tmp_slots}
        self._tmp_nb = nb
        self._tmp_dev = dev

    def release(self):
        self._tmp = None
        self._tmp_nb = 0
        self._ws = None

    def _ops_for(self, nb):
        if nb not in self._ops:
            arr = self.low.finalize(nb)
            ws = _lib.load().kb200_plan_workspace_bytes(arr, len(arr))
            self._ops[nb] = (arr, int(ws))
        return self._ops[nb]

    # -- run -----------------------------------------------------------------
    def run(self, tensors, ng, chunk=None, timings=None):
        """tensors: slot -> CUDA float64 tensor (batched slots: leading axis ng).
        Scratch is allocated for `chunk` grid points at a time (default: all)."""
        lib = _lib.load()
        dev = _lib.device()
        nb_max = ng if chunk is None else max(1, min(int(chunk), ng))
        self._ensure_tmp(nb_max, dev)
        names = self.low.slot_names
        for s in self.inputs + self.outputs:
            t = tensors[s]
            want = ((ng,) if self.batched[s] else ()) + tuple(self.shapes[s])
            if tuple(t.shape) != want or t.dtype != torch.float64 or not t.is_cuda \
                    or not t.is_contiguous():
                raise Exception("plan %s: slot %s expects contiguous cuda float64 %s, got %s %s"
                                % (self.name, s, want, tuple(t.shape), t.dtype))
        for name, (src, perm) in self.derived.items():
            st = tensors[src]
            key = (st.data_ptr(), st._version)
            have = self._derived_bufs.get(name)
            # the entry keeps its source tensor alive, so an equal (address, version) can only
            # be that same tensor (a freed tensor's address could be reused by another one)
            if have is None or have[0] != key or have[2] is not st:
                self._derived_bufs[name] = (key, permute_copy(st, perm), st)
        y0 = 0
        while y0 < ng:
            nb = min(nb_max, ng - y0)
            ops, wsb = self._ops_for(nb)
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

"""NumPy executor for lowered contraction plans (TEST INFRASTRUCTURE).

Runs the exact ``kb200_op`` descriptors and offset tables the CUDA library
would receive, with NumPy gathers, so that the plan compiler (spin expansion,
merging, reverse mode, table construction) is verified on CPU boxes.  The
product never imports this module.
"""
import numpy


def run_lowered(low, arrays, nbatch, part=None):
    """arrays: slot name -> ndarray; batched slots have a leading tau axis.
    part = (rank, world): row-slabbed contractions do this rank's rows only."""
    ops = low.finalize(nbatch, None, part)
    tabs = low.tables.astype(numpy.int64)
    arrays = dict(arrays)
    for nm, (src, perm) in low.derived.items():
        arrays[nm] = numpy.ascontiguousarray(arrays[src].transpose(perm))
    flat = {}
    padded = {}
    for nm in low.slot_names:
        a = arrays[nm]
        assert a.flags.c_contiguous, nm
        if nm in low.pad:
            # slots the plan stores with padded strides (plan.padded_strides): lay the data out
            # that way in a flat buffer; the pads are poisoned -- nothing may ever read them
            st, size = low.strides_of(nm), low.slot_size(nm)
            lead = a.shape[:a.ndim - len(st)]
            nb_ = int(numpy.prod(lead)) if lead else 1
            buf = numpy.full(nb_*size, numpy.nan)
            view = numpy.lib.stride_tricks.as_strided(
                buf, shape=(nb_,) + tuple(a.shape[a.ndim - len(st):]),
                strides=tuple(8*x for x in [size] + list(st)))
            view[...] = a.reshape(view.shape)
            flat[nm] = buf
            padded[nm] = (view, a)
        else:
            flat[nm] = a.reshape(-1)
    for o in ops:
        A = flat[low.slot_names[o.a]]
        C = flat[low.slot_names[o.c]]
        cm = tabs[o.tCm:o.tCm + o.M]
        cn = tabs[o.tCn:o.tCn + o.N]
        am = tabs[o.tAm:o.tAm + o.M]
        for b in range(o.batch):
            cidx = (o.c_off + b * o.bsC + cm[:, None] + cn[None, :]).reshape(-1)
            if o.kind == 1:
                an = tabs[o.tAk:o.tAk + o.N]
                val = A[(o.a_off + b * o.bsA + am[:, None] + an[None, :]).reshape(-1)]
            elif o.kind == 3:
                # one term of a fused elementwise sum; entries after the first carry beta = 1,
                # so running them one after the other equals the fused pass
                an = tabs[o.tAk:o.tAk + o.N]
                val = A[(o.a_off + b * o.bsA + am[:, None] + an[None, :]).reshape(-1)]
                if o.b >= 0:
                    Y = flat[low.slot_names[o.b]]
                    ym = tabs[o.tBk:o.tBk + o.M]
                    yn = tabs[o.tBn:o.tBn + o.N]
                    val = val * Y[(o.b_off + b * o.bsB + ym[:, None] + yn[None, :]).reshape(-1)]
            else:
                B = flat[low.slot_names[o.b]]
                ak = tabs[o.tAk:o.tAk + o.K]
                bk = tabs[o.tBk:o.tBk + o.K]
                bn = tabs[o.tBn:o.tBn + o.N]
                Am = A[o.a_off + b * o.bsA + am[:, None] + ak[None, :]]
                Bm = B[o.b_off + b * o.bsB + bk[:, None] + bn[None, :]]
                val = (Am @ Bm).reshape(-1)
            if o.beta == 0.0:
                C[cidx] = o.alpha * val
            else:
                C[cidx] = o.beta * C[cidx] + o.alpha * val
    for nm, (view, a) in padded.items():
        if nm not in low.derived:
            a[...] = view.reshape(a.shape)


def run_hybrid(hp, arrays_per_rank, nbatch, batched):
    """Execute a plan.HybridProgram for every simulated rank (arrays_per_rank[r]: slot ->
    ndarray holding the integrals, the inputs and the zero-filled distributed buffers) with the
    exchanges done as sums over the simulated ranks."""
    from kelvin_b200 import plan
    world = len(arrays_per_rank)
    have = set(arrays_per_rank[0])
    for p, rops in enumerate(hp.phases):
        if rops:
            slots = []
            for op in rops:
                for sl, _ in [op.out] + list(op.ins):
                    if sl not in slots:
                        slots.append(sl)
            sh = {s_: hp.shapes[s_] for s_ in slots}
            preset = [s_ for s_ in slots if s_ in have]
            for r in range(world):
                arr = arrays_per_rank[r]
                for s_ in slots:
                    if s_ not in arr:
                        fill = 0.0 if s_.startswith(plan.TRI_PREFIX) else numpy.nan
                        arr[s_] = numpy.full(((nbatch,) if batched(s_) else ()) + tuple(sh[s_]), fill)
                low = plan.Lowered(rops, sh, {s_: batched(s_) for s_ in slots}, preset)
                run_lowered(low, arr, nbatch, part=(r, world))
            have.update(slots)
        for d in hp.exchange[p]:
            tot = sum(arrays_per_rank[r][d] for r in range(world))
            for r in range(world):
                arrays_per_rank[r][d] = tot.copy()

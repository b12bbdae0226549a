"""Worker of tests/test_gpu_multirank.py: one FT-UCCSD calculation through the public API under
whatever torch.distributed world it is launched in; rank 0 writes the numbers to a JSON file."""
import json
import logging
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    from kelvin_b200 import cc_utils, parallel
    from kelvin_b200.ccsd import ccsd
    from kelvin_b200.ueg_system import UEGSystem

    class Traj(logging.Handler):
        def __init__(self):
            super().__init__(level=logging.INFO)
            self.E, self.res = [], []

        def emit(self, record):
            f = record.getMessage().split()
            if len(f) == 3 and f[0].isdigit():
                self.E.append(float(f[1]))
                self.res.append(float(f[2]))
    h = Traj()
    logging.getLogger().setLevel(logging.INFO)
    logging.getLogger().addHandler(h)
    # 19 plane waves, ngrid 6: tau_0 skipped -> 5 evaluated points: 2 per rank + 1 leftover;
    # the m^6 contractions (361^3) are large enough to be dealt out by rows
    T, mu, ng = 0.5, 7.0, int(os.environ.get("KB200_TEST_NG", "6"))
    s = UEGSystem(T, 1.942, 30.0, mu=mu, norb=19, orbtype='u')
    cc = ccsd(s, T=T, mu=mu, ngrid=ng, max_iter=3, damp=0.0)
    om = cc.run()
    traj_E, traj_res = list(h.E), list(h.res)
    cc.max_iter = 4
    cc.compute_ESN()
    sh = parallel.Shards(ng, 1)
    res = {"world": world, "sharded": parallel.active(),
           "hybrid_used": bool(parallel.active() and sh.use_hybrid()),
           "traj_E": traj_E, "traj_res": traj_res, "omega": om[0],
           "E": cc.E, "S": cc.S, "N": cc.N,
           "lam_norm": float(sum(float(x.norm()) for x in cc.L2)),
           "t2_checksum": float(sum(float((x*x).sum()) for x in cc.T2)),
           "n1rdm_trace": float(sum(float(x.diagonal().sum()) for x in cc.n1rdm))}
    agree = True
    if world > 1:
        v = torch.tensor([res["omega"], res["E"], res["S"], res["N"], res["t2_checksum"]],
                         dtype=torch.float64, device="cuda")
        lo, hi = v.clone(), v.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        agree = bool(torch.equal(lo, hi))
        dist.barrier()
    res["ranks_agree"] = agree
    if rank == 0:
        json.dump(res, open(out, "w"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Run under torchrun (one rank per GPU): band-sharded and atmosphere-sharded TUD must equal the single-GPU
results bit for bit.  Used by tests/test_gpu_multi.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radtxfr_b200 import distributed as D, radiative_transfer as rt, synthetic as syn  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables(scale=0.01)
    opts = dict(rt.options)
    rt.set_line_database(tabs, opts=opts)
    kw = dict(Zs=atm["Zs"], Ts=atm["Ts"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_VAL=atm["MFs_VAL"], MFs_ID=atm["MFs_ID"],
              DVOUT=0.01, Altitudes=np.array([3.0, 500.0]))
    X, tau, Lu, Ld = D.compute_TUD_banded(700.0, 1500.0, dict(opts), **kw)
    o1 = dict(opts); o1["voigt"] = {}
    # single-GPU reference on this rank, same pinned tile size
    from radtxfr_b200 import engine
    p, w = rt.layer_weights(opts["LINE_DB"], atm["Ts"], atm["Ps"], atm["PLs"], atm["MFs_VAL"], atm["MFs_ID"])
    grid = engine.to_device(X, opts["LINE_DB"].device)
    od = opts["LINE_DB"].voigt_sum(atm["Ts"], p, w, grid, flags=D.FORCE_K8)
    t1, u1, d1 = engine.tud(od, atm["Ts"][None, :], grid, atm["Zs"], [3.0, 500.0], [1.0], 30)
    ok = (np.array_equal(tau, t1[0, :, :, 0].cpu().numpy()) and np.array_equal(Lu, u1[0, :, :, 0].cpu().numpy())
          and np.array_equal(Ld, d1[0].cpu().numpy()))
    # atmosphere sharding
    Ts, MF = syn.perturbed_atmospheres(5, 5)
    kb = {k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL", "Altitudes")}
    Xb, tb, ub, db_ = D.compute_TUD_batch_sharded(700.0, 900.0, Ts, MF, dict(opts), **kb)
    Xs, ts, us, ds = rt.compute_TUD_batch(700.0, 900.0, Ts, MF, opts=dict(opts), **kb)
    ok2 = torch.equal(tb, ts) and torch.equal(ub, us) and torch.equal(db_, ds)
    flag = torch.tensor([int(ok), int(ok2)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_RESULT world=%d banded_bit_identical=%d atmospheres_bit_identical=%d" % (world, flag[0].item(), flag[1].item()),
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.min().item() == 1 else 1)


if __name__ == "__main__":
    main()

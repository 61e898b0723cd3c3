"""Run under torchrun (one rank per GPU): band-sharded and atmosphere-sharded TUD must equal the single-GPU
results bit for bit.  Used by tests/test_gpu_multi.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_worker.py
With fewer GPUs than ranks every rank uses cuda:0 and the exchange goes over gloo (NCCL refuses two ranks on one
device): the sharded computation -- band plan, pinned tile size, line halo, assembly -- is the same code."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radtxfr_b200 import distributed as D, radiative_transfer as rt, synthetic as syn  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    shared = torch.cuda.device_count() < world_env
    dev = torch.device("cuda", 0 if shared else local)
    torch.cuda.set_device(dev)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables(scale=0.01)
    opts = dict(rt.options)
    rt.set_line_database(tabs, opts=opts)
    kw = dict(Zs=atm["Zs"], Ts=atm["Ts"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_VAL=atm["MFs_VAL"], MFs_ID=atm["MFs_ID"],
              DVOUT=0.01, Altitudes=np.array([3.0, 500.0]))
    ok = True
    for balance in (True, False):
        X, tau, Lu, Ld = D.compute_TUD_banded(700.0, 1500.0, dict(opts), balance=balance, **kw)
        # single-GPU call of the public API on this rank with the same pinned tile size
        o1 = dict(opts); o1["voigt"] = {"tile_k": 8}
        X1, t1, u1, d1 = rt.compute_TUD(700.0, 1500.0, opts=o1, **kw)
        ok = ok and np.array_equal(X, X1) and np.array_equal(tau, t1) and np.array_equal(Lu, u1) and np.array_equal(Ld, d1)
    # one altitude (the k == 1 layout of the shard) on a fine grid where far-field blocks are active
    kw1 = dict(kw, Altitudes=np.array([500.0]), DVOUT=0.001)
    X, tau, Lu, Ld = D.compute_TUD_banded(900.0, 930.0, dict(opts), **kw1)
    o1 = dict(opts); o1["voigt"] = {"tile_k": 8}
    X1, t1, u1, d1 = rt.compute_TUD(900.0, 930.0, opts=o1, **kw1)
    ok = ok and np.array_equal(tau, t1) and np.array_equal(Lu, u1) and np.array_equal(Ld, d1)
    # atmosphere sharding (5 atmospheres; then 1 atmosphere, so that the last ranks own none)
    kb = {k: v for k, v in kw.items() if k not in ("Ts", "MFs_VAL", "Altitudes")}
    ok2 = True
    for n_atm in (5, 1):
        Ts, MF = syn.perturbed_atmospheres(5, n_atm)
        Xb, tb, ub, db_ = D.compute_TUD_batch_sharded(700.0, 900.0, Ts, MF, dict(opts), **kb)
        Xs, ts, us, ds = rt.compute_TUD_batch(700.0, 900.0, Ts, MF, opts=dict(opts), **kb)
        ok2 = ok2 and torch.equal(tb, ts) and torch.equal(ub, us) and torch.equal(db_, ds)
    flag = torch.tensor([int(ok), int(ok2)])
    if not shared:
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_RESULT world=%d backend=%s banded_bit_identical=%d atmospheres_bit_identical=%d"
              % (world, dist.get_backend(), flag[0].item(), flag[1].item()), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.min().item() == 1 else 1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        print("WORKER FAILED rank=%s\n%s" % (os.environ.get("RANK"), traceback.format_exc()), flush=True)
        raise

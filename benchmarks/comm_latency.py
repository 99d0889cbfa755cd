"""Latency of the slab collectives inside libjfnk.so: one-shot peer-memory all-reduce / direct-store halo exchange
(default) against the NCCL path (JFNK_P2P=0).  Launch with torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/comm_latency.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import jfnk_b200 as jf


def main():
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = jf.SlabComm()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    F = jf.SHResidual(N=N, d=0.625 * N, comm=comm, inner_m=1, outer_k=0)
    ctx = F.context()
    x = torch.randn(ctx.n, dtype=torch.float64, device="cuda")
    out = {"n_gpus": comm.size, "N": N, "peer_memory": ctx.peer_memory()}
    for what, kw in (("allreduce", dict(count=1)), ("allreduce", dict(count=32)), ("halo", dict(field=x))):
        ctx.comm_bench(what, reps=20, **kw)
        dist.barrier()
        us = ctx.comm_bench(what, reps=500, **kw)
        t = torch.tensor([us], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[what + ("_%d" % kw["count"] if "count" in kw else "_2rows") + "_us"] = round(float(t.item()), 2)
    if comm.rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

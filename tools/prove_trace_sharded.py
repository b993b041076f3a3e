#!/usr/bin/env python
"""Stage-by-stage wall clock of ONE proof over N GPUs (machine.prove_sharded); rank 0's library prints its stages (NB200_TRACE).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/prove_trace_sharded.py --log-rows 20
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    if rank == 0:
        os.environ["NB200_TRACE"] = "1"
    import torch
    import torch.distributed as dist
    import nexus_zkvm_b200 as nb
    from nexus_zkvm_b200 import machine as M
    from nexus_zkvm_b200.nexus_v1 import NexusV1Machine
    from nexus_zkvm_b200.prover import CudaBackend
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = nb.Context(local)
    ctx.comm_init_from_torch(dist, dev)
    m = NexusV1Machine(a.log_rows)
    cols = m.fill_main_trace(seed=0)
    be = CudaBackend(ctx)
    config = dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3)   # bench.py CONFIG
    pr = be.prover(m.words, config)
    res = [M.shard_host_tree(m, pr, m.preprocessed_columns(), rank, world), M.shard_host_tree(m, pr, list(cols), rank, world)]
    del pr
    for r in range(a.reps):
        dist.barrier(); torch.cuda.synchronize()
        if rank == 0:
            print(f"---- rep {r}", file=sys.stderr, flush=True)
        t = time.perf_counter()
        proof, claimed, aux = M.prove_sharded(m, be, None, None, rank, world, config=config, resident=res)
        if rank == 0:
            print(f"[host] sharded prove total {1e3 * (time.perf_counter() - t):.1f} ms, proof {len(proof)} bytes", file=sys.stderr, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

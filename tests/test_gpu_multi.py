"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the NCCL column-sharded -> all-to-all -> row-sharded
commit of nexus_zkvm_b200.parallel reproduces the single-GPU Merkle root bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = (1 << 31) - 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cols, log_size, q):
    import torch
    import torch.distributed as dist
    import nexus_zkvm_b200 as nb
    from nexus_zkvm_b200 import parallel as par
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = nb.Context(rank, stream=torch.cuda.current_stream().cuda_stream)
        rng = np.random.default_rng(7)
        full = rng.integers(0, P, (n_cols, 1 << log_size), dtype=np.uint32)
        lo, hi = par.column_ranges(n_cols, world)[rank]
        mine = torch.from_numpy(full[lo:hi].view(np.int32).copy()).cuda()
        root, rows, caps = par.sharded_commit(par.CudaEngine(ctx), dist, mine, n_cols, log_size, 1)
        single = None
        if rank == 0:
            ev = ctx.upload(full)
            _, _, tree = ctx.commit_evals([ev], 1)
            single = tree.root
        q.put((rank, root, single))
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def _lib_worker(rank, world, port, n_cols, log_size, small, q):
    """The in-library path: nb200_comm_init + nb200_commit_sharded (NCCL inside libnexus_b200.so, no torch collectives on the data path)."""
    import torch
    import torch.distributed as dist
    import nexus_zkvm_b200 as nb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = nb.Context(rank)      # its own non-blocking stream: nothing here depends on torch's stream
        ctx.comm_init_from_torch(dist, torch.device("cuda", rank))
        rng = np.random.default_rng(7)
        full = rng.integers(0, P, (n_cols, 1 << log_size), dtype=np.uint32)
        smalls = [rng.integers(0, P, (nc, 1 << lg), dtype=np.uint32) for nc, lg in small]
        first, count = nb.Context.shard_range(n_cols, world, rank)
        shard = ctx.upload(full[first:first + count]) if count else None
        rep = [ctx.upload(s) for s in smalls]
        co, rows, sub, caps, root = ctx.commit_sharded(shard, n_cols, log_size, 1, rep)
        ok_coeffs = True
        single = None
        if rank == 0:
            ev = [ctx.upload(full)] + [ctx.upload(s) for s in smalls]
            coeffs1, ldes1, tree = ctx.commit_evals(ev, 1)
            single = tree.root
            ok_coeffs = bool(np.array_equal(co.download(), coeffs1[0].download()[first:first + count]))
            # the row-slice batch holds every column's rows [rank * S, (rank + 1) * S)
            S = (1 << (log_size + 1)) // world
            ok_coeffs = ok_coeffs and bool(np.array_equal(rows.download(), ldes1[0].download()[:, rank * S:(rank + 1) * S]))
        q.put((rank, root, single, ok_coeffs, caps))
        ctx.close()
    finally:
        dist.destroy_process_group()


def _proof_worker(rank, world, port, kind, log_size, q):
    """ONE proof by `world` GPUs (machine.prove_sharded: sharded commits, interaction trace, constraint rows, OODS, DEEP, decommit — every
    collective inside libnexus_b200.so) against the same proof by one GPU."""
    import torch
    import torch.distributed as dist
    import nexus_zkvm_b200 as nb
    from nexus_zkvm_b200 import machine as M
    from nexus_zkvm_b200.prover import CudaBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = nb.Context(rank)
        ctx.comm_init_from_torch(dist, torch.device("cuda", rank))
        if kind == "nexus_v1":
            from nexus_zkvm_b200.nexus_v1 import NexusV1Machine
            m = NexusV1Machine(log_size)
            cols, mult = m.fill_main_trace(seed=log_size), None
        else:
            m = M.AddMachine(log_size=log_size, n_lanes=21)
            cols, mult = m.fill_main_trace(seed=log_size, n_padding=37)
        proof, claimed, aux = M.prove_sharded(m, CudaBackend(ctx), cols, mult, rank, world, associated_data=b"ng")
        single = None
        if rank == 0:
            ctx1 = nb.Context(rank)           # no communicator: the ordinary single-GPU proof
            single, claimed1, aux1 = M.prove(m, CudaBackend(ctx1), cols, mult, associated_data=b"ng")
            assert aux1["roots"] == aux["roots"], "roots differ"
            assert claimed1 == claimed
        q.put((rank, proof, single))
        ctx.sync()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,log_size", [("add", 10), ("nexus_v1", 12), ("nexus_v1", 16)])
def test_one_proof_over_two_gpus_matches_single_gpu_bytes(kind, log_size):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_proof_worker, args=(r, world, port, kind, log_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = [s for _, _, s in res if s is not None][0]
    for rank, proof, _ in res:
        assert proof == single, f"rank {rank}: proof bytes differ (len {len(proof)} vs {len(single)})"


@pytest.mark.parametrize("n_cols,log_size,small", [(50, 12, []), (339, 14, [(1, 8)]), (1012, 13, [(4, 8), (3, 0)])])
def test_library_sharded_commit_matches_single_gpu_root(n_cols, log_size, small):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lib_worker, args=(r, world, port, n_cols, log_size, small, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = [s for _, _, s, _, _ in res if s is not None][0]
    for rank, root, _, ok, caps in res:
        assert root == single, f"rank {rank}"
        assert ok and len(caps) == world


@pytest.mark.parametrize("n_cols,log_size", [(50, 12), (1386, 14)])
def test_sharded_commit_matches_single_gpu_root(n_cols, log_size):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, log_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = [s for _, _, s in res if s is not None][0]
    for rank, root, _ in res:
        assert root == single, f"rank {rank}"

"""N>1 path on CPU: world_size-2 (and 4) `gloo` runs of nexus_zkvm_b200.parallel.sharded_commit with a CPU checker
engine.  Property: the root of the column-sharded -> all-to-all -> row-sharded -> caps commit equals the root of the
single-process commit of the same columns (bit-exact)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nexus_zkvm_b200 import parallel as par
from nexus_zkvm_b200 import build as nb_build
from oracle import pyoracle as orc

P = (1 << 31) - 1


class CpuCheckerEngine:
    """Same protocol as parallel.CudaEngine, computed by the oracle on CPU tensors (tests only)."""

    def lde(self, evals, log_blowup):
        if evals.shape[0] == 0:
            return torch.empty((0, evals.shape[1] << log_blowup), dtype=torch.int32)
        _, lde = orc.interpolate_evaluate_batch(evals.numpy().view(np.uint32), log_blowup)
        return torch.from_numpy(lde.view(np.int32).copy())

    def subtree_root(self, cols):
        return orc.merkle_commit(list(cols.numpy().view(np.uint32)))

    def hash_node(self, left, right):
        import ctypes as C
        import nexus_zkvm_b200 as nb
        out = (C.c_uint8 * 32)()
        st = nb.lib().nb200_hash_node(C.c_int(0), (C.c_uint8 * 32).from_buffer_copy(left), (C.c_uint8 * 32).from_buffer_copy(right), None, C.c_size_t(0), out)
        assert st == 0
        return bytes(out)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cols, log_size, blow, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        full = rng.integers(0, P, (n_cols, 1 << log_size), dtype=np.uint32)
        lo, hi = par.column_ranges(n_cols, world)[rank]
        mine = torch.from_numpy(full[lo:hi].view(np.int32).copy())
        root, rows, caps = par.sharded_commit(CpuCheckerEngine(), dist, mine, n_cols, log_size, blow)
        q.put((rank, root, rows.shape, caps))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cols,log_size", [(2, 37, 6), (2, 16, 5), (4, 50, 7)])
def test_sharded_commit_root_equals_single_process_root(world, n_cols, log_size):
    nb_build.build()
    blow = 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cols, log_size, blow, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    full = rng.integers(0, P, (n_cols, 1 << log_size), dtype=np.uint32)
    _, lde = orc.interpolate_evaluate_batch(full, blow)
    expect = orc.merkle_commit(list(lde))
    for rank, root, shape, caps in res:
        assert root == expect, f"rank {rank}"
        assert shape == (n_cols, (1 << (log_size + blow)) // world)
        assert len(caps) == world


def test_column_ranges_are_block_aligned_and_cover():
    for n_cols, world in [(1386, 8), (347, 8), (27, 4), (5, 2), (16, 8), (1012, 2)]:
        r = par.column_ranges(n_cols, world)
        assert r[0][0] == 0 and r[-1][1] == n_cols
        for (a, b), (c, d) in zip(r, r[1:]):
            assert b == c
        assert all(a % 16 == 0 for a, _ in r if a < n_cols)
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) < 32 or n_cols < 16 * world  # one block of imbalance + the ragged last block

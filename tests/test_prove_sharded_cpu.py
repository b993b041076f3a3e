"""N>1 path of the PROOF on CPU: a world_size-2 `gloo` run of nexus_zkvm_b200.machine.prove_sharded — the driver every rank executes for one
proof over N GPUs (shard ranges, which columns are replicated, the placeholder batches of the replicated components, parameter / claimed-sum
bookkeeping, transcript order) — against a checker backend that gathers the column shards over gloo and hands the assembled columns to the oracle.
Property: every rank returns exactly the bytes of the single-process proof.  (The device side of the same sequence — NCCL / NVLink inside
libnexus_b200.so — is tests/test_gpu_multi.py.)"""
import os
import socket

import numpy as np
import pytest

from nexus_zkvm_b200 import machine as M
from tests.oracle_backend import OracleBackend, _OracleProver
from oracle import pyoracle as orc


class _Batch:
    """stands for a device batch: host columns in finalized (bit-reversed) order"""
    def __init__(self, arr, n_cols=None, log_size=None):
        self.arr = arr
        self.n_cols = arr.shape[0] if arr is not None else n_cols
        self.log_size = (arr.shape[1].bit_length() - 1) if arr is not None else log_size
        self.device_ptr = 0


class _Ctx:
    def upload(self, host2d, coset_order=False):
        a = np.ascontiguousarray(host2d, dtype=np.uint32)
        return _Batch(np.stack([orc.finalize_column(c) for c in a]) if coset_order else a)

    def wrap_device(self, ptr, n_cols, log_size):
        return _Batch(None, n_cols, log_size)       # the placeholder a replicated component never reads


class _ShardedChecker(_OracleProver):
    """commit_sharded / gen_interaction_sharded with the semantics of the CUDA backend, computed by the oracle on the gathered columns"""
    def __init__(self, words, config, dist, rank, world):
        super().__init__(words, config)
        self.ctx, self.dist, self.rank, self.world = _Ctx(), dist, rank, world
        self.seen_replicate = []

    def _batches_from_host(self, cols, coset_order):
        out, i = [], 0
        while i < len(cols):
            j = i
            while j < len(cols) and len(cols[j]) == len(cols[i]):
                j += 1
            out.append(self.ctx.upload(np.stack(cols[i:j]), coset_order))
            i = j
        return out

    def commit_sharded(self, big_shard, total_big, log_size, small, replicate, keep_eval_rows, ch):
        from nexus_zkvm_b200 import Context
        first, count = Context.shard_range(total_big, self.world, self.rank)
        assert (big_shard.n_cols if big_shard is not None else 0) == count, "a rank must hand over exactly its shard_range"
        parts = [None] * self.world
        self.dist.all_gather_object(parts, big_shard.arr if big_shard is not None else None)
        big = [c for p in parts if p is not None for c in p]
        assert len(big) == total_big and all(len(c) == 1 << log_size for c in big)
        assert all(0 <= g < total_big for g in replicate)
        self.seen_replicate.append(sorted(replicate))
        cols = big + [c for b in small for c in b.arr]
        return self.p.commit(cols, ch, self.config["log_blowup"])

    def gen_interaction_sharded(self, comp, params):
        from nexus_zkvm_b200 import Context
        n_cols = self.n_logup[comp]
        cols, cs = self.p.gen_interaction(comp, self.log_sizes[comp], n_cols, np.array(params, dtype=np.uint32))
        first, count = Context.shard_range(4 * n_cols, self.world, self.rank)
        return _Batch(cols[first:first + count]) if count else _Batch(None, 0, self.log_sizes[comp]), cs

    def gen_interaction_replicated(self, comp, log_size, n_logup_cols, params, tree0, tree1):
        assert tree0[0].arr is None and tree1[0].arr is None      # the placeholders stand first, the real small batches follow
        cols, cs = self.p.gen_interaction(comp, log_size, n_logup_cols, np.array(params, dtype=np.uint32))
        return _Batch(cols), cs


class _ShardedBackend(OracleBackend):
    def __init__(self, machine, dist, rank, world):
        self.ctx, self.machine, self.dist, self.rank, self.world = _Ctx(), machine, dist, rank, world
        self.last = None

    def prover(self, words, config):
        p = _ShardedChecker(words, config, self.dist, self.rank, self.world)
        p.n_logup = [max(c.batching) + 1 for c in self.machine.air.components]
        p.log_sizes = [c.log_size for c in self.machine.air.components]
        self.last = p
        return p


def _machine(kind):
    if kind == "nexus_v1":
        from nexus_zkvm_b200.nexus_v1 import NexusV1Machine
        m = NexusV1Machine(8)
        return m, m.fill_main_trace(seed=4), None
    m = M.AddMachine(log_size=8, n_lanes=2)
    cols, mult = m.fill_main_trace(seed=4, n_padding=9)
    return m, cols, mult


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, kind, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, cols, mult = _machine(kind)
        be = _ShardedBackend(m, dist, rank, world)
        proof, claimed, aux = M.prove_sharded(m, be, cols, mult, rank, world, associated_data=b"cpu")
        q.put((rank, proof, claimed, aux["roots"], be.last.seen_replicate))
    except BaseException:   # noqa: BLE001 - reported to the parent
        import traceback
        q.put((rank, "worker-error", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world", [("add", 2), ("nexus_v1", 2), ("nexus_v1", 4)])
def test_prove_sharded_driver_reproduces_the_single_process_proof(kind, world):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] != "worker-error", r[2]
    m, cols, mult = _machine(kind)
    o_proof, o_claimed, o_aux = M.prove(m, OracleBackend(), cols, mult, associated_data=b"cpu")
    main = m.air.components[0]
    want_rep = [sorted({c for (t, c, off) in main.masks if t == tree and off != 0}) for tree in (0, 1, 2)]
    for rank, proof, claimed, roots, seen_rep in res:
        assert roots == o_aux["roots"], f"rank {rank}"
        assert claimed == o_claimed
        assert proof == o_proof, f"rank {rank}: proof bytes differ"
        assert seen_rep == want_rep          # exactly the columns the AIR reads at a row offset are asked to be replicated
    assert want_rep[2], "the LogUp constraint reads the last interaction column at the previous row"

"""Multi-GPU sharding of ONE commitment over the ranks of a `torch.distributed` group.

The commit of a tree (TreeBuilder::extend_evals + commit, /root/reference prover/src/machine.rs:208-263) shards at two
granularities with one exchange between them:

  1. column-sharded  — iFFT / LDE of a column are independent: rank r transforms the contiguous column range
     `column_ranges(n_cols, world)[r]` (multiples of 16 columns = one 64-byte Blake2s block);
  2. exchange        — one all-to-all re-shards the LDE from column-major to row-slice-major: rank j receives, from
     every rank, rows [j*rows/world, (j+1)*rows/world) of that rank's columns (in bit-reversed order a contiguous
     1/world slice of every column is exactly one depth-log2(world) sub-tree of the Merkle tree);
  3. row-sharded     — rank j hashes its sub-tree (all columns, its rows) and the 32-byte sub-tree roots ("caps") are
     all-gathered; every rank finishes the top log2(world) levels on the host with `nb200_hash_node`.

The resulting root is bit-identical to the single-GPU root.  The module is written against a tiny engine protocol
(`lde`, `subtree_root`, `hash_node`) and plain torch tensors, so the same code runs over NCCL with the CUDA engine and,
in the CPU test-suite, over gloo with a CPU checker engine.
"""
import ctypes as C

import numpy as np


def column_ranges(n_cols, world, align=16):
    """Contiguous, block-aligned column ranges, as even as possible; the last ranks may be empty for tiny trees."""
    blocks = (n_cols + align - 1) // align
    out, start = [], 0
    for r in range(world):
        nb = blocks // world + (1 if r < blocks % world else 0)
        end = min(n_cols, start + nb * align)
        out.append((start, end))
        start = end
    return out


def cap_root(caps, hash_node):
    """Finish the top levels from the world sub-tree roots (no column values live above the cap layer here)."""
    level = list(caps)
    while len(level) > 1:
        level = [hash_node(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    return level[0]


class CudaEngine:
    """Device engine: LDE and sub-tree hashing through the C ABI on torch CUDA tensors (int32 storage of u32 words)."""

    def __init__(self, ctx, merkle_hash=0):
        self.ctx, self.merkle_hash = ctx, merkle_hash

    def _fence(self):
        """The ctx may run on its own non-blocking stream, which has no implicit ordering with torch's current stream: drain both sides around
        every ABI call that touches torch tensors (this path is the older torch-collective variant; nb200_commit_sharded needs none of this)."""
        import torch
        torch.cuda.current_stream().synchronize()
        self.ctx.sync()

    def lde(self, evals, log_blowup):
        """evals: (n_cols, 2^log) int32 device tensor -> (n_cols, 2^(log+blowup)) int32 device tensor."""
        import torch
        from . import lib
        n_cols, n = evals.shape
        log = n.bit_length() - 1
        out = torch.empty((n_cols, n << log_blowup), dtype=torch.int32, device=evals.device)
        if n_cols == 0:
            return out
        co = torch.empty_like(evals)
        ev = self.ctx.wrap_device(evals.data_ptr(), n_cols, log)
        cob = self.ctx.wrap_device(co.data_ptr(), n_cols, log)
        ldb = self.ctx.wrap_device(out.data_ptr(), n_cols, log + log_blowup)
        co.copy_(evals)
        self._fence()
        self.ctx.interpolate(cob)
        self.ctx._chk(lib().nb200_evaluate(self.ctx._h, cob._h, C.c_uint32(log_blowup), ldb._h))
        self.ctx.sync()
        del ev
        return out

    def subtree_root(self, cols):
        """cols: (n_cols, rows) int32 device tensor -> 32-byte root of the Merkle tree over these rows."""
        n_cols, rows = cols.shape
        self._fence()   # `cols` comes out of a torch collective on torch's stream
        b = self.ctx.wrap_device(cols.data_ptr(), n_cols, rows.bit_length() - 1)
        return self.ctx.merkle_commit([b]).root

    def hash_node(self, left, right):
        from . import lib
        out = (C.c_uint8 * 32)()
        st = lib().nb200_hash_node(C.c_int(self.merkle_hash), (C.c_uint8 * 32).from_buffer_copy(left), (C.c_uint8 * 32).from_buffer_copy(right),
                                   None, C.c_size_t(0), out)
        assert st == 0
        return bytes(out)


def sharded_commit(engine, dist, my_evals, n_cols_total, log_size, log_blowup, group=None):
    """Commit one tree of `n_cols_total` columns of 2^log_size rows spread over the ranks of `dist`.

    my_evals: this rank's column range (column_ranges(..)[rank]) as an (n_my, 2^log_size) int32 tensor on the engine's
    device.  Returns (root bytes, my LDE row-slice (n_cols_total, rows/world) tensor, caps list).
    """
    import torch
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert world & (world - 1) == 0, "world size must be a power of two"
    ranges = column_ranges(n_cols_total, world)
    assert my_evals.shape[0] == ranges[rank][1] - ranges[rank][0]
    rows = 1 << (log_size + log_blowup)
    slice_rows = rows // world
    assert slice_rows >= 1
    lde = engine.lde(my_evals, log_blowup)                              # (n_my, rows), column-sharded phase
    # all-to-all: send rows-slice j of my columns to rank j
    send = lde.reshape(lde.shape[0], world, slice_rows).permute(1, 0, 2).contiguous()   # (world, n_my, slice_rows)
    recv_counts = [(ranges[j][1] - ranges[j][0]) * slice_rows for j in range(world)]
    send_counts = [lde.shape[0] * slice_rows] * world
    recv = torch.empty(sum(recv_counts), dtype=lde.dtype, device=lde.device)
    dist.all_to_all_single(recv, send.reshape(-1), output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    mine = recv.reshape(n_cols_total, slice_rows)                       # rows are already grouped by source rank = column order
    cap = engine.subtree_root(mine)                                     # row-sharded phase
    cap_t = torch.frombuffer(bytearray(cap), dtype=torch.uint8).to(lde.device)
    gathered = [torch.empty_like(cap_t) for _ in range(world)]
    dist.all_gather(gathered, cap_t, group=group)
    caps = [bytes(g.cpu().numpy().tobytes()) for g in gathered]
    return cap_root(caps, engine.hash_node), mine, caps

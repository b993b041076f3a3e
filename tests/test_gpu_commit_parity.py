"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every call goes through the C ABI
(libnexus_b200.so via ctypes); the oracle is only the checker.  Bar: bit-exact."""
import numpy as np
import pytest

import nexus_zkvm_b200 as nb
from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu
P = (1 << 31) - 1


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


def _rand_cols(rng, n_cols, log):
    return rng.integers(0, P, (n_cols, 1 << log), dtype=np.uint32)


@pytest.mark.parametrize("log", [1, 2, 3, 5, 8, 12, 16])
def test_twiddle_bank_matches_oracle(log):
    c = nb.Context(0)
    tw, itw = c.twiddles(log)
    otw, oitw = orc.twiddles(log)
    assert np.array_equal(tw, otw)
    assert np.array_equal(itw, oitw)
    c.close()


@pytest.mark.parametrize("log", list(range(0, 19)))
def test_interpolate_bit_exact(ctx, log):
    rng = np.random.default_rng(100 + log)
    n_cols = 5 if log <= 14 else 3
    v = _rand_cols(rng, n_cols, log)
    # structured edge cases the reference's traces contain: zeros, all P-1, impulse, byte-valued
    v[0, :] = 0
    if n_cols > 1:
        v[1, :] = P - 1
    if n_cols > 2:
        v[2, :] = 0; v[2, 0] = 1
    if n_cols > 3:
        v[3, :] = rng.integers(0, 256, 1 << log)
    cols = ctx.upload(v)
    ctx.interpolate(cols)
    got = cols.download()
    for c in range(n_cols):
        exp = orc.interpolate(v[c]) if log >= 1 else v[c]
        assert np.array_equal(got[c], exp), f"log={log} col={c}"


@pytest.mark.parametrize("log,blow", [(0, 1), (1, 1), (3, 1), (4, 2), (7, 1), (8, 1), (9, 1), (11, 2), (12, 1), (13, 1), (15, 1), (16, 2), (18, 1)])
def test_evaluate_lde_bit_exact(ctx, log, blow):
    rng = np.random.default_rng(200 + log)
    n_cols = 3
    coeffs = _rand_cols(rng, n_cols, log)
    cols = ctx.upload(coeffs)
    lde = ctx.evaluate(cols, blow)
    got = lde.download()
    for c in range(n_cols):
        assert np.array_equal(got[c], orc.evaluate(coeffs[c], log + blow)), f"log={log} col={c}"


@pytest.mark.parametrize("log", [20, 22])
def test_large_roundtrip_and_sampled_parity(ctx, log):
    # full-size property: evaluate(interpolate(v)) == v, plus oracle parity on one column
    rng = np.random.default_rng(300 + log)
    n_cols = 6 if log == 20 else 2
    v = _rand_cols(rng, n_cols, log)
    cols = ctx.upload(v)
    ctx.interpolate(cols)
    coeffs = cols.download()
    assert np.array_equal(coeffs[0], orc.interpolate(v[0]))
    back = ctx.evaluate(cols, 0).download()
    assert np.array_equal(back, v)
    # linearity of the transform: I(a) + I(b) == I(a+b)
    s = ((v[0].astype(np.uint64) + v[1]) % P).astype(np.uint32)
    cs = ctx.upload(s[None, :]); ctx.interpolate(cs)
    assert np.array_equal(cs.download()[0], ((coeffs[0].astype(np.uint64) + coeffs[1]) % P).astype(np.uint32))


def test_finalize_order_matches_reference_permutation(ctx):
    rng = np.random.default_rng(7)
    for log in (3, 4, 10, 13):
        v = _rand_cols(rng, 3, log)
        up = ctx.upload(v, coset_order=True).download()
        for c in range(3):
            assert np.array_equal(up[c], orc.finalize_column(v[c]))
        cols = ctx.upload(v)
        cols.finalize_order()
        assert np.array_equal(cols.download(), up)


@pytest.mark.parametrize("variant", [0, 1])
def test_merkle_mixed_sizes_bit_exact(variant):
    c = nb.Context(0)
    c.set_flavor(merkle_hash=variant)
    orc.set_flavor(merkle_hash=variant)
    try:
        rng = np.random.default_rng(11 + variant)
        shapes = [(7, 10), (3, 4), (20, 10), (1, 0), (5, 6), (16, 10), (33, 8)]
        host = [_rand_cols(rng, n, lg) for n, lg in shapes]
        batches = [c.upload(h) for h in host]
        tree = c.merkle_commit(batches)
        flat = [col for h in host for col in h]
        root, layers = orc.merkle_commit(flat, want_layers=True)
        assert tree.root == root
        off = 0
        for lg in range(0, 11):
            assert np.array_equal(tree.layer(lg), layers[off:off + (32 << lg)]), f"layer {lg}"
            off += 32 << lg
        queries = {10: [0, 5, 6, 700, 1023], 8: [1, 200], 4: [3], 0: [0]}
        qv, hw, cw = tree.decommit(queries)
        oqv, ohw, ocw = orc.merkle_decommit(flat, queries)
        assert np.array_equal(qv, oqv) and hw == ohw and np.array_equal(cw, ocw)
    finally:
        orc.set_flavor()
        c.close()


def test_merkle_empty_and_single(ctx):
    tree = ctx.merkle_commit([])
    assert tree.root == orc.merkle_commit([])
    v = np.array([[5]], dtype=np.uint32)
    assert ctx.merkle_commit([ctx.upload(v)]).root == orc.merkle_commit([v[0]])


@pytest.mark.parametrize("log", [0, 1, 4, 9, 14, 17])
def test_eval_at_points_bit_exact(ctx, log):
    rng = np.random.default_rng(400 + log)
    coeffs = _rand_cols(rng, 4, log)
    pts = rng.integers(0, P, (3, 2, 4), dtype=np.uint32)
    got = ctx.eval_at_points(ctx.upload(coeffs), pts)
    for c in range(4):
        for p in range(3):
            assert np.array_equal(got[c, p], orc.eval_at_point(coeffs[c], pts[p, 0], pts[p, 1])), (c, p)


@pytest.mark.parametrize("logs", [[10], [12, 5, 12, 4], [14, 8]])
def test_commit_evals_matches_oracle_pipeline(ctx, logs):
    rng = np.random.default_rng(sum(logs))
    host = [_rand_cols(rng, 19 if lg > 6 else 3, lg) for lg in logs]
    batches = [ctx.upload(h) for h in host]
    coeffs, ldes, tree = ctx.commit_evals(batches, 1)
    flat = []
    for h, lde, b, cf in zip(host, ldes, batches, coeffs):
        co, ol = orc.interpolate_evaluate_batch(h, 1, want_coeffs=True)
        assert np.array_equal(b.download(), h)  # evaluations are left untouched
        assert np.array_equal(cf.download(), co)
        assert np.array_equal(lde.download(), ol)
        flat += list(ol)
    assert tree.root == orc.merkle_commit(flat)


@pytest.mark.parametrize("coset_order", [False, True])
def test_commit_host_pipelined_matches_oracle(ctx, coset_order):
    # host columns in pinned memory -> chunked H2D overlapped with the transforms (2 chunks at this size) -> Merkle
    rng = np.random.default_rng(99)
    log, n_cols = 18, 300
    host = ctx.host_alloc(n_cols, log)
    host[:] = rng.integers(0, P, (n_cols, 1 << log), dtype=np.uint32)
    small = rng.integers(0, P, (3, 1 << 6), dtype=np.uint32)
    evals, coeffs, ldes, tree = ctx.commit_host([host, small], 1, coset_order=coset_order)
    ref = [np.stack([orc.finalize_column(c) for c in b]) if coset_order else b for b in (host, small)]
    flat = []
    for r, ev, lde in zip(ref, evals, ldes):
        assert np.array_equal(ev.download(), r)
        _, ol = orc.interpolate_evaluate_batch(r, 1)
        assert np.array_equal(lde.download(), ol)
        flat += list(ol)
    assert tree.root == orc.merkle_commit(flat)

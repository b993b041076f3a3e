"""GPU parity of the backend-trait level entry points of the C ABI (FriOps, QuotientOps, AccumulationOps, GrindOps,
ComponentProver::evaluate_constraint_quotients_on_domain) against the oracle's restatement of the same functions.
Bit-exact; inputs are seeded."""
import numpy as np
import pytest

import nexus_zkvm_b200 as nb
from nexus_zkvm_b200 import machine as M
from nexus_zkvm_b200 import prover as NP
from nexus_zkvm_b200.prover import CudaBackend
from oracle import pyoracle as orc
from tests.oracle_backend import OracleBackend

pytestmark = pytest.mark.gpu
P = (1 << 31) - 1


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


def component_headers(words):
    """(log_size, log_expand, n_constraints) per component, parsed from the AIR bytecode (layout: nexus_zkvm_b200/air.py)."""
    w = [int(x) for x in words]
    out, i = [], 4
    for _ in range(w[3]):
        out.append((w[i], w[i + 1], w[i + 2]))
        i += 3
        i += 1 + 3 * w[i]            # masks
        i += 2                       # register counts
        i += 1 + 4 * w[i]            # constraint program
        n_fracs = w[i]; i += 3       # n_fracs + logup register counts
        i += 1 + 4 * w[i]            # logup program
        i += n_fracs + 2
    assert i == len(w)
    return out


def rnd(rng, *shape):
    return rng.integers(0, P, size=shape, dtype=np.uint32)


@pytest.mark.parametrize("k", [1, 2, 5, 12, 16])
def test_fold_line(ctx, k):
    rng = np.random.default_rng(100 + k)
    src = rnd(rng, 4, 1 << k)
    alpha = rnd(rng, 4)
    d = NP.fold_line(ctx, ctx.upload(src), alpha)
    assert d.n_cols == 4 and d.log_size == k - 1
    assert np.array_equal(d.download(), orc.fold_line(src, alpha))


@pytest.mark.parametrize("k", [3, 4, 9, 15])
def test_fold_circle_into_line(ctx, k):
    rng = np.random.default_rng(200 + k)
    src, dst, alpha = rnd(rng, 4, 1 << k), rnd(rng, 4, 1 << (k - 1)), rnd(rng, 4)
    g = ctx.upload(dst)
    NP.fold_circle_into_line(ctx, g, ctx.upload(src), alpha)
    assert np.array_equal(g.download(), orc.fold_circle_into_line(dst, src, alpha))


def test_fold_shape_errors(ctx):
    rng = np.random.default_rng(1)
    with pytest.raises(nb.Nb200Error):
        NP.fold_line(ctx, ctx.upload(rnd(rng, 3, 16)), rnd(rng, 4))           # not a secure column
    with pytest.raises(nb.Nb200Error):
        NP.fold_circle_into_line(ctx, ctx.upload(rnd(rng, 4, 16)), ctx.upload(rnd(rng, 4, 16)), rnd(rng, 4))  # dst not half the size


def test_accumulate(ctx):
    rng = np.random.default_rng(3)
    a, b = rnd(rng, 4, 1 << 11), rnd(rng, 4, 1 << 11)
    a[0, :4] = P - 1; b[0, :4] = [0, 1, P - 1, 2]          # wrap-around cases
    g = ctx.upload(a)
    NP.accumulate(ctx, g, ctx.upload(b))
    want = ((a.astype(np.uint64) + b.astype(np.uint64)) % P).astype(np.uint32)
    assert np.array_equal(g.download(), want)
    with pytest.raises(nb.Nb200Error):
        NP.accumulate(ctx, g, ctx.upload(rnd(rng, 4, 1 << 10)))


@pytest.mark.parametrize("pow_bits", [0, 1, 5, 12, 18])
def test_grind(ctx, pow_bits):
    rng = np.random.default_rng(400 + pow_bits)
    for _ in range(3):
        digest = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
        assert NP.grind(ctx, digest, pow_bits) == orc.grind(digest, pow_bits)


@pytest.mark.parametrize("log_size,splits", [(4, [3]), (10, [2, 5]), (14, [1, 1, 6])])
def test_fri_quotients(ctx, log_size, splits):
    rng = np.random.default_rng(500 + log_size)
    blocks = [rnd(rng, n, 1 << log_size) for n in splits]
    cols = [c for b in blocks for c in b]
    n = len(cols)
    # three sample batches: every column at the OODS point, two columns also at a shifted point, one column alone
    pts = [rnd(rng, 8) for _ in range(3)]
    batches = [(pts[0], [(i, rnd(rng, 4)) for i in range(n)]),
               (pts[1], [(0, rnd(rng, 4)), (n - 1, rnd(rng, 4))]),
               (pts[2], [(n // 2, rnd(rng, 4))])]
    rc = rnd(rng, 4)
    g = NP.fri_quotients(ctx, [ctx.upload(b) for b in blocks], log_size, batches, rc)
    assert g.n_cols == 4 and g.log_size == log_size
    assert np.array_equal(g.download(), orc.accumulate_quotients(cols, rc, batches))


def test_fri_quotients_rejects_bad_column_index(ctx):
    rng = np.random.default_rng(9)
    with pytest.raises(nb.Nb200Error):
        NP.fri_quotients(ctx, [ctx.upload(rnd(rng, 2, 16))], 4, [(rnd(rng, 8), [(2, rnd(rng, 4))])], rnd(rng, 4))


@pytest.mark.parametrize("log_size,lanes,pairs", [(8, 1, False), (9, 2, True)])
def test_constraint_quotients_per_component(ctx, log_size, lanes, pairs):
    """evaluate_constraint_quotients_on_domain, one component at a time, with arbitrary coefficients and a non-zero
    accumulator going in."""
    m = M.AddMachine(log_size=log_size, n_lanes=lanes, logup_in_pairs=pairs)
    cols, mult = m.fill_main_trace(seed=5 + log_size, n_padding=2)
    cfg = dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3)
    provers, params = [], None
    for be in (CudaBackend(ctx), OracleBackend()):
        ch = be.channel()
        p = be.prover(m.words, cfg)
        p.commit(m.preprocessed_columns(), ch, coset_order=True)
        p.commit(list(cols) + [mult], ch, coset_order=True)
        prm = [(0, 0, 0, 0)] * m.air.n_params
        m.range256.draw(ch, prm)
        inter = []
        for k, comp in enumerate(m.air.components):
            c, cs = p.gen_interaction(k, comp.log_size, max(comp.batching) + 1, prm)
            inter.append(c)
            prm[comp.cumsum_shift_param] = M.F.qm31_mul_m31(cs, M.F.m31_inv((1 << comp.log_size) % P))
        p.commit_interaction(inter, ch)
        params = params or prm
        assert prm == params
        provers.append(p)
    gpu, oracle = provers
    rng = np.random.default_rng(77)
    for k, (lg, expand, n_constraints) in enumerate(component_headers(m.words)):
        elog = lg + expand
        coeffs = rnd(rng, n_constraints, 4)
        acc0 = rnd(rng, 4, 1 << elog)
        g_acc = ctx.upload(acc0)
        gpu.constraint_quotients(k, params, coeffs, g_acc)
        want = oracle.p.constraint_quotients(k, elog, np.array(params, dtype=np.uint32), coeffs, acc0)
        assert np.array_equal(g_acc.download(), want), f"component {k}"
        # wrong coefficient count is an argument error, not a crash
        with pytest.raises(nb.Nb200Error):
            gpu.constraint_quotients(k, params, coeffs[:-1], g_acc)

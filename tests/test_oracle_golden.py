"""Pins the oracle against every external anchor available for this path:
RFC 7693 Blake2s vectors (+ hashlib as an independent implementation), the field definitions of the spec
PDF §3.1, the circle-group constants, and the reference's in-repo ordering identity
(/root/reference prover/src/trace/utils.rs:110-128, utils_external.rs:24-39).
Everything Stwo-internal beyond that is PARITY UNPINNED and only checked for mathematical self-consistency.
"""
import hashlib

import numpy as np
import pytest

from oracle import pyoracle as orc

P = (1 << 31) - 1
rng = np.random.default_rng(0xB200)


def test_blake2s_rfc7693_abc():
    # RFC 7693 Appendix B
    exp = bytes.fromhex("508C5E8C327C14E2E1A72BA34EEB452F37458B209ED63A294D999B4C86675982")
    assert orc.blake2s(b"abc") == exp


@pytest.mark.parametrize("n", [0, 1, 31, 32, 63, 64, 65, 127, 128, 129, 1000, 5544])
def test_blake2s_vs_hashlib(n):
    data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
    assert orc.blake2s(data) == hashlib.blake2s(data, digest_size=32).digest()


def test_blake2s_compress_is_rfc_F():
    # one-block message: Blake2s(m) == F(IV ^ param, m, t=len, f0=~0)
    iv = np.array([0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19], dtype=np.uint32)
    h0 = iv.copy(); h0[0] ^= 0x01010020
    for ln in (1, 17, 64):
        msg = bytes(rng.integers(0, 256, ln, dtype=np.uint8))
        m = np.frombuffer(msg.ljust(64, b"\0"), dtype="<u4")
        h = orc.blake2s_compress(h0, m, t0=ln, f0=0xFFFFFFFF)
        assert h.astype("<u4").tobytes() == hashlib.blake2s(msg, digest_size=32).digest()


def test_m31_field():
    for a, b in [(0, 0), (1, P - 1), (P - 1, P - 1), (2, 1 << 30), (123456789, 987654321)]:
        assert orc.m31_mul(a, b) == (a * b) % P
    for a in [1, 2, P - 1, 1268011823, 12345]:
        assert orc.m31_mul(a, orc.m31_inv(a)) == 1


def _qm31_mul_py(x, y):
    # spec PDF §3.1: (a+bi) + (c+di)u with i^2=-1, u^2=2+i
    def cmul(p, q): return ((p[0] * q[0] - p[1] * q[1]) % P, (p[0] * q[1] + p[1] * q[0]) % P)
    def cadd(p, q): return ((p[0] + q[0]) % P, (p[1] + q[1]) % P)
    a, b, c, d = (x[0], x[1]), (x[2], x[3]), (y[0], y[1]), (y[2], y[3])
    r = (2, 1)
    lo = cadd(cmul(a, c), cmul(r, cmul(b, d)))
    hi = cadd(cmul(a, d), cmul(b, c))
    return [lo[0], lo[1], hi[0], hi[1]]


def test_qm31_field():
    for _ in range(50):
        x = [int(v) for v in rng.integers(0, P, 4)]
        y = [int(v) for v in rng.integers(0, P, 4)]
        assert list(orc.qm31_mul(x, y)) == _qm31_mul_py(x, y)
        assert list(orc.qm31_mul(x, orc.qm31_inv(x))) == [1, 0, 0, 0]
    # u^2 == 2 + i
    assert list(orc.qm31_mul([0, 0, 1, 0], [0, 0, 1, 0])) == [2, 1, 0, 0]


def test_circle_generator_order():
    # G = (2, 1268011823) on x^2+y^2=1, order 2^31
    x, y = 2, 1268011823
    assert (x * x + y * y) % P == 1
    for k in range(30):
        x, y = (2 * x * x - 1) % P, (2 * x * y) % P
    assert (x, y) == (P - 1, 0)  # 2^30 * G = (-1, 0)


def test_reference_ordering_identity():
    # /root/reference prover/src/trace/utils.rs:110-128 (test_order), for several sizes
    for log_size in (3, 4, 6, 9):
        n = 1 << log_size
        vals = np.arange(n, dtype=np.uint32)
        col = orc.finalize_column(vals)
        for i in range(n):
            idx = orc.bit_reverse_index(orc.coset_index_to_circle_domain_index(i, log_size), log_size)
            assert col[i] == vals[idx]


def test_twiddles_are_domain_points():
    # line twiddles of layer 1 are the x coordinates, circle twiddles the y coordinates, of the
    # bit-reversed canonic circle domain — ties the buffer recipe to the geometry.
    for log_size in (3, 5, 8):
        tw, itw = orc.twiddles(log_size)
        n = 1 << log_size
        assert all(orc.m31_mul(int(a), int(b)) == 1 for a, b in zip(tw, itw))
        first_line = tw[: n // 4]
        for h in range(n // 4):
            x, _ = orc.circle_domain_at(log_size, orc.bit_reverse_index(4 * h, log_size))
            assert int(first_line[h]) == x


@pytest.mark.parametrize("log_size", [1, 2, 3, 4, 7, 10])
def test_interpolate_matches_pointwise_evaluation(log_size):
    n = 1 << log_size
    v = rng.integers(0, P, n, dtype=np.uint32)
    coeffs = orc.interpolate(v)
    # f(domain.at(bitrev(i))) == v[i]
    for i in rng.integers(0, n, min(n, 16)):
        x, y = orc.circle_domain_at(log_size, orc.bit_reverse_index(int(i), log_size))
        got = orc.eval_at_point(coeffs, [x, 0, 0, 0], [y, 0, 0, 0])
        assert list(got) == [int(v[i]), 0, 0, 0]
    assert np.array_equal(orc.evaluate(coeffs, log_size), v)


@pytest.mark.parametrize("log_size,blow", [(3, 1), (5, 1), (6, 2), (9, 1)])
def test_lde_agrees_with_eval_at_point(log_size, blow):
    n = 1 << log_size
    v = rng.integers(0, P, n, dtype=np.uint32)
    coeffs = orc.interpolate(v)
    lde = orc.evaluate(coeffs, log_size + blow)
    for i in rng.integers(0, n << blow, 12):
        x, y = orc.circle_domain_at(log_size + blow, orc.bit_reverse_index(int(i), log_size + blow))
        assert list(orc.eval_at_point(coeffs, [x, 0, 0, 0], [y, 0, 0, 0])) == [int(lde[i]), 0, 0, 0]
    _, lde2 = orc.interpolate_evaluate_batch(v[None, :], blow)
    assert np.array_equal(lde2[0], lde)


def _hash_node_py(left, right, vals, variant):
    if variant == 1:
        h = hashlib.blake2s(digest_size=32)
        if left is not None:
            h.update(left); h.update(right)
        h.update(np.asarray(vals, dtype="<u4").tobytes())
        return h.digest()
    st = np.zeros(8, np.uint32)
    if left is not None:
        st = orc.blake2s_compress(st, np.frombuffer(left + right, dtype="<u4"))
    vals = list(vals)
    for i in range(0, len(vals), 16):
        blk = vals[i:i + 16] + [0] * (16 - len(vals[i:i + 16]))
        st = orc.blake2s_compress(st, np.array(blk, dtype=np.uint32))
    return st.astype("<u4").tobytes()


@pytest.mark.parametrize("variant", [0, 1])
def test_merkle_mixed_sizes_against_python_restatement(variant):
    orc.set_flavor(merkle_hash=variant)
    try:
        cols = [rng.integers(0, P, 1 << l, dtype=np.uint32) for l in (4, 2, 4, 0, 3, 4)]
        root = orc.merkle_commit(cols)
        # independent restatement of MerkleProver::commit (stable sort by length desc, layer by layer)
        order = sorted(range(len(cols)), key=lambda i: -cols[i].size)
        prev = None
        for log in range(4, -1, -1):
            here = [cols[i] for i in order if cols[i].size == 1 << log]
            layer = []
            for r in range(1 << log):
                vals = [int(c[r]) for c in here]
                if prev is None:
                    layer.append(_hash_node_py(None, None, vals, variant))
                else:
                    layer.append(_hash_node_py(prev[2 * r], prev[2 * r + 1], vals, variant))
            prev = layer
        assert root == prev[0]
    finally:
        orc.set_flavor()


def test_merkle_decommit_reconstructs_root():
    cols = [rng.integers(0, P, 1 << l, dtype=np.uint32) for l in (5, 5, 3, 5, 3)]
    root, layers = orc.merkle_commit(cols, want_layers=True)
    queries = {5: [3, 4, 17], 3: [0, 4]}
    qv, hw, cw = orc.merkle_decommit(cols, queries)
    # 3 queried rows x 3 cols at log 5, 2 rows x 2 cols at log 3
    assert len(qv) == 3 * 3 + 2 * 2
    assert list(qv[:3]) == [int(cols[0][3]), int(cols[1][3]), int(cols[3][3])]
    assert len(hw) > 0 and all(len(h) == 32 for h in hw)
    assert root == bytes(layers[:32])


def test_channel_semantics():
    ch = orc.Channel()
    assert ch.digest() == bytes(32)
    ch.mix_u64(0x0123456789ABCDEF)
    exp = hashlib.blake2s(bytes(32) + (0x0123456789ABCDEF).to_bytes(8, "little"), digest_size=32).digest()
    assert ch.digest() == exp
    r = ch.draw_random_bytes()
    assert r == hashlib.blake2s(exp + bytes(32), digest_size=32).digest()
    r2 = ch.draw_random_bytes()
    assert r2 == hashlib.blake2s(exp + (1).to_bytes(8, "little") + bytes(24), digest_size=32).digest()
    f = ch.draw_felt()
    assert all(int(x) < P for x in f)
    ch.mix_felts([[1, 2, 3, 4], [5, 6, 7, 8]])
    ch.mix_root(bytes(range(32)))
    assert ch.digest() != exp

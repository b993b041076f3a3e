"""CPU-only checks of the library's host logic (no kernels run): the Blake2sChannel of libnexus_b200.so against the
oracle's channel and hashlib, the AIR bytecode loader's validation, and the AIR builder's structural invariants."""
import hashlib

import numpy as np
import pytest

import nexus_zkvm_b200 as nb
from nexus_zkvm_b200 import air as A
from nexus_zkvm_b200 import machine as M
from nexus_zkvm_b200 import build as nb_build
from nexus_zkvm_b200.prover import Channel
from oracle import pyoracle as orc

P = (1 << 31) - 1


@pytest.fixture(scope="module", autouse=True)
def built():
    nb_build.build()


def test_channel_matches_oracle_and_hashlib():
    a, b = Channel(None), orc.Channel()
    assert a.digest() == bytes(32)
    rng = np.random.default_rng(5)
    for step in range(40):
        k = step % 5
        if k == 0:
            v = int(rng.integers(0, 1 << 63)); a.mix_u64(v); b.mix_u64(v)
        elif k == 1:
            f = rng.integers(0, P, (3, 4), dtype=np.uint32); a.mix_felts(f); b.mix_felts(f)
        elif k == 2:
            r = bytes(rng.integers(0, 256, 32, dtype=np.uint8)); a.mix_root(r); b.mix_root(r)
        elif k == 3:
            assert np.array_equal(a.draw_felts(3), b.draw_felts(3))
            assert np.array_equal(a.draw_felt(), b.draw_felt())
        else:
            assert a.draw_random_bytes() == b.draw_random_bytes()
        assert a.digest() == b.digest()
    c = Channel(None)
    c.mix_u64(7)
    assert c.digest() == hashlib.blake2s(bytes(32) + (7).to_bytes(8, "little"), digest_size=32).digest()
    d = c.clone(); d.mix_u64(1)
    assert d.digest() != c.digest()


def test_air_builder_invariants():
    m = M.AddMachine(log_size=8, n_lanes=3)
    main = m.main
    # declaration order of offsets is what defines the sample-point order of a column (Pc: [0, 1]; last logup col: [-1, 0])
    pc_offs = [o for (t, c, o) in main.masks if (t, c) == (1, 0)]
    assert pc_offs == [0, 1]
    last_cols = list(range(main.interaction_col0 + 4 * 35, main.interaction_col0 + 4 * 36))
    for c in last_cols:
        assert [o for (t, cc, o) in main.masks if (t, cc) == (2, c)] == [-1, 0]
    # extension masks are 4 consecutive slots
    words = m.words
    assert words[0] == 0x5241424E and words[3] == 2
    assert m.air.n_params == 2 + 2  # z, alpha^0, two cumsum shifts


def test_air_loader_rejects_malformed_bytecode():
    m = M.AddMachine(log_size=8, n_lanes=1)
    L = nb.lib()
    import ctypes as C

    def load(words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        h = C.c_void_p()
        st = L.nb200_air_load(None, w.ctypes.data_as(nb.u32p), C.c_size_t(w.size), C.byref(h))
        if st == 0:
            L.nb200_air_free(h)
        return st

    assert load(m.words) == 0
    bad = m.words.copy(); bad[0] ^= 1
    assert load(bad) != 0
    assert load(m.words[:-1]) != 0
    assert load(np.concatenate([m.words, [0]])) != 0
    # corrupt an opcode
    bad = m.words.copy()
    idx = int(np.where(bad == 0x5241424E)[0][0]) + 4 + 3 + 1 + 3 * len(m.main.masks) + 3  # first instruction's opcode
    bad[idx] = 99
    assert load(bad) != 0


def test_oracle_and_builder_agree_on_register_counts():
    # both parsers accept the same program and the oracle's interpreter stays inside the declared register file
    m = M.AddMachine(log_size=8, n_lanes=2)
    orc.Prover(m.words)


# ---- arithmetic identities the generated kernels rely on (csrc/jit.cu prelude, m31.cuh m31_red64) -----------------------------
def _red64_model(x):
    """Bit-level model of m31_red64: y = 2 * hi + lo (mad.wide), s = (yl & P) + funnelshift_r(yl, yh, 31), r = umin(s, s - P)."""
    P = (1 << 31) - 1
    M32 = (1 << 32) - 1
    hi, lo = x >> 32, x & M32
    y = 2 * hi + lo
    assert y < 3 << 32
    yl, yh = y & M32, y >> 32
    s = (yl & P) + (((yh << 32 | yl) >> 31) & M32)
    assert s <= P + 5 and s <= M32
    return min(s, (s - P) & M32)


def test_red64_is_the_canonical_residue_of_any_u64():
    import random
    P = (1 << 31) - 1
    rnd = random.Random(7)
    worst = 4 * (P - 1) * P + P                       # four products of a canonical by a value <= P, plus a canonical carry-in
    assert worst < 1 << 64
    cases = [0, 1, P - 1, P, P + 1, 2 * P, (1 << 32) - 1, 1 << 32, (1 << 62), (1 << 63) - 1, (1 << 64) - 1, worst, worst - 1]
    cases += [rnd.getrandbits(64) for _ in range(20000)]
    cases += [k * P + d for k in (1, 2, 3, (1 << 33) - 1, (1 << 33)) for d in (-1, 0, 1) if 0 <= k * P + d < 1 << 64]
    for x in cases:
        assert _red64_model(x) == x % P, hex(x)


def test_qmac_regrouping_equals_qm31_multiplication():
    """r_j = sum of four products with the derived multipliers (P - y1, P - y3, 2 y2 - y3, y2 + 2 y3, P - (y2 + 2 y3)) is (acc + x * y)_j."""
    import random
    from nexus_zkvm_b200 import field as F
    P = (1 << 31) - 1
    rnd = random.Random(11)
    edge = [0, 1, P - 1]
    for trial in range(3000):
        pick = (lambda: rnd.choice(edge)) if trial < 300 else (lambda: rnd.randrange(P))
        x, y, acc = tuple(pick() for _ in range(4)), tuple(pick() for _ in range(4)), tuple(pick() for _ in range(4))
        ny1, ny3 = P - y[1], P - y[3]
        gp = (2 * y[3] + y[2]) % P
        g = (2 * y[2] - y[3]) % P
        h = P - gp
        r = (acc[0] + x[0] * y[0] + x[1] * ny1 + x[2] * g + x[3] * h,
             acc[1] + x[0] * y[1] + x[1] * y[0] + x[2] * gp + x[3] * g,
             acc[2] + x[0] * y[2] + x[1] * ny3 + x[2] * y[0] + x[3] * ny1,
             acc[3] + x[0] * y[3] + x[1] * y[2] + x[2] * y[1] + x[3] * y[0])
        assert all(v < 1 << 64 for v in r)
        want = F.qm31_add(acc, F.qm31_mul(x, y))
        assert tuple(_red64_model(v) for v in r) == tuple(want)

"""Exhaustive bank-conflict check of the FFT tile kernel's shared-memory layout (nexus_zkvm_b200/csrc/fft.cu).

Model: 32 banks x 4 bytes.  A 32-bit warp access is conflict-free iff the 32 lanes hit 32 distinct banks; a 128-bit
access is processed per quarter-warp (8 lanes) and is conflict-free iff the 8 lanes hit 8 distinct 16-byte bank groups.
The kernel uses swz2 for the specialised tiles (128-bit staging, 128-bit round at bit 0, 32-bit rounds elsewhere) and
swz for the generic fallback (32-bit everywhere)."""
import pytest


def swz2(s):
    return s ^ (((s >> 5) & 3) << 2) ^ (((s >> 8) & 1) << 4)


def swz(s):
    return s ^ ((s >> 4) & 31)


def insert(tau, b, k):
    return ((tau >> b) << (b + 4)) | (k << b) | (tau & ((1 << b) - 1))


def wf32(addrs):
    banks = {}
    for a in addrs:
        banks.setdefault(a % 32, set()).add(a)
    return max(len(v) for v in banks.values())


def wf128(addrs):
    worst = 0
    for q in range(0, len(addrs), 8):
        grp = {}
        for a in addrs[q:q + 8]:
            grp.setdefault((a // 4) % 8, set()).add(a)
        worst = max(worst, max(len(v) for v in grp.values()))
    return worst


def rounds(T, W):
    L = T - W
    r = [W + 4 * i for i in range(L // 4)]
    if L % 4:
        r.append(T - 4)
    return r


FAST = [(9, 0), (10, 0), (11, 0), (12, 0), (13, 0), (12, 4), (12, 5), (12, 6), (12, 7), (12, 8), (13, 4)]


@pytest.mark.parametrize("T,W", FAST)
def test_specialised_tiles_are_conflict_free(T, W):
    nthreads = 1 << (T - 4)
    for b in rounds(T, W):
        for warp in range(0, nthreads, 32):
            lanes = range(warp, min(warp + 32, nthreads))
            if b == 0:
                for q in range(4):
                    assert wf128([swz2(insert(t, 0, 4 * q)) for t in lanes]) == 1
            else:
                for k in range(16):
                    assert wf32([swz2(insert(t, b, k)) for t in lanes]) == 1
    for base in range(0, 1 << T, 128):
        assert wf128([swz2(s) for s in range(base, base + 128, 4)]) == 1
    # swz2 is a bijection that keeps aligned groups of 4 words together
    assert sorted(swz2(s) for s in range(1 << T)) == list(range(1 << T))
    assert all(swz2(s) // 4 == swz2(s + 3) // 4 and swz2(s + 1) == swz2(s) + 1 for s in range(0, 1 << T, 4))


@pytest.mark.parametrize("T,W", [(12, 0), (13, 0), (14, 0), (12, 3), (12, 9), (12, 10), (13, 5)])
def test_generic_fallback_is_conflict_free(T, W):
    nthreads = 1 << (T - 4)
    for b in rounds(T, W):
        for warp in range(0, nthreads, 32):
            for k in range(16):
                assert wf32([swz(insert(t, b, k)) for t in range(warp, min(warp + 32, nthreads))]) == 1
    for base in range(0, 1 << T, 32):
        assert wf32([swz(s) for s in range(base, base + 32)]) == 1


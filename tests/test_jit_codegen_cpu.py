"""The AIR code generator (csrc/jit.cu) checked WITHOUT a GPU: the CUDA C it emits for a component's constraint program is compiled
as host C++ behind a thin shim (thread indices, __ldg, funnel shift; the one PTX statement is replaced by its C meaning) and run row
by row; the accumulated quotients must match the oracle's evaluate_constraint_quotients_on_domain on the same extended columns.
This exercises the 64-bit lazy multiply-accumulate, the coefficient tables, the next-row offsets and the chunking on every CPU run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import nexus_zkvm_b200 as nb
from nexus_zkvm_b200 import machine as M
from oracle import pyoracle as orc
from tests.oracle_backend import OracleBackend

P = (1 << 31) - 1

SHIM = r'''
#include <cstdint>
#include <cstring>
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __global__
#define __restrict__
#define __launch_bounds__(...)
#define __constant__
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((unsigned long long)hi) << 32) | lo) >> (s & 31)); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
struct Idx { unsigned x; };
static Idx blockIdx, blockDim, threadIdx;
static inline void __syncthreads() {}
'''

DRIVER = r'''
extern "C" void run_rows(const unsigned* const* cols, const unsigned* params, const unsigned* coeff, const unsigned* dinv,
                         unsigned* a0, unsigned* a1, unsigned* a2, unsigned* a3, unsigned EL, unsigned rows) {
  blockDim.x = 1; threadIdx.x = 0;
  for (unsigned i = 0; i < NB_NMASKS; ++i) ccols[i] = cols[i];   // the library fills the __constant__ table before each launch
  for (unsigned r = 0; r < rows; ++r) { blockIdx.x = r; nbjit(cols, params, coeff, dinv, a0, a1, a2, a3, EL, 0u); }
}
'''


def host_build(tmp_path, source, driver=None):
    src = source.decode()
    asm = 'asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(y) : "r"((u32)(x >> 32)), "l"((u64)(u32)x));'
    assert src.count(asm) == 1, "the generated prelude changed: update the host shim of this test"
    src = src.replace(asm, "y = (u64)(u32)(x >> 32) * 2ull + (u64)(u32)x;")
    cu = tmp_path / "kernel_host.cc"
    cu.write_text(SHIM + src + (driver or DRIVER))
    so = tmp_path / "kernel_host.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-w", str(cu), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def parse_component(words, k):
    """(log_size, log_expand, n_constraints, masks) of component k (bytecode layout: nexus_zkvm_b200/air.py)."""
    w = [int(x) for x in words]
    i = 4
    for c in range(w[3]):
        log_size, log_expand, n_constraints = w[i:i + 3]
        i += 3
        n_masks = w[i]
        masks = [(w[i + 1 + 3 * m], w[i + 2 + 3 * m], np.int32(np.uint32(w[i + 3 + 3 * m])).item()) for m in range(n_masks)]
        i += 1 + 3 * n_masks + 2
        i += 1 + 4 * w[i]
        n_fracs = w[i]
        i += 3
        i += 1 + 4 * w[i]
        i += n_fracs + 2
        if c == k:
            return log_size, log_expand, n_constraints, masks
    raise IndexError(k)


def coeff_table(coeffs):
    t = np.zeros((len(coeffs), 12), np.uint32)
    for k, y in enumerate(coeffs):
        y = [int(v) for v in y]
        gp = (2 * y[3] + y[2]) % P
        g = (2 * y[2] - y[3]) % P
        t[k, :9] = [y[0], y[1], y[2], y[3], P - y[1], P - y[3], g, gp, P - gp]
    return t


def _machine(kind, seed):
    """(machine, tree-1 host columns): the synthetic ADD machine, or the recorded v1 main component (19 code chunks, 26 + 15 live virtual
    registers: the case the chunk-boundary liveness analysis of csrc/jit.cu exists for)."""
    if kind == "nexus_v1":
        from nexus_zkvm_b200.nexus_v1 import NexusV1Machine
        m = NexusV1Machine(8)
        return m, m.fill_main_trace(seed=seed)
    lanes, pairs = kind
    m = M.AddMachine(log_size=8, n_lanes=lanes, logup_in_pairs=pairs)
    cols, mult = m.fill_main_trace(seed=seed + lanes, n_padding=2)
    return m, list(cols) + [mult]


def _flat(cols):
    out = []
    for c in cols:
        a = np.asarray(c)
        out += list(a.astype(np.uint32)) if a.ndim == 2 else [np.ascontiguousarray(a, dtype=np.uint32)]
    return out


def _draw(m, ch, params):
    for rel in (getattr(m, "relations", None) or [m.range256]):
        rel.draw(ch, params)


@pytest.mark.parametrize("kind", [(1, False), (2, True), "nexus_v1"])
def test_generated_constraint_kernel_matches_the_oracle_on_the_cpu(tmp_path, kind):
    m, t1 = _machine(kind, 9)
    # the oracle side: commit the three trees exactly as Machine::prove does
    be = OracleBackend()
    ch = be.channel()
    p = be.prover(m.words, dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3))
    tree0 = [orc.finalize_column(c) for c in _flat(m.preprocessed_columns())]
    tree1 = [orc.finalize_column(c) for c in _flat(t1)]
    p.commit(m.preprocessed_columns(), ch, coset_order=True)
    p.commit(t1, ch, coset_order=True)
    params = [(0, 0, 0, 0)] * m.air.n_params
    _draw(m, ch, params)
    inter = []
    for k, comp in enumerate(m.air.components):
        c, cs = p.gen_interaction(k, comp.log_size, max(comp.batching) + 1, params)
        inter.append(c)
        params[comp.cumsum_shift_param] = M.F.qm31_mul_m31(cs, M.F.m31_inv((1 << comp.log_size) % P))
    p.commit_interaction(inter, ch)
    tree2 = [c for block in inter for c in block]
    trees = [tree0, tree1, tree2]

    log_size, log_expand, n_constraints, masks = parse_component(m.words, 0)
    elog = log_size + log_expand
    rng = np.random.default_rng(3)
    coeffs = rng.integers(0, P, size=(n_constraints, 4), dtype=np.uint32)
    want = p.p.constraint_quotients(0, elog, np.array(params, dtype=np.uint32), coeffs)

    # the generated kernel on the host: every mask's column extended to the evaluation domain with the oracle's transforms
    words = np.ascontiguousarray(m.words, dtype=np.uint32)
    h = C.c_void_p()
    assert nb.lib().nb200_air_load(None, words.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(words.size), C.byref(h)) == 0
    src = C.c_void_p()
    assert nb.lib().nb200_air_kernel_source(h, C.c_uint32(0), C.c_int(0), C.byref(src)) == 0
    lib = host_build(tmp_path, C.string_at(src))
    nb.lib().nb200_free(src)
    nb.lib().nb200_air_free(h)
    ext = {}
    for (t, c, _off) in masks:
        if (t, c) not in ext:
            ext[(t, c)] = np.ascontiguousarray(orc.evaluate(orc.interpolate(trees[t][c]), elog), dtype=np.uint32)
    col_ptrs = (C.POINTER(C.c_uint32) * len(masks))(*[ext[(t, c)].ctypes.data_as(C.POINTER(C.c_uint32)) for (t, c, _o) in masks])
    prm = np.ascontiguousarray(np.array(params, dtype=np.uint32).reshape(-1, 4))
    tab = coeff_table(coeffs)
    rows = 1 << elog
    dinv = np.ones(1 << log_expand, np.uint32)     # the vanishing inverse is applied by the comparison below
    acc = [np.zeros(rows, np.uint32) for _ in range(4)]
    u32p = C.POINTER(C.c_uint32)
    lib.run_rows(col_ptrs, prm.ctypes.data_as(u32p), tab.ctypes.data_as(u32p), dinv.ctypes.data_as(u32p),
                 *[a.ctypes.data_as(u32p) for a in acc], C.c_uint32(elog), C.c_uint32(rows))
    got = np.stack(acc)                              # sum_k coeff_k * constraint_k per row, without 1 / vanishing
    # want = got * dinv[row >> log_size] with ONE field element per block of 2^log_size rows: recover it from one row, check all rows
    n_blk = 1 << log_expand
    for blk in range(n_blk):
        sl = slice(blk << log_size, (blk + 1) << log_size)
        g, w_ = got[:, sl].astype(object), want[:, sl].astype(object)
        nz = np.argwhere(g[0] != 0)
        assert nz.size, "degenerate block"
        r0 = int(nz[0][0])
        d = int(w_[0][r0]) * pow(int(g[0][r0]), P - 2, P) % P
        assert d != 0
        assert np.array_equal((g * d) % P, w_), f"block {blk}"


LOGUP_DRIVER = r'''
extern "C" void run_rows(const unsigned* const* cols, const unsigned* params, unsigned* out, unsigned LS, unsigned rows) {
  blockDim.x = 1; threadIdx.x = 0;
  for (unsigned i = 0; i < NB_NMASKS; ++i) ccols[i] = cols[i];
  for (unsigned r = 0; r < rows; ++r) { blockIdx.x = r; nbjit(cols, params, out, LS); }
}
'''


@pytest.mark.parametrize("kind", [(1, False), (2, True), "nexus_v1"])
def test_generated_logup_kernel_matches_the_oracle_on_the_cpu(tmp_path, kind):
    """The generated interaction-trace kernel (batched QM31 inverses, running row sums): every logup column except the last secure
    column (which additionally gets the coset-order prefix sum outside the kernel) must equal the oracle's LogupTraceGenerator."""
    m, t1 = _machine(kind, 4)
    be = OracleBackend()
    ch = be.channel()
    p = be.prover(m.words, dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3))
    tree0 = [orc.finalize_column(c) for c in _flat(m.preprocessed_columns())]
    tree1 = [orc.finalize_column(c) for c in _flat(t1)]
    p.commit(m.preprocessed_columns(), ch, coset_order=True)
    p.commit(t1, ch, coset_order=True)
    params = [(0, 0, 0, 0)] * m.air.n_params
    _draw(m, ch, params)
    comp = m.air.components[0]
    n_logup = max(comp.batching) + 1
    want, _cs = p.gen_interaction(0, comp.log_size, n_logup, params)
    want = np.asarray(want, dtype=np.uint32).reshape(4 * n_logup, -1)

    words = np.ascontiguousarray(m.words, dtype=np.uint32)
    h = C.c_void_p()
    assert nb.lib().nb200_air_load(None, words.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(words.size), C.byref(h)) == 0
    src = C.c_void_p()
    assert nb.lib().nb200_air_kernel_source(h, C.c_uint32(0), C.c_int(1), C.byref(src)) == 0
    lib = host_build(tmp_path, C.string_at(src), LOGUP_DRIVER)
    nb.lib().nb200_free(src)
    nb.lib().nb200_air_free(h)
    _ls, _le, _nc, masks = parse_component(m.words, 0)
    trees = [tree0, tree1]
    zero = np.zeros(1 << comp.log_size, np.uint32)
    keep = [trees[t][c] if (t < 2 and off == 0) else zero for (t, c, off) in masks]   # masks the logup program never reads stay zero
    u32p = C.POINTER(C.c_uint32)
    col_ptrs = (u32p * len(masks))(*[np.ascontiguousarray(a, dtype=np.uint32).ctypes.data_as(u32p) for a in keep])
    prm = np.ascontiguousarray(np.array(params, dtype=np.uint32).reshape(-1, 4))
    rows = 1 << comp.log_size
    out = np.zeros((4 * n_logup, rows), np.uint32)
    lib.run_rows(col_ptrs, prm.ctypes.data_as(u32p), out.ctypes.data_as(u32p), C.c_uint32(comp.log_size), C.c_uint32(rows))
    assert np.array_equal(out[:4 * (n_logup - 1)], want[:4 * (n_logup - 1)])

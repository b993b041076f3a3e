"""GPU parity of the full proving path (run on the B200 box: `pytest -m gpu`): the proof bytes produced through the
C ABI (nb200_scheme_commit / nb200_gen_interaction_trace / nb200_prove) must equal the oracle's bytes for the same
machine, trace and transcript, and the oracle's independent verifier must accept them."""
import numpy as np
import pytest

import nexus_zkvm_b200 as nb
from nexus_zkvm_b200 import machine as M
from nexus_zkvm_b200.prover import CudaBackend
from oracle import pyoracle as orc
from tests.oracle_backend import OracleBackend, verify

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    b = CudaBackend(nb.Context(0))
    yield b
    b.ctx.close()


def test_interaction_trace_bit_exact(backend):
    m = M.AddMachine(log_size=9, n_lanes=2)
    cols, mult = m.fill_main_trace(seed=11, n_padding=3)
    cfg = dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3)
    params = None
    outs = []
    for be in (backend, OracleBackend()):
        ch = be.channel()
        p = be.prover(m.words, cfg)
        p.commit(m.preprocessed_columns(), ch, coset_order=True)
        p.commit(list(cols) + [mult], ch, coset_order=True)
        prm = [(0, 0, 0, 0)] * m.air.n_params
        m.range256.draw(ch, prm)
        if params is None:
            params = prm
        assert prm == params  # same transcript so far => same lookup elements
        res = []
        for k, comp in enumerate(m.air.components):
            c, cs = p.gen_interaction(k, comp.log_size, max(comp.batching) + 1, prm)
            res.append((c.download() if hasattr(c, "download") else c, cs))
        outs.append(res)
    for (gc, gcs), (oc, ocs) in zip(*outs):
        assert gcs == ocs
        assert np.array_equal(gc, oc)
    assert M.verify_claimed_sums([cs for _, cs in outs[0]])


@pytest.mark.parametrize("log_size,lanes,pad", [(8, 1, 0), (8, 2, 5), (10, 1, 0), (12, 3, 100)])
def test_proof_bytes_match_oracle(backend, log_size, lanes, pad):
    m = M.AddMachine(log_size=log_size, n_lanes=lanes)
    cols, mult = m.fill_main_trace(seed=log_size * 10 + lanes, n_padding=pad)
    ad = bytes([log_size, lanes])
    g_proof, g_claimed, g_aux = M.prove(m, backend, cols, mult, associated_data=ad)
    o_proof, o_claimed, o_aux = M.prove(m, OracleBackend(), cols, mult, associated_data=ad)
    assert g_claimed == o_claimed
    assert g_aux["params"] == o_aux["params"]
    assert g_proof == o_proof, f"proof bytes differ (len {len(g_proof)} vs {len(o_proof)})"
    # the oracle's verifier accepts the GPU's proof (transcript replayed with the oracle's channel)
    verify(m, g_proof, o_aux)


@pytest.mark.parametrize("sizes", [list(range(4, 12)), list(range(4, 18)), list(range(4, 22))])
def test_multi_component_machine_proof_bytes(backend, sizes):
    """a prover2-shaped machine (machine.MultiMachine) — up to 20 components of distinct log sizes 4..21 (plus the 2^8
    and 2^16 lookup tables), degree bounds 1 and 2 mixed, fractions one per column and in pairs, a 3-ary BitOp relation against a
    preprocessed 2^16-row truth table (the keccak extension's table shape).  Trees with 20 column sizes, 20 DEEP-quotient sizes, FRI fed
    at every layer: the proof bytes equal the oracle's and the oracle's verifier accepts them."""
    import os
    orc.set_num_threads(os.cpu_count() or 1)
    m = M.MultiMachine(sizes)
    cols = m.fill_main_trace(seed=len(sizes))
    g_proof, g_claimed, g_aux = M.prove(m, backend, cols, None, associated_data=b"p2")
    o_proof, o_claimed, o_aux = M.prove(m, OracleBackend(), cols, None, associated_data=b"p2")
    assert g_aux["roots"] == o_aux["roots"]
    assert g_claimed == o_claimed and M.verify_claimed_sums(g_claimed)
    assert g_proof == o_proof, f"proof bytes differ (len {len(g_proof)} vs {len(o_proof)})"
    verify(m, g_proof, o_aux)


@pytest.mark.parametrize("log_size", [9, 12])
def test_nexus_v1_main_component_proof_bytes(backend, log_size):
    """The reference's v1 main component recorded as data (nexus_zkvm_b200/nexus_v1.py: 27/347/1012 columns, 413 constraints, 253 LogUp
    fractions over 9 relations with tuples of 1..9 values, next-row masks on Pc / IsPadding) with its padding-only witness: the GPU's proof
    bytes equal the oracle's and the oracle's verifier accepts them; an un-padded row is rejected."""
    from nexus_zkvm_b200.nexus_v1 import NexusV1Machine, MAIN_COLUMNS
    m = NexusV1Machine(log_size)
    cols = m.fill_main_trace(seed=log_size)
    g_proof, g_claimed, g_aux = M.prove(m, backend, cols, None, associated_data=b"v1")
    o_proof, o_claimed, o_aux = M.prove(m, OracleBackend(), cols, None, associated_data=b"v1")
    assert g_aux["roots"] == o_aux["roots"]
    assert g_claimed == o_claimed and M.verify_claimed_sums(g_claimed)
    assert g_proof == o_proof, f"proof bytes differ (len {len(g_proof)} vs {len(o_proof)})"
    verify(m, g_proof, o_aux)
    off = sum(s for n, s in MAIN_COLUMNS[:[n for n, _ in MAIN_COLUMNS].index("IsPadding")])
    cols[off] = cols[off].copy(); cols[off][5] = 0           # row 5 claims to execute something: the one-hot opcode sum (cpu.rs:379-420) fails
    with pytest.raises(nb.Nb200Error, match="status 5"):
        M.prove(m, backend, cols, None)


def test_non_default_config(backend):
    m = M.AddMachine(log_size=9, n_lanes=1)
    cols, mult = m.fill_main_trace(seed=77)
    cfg = dict(pow_bits=10, log_blowup=2, log_last=1, n_queries=7)
    g_proof, _, _ = M.prove(m, backend, cols, mult, config=cfg)
    o_proof, _, o_aux = M.prove(m, OracleBackend(), cols, mult, config=cfg)
    assert g_proof == o_proof
    verify(m, g_proof, o_aux)


def test_constraints_not_satisfied_is_reported(backend):
    m = M.AddMachine(log_size=8, n_lanes=1)
    cols, mult = m.fill_main_trace(seed=4)
    cols[2 + 8][17] = (int(cols[2 + 8][17]) + 1) % 256
    with pytest.raises(nb.Nb200Error, match="status 5"):
        M.prove(m, backend, cols, mult)


def test_paired_logup_matches_oracle(backend):
    # prover2-style batching: two fractions per secure column (finalize_logup_in_pairs)
    m = M.AddMachine(log_size=9, n_lanes=2, logup_in_pairs=True)
    assert m.air.n_columns()[2] == 4 * 12 + 4
    cols, mult = m.fill_main_trace(seed=21, n_padding=2)
    g_proof, g_claimed, _ = M.prove(m, backend, cols, mult)
    o_proof, o_claimed, o_aux = M.prove(m, OracleBackend(), cols, mult)
    assert g_claimed == o_claimed and g_proof == o_proof
    verify(m, g_proof, o_aux)


@pytest.mark.parametrize("merkle_hash,draw_sep", [(1, 0), (0, 1), (1, 1)])
def test_transcript_flavours_match_oracle(merkle_hash, draw_sep):
    # the parity-risk switches change the proof bytes consistently on both sides
    ctx = nb.Context(0)
    ctx.set_flavor(merkle_hash=merkle_hash, draw_domain_sep=draw_sep)
    orc.set_flavor(merkle_hash=merkle_hash, draw_domain_sep=draw_sep)
    try:
        m = M.AddMachine(log_size=8, n_lanes=1)
        cols, mult = m.fill_main_trace(seed=31)
        g_proof, _, _ = M.prove(m, CudaBackend(ctx), cols, mult)
        o_proof, _, o_aux = M.prove(m, OracleBackend(), cols, mult)
        assert g_proof == o_proof
        verify(m, g_proof, o_aux)
        orc.set_flavor()
        d_proof, _, _ = M.prove(m, OracleBackend(), cols, mult)
        assert d_proof != o_proof  # the switch really changes the transcript
    finally:
        orc.set_flavor()
        ctx.close()


def test_interpreter_and_jit_kernels_agree(backend, monkeypatch):
    # the constraint program runs either through the NVRTC-specialised kernel (default when libnvrtc is present) or the
    # bytecode interpreter (NB200_JIT=0); both must give the oracle's bytes
    m = M.AddMachine(log_size=9, n_lanes=2)
    cols, mult = m.fill_main_trace(seed=41)
    o_proof, _, _ = M.prove(m, OracleBackend(), cols, mult)
    monkeypatch.setenv("NB200_JIT", "0")
    i_proof, _, _ = M.prove(m, backend, cols, mult)
    monkeypatch.setenv("NB200_JIT", "1")
    j_proof, _, _ = M.prove(m, backend, cols, mult)
    assert i_proof == o_proof and j_proof == o_proof


def test_nvrtc_path_with_an_empty_cubin_cache(monkeypatch, tmp_path):
    """The shipped jit_cache/ holds nvcc-built cubins for the machines used here; point the cache at an empty directory so
    that this proof compiles its kernels with NVRTC at run time, fills the cache, and a second context loads them back."""
    monkeypatch.setenv("NB200_JIT_CACHE", str(tmp_path))
    m = M.AddMachine(log_size=8, n_lanes=1)
    cols, mult = m.fill_main_trace(seed=5)
    o_proof, _, _ = M.prove(m, OracleBackend(), cols, mult)
    for _ in range(2):  # first: NVRTC compile + store; second (new AIR handle): load from the cache
        ctx = nb.Context(0)
        try:
            g_proof, _, _ = M.prove(m, CudaBackend(ctx), cols, mult)
        finally:
            ctx.close()
        assert g_proof == o_proof
    assert len(list(tmp_path.glob("*.cubin"))) >= 1


@pytest.mark.parametrize("log_expand,blowup", [(1, 1), (3, 1), (2, 2), (3, 2), (1, 2)])
def test_every_quotient_domain_route(backend, log_expand, blowup):
    """Constraint quotients are evaluated on the committed LDE alone (degree bound = blow-up), on the LDE plus one half-size coset
    (degree bound = blow-up + 1) or on the reference's full domain (anything else); all three must give the oracle's bytes."""
    m = M.AddMachine(log_size=9, n_lanes=1, log_expand=log_expand)
    cols, mult = m.fill_main_trace(seed=60 + log_expand)
    cfg = dict(pow_bits=5, log_blowup=blowup, log_last=0, n_queries=3)
    g_proof, _, _ = M.prove(m, backend, cols, mult, config=cfg)
    o_proof, _, o_aux = M.prove(m, OracleBackend(), cols, mult, config=cfg)
    assert g_proof == o_proof
    verify(m, g_proof, o_aux)


def test_degree_hint_only_moves_work(backend):
    """nb200_scheme_set_constraint_log_degree lets host commits pre-compute the half-coset evaluations under the PCIe copy; the
    proof must not depend on it (0 = unknown, a wrong value, the right value)."""
    import ctypes as C
    m = M.AddMachine(log_size=9, n_lanes=2)
    cols, mult = m.fill_main_trace(seed=8)
    o_proof, _, _ = M.prove(m, OracleBackend(), cols, mult)

    class Hinted(CudaBackend):
        def __init__(self, ctx, hint):
            super().__init__(ctx)
            self.hint = hint

        def prover(self, words, config):
            p = super().prover(words, config)
            self.ctx._chk(nb.lib().nb200_scheme_set_constraint_log_degree(p._h, C.c_uint32(self.hint)))
            return p

    for hint in (0, 1, 2, 5):
        g_proof, _, _ = M.prove(m, Hinted(backend.ctx, hint), cols, mult)
        assert g_proof == o_proof, f"hint {hint}"

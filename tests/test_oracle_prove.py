"""Self-consistency of the oracle's full prover (CPU): a proof of the Nexus-shaped synthetic machine is accepted by
the oracle's independent verifier; tampering and unsatisfied constraints are rejected.  (The reference's own
proving tests are exactly this kind of prove -> verify round trip: prover/src/machine.rs:505-533.)"""
import numpy as np
import pytest

from nexus_zkvm_b200 import machine as M
from oracle import pyoracle as orc
from tests.oracle_backend import OracleBackend, verify

P = (1 << 31) - 1


@pytest.fixture(scope="module")
def proved():
    m = M.AddMachine(log_size=8, n_lanes=2)
    cols, mult = m.fill_main_trace(seed=3, n_padding=5)
    proof, claimed, aux = M.prove(m, OracleBackend(), cols, mult, associated_data=b"\x01\x02")
    return m, proof, claimed, aux


def test_air_shape():
    m = M.AddMachine(log_size=8, n_lanes=21)
    assert m.air.n_columns() == [3, 2 + 16 * 21 + 1, 4 * 12 * 21 + 4]  # 3 / 339 / 1012 columns
    assert len(m.main.constraints) == 3 + 8 * 21 + 12 * 21


def test_oracle_proof_verifies(proved):
    m, proof, claimed, aux = proved
    assert M.verify_claimed_sums(claimed)
    verify(m, proof, aux)
    assert 1000 < len(proof) < 200000


def test_tampered_proof_rejected(proved):
    m, proof, claimed, aux = proved
    rng = np.random.default_rng(0)
    rejected = 0
    for _ in range(12):
        i = int(rng.integers(8, len(proof)))
        bad = bytearray(proof); bad[i] ^= 1 << int(rng.integers(0, 7))
        try:
            verify(m, bytes(bad), aux)
        except orc.OracleError:
            rejected += 1
    assert rejected == 12


def test_unsatisfied_constraint_is_reported():
    m = M.AddMachine(log_size=8, n_lanes=1)
    cols, mult = m.fill_main_trace(seed=4)
    cols[2 + 8][17] = (int(cols[2 + 8][17]) + 1) % 256  # break c[0] of the ADD at row 17 (also breaks the range multiset)
    with pytest.raises(orc.OracleError, match="ConstraintsNotSatisfied"):
        M.prove(m, OracleBackend(), cols, mult)


def test_larger_trace_with_mixed_sizes():
    m = M.AddMachine(log_size=10, n_lanes=1)
    cols, mult = m.fill_main_trace(seed=5)
    proof, claimed, aux = M.prove(m, OracleBackend(), cols, mult)
    assert M.verify_claimed_sums(claimed)
    verify(m, proof, aux)


def test_paired_logup_proof_verifies():
    m = M.AddMachine(log_size=8, n_lanes=2, logup_in_pairs=True)
    cols, mult = m.fill_main_trace(seed=6)
    proof, claimed, aux = M.prove(m, OracleBackend(), cols, mult)
    assert M.verify_claimed_sums(claimed)
    verify(m, proof, aux)

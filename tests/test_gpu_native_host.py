"""The C ABI driven by a native C++ host (examples/prove_demo.cc) — no Python between the host columns and the proof bytes.
The proof must equal the oracle's (and therefore the Python harness's) byte for byte."""
import subprocess

import pytest

from nexus_zkvm_b200 import machine as M
from tests.native_job import build_demo, write_job
from tests.oracle_backend import OracleBackend, verify

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_size,lanes,pairs,ad", [(8, 1, False, b""), (10, 2, True, b"\x01\x7f")])
def test_native_host_produces_the_oracle_proof(tmp_path, log_size, lanes, pairs, ad):
    exe = build_demo()
    m = M.AddMachine(log_size=log_size, n_lanes=lanes, logup_in_pairs=pairs)
    cols, mult = m.fill_main_trace(seed=21 + lanes, n_padding=3)
    cfg = dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3)
    job, out = tmp_path / "job.bin", tmp_path / "proof.bin"
    write_job(job, m, cols, mult, cfg, ad)
    r = subprocess.run([exe, str(job), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    o_proof, _, o_aux = M.prove(m, OracleBackend(), cols, mult, config=cfg, associated_data=ad)
    assert out.read_bytes() == o_proof
    verify(m, out.read_bytes(), o_aux)

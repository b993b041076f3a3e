"""Test-only adapter: lets nexus_zkvm_b200.machine.prove() drive the ORACLE through the same backend protocol the
CUDA backend implements.  Lives under tests/ because only tests may touch oracle/."""
import numpy as np

from oracle import pyoracle as orc


class _OracleProver:
    def __init__(self, words, config):
        self.p = orc.Prover(words)
        self.config = config

    def commit(self, cols, ch, coset_order=False):
        flat = []
        for c in cols:  # 2-D blocks of columns are accepted like the product backend does
            a = np.asarray(c)
            flat += list(a) if a.ndim == 2 else [a]
        cols = [np.ascontiguousarray(c, dtype=np.uint32) for c in flat]
        if coset_order:
            cols = [orc.finalize_column(c) for c in cols]
        return self.p.commit(cols, ch, self.config["log_blowup"])

    def gen_interaction(self, comp, log_size, n_logup_cols, params):
        return self.p.gen_interaction(comp, log_size, n_logup_cols, np.array(params, dtype=np.uint32))

    def commit_interaction(self, inter, ch):
        cols = [c for block in inter for c in block]
        return self.p.commit(cols, ch, self.config["log_blowup"])

    def prove(self, ch, params):
        c = self.config
        return self.p.prove(ch, np.array(params, dtype=np.uint32), c["pow_bits"], c["log_blowup"], c["log_last"], c["n_queries"])


class OracleBackend:
    def channel(self):
        return orc.Channel()

    def prover(self, words, config):
        return _OracleProver(words, config)


def verify(machine, proof, aux):
    orc.verify(machine.words, np.array(aux["params"], dtype=np.uint32), proof, aux["channel_at_prove"], machine.column_log_sizes())


def verify_with_replayed_transcript(machine, proof, claimed, aux):
    """The oracle's verifier on a proof produced elsewhere (the GPU): the transcript is replayed with the ORACLE's channel from the values the
    prover returned (associated data, log sizes, the three roots, claimed sums) in Machine::prove's order (machine.rs:197-263)."""
    ch = orc.Channel()
    for byte in aux["associated_data"]:
        ch.mix_u64(int(byte))
    for ls in aux["log_sizes"]:
        ch.mix_u64(ls)
    ch.mix_root(aux["roots"][0])
    ch.mix_root(aux["roots"][1])
    for _ in (getattr(machine, "relations", None) or [None]):
        ch.draw_felts(2)                      # LookupElements::draw per relation
    ch.mix_felts(claimed)
    ch.mix_root(aux["roots"][2])
    orc.verify(machine.words, np.array(aux["params"], dtype=np.uint32), proof, ch, machine.column_log_sizes())

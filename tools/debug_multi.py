"""Localise a GPU/oracle difference on machine.MultiMachine: per size set, compare roots, interaction traces, claimed sums, proof bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nexus_zkvm_b200 as nb
from nexus_zkvm_b200 import machine as M
from nexus_zkvm_b200.prover import CudaBackend
from tests.oracle_backend import OracleBackend
from oracle import pyoracle as orc

be = CudaBackend(nb.Context(0))
cfg = dict(pow_bits=5, log_blowup=1, log_last=0, n_queries=3)
for sizes in ([8, 9], [9, 8], [8], [9], [4], [5], [6, 7], [4, 5, 6, 7], [8, 9, 10, 11], list(range(4, 12))):
    m = M.MultiMachine(sizes)
    cols = m.fill_main_trace(seed=1)
    outs = []
    for b in (be, OracleBackend()):
        ch = b.channel()
        for ls in m.log_sizes:
            ch.mix_u64(ls)
        p = b.prover(m.words, cfg)
        r0 = p.commit(m.preprocessed_columns(), ch, coset_order=True)
        r1 = p.commit(list(cols), ch, coset_order=True)
        prm = [(0, 0, 0, 0)] * m.air.n_params
        for rel in m.relations:
            rel.draw(ch, prm)
        res = []
        for k, comp in enumerate(m.air.components):
            c, cs = p.gen_interaction(k, comp.log_size, max(comp.batching) + 1, prm)
            res.append((c.download() if hasattr(c, "download") else np.asarray(c), cs))
        outs.append((r0, r1, prm, res))
    g, o = outs
    msg = [f"sizes {sizes}: roots {'ok' if g[0] == o[0] and g[1] == o[1] else 'DIFF'} params {'ok' if g[2] == o[2] else 'DIFF'}"]
    for k, ((gc, gcs), (oc, ocs)) in enumerate(zip(g[3], o[3])):
        if gcs != ocs or not np.array_equal(gc, oc):
            bad = [i for i in range(gc.shape[0]) if not np.array_equal(gc[i], oc[i])]
            msg.append(f"comp {k} (log {m.air.components[k].log_size}): claimed {'ok' if gcs == ocs else 'DIFF'} bad interaction columns {bad[:8]} of {gc.shape[0]}")
    try:
        gp, _, _ = M.prove(m, be, cols, None)
        op, _, _ = M.prove(m, OracleBackend(), cols, None)
        msg.append("proof " + ("ok" if gp == op else "BYTES DIFF"))
    except Exception as e:
        msg.append("prove failed: " + str(e)[-60:])
    print(" | ".join(msg), flush=True)

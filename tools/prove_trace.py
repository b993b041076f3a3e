#!/usr/bin/env python
"""Stage-by-stage wall clock of one full proof on the GPU (NB200_TRACE=1 makes the library print its own stages).

    NB200_TRACE=1 python tools/prove_trace.py --log-rows 20 --lanes 21
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-rows", type=int, default=20)
    ap.add_argument("--lanes", type=int, default=21)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import nexus_zkvm_b200 as nb
    from nexus_zkvm_b200 import machine as M
    from nexus_zkvm_b200.prover import CudaBackend
    t0 = time.perf_counter()
    m = M.AddMachine(log_size=a.log_rows, n_lanes=a.lanes)
    t1 = time.perf_counter()
    ctx = nb.Context(0)
    cols, mult = m.fill_main_trace(seed=1, out=ctx.host_alloc(m.n_main_columns(), a.log_rows))
    t2 = time.perf_counter()
    print(f"[host] build AIR {1e3 * (t1 - t0):.1f} ms, fill trace {1e3 * (t2 - t1):.1f} ms, bytecode {m.words.size} words", file=sys.stderr)
    be = CudaBackend(ctx)
    for r in range(a.reps):
        print(f"---- rep {r}", file=sys.stderr)
        t = time.perf_counter()
        proof, claimed, aux = M.prove(m, be, cols, mult)
        print(f"[host] prove total {1e3 * (time.perf_counter() - t):.1f} ms, proof {len(proof)} bytes", file=sys.stderr)


if __name__ == "__main__":
    main()

"""Test helper: build examples/prove_demo (a C++ host for the C ABI) and write its job file for an AddMachine trace."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "prove_demo")


def build_demo():
    src = os.path.join(ROOT, "examples", "prove_demo.cc")
    lib = os.path.join(ROOT, "nexus_zkvm_b200", "libnexus_b200.so")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), src, "-L" + os.path.join(ROOT, "nexus_zkvm_b200"),
               "-lnexus_b200", "-Wl,-rpath," + os.path.join(ROOT, "nexus_zkvm_b200"), "-o", EXE]
        subprocess.run(cmd, check=True)
    return EXE


def _batches(cols):
    """Same grouping as CommitmentSchemeProver._host_batches: consecutive equal-length columns form one batch."""
    out, run = [], []
    for c in cols:
        a = np.ascontiguousarray(c, dtype=np.uint32)
        if a.ndim == 2:
            if run:
                out.append(np.stack(run)); run = []
            out.append(a)
            continue
        if run and run[0].size != a.size:
            out.append(np.stack(run)); run = []
        run.append(a)
    if run:
        out.append(np.stack(run))
    return out


def write_job(path, machine, main_cols, mult, config, associated_data=b""):
    w = [0x424A424E, 1, int(machine.words.size)]
    parts = [np.array(w, np.uint32), np.ascontiguousarray(machine.words, dtype=np.uint32)]
    hdr = [config["pow_bits"], config["log_blowup"], config["log_last"], config["n_queries"], len(associated_data), *associated_data,
           2, machine.log_size, 8]
    rels = [machine.range256]
    hdr += [len(rels)]
    for r in rels:
        hdr += [r.z, r.size, *r.alpha_powers]
    comps = machine.air.components
    hdr += [len(comps)]
    for c in comps:
        hdr += [c.log_size, max(c.batching) + 1, c.cumsum_shift_param]
    parts.append(np.array(hdr, np.uint32))
    main_part = [main_cols] if getattr(main_cols, "ndim", 1) == 2 else list(main_cols)
    for tree in (machine.preprocessed_columns(), main_part + [mult]):
        bs = _batches(tree)
        parts.append(np.array([len(bs)], np.uint32))
        for b in bs:
            parts.append(np.array([b.shape[0], int(b.shape[1]).bit_length() - 1], np.uint32))
            parts.append(b.reshape(-1))
    with open(path, "wb") as f:
        for p in parts:
            f.write(np.ascontiguousarray(p, dtype=np.uint32).tobytes())

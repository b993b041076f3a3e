#!/usr/bin/env python
"""Pre-compile the generated AIR kernels of the v1 main component at 2^log_rows rows for alternative code-generation settings (NB200_JIT_CHUNK =
bytecode instructions per CTA-synchronised chunk, NB200_JIT_BLOCK = threads per CTA), so that a sweep on the GPU box costs no NVRTC time:

    python tools/jit_variants.py 20 125:512 500:512 1000:512 250:512:512           # chunk:block[:bound], here (no GPU needed)
    NB200_JIT_CHUNK=500 NB200_JIT_BLOCK=512 NB200_TRACE=1 python bench.py ...      # on the box: the library finds the variant's cubin by source hash
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r"""
import os, subprocess, sys
sys.path.insert(0, %r)
import nexus_zkvm_b200.build as B
from nexus_zkvm_b200.nexus_v1 import NexusV1Machine
m = NexusV1Machine(int(sys.argv[1]))
for key, src in B.kernel_sources(m.words):
    path = os.path.join(B.JIT_CACHE, f"{key:016x}.cubin")
    if os.path.exists(path):
        print("have", path); continue
    cu = path[:-6] + ".cu"
    open(cu, "wb").write(src)
    r = subprocess.run([B.NVCC] + B.ARCH + ["-O3", "-std=c++17", "-lineinfo", "-ccbin", "/usr/bin/g++", "-cubin", cu, "-o", path], capture_output=True, text=True)
    os.remove(cu)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-2000:])
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    print("built", os.path.basename(path), [l.strip() for l in res.splitlines() if "REG:" in l][:1])
""" % ROOT


def main():
    log_rows = sys.argv[1]
    procs = []
    for spec in sys.argv[2:]:
        chunk, block, *bound = spec.split(":")
        env = dict(os.environ, NB200_JIT_CHUNK=chunk, NB200_JIT_BLOCK=block)
        if bound:
            env["NB200_JIT_BOUND"] = bound[0]
        procs.append((spec, subprocess.Popen([sys.executable, "-c", CHILD, log_rows], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for spec, p in procs:
        out, _ = p.communicate()
        print(f"--- chunk:block = {spec}\n{out.strip()}")


if __name__ == "__main__":
    main()

"""Build provenance of libnexus_b200.so for profiles/: toolchain, flags, per-kernel ptxas resource usage (registers / spills / shared),
and the SASS evidence for the async-copy staging (LDGSTS) and the absence of local-memory spills in the FFT kernels.

    python tools/build_provenance.py > profiles/build_rNN.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from nexus_zkvm_b200 import build as B  # noqa: E402


def sh(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def short(d):
    """Demangled kernel name without its parameter list (template arguments kept, casts like (bool)1 -> 1)."""
    d = re.sub(r">\(.*$", ">", d) if ">(" in d else re.sub(r"\(.*", "", d)
    d = re.sub(r"\((?:bool|int|unsigned int)\)", "", d)
    return d.replace("void ", "").replace("nb::", "").replace("unsigned int", "u32")


def main():
    lib = B.build()
    print("# build provenance: libnexus_b200.so")
    print("nvcc      :", sh([B.NVCC, "--version"]).strip().splitlines()[-2])
    print("host cc   :", sh(["/usr/bin/g++", "--version"]).splitlines()[0])
    print("flags     :", " ".join(B.ARCH + B.CFLAGS))
    print("library   : %d bytes, sha256 %s" % (os.path.getsize(lib), hashlib.sha256(open(lib, "rb").read()).hexdigest()[:16]))
    print("sources   :")
    for s in B._sources():
        print("   %-22s sha256 %s" % (os.path.relpath(s, B.CSRC), hashlib.sha256(open(s, "rb").read()).hexdigest()[:16]))
    print("embedded  :", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sh(["cuobjdump", "-lelf", lib]) + sh(["cuobjdump", lib]))))) or "sm_100a")
    # ---- resource usage per kernel (cuobjdump -res-usage reads the cubin's own records: what actually ships)
    res = sh(["cuobjdump", "-res-usage", lib])
    rows = []
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((name, int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))))
            name = None
    dem = sh(["cu++filt"] + [r[0] for r in rows]).splitlines() if rows else []
    print("\n# kernels: registers / stack bytes / static shared bytes / local bytes   (%d kernels)" % len(rows))
    for (mn, reg, stack, shared, local), d in sorted(zip(rows, dem), key=lambda x: x[1]):
        d = short(d)
        print("  %-110s REG %3d  STACK %4d  SHARED %6d  LOCAL %4d" % (d[:110], reg, stack, shared, local))
    # ---- SASS mnemonics of the FFT / Merkle / quotient kernels
    sass = sh(["cuobjdump", "-sass", lib])
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    names = list(counts)
    dem = dict(zip(names, sh(["cu++filt"] + names).splitlines()))
    print("\n# SASS instruction mix of the hot kernels (static counts; LDGSTS = cp.async global->shared staging, IMAD.WIDE folded into IMAD)")
    for mn in sorted(names, key=lambda x: dem[x]):
        d = short(dem[mn])
        if not re.search(r"fft_tile_async|fft_mid|merkle_layer|blake2s_leaf|quotients_kernel|eval_points|reorder_kernel|fold_", d):
            continue
        c = counts[mn]
        tot = sum(c.values())
        top = ", ".join("%s %d" % kv for kv in c.most_common(8))
        print("  %-100s %5d instr; LDGSTS %3d  STL %d LDL %d | %s" % (d[:100], tot, c.get("LDGSTS", 0), c.get("STL", 0), c.get("LDL", 0), top))
    # ---- the shipped cubin cache of the generated AIR kernels
    cache = sorted(f for f in os.listdir(B.JIT_CACHE) if f.endswith(".cubin")) if os.path.isdir(B.JIT_CACHE) else []
    print("\n# jit_cache: %d cubins (generated AIR kernels, compiled by NVRTC for sm_100a at build time; keyed by source hash)" % len(cache))
    for f in cache:
        p = os.path.join(B.JIT_CACHE, f)
        r = sh(["cuobjdump", "-res-usage", p])
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", r)
        print("  %s  %8d bytes  %s" % (f, os.path.getsize(p), m.group(0) if m else ""))


if __name__ == "__main__":
    main()

"""The Rust shim (shim/nexus-b200) cannot be compiled in this image (no cargo).  What CAN be checked here: its
`extern "C"` block declares exactly the entry points of include/nb200.h, with the same number of parameters, and the recorder's
opcode table equals air.py's (the bytecode both emit is what nb200_air_load parses)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_args(s):
    s = s.strip()
    if s in ("", "void"):
        return []
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([<":
            depth += 1
        elif ch in ")]>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return [a for a in out if a.strip()]


def _header_decls():
    h = open(os.path.join(ROOT, "include", "nb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"\b(nb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S)}


def _rust_decls():
    r = open(os.path.join(ROOT, "shim", "nexus-b200", "src", "ffi.rs")).read()
    r = re.sub(r"//.*", "", r)
    return {m.group(1): len(_split_args(m.group(2))) for m in re.finditer(r"pub fn (nb200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(?:->[^;]*)?;", r, flags=re.S)}


def test_ffi_declares_every_entry_point_with_the_same_arity():
    c, r = _header_decls(), _rust_decls()
    assert len(c) > 60
    assert set(c) == set(r), f"only in nb200.h: {sorted(set(c) - set(r))}; only in ffi.rs: {sorted(set(r) - set(c))}"
    bad = {k: (c[k], r[k]) for k in c if c[k] != r[k]}
    assert not bad, f"parameter counts differ (header, rust): {bad}"


def test_recorder_opcodes_match_air_py():
    from nexus_zkvm_b200 import air as A
    r = open(os.path.join(ROOT, "shim", "nexus-b200", "src", "recorder.rs")).read()
    ops = {m.group(1): int(m.group(2)) for m in re.finditer(r"const (OP_[A-Z]+): u32 = (\d+);", r)}
    assert len(ops) >= 19
    for name, val in ops.items():
        assert getattr(A, name) == val, name
    assert "0x5241_424E" in r and A.NO_PARAM == 0xFFFFFFFF


def test_shim_files_present():
    for f in ("shim/nexus-b200/Cargo.toml", "shim/nexus-b200/build.rs", "shim/nexus-b200/src/lib.rs", "shim/nexus-b200/src/context.rs",
              "shim/nexus-b200/src/recorder.rs", "shim/prover-patch/src/cuda/mod.rs", "shim/prover-patch/src/cuda/lookups.rs",
              "shim/prover-patch/tests/differential.rs", "shim/prover-patch/apply.md"):
        assert os.path.getsize(os.path.join(ROOT, f)) > 200, f

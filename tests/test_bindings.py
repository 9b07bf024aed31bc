"""The boundary's three descriptions of itself must agree: the C header (include/lorikeet_b200.h), the Python ctypes mirror
(lorikeet_b200/_lib.py) and the Rust binding crate (bindings/lorikeet_b200-sys/src/lib.rs, which cannot be compiled here: no
Rust toolchain in the build container).  tests/c_abi_smoke.c is compiled as strict C11 against the header and reports struct
layouts; the Rust file is checked declaration by declaration."""
import ctypes as C
import json
import os
import re
import subprocess

from conftest import ROOT
from lorikeet_b200 import _lib, synth

HDR = os.path.join(ROOT, "include", "lorikeet_b200.h")
RUST_SYS = os.path.join(ROOT, "bindings", "lorikeet_b200-sys", "src", "lib.rs")
RUST_SAFE = os.path.join(ROOT, "bindings", "lorikeet_b200", "src", "lib.rs")


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def _c_functions():
    """name -> number of parameters, from the header"""
    src = _strip_comments(open(HDR).read())
    out = {}
    for m in re.finditer(r"\b(lk_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def _c_structs():
    src = _strip_comments(open(HDR).read())
    out = {}
    for m in re.finditer(r"typedef struct (lk_\w+) \{(.*?)\} \1;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if decl:
                for part in decl.split(","):
                    fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
        out[m.group(1)] = fields
    return out


def test_header_is_strict_c11_and_layouts_match_the_ctypes_mirror(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe], check=True)
    info = json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    mirrors = {"lk_sw_params": _lib.SwParams, "lk_pairhmm_batch": _lib.PairHmmBatchStruct, "lk_realign_batch": _lib.RealignBatchStruct,
               "lk_engine_args": _lib.EngineArgs, "lk_synth_params": synth.SynthParams}
    for name, cls in mirrors.items():
        assert C.sizeof(cls) == info[name]["size"], name
        assert [f[0] for f in cls._fields_] == [f[0] for f in info[name]["fields"]], name
        for fname, off in info[name]["fields"]:
            assert getattr(cls, fname).offset == off, (name, fname)
    e = info["enums"]
    assert (e["LK_OK"], e["LK_ERR_INVALID"], e["LK_ERR_CUDA"], e["LK_ERR_NOMEM"], e["LK_ERR_CAPACITY"]) == \
           (_lib.LK_OK, _lib.LK_ERR_INVALID, _lib.LK_ERR_CUDA, _lib.LK_ERR_NOMEM, _lib.LK_ERR_CAPACITY)
    # c_abi_smoke.c lists every field of every struct the header defines
    for name, fields in _c_structs().items():
        assert [f[0] for f in info[name]["fields"]] == fields, name


def test_rust_sys_crate_declares_exactly_the_header():
    rust = _strip_comments(open(RUST_SYS).read())
    block = rust[rust.index('extern "C" {'):]
    rust_fns = {}
    for m in re.finditer(r"pub fn (lk_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S):
        params = m.group(2).strip().rstrip(",")
        rust_fns[m.group(1)] = 0 if not params else len([p for p in params.split(",") if p.strip()])
    c_fns = _c_functions()
    assert set(rust_fns) == set(c_fns), (sorted(set(c_fns) - set(rust_fns)), sorted(set(rust_fns) - set(c_fns)))
    for name, n in c_fns.items():
        assert rust_fns[name] == n, (name, n, rust_fns[name])
    # structs: same fields in the same order
    for name, fields in _c_structs().items():
        m = re.search(r"pub struct %s \{(.*?)\n\}" % name, rust, flags=re.S)
        assert m, name
        rfields = re.findall(r"pub (\w+):", m.group(1))
        assert rfields == fields, (name, rfields, fields)
    # enum constants
    hdr = _strip_comments(open(HDR).read())
    for const, val in re.findall(r"\b(LK_[A-Z0-9_]+)\s*=\s*(\d+)", hdr):
        m = re.search(r"pub const %s: \w+ = (\d+);" % const, rust)
        assert m and int(m.group(1)) == int(val), const


def test_safe_wrapper_uses_only_declared_functions_and_covers_the_seams():
    safe = _strip_comments(open(RUST_SAFE).read())
    used = set(re.findall(r"sys::(lk_[a-z0-9_]+)\s*\(", safe))
    assert used <= set(_c_functions())
    # the two gkl seams, the region driver, the callers either side, the pool and the batcher are all wrapped
    for fn in ("lk_pairhmm_forward", "lk_sw_align", "lk_pairhmm_batch_host", "lk_engine_compute_read_likelihoods", "lk_sw_batch_host",
               "lk_calculate_cigar_batch", "lk_realign_reads_batch", "lk_dangling_end_sw_batch", "lk_pool_create",
               "lk_pool_compute_read_likelihoods", "lk_pool_batcher_create", "lk_batcher_submit_async", "lk_batcher_wait"):
        assert fn in used, fn

"""The C-ABI library loads on a CPU-only box, exports every symbol include/ganrev.h declares,
and refuses to run without an sm_100 GPU (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ganrev.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ganrev_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    names = header_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(pkg._lib.SO_PATH)
    for n in names:
        assert hasattr(lib, n), f"libganrev_cuda.so does not export {n}"
    assert sorted(pkg._lib.EXPORTS) == names, "ctypes binding and header disagree"


def test_no_torch_in_the_abi(pkg):
    import subprocess
    out = subprocess.run(["ldd", pkg._lib.SO_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "cudnn" not in out and "cublas" not in out, out


def test_create_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.GanrevError):
        pkg.Context(0)
    with pytest.raises(NotImplementedError):
        pkg.models._Module().float()          # MODEL:float() has no CPU path to fall back to


def test_oracle_is_not_reachable_from_the_product():
    """Nothing under gan-reverser_b200/ or include/ may import, include, link or load the oracle
    (comments may cite it as the mirror of a kernel's arithmetic)."""
    bad = re.compile(r"import\s+oracle|from\s+oracle|libganrev_oracle|ganrev_oracle\.h|\borc_[a-z0-9_]+\s*\(|dlopen\([^)]*oracle")
    for base in ("gan-reverser_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".lua", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert not bad.search(txt), (dp, f, bad.search(txt).group(0))


def test_lua_shim_declarations_match_header():
    """The LuaJIT shim cannot run here, so at least its ffi.cdef block must agree with include/ganrev.h: every function
    it declares exists in the header with the same return type and parameter types, and every lib.<fn> it calls is declared."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ganrev.h")).read()
    lua = open(os.path.join(root, "gan-reverser_b200", "lua", "ganrev.lua")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)

    def protos(text):
        out = {}
        for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(ganrev_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
            ret, name, args = m.group(1), m.group(2), m.group(3)
            types = []
            for a in args.split(","):
                a = " ".join(a.split())
                if a in ("void", ""):
                    continue
                a = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                types.append(a.replace(" *", "*").replace("* ", "*").strip())
            out[name] = (" ".join(ret.split()).replace(" *", "*"), types)
        return out

    h = protos(hdr)
    cdef = re.search(r"ffi\.cdef\[\[(.*?)\]\]", lua, flags=re.S).group(1)
    l = protos(cdef)
    assert len(l) >= 20
    for name, sig in l.items():
        assert name in h, f"{name} is declared in ganrev.lua but not in ganrev.h"
        assert sig == h[name], f"{name}: ganrev.lua says {sig}, ganrev.h says {h[name]}"
    called = set(re.findall(r"\blib\.(ganrev_[A-Za-z0-9_]+)", lua))
    assert called <= set(l), f"ganrev.lua calls undeclared functions: {sorted(called - set(l))}"

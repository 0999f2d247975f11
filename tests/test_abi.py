"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/gumbi_b200.h declares, the ctypes
struct mirrors have the C layout, argument validation happens host-side, and the product path fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "gumbi_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gb2_[A-Za-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(lib_built):
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib_built, s), f"{s} declared in include/gumbi_b200.h but not exported"
    from gumbi_b200 import _lib

    assert sorted(_lib.EXPORTS) == syms, "gumbi_b200/_lib.py EXPORTS out of sync with the header"
    assert lib_built.gb2_abi_version() == 1


def test_struct_layout_matches_c(lib_built, tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from gumbi_b200 import _lib

    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "gumbi_b200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(gb2_term), sizeof(gb2_kernel),"
        " offsetof(gb2_term, ls), offsetof(gb2_term, coreg_B), offsetof(gb2_kernel, sigma),"
        " offsetof(gb2_kernel, noise_B), offsetof(gb2_kernel, jitter)); return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    want = [C.sizeof(_lib.Term), C.sizeof(_lib.Kernel), _lib.Term.ls.offset, _lib.Term.coreg_B.offset,
            _lib.Kernel.sigma.offset, _lib.Kernel.noise_B.offset, _lib.Kernel.jitter.offset]
    assert got == want


def test_no_gpu_fails_loudly(lib_built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from gumbi_b200 import BackendUnavailable, GPEngine

    with pytest.raises(BackendUnavailable, match="no CPU fallback"):
        GPEngine(0)


def test_missing_library_fails_loudly(monkeypatch):
    from gumbi_b200 import BackendUnavailable, _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", "/nonexistent/libgumbi_b200.so")
    with pytest.raises(BackendUnavailable, match="no CPU fallback"):
        _lib.load()


def test_null_handle_and_bad_create(lib_built):
    assert lib_built.gb2_factorize(None) == -1
    assert lib_built.gb2_destroy(None) == 0
    h = C.c_void_p()
    assert lib_built.gb2_create(C.byref(h), 0, 7) == -1
    assert b"precision" in lib_built.gb2_last_error(None)


def test_kernel_struct_validation():
    from gumbi_b200 import _lib

    base = {"terms": [{"kind": "ExpQuad", "cont_idx": [0, 1], "ls": [1.0, 2.0], "eta": 1.0}], "sigma": 0.1}
    k, keep = _lib.make_kernel_struct(base)
    assert k.n_terms == 1 and k.terms[0].d == 2 and k.noise_col == -1 and k.jitter == 1e-6
    k, _ = _lib.make_kernel_struct({"terms": [{"kind": "Matern52", "cont_idx": [0, 1, 2], "ls": [1.5], "eta": 1.0}], "sigma": 0.1})
    assert [k.terms[0].ls[i] for i in range(3)] == [1.5, 1.5, 1.5]  # ARD=False: shared lengthscale (GP.py:400)
    with pytest.raises(ValueError, match="Continuous kernel must be one of"):
        _lib.make_kernel_struct({"terms": [{"kind": "RatQuad", "cont_idx": [0], "ls": [1.0], "eta": 1.0}], "sigma": 0.1})
    with pytest.raises(ValueError):
        _lib.make_kernel_struct({"terms": [{"kind": "ExpQuad", "cont_idx": [0, 1], "ls": [1.0, 2.0, 3.0], "eta": 1.0}], "sigma": 0.1})
    with pytest.raises(ValueError):
        _lib.make_kernel_struct({"terms": [], "sigma": 0.1})
    with pytest.raises(ValueError, match="levels"):
        _lib.make_kernel_struct({"terms": [{"kind": "ExpQuad", "cont_idx": [0], "ls": [1.0], "eta": 1.0,
                                            "coreg": [{"col": 1, "W": np.zeros((3, 2)), "kappa": [1.0, 1.0]}]}], "sigma": 0.1})
    B = _lib.coregion_B([[1.0, 2.0], [0.5, -1.0]], [0.3, 0.7])
    np.testing.assert_allclose(B, [[5.3, -1.5], [-1.5, 1.95]])


def test_product_package_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gumbi_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gumbi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "gp_oracle" not in txt, f
    # dev tools measure the product path only: checker programs that need the oracle live under tests/
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "tools", f), encoding="utf-8").read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|gen_golden|test_backend_host)\b", txt, flags=re.M), f
            assert "gp_oracle" not in txt and "OracleEngine" not in txt, f

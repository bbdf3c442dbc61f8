"""CPU-side checks of the drop-in boundary: librae.so builds/loads, exports every symbol include/rae.h declares, the
ctypes structs match the C layout, and the product refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from relation_autoencoder_b200 import _lib as L
from relation_autoencoder_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return L.load()


def test_library_exports_every_header_symbol(lib):
    names = L.header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "librae.so does not export %s declared in include/rae.h" % n
    # and the binding covers exactly the header
    assert sorted(L._SIGNATURES) == names


def test_abi_version(lib):
    assert lib.rae_abi_version() == L.RAE_ABI_VERSION


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rae.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(rae_config),offsetof(rae_config,F),offsetof(rae_config,lr),offsetof(rae_config,z_total),'
                   'sizeof(rae_step_stats),offsetof(rae_step_stats,algorithmic_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [C.sizeof(L.RaeConfig), L.RaeConfig.F.offset, L.RaeConfig.lr.offset, L.RaeConfig.z_total.offset,
           C.sizeof(L.RaeStepStats), L.RaeStepStats.algorithmic_bytes.offset]
    assert [int(x) for x in out] == got


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "h.c"
    src.write_text('#include "rae.h"\nint main(void){return RAE_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "h.o")])


def test_header_cites_reference_interfaces():
    txt = open(L.HEADER_PATH).read()
    for cite in ("OieInduction.py:146-149", "OieInduction.py:151-155", "Optimizers.py", "Decoder.py:84-93"):
        assert cite in txt


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = L.RaeConfig()
    cfg.abi_version = L.RAE_ABI_VERSION
    cfg.model, cfg.K, cfg.d, cfg.S, cfg.B, cfg.F, cfg.N = L.RAE_MODEL_AC, 4, 4, 2, 8, 16, 16
    cfg.lr, cfg.alpha, cfg.adj = 0.1, 1.0, 1.0
    h = C.c_void_p(0)
    rc = lib.rae_create(C.byref(cfg), C.byref(h))
    assert rc == L.RAE_ENODEVICE and not h.value
    assert b"no CPU fallback" in lib.rae_last_error(None)
    from relation_autoencoder_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine("rescal+sp", 4, 4, 2, 8, 16, 16, 8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "relation_autoencoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f

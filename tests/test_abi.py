"""The C-ABI shared library loads and exports every symbol include/bf_cuda.h declares; without a
GPU every compute entry point fails loudly (there is no CPU fallback in the product path)."""
import ctypes as C
import os
import re

import pytest

import better_flow_b200 as bf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "bf_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = bf.load()
    names = header_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding knows about all of them
    assert sorted(bf.ABI_SYMBOLS) == names


def test_struct_layouts():
    assert C.sizeof(bf.Model) == 88            # ObjectModel's 11 scalars (object_model.h:10-13)
    assert C.sizeof(bf.SliceResult) == 160
    assert bf.EVENT_DTYPE.itemsize == 8


def test_version_and_error_strings():
    lib = bf.load()
    assert b"sm_100a" in lib.bf_version()
    assert isinstance(lib.bf_last_error(), bytes)


def test_no_cpu_fallback():
    lib = bf.load()
    if lib.bf_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bf.BfError) as e:
        bf.Context(180, 240, 3, 1 << 16, 4)
    assert "no CPU fallback" in str(e.value)
    assert lib.bf_cuda_init(0) < 0


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """The product (package + C sources + bench's own arm) must not import or link oracle/."""
    offenders = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "better_flow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|libbf_oracle|libbf_ref", txt, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders

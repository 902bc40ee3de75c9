"""The C-ABI library loads on a CPU-only box, exports every symbol include/sdrgpu.h declares, and its compute
entry points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdrgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdrgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from sdrtrunk_b200 import native
    lib = ctypes.CDLL(native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_prototypes_cover_header():
    from sdrtrunk_b200 import native
    assert sorted(native.PROTOTYPES) == declared_symbols()


def test_no_cpu_fallback_without_gpu():
    from sdrtrunk_b200 import native
    n = ctypes.c_int(-1)
    status = native.lib().sdrgpu_device_count(ctypes.byref(n))
    if status == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(native.CudaError):
        native.init(0)
    from sdrtrunk_b200.dsp import ComplexPolyphaseChannelizerM2
    with pytest.raises(native.CudaError):
        ComplexPolyphaseChannelizerM2(np.ones(18, np.float32), 50000, 2)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sdrtrunk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M) or "liboracle" in src \
                        or re.search(r'#include\s+"[^"]*oracle', src):
                    bad.append(f)
    assert not bad, bad


def test_pack_dibits_host_helper():
    from sdrtrunk_b200 import native
    import oracle
    d = np.random.default_rng(0).integers(0, 4, 41).astype(np.uint8)
    out = np.zeros(16, np.uint8)
    n = native.lib().sdrgpu_pack_dibits(d.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), d.size,
                                        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    assert n == 10
    assert np.array_equal(out[:n], oracle.pack_dibits(d))
    assert out[0] == (d[0] << 6 | d[1] << 4 | d[2] << 2 | d[3])


def test_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/sdrgpu.h must compile as C99 (what jextract / cgo / ctypesgen read) and
    as C++, with no torch / CUDA types in it"""
    import shutil
    import subprocess
    header = os.path.join(ROOT, "include", "sdrgpu.h")
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)      # declarations only, comments stripped
    assert "torch" not in text and "cudaStream_t" not in text and "#include <cuda" not in text
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", header])
    subprocess.check_call([gcc, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", header])


def test_null_handles_are_refused_not_dereferenced():
    """argument checks run before anything touches the device: every pipeline entry point answers INVALID_ARG to a NULL
    handle (the Java shim maps it to IllegalArgumentException), also without a GPU"""
    from sdrtrunk_b200 import native
    L = native.lib()
    invalid = 1   # SDRGPU_ERR_INVALID_ARG
    assert L.sdrgpu_pipeline_wait(None) == invalid
    assert L.sdrgpu_pipeline_submit_multi(None, None, 0, None, 0, None) == invalid
    assert L.sdrgpu_pipeline_process_multi(None, None, 0, native.HOST, None, 0, None, 0, None, native.HOST) == invalid
    assert L.sdrgpu_pipeline_set_chunks(None, 4) == invalid
    assert L.sdrgpu_pipeline_set_device_chunks(None, 4) == invalid
    assert L.sdrgpu_pipeline_wait(None) == invalid
    assert b"NULL" in L.sdrgpu_last_error()

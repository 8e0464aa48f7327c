"""The C-ABI library loads on a CPU-only box and exports every symbol include/cco_b200.h declares; without a GPU
every compute entry fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cco_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cco_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    from universal_recommender_b200 import _native
    assert declared_symbols() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    from universal_recommender_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/cco_b200.h but not exported"
    assert lib.cco_abi_version() == 2


def test_status_strings():
    from universal_recommender_b200 import _native
    L = _native.lib()
    assert L.cco_status_string(0) == b"ok"
    assert L.cco_status_string(_native.E_SHAPE_MISMATCH) == b"shape mismatch"


def test_no_cpu_fallback_without_gpu():
    """On the CPU box cco_create must fail with CCO_E_CUDA -- the product path never routes through a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only box")
    import universal_recommender_b200 as ur
    with pytest.raises(ur.CcoError) as e:
        ur.CcoContext(device=0)
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/.  Python files of the package must not mention
    it at all; CUDA sources may name the restatement they are checked against in comments, nothing else."""
    pkg = os.path.join(ROOT, "universal_recommender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            txt = open(os.path.join(dirpath, f), errors="replace").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
            elif txt:
                for line in txt.splitlines():
                    if "oracle" in line.lower():
                        assert line.lstrip().startswith(("//", "*", "/*")) and "#include" not in line, f"{f}: {line.strip()}"


def test_plain_c_program_links_against_the_abi(tmp_path):
    """gcc (not nvcc, not g++) compiles a C consumer of the header and links the shared library."""
    import subprocess
    from universal_recommender_b200 import _native
    exe = tmp_path / "c_abi_check"
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi", "c_abi_check.c"), "-o", str(exe), "-L", libdir, "-lcco_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)


def _build_c_consumer(tmp_path):
    import subprocess
    from universal_recommender_b200 import _native
    exe = tmp_path / "c_abi_check"
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi", "c_abi_check.c"), "-o", str(exe), "-L", libdir, "-lcco_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.gpu
def test_plain_c_program_trains_on_the_b200(tmp_path):
    """the non-Python client of the boundary, on the GPU box: its B200 branch (one train through the C ABI) must run"""
    import subprocess
    p = subprocess.run([str(_build_c_consumer(tmp_path))], capture_output=True, text=True)
    assert p.returncode == 0 and "gpu ok" in p.stdout, (p.returncode, p.stdout, p.stderr)


def test_jni_shim_compiles():
    """jni/cco_jni.c is shipped as source (no JDK here): type-check it against the C ABI header and the minimal JNI
    declarations of tests/abi/jni_min/jni.h, so that the shim cannot rot into pseudo-code."""
    import subprocess
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "abi", "jni_min"),
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "cco_jni.c")], check=True)

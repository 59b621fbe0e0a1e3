"""CPU: the C-ABI library builds, loads and exports every symbol include/sfm_b200.h declares.
No compute calls: without a GPU sfm_create must fail loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from sfm_opencv_b200 import build
    build.build()
    import sfm_opencv_b200 as sfm
    return sfm.load()


def test_header_symbols_are_exported(lib):
    import sfm_opencv_b200 as sfm
    hdr = open(os.path.join(ROOT, "include", "sfm_b200.h")).read()
    declared = set(re.findall(r"SFM_API\s+[\w\s\*]+?\b(sfm_\w+)\s*\(", hdr))
    assert len(declared) >= 17
    assert declared == set(sfm.SYMBOLS), declared ^ set(sfm.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_strerror(lib):
    assert lib.sfm_abi_version() == 2
    assert lib.sfm_strerror(-7) == b"a pair has fewer than 2 train descriptors"


def test_no_cpu_fallback_without_gpu():
    import torch
    import sfm_opencv_b200 as sfm
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sfm.SfmError) as e:
        sfm.Context(0)
    assert e.value.code == -2


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sfm_opencv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "oracle/" not in src, f


@pytest.mark.parametrize("compiler,std", [("gcc", "-std=c99"), ("g++", "-std=c++11")])
def test_header_compiles_and_links_from_c_and_cpp(lib, tmp_path, compiler, std):
    """The integration route of INTEGRATION.md: #include "sfm_b200.h" + link libsfm_b200.so,
    from plain C and from C++ (the reference's language), no CUDA headers involved."""
    import shutil
    import subprocess
    if shutil.which(compiler) is None:
        pytest.skip(f"{compiler} not installed")
    src = os.path.join(ROOT, "examples", "link_check.c")
    if compiler == "g++":                                   # same source, compiled as C++
        cpp = tmp_path / "link_check.cpp"
        cpp.write_text(open(src).read())
        src = str(cpp)
    exe = str(tmp_path / "link_check")
    libdir = os.path.join(ROOT, "sfm_opencv_b200")
    subprocess.run([compiler, std, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-L", libdir, "-l:libsfm_b200.so", "-Wl,-rpath," + libdir], check=True)
    hdr = open(os.path.join(ROOT, "include", "sfm_b200.h")).read()
    declared = set(re.findall(r"SFM_API\s+[\w\s\*]+?\b(sfm_\w+)\s*\(", hdr))
    used = set(re.findall(r"\(const void\*\)(sfm_\w+)", open(os.path.join(ROOT, "examples", "link_check.c")).read()))
    assert used == declared, used ^ declared               # the example covers the whole header
    yml = tmp_path / "s.yml"
    out = subprocess.run([exe, str(yml)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "entry points" in out.stdout
    assert yml.read_text().startswith("%YAML:1.0\n---\nCamera Count: 1\nPoint Count: 2\n")

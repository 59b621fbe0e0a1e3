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
    assert lib.sfm_abi_version() == 1
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

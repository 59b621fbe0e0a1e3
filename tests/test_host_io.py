"""CPU: save_structure() / write_ply_binary() (NViewReconstuct.cpp:186-227, :229-294) through
the C ABI.  The golden vectors are the reference's OWN bundled outputs (Viewer/structure.yml,
structure_ba.yml, structure_ba.ply, structure_ba_crazyhorse.ply), parsed by
tests/golden/make_golden_io.py: rewriting them must give the same bytes (sha256)."""
import hashlib
import os

import numpy as np
import pytest

import sfm_opencv_b200 as sfm

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "viewer_outputs.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("name", ["structure.yml", "structure_ba.yml"])
def test_structure_yml_is_byte_identical(gold, tmp_path, name):
    k = name.replace(".", "_")
    out = tmp_path / name
    sfm.save_structure(out, gold[k + "_R"], gold[k + "_T"], gold[k + "_X"], gold[k + "_c"])
    assert os.path.getsize(out) == int(gold[k + "_bytes"])
    assert _sha(out) == str(gold[k + "_sha256"])
    ref = os.path.join("/root/reference/Viewer", name)
    if os.path.exists(ref):                       # in the build container: compare the bytes too
        assert open(out, "rb").read() == open(ref, "rb").read()


@pytest.mark.parametrize("name", ["structure_ba.ply", "structure_ba_crazyhorse.ply"])
def test_ply_is_byte_identical(gold, tmp_path, name):
    k = name.replace(".", "_")
    v = gold[k + "_v"]
    out = tmp_path / name
    sfm.write_ply_binary(out, v[:, :3], v[:, 3:], gold[k + "_c"], crlf=bool(gold[k + "_crlf"]))
    assert os.path.getsize(out) == int(gold[k + "_bytes"])
    assert _sha(out) == str(gold[k + "_sha256"])


def test_ply_skips_nan_vertices(tmp_path):
    xyz = np.arange(12, dtype=np.float32).reshape(4, 3)
    nrm = np.ones((4, 3), np.float32)
    xyz[1, 2] = np.nan
    nrm[3, 0] = np.nan
    rgb = np.arange(12, dtype=np.uint8).reshape(4, 3)
    out = tmp_path / "a.ply"
    sfm.write_ply_binary(out, xyz, nrm, rgb, crlf=False)
    raw = open(out, "rb").read()
    head, body = raw.split(b"end_header\n")
    assert b"element vertex 2\n" in head and len(body) == 2 * 27
    rec = np.frombuffer(body, np.dtype([("v", "<f4", 6), ("c", "u1", 3)]))
    assert np.array_equal(rec["v"][:, :3], xyz[[0, 2]]) and np.array_equal(rec["c"], rgb[[0, 2]])


def test_yaml_number_and_wrap_rules(tmp_path):
    """Against cv::FileStorage itself (the image's cv2 4.13) on everything whose text does not
    depend on the OpenCV version: structure, indentation, flow wrapping, integer-valued and
    non-finite doubles, empty collections.  (Non-integer doubles are printed with "%.16e" by
    the reference's OpenCV 4.4 and with a shortest-round-trip format by 4.13; that format is
    pinned by the reference's own bundled files above.)"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    R = [np.eye(3), np.rint(rng.normal(size=(3, 3)) * 1e6), -np.rint(rng.normal(size=(3, 3)) * 1e8)]
    T = [np.zeros((3, 1)), np.array([[np.inf], [-np.inf], [np.nan]]), np.array([[-0.0], [7.0], [-2147483647.0]])]
    X = np.clip(np.rint(rng.normal(size=(52, 3)) * 10 ** rng.integers(0, 9, (52, 3))), -2e9, 2e9)
    c = rng.integers(0, 256, (52, 3), dtype=np.uint8)
    for pts, cols in ((X, c), (X[:0], c[:0])):
        ours, theirs = str(tmp_path / "o.yml"), str(tmp_path / "t.yml")
        sfm.save_structure(ours, R, T, pts, cols)
        fs = cv2.FileStorage(theirs, cv2.FILE_STORAGE_WRITE)
        fs.write("Camera Count", len(R))
        fs.write("Point Count", int(pts.shape[0]))
        for key, mats in (("Rotations", R), ("Motions", T)):
            fs.startWriteStruct(key, cv2.FILE_NODE_SEQ)
            for m in mats:
                fs.write("", np.ascontiguousarray(m, np.float64))
            fs.endWriteStruct()
        fs.startWriteStruct("Points", cv2.FILE_NODE_SEQ)
        for p in pts:
            fs.startWriteStruct("", cv2.FILE_NODE_SEQ | cv2.FILE_NODE_FLOW)
            for v in p:
                fs.write("", float(v))
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.startWriteStruct("Colors", cv2.FILE_NODE_SEQ)
        for p in cols:
            fs.startWriteStruct("", cv2.FILE_NODE_SEQ | cv2.FILE_NODE_FLOW)
            for v in p:
                fs.write("", int(v))
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.release()
        assert open(ours, "rb").read() == open(theirs, "rb").read()


def test_bad_arguments():
    with pytest.raises(sfm.SfmError):
        sfm.save_structure("/nonexistent-dir/x.yml", [np.eye(3)], [np.zeros(3)], np.zeros((1, 3)),
                           np.zeros((1, 3), np.uint8))
    with pytest.raises(sfm.SfmError):
        sfm.save_structure("/tmp/x.yml", [np.eye(3)], [], np.zeros((1, 3)), np.zeros((1, 3), np.uint8))

"""Parses the reference's bundled OUTPUT files (Viewer/structure.yml, structure_ba.yml,
structure_ba.ply, structure_ba_crazyhorse.ply) into arrays + the sha256 of the original
bytes.  tests/test_host_io.py rewrites them with the library's writers and compares hashes:
the writers must reproduce the reference's own artefacts byte for byte.

Run HERE (container with /root/reference and cv2):  python tests/golden/make_golden_io.py
"""
import hashlib
import os

import cv2
import numpy as np

REF = "/root/reference/Viewer"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_yml(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    n = int(fs.getNode("Camera Count").real())
    R = np.stack([fs.getNode("Rotations").at(i).mat() for i in range(n)])
    T = np.stack([fs.getNode("Motions").at(i).mat() for i in range(n)])
    P, C = fs.getNode("Points"), fs.getNode("Colors")
    X = np.array([[P.at(i).at(k).real() for k in range(3)] for i in range(P.size())], np.float64)
    c = np.array([[int(C.at(i).at(k).real()) for k in range(3)] for i in range(C.size())], np.uint8)
    assert int(fs.getNode("Point Count").real()) == X.shape[0]
    return R, T, X, c


def parse_ply(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header") + len(b"end_header")
    crlf = raw[end:end + 2] == b"\r\n"
    body = raw[end + (2 if crlf else 1):]
    n = int([l for l in raw[:end].split(b"\n") if l.startswith(b"element vertex")][0].split()[2])
    rec = np.frombuffer(body, np.dtype([("v", "<f4", 6), ("c", "u1", 3)]), n)
    assert len(body) == 27 * n
    return rec["v"].copy(), rec["c"].copy(), crlf


def main():
    out = {}
    for name in ("structure.yml", "structure_ba.yml"):
        R, T, X, c = parse_yml(os.path.join(REF, name))
        key = name.replace(".", "_")
        out[key + "_R"], out[key + "_T"], out[key + "_X"], out[key + "_c"] = R, T, X, c
        raw = open(os.path.join(REF, name), "rb").read()
        out[key + "_sha256"] = np.array(hashlib.sha256(raw).hexdigest())
        out[key + "_bytes"] = np.int64(len(raw))
        print(name, R.shape, X.shape, c.shape, len(raw))
    for name in ("structure_ba.ply", "structure_ba_crazyhorse.ply"):
        v, c, crlf = parse_ply(os.path.join(REF, name))
        key = name.replace(".", "_")
        out[key + "_v"], out[key + "_c"], out[key + "_crlf"] = v, c, np.bool_(crlf)
        raw = open(os.path.join(REF, name), "rb").read()
        out[key + "_sha256"] = np.array(hashlib.sha256(raw).hexdigest())
        out[key + "_bytes"] = np.int64(len(raw))
        print(name, v.shape, crlf, len(raw))
    np.savez_compressed(os.path.join(OUT, "viewer_outputs.npz"), **out)


if __name__ == "__main__":
    main()

"""CPU: host-side pieces of the Python mirror that need no device."""
import numpy as np
import pytest

from sfm_opencv_b200 import api


def test_pair_lists_behave_like_a_list_of_arrays():
    flat = np.arange(10)
    off = np.array([0, 3, 3, 7, 10])
    pl = api.PairLists(flat, off)
    assert len(pl) == 4
    assert [x.tolist() for x in pl] == [[0, 1, 2], [], [3, 4, 5, 6], [7, 8, 9]]
    assert pl[-1].tolist() == [7, 8, 9] and pl[1].size == 0
    assert [x.tolist() for x in pl[1:3]] == [[], [3, 4, 5, 6]]
    acc = []
    acc += pl                                           # used by the sharded gather
    assert len(acc) == 4 and acc[2].base is not None    # views, not copies
    with pytest.raises(IndexError):
        pl[4]


def test_build_projection_matches_cv_gemm_bitwise():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    K = np.array([[2759.48, 0, 1520.69], [0, 2764.16, 1006.81], [0, 0, 1]])
    for _ in range(100):
        R, _ = cv2.Rodrigues(rng.normal(0, 0.4, 3))
        T = rng.normal(0, 2, 3)
        RT = np.concatenate([R.astype(np.float32), T.astype(np.float32).reshape(3, 1)], 1)
        assert np.array_equal(api.build_projection(K, R, T), cv2.gemm(K.astype(np.float32), RT, 1.0, None, 0.0))


def test_enumerate_observations_is_camera_major():
    ids = [np.array([-1, 4, 2]), np.array([0, -1]), np.array([], int)]
    kps = [np.arange(6, dtype=np.float32).reshape(3, 2), np.arange(4, dtype=np.float32).reshape(2, 2) + 10,
           np.zeros((0, 2), np.float32)]
    cam, pt, obs = api.enumerate_observations(ids, kps)
    assert cam.tolist() == [0, 0, 1] and pt.tolist() == [4, 2, 0]
    assert obs.tolist() == [[2, 3], [4, 5], [10, 11]]


def test_cpp_enumerate_observations_equals_the_python_mirror(tmp_path):
    """include/sfm_b200.hpp::enumerate_observations (what a C++ caller of the reference's
    bundle_adjustment, NViewReconstuct.cpp:1187-1211, uses) against api.enumerate_observations."""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not installed")
    from sfm_opencv_b200 import build
    build.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "sfm_opencv_b200")
    exe = str(tmp_path / "enumerate_check")
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "enumerate_check.cpp"), "-o", exe,
                    "-L", libdir, "-l:libsfm_b200.so", "-Wl,-rpath," + libdir], check=True)
    rng = np.random.default_rng(4)
    ids = [rng.integers(-1, 50, n) for n in (0, 7, 300, 1, 64)]
    ids[3][0] = -1                                           # an image without any structure point
    kps = [rng.uniform(0, 3000, (len(i), 2)).astype(np.float32) for i in ids]
    text = [str(len(ids))]
    for i, k in zip(ids, kps):
        text.append(str(len(i)))
        text += [f"{a} {float(x):.9g} {float(y):.9g}" for a, (x, y) in zip(i, k)]
    out = subprocess.run([exe], input="\n".join(text) + "\n", capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    got = np.array([l.split() for l in out.stdout.splitlines()], np.float64).reshape(-1, 4)
    cam, pt, obs = api.enumerate_observations(ids, kps)
    assert np.array_equal(got[:, 0], cam) and np.array_equal(got[:, 1], pt)
    assert np.array_equal(got[:, 2:].astype(np.float32), obs)
    bad = subprocess.run([exe], input="1\n0\n".replace("1\n0", "2\n0"), capture_output=True, text=True, timeout=60)
    assert bad.returncode == 2                               # truncated input is refused, not guessed

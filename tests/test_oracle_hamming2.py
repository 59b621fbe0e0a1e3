"""CPU: the NORM_HAMMING2 restatement (the LIVE reference path, NViewReconstuct.cpp:797,876)
against the reference's own library call (cv2.batchDistance / BFMatcher) and the AKAZE fixture
generated from the bundled dataset/desktop images (tests/golden/make_golden.py)."""
import numpy as np

from oracle import matching as M

# match counts of the four consecutive desktop pairs that reproduce the reference's bundled
# output Viewer/structure.yml (3190 fused points)
STRUCTURE_YML_MATCH_COUNTS = [2186, 1063, 230, 553]


def _rand_bin(n, width, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, width), dtype=np.uint8)


def test_popcount_table_is_two_bit_cells():
    for x in (0b00000000, 0b00000001, 0b00000010, 0b00000011, 0b01010101, 0b10101010, 0xFF, 0x90):
        cells = sum(1 for c in range(4) if (x >> (2 * c)) & 3)
        assert M._POP2[x] == cells


def test_restatement_equals_cv2_random():
    q, t = _rand_bin(300, 61, 1), _rand_bin(777, 61, 2)
    d, i = M.knn2_hamming2_int(q, t)
    dc, ic = M.knn2_cv_hamming2(q, t)
    assert np.array_equal(i, ic) and np.array_equal(d, dc)


def test_restatement_equals_bfmatcher():
    import cv2
    q, t = _rand_bin(40, 61, 3), _rand_bin(90, 61, 4)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING2).knnMatch(q, t, 2)
    d, i = M.knn2_hamming2_int(q, t)
    assert [[m[0].trainIdx, m[1].trainIdx] for m in knn] == i.tolist()
    assert [[m[0].distance, m[1].distance] for m in knn] == d.tolist()


def test_ties_go_to_the_lower_index():
    t = _rand_bin(100, 61, 5)
    t[40] = t[3]; t[70] = t[3]
    q = t[3:4].copy()
    for knn in (M.knn2_hamming2_int, M.knn2_cv_hamming2):
        d, i = knn(q, t)
        assert i[0].tolist() == [3, 40] and d[0].tolist() == [0.0, 0.0]


def test_golden_akaze_fixture(golden):
    g = golden("desktop", "akaze")
    n = int(g["n_img"])
    assert [g[f"desc_{i}"].shape[0] for i in range(n)] == [13287, 12697, 8903, 1796, 1572]
    counts = []
    for p in range(n - 1):
        q, t = g[f"desc_{p}"], g[f"desc_{p + 1}"]
        sel = np.linspace(0, q.shape[0] - 1, 400).astype(int)     # a few seconds of numpy
        d, i = M.knn2_hamming2_int(q[sel], t)
        assert np.array_equal(i, g[f"knn_idx_{p}"][sel])
        assert np.array_equal(d, g[f"knn_dist_{p}"][sel])
        m, md0, md = M.filter_matches(g[f"knn_dist_{p}"], g[f"knn_idx_{p}"])
        assert np.array_equal(m, g[f"match_{p}"]) and md == g[f"min_dist_{p}"]
        counts.append(len(m))
    assert counts == STRUCTURE_YML_MATCH_COUNTS

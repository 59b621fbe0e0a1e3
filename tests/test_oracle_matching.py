"""CPU: the integer restatement of knnMatch + filter against the reference's own library
(cv2 BFMatcher / batchDistance) and the committed golden fixtures."""
import numpy as np
import pytest

from oracle import matching as M
from oracle import synth


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_int_oracle_equals_cv2_random():
    b = synth.image_bank(2, 700, seed0=3)
    d, i = M.knn2_int(b[0][:500], b[1])
    dc, ic = M.knn2_cv(b[0][:500], b[1])
    assert np.array_equal(i, ic)
    assert np.array_equal(_bits(d), _bits(dc))


def test_bfmatcher_equals_batchdistance():
    b = synth.image_bank(2, 300, seed0=5)
    d, i = M.knn2_cv(b[0], b[1])
    db, ib = M.knn2_bfmatcher(b[0], b[1])
    assert np.array_equal(i, ib) and np.array_equal(_bits(d), _bits(db))


def test_tie_break_lower_index_first():
    t = synth.sift_like(64, 11)
    t[10] = t[0]
    t[20] = t[0]
    q = t[:1].copy()
    for knn in (M.knn2_int, M.knn2_cv):
        d, i = knn(q, t)
        assert list(i[0]) == [0, 10] and d[0, 0] == 0 and d[0, 1] == 0


def _collision_rows():
    """Two train rows whose exact d2 to the zero query differ by 1 but share one float sqrt
    (possible only for d2 >= 2^22), plus a far row."""
    for a in range(1, 200):
        t = np.zeros((3, 128), np.uint8)
        t[0, :100] = 250; t[0, 100] = a
        t[1] = t[0]; t[1, 101] = 1
        t[2] = 255
        d2 = M.sq_dist_int(np.zeros((1, 128), np.uint8), t)[0]
        if np.sqrt(np.float32(d2[0])) == np.sqrt(np.float32(d2[1])):
            assert d2[1] == d2[0] + 1 and d2[0] >= 1 << 22
            return t
    raise AssertionError("no collision found")


def test_sqrt_collision_resolved_by_index():
    # beyond 2^22 float sqrt collides: OpenCV orders such rows by index, not by d2
    q = np.zeros((1, 128), np.uint8)
    t = _collision_rows()[[1, 0, 2]]                   # the larger d2 first
    d, i = M.knn2_int(q, t)
    dc, ic = M.knn2_cv(q, t)
    assert list(i[0]) == [0, 1] and np.array_equal(i, ic) and np.array_equal(_bits(d), _bits(dc))
    assert d[0, 0] == d[0, 1]


def test_filter_semantics():
    dist = np.array([[6.0, 10.0], [6.1, 10.0], [3.0, 100.0], [70.0, 200.0]], np.float32)
    idx = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.int32)
    m, d, md = M.filter_matches(dist, idx)
    # row 0: 6.0 > 0.6*10.0 (=6.000000000000001 in double? 0.6*10 = 6.0) -> not greater: passes
    assert md == np.float32(3.0)
    # gate = 5*max(3,10) = 50: row 3 (70) rejected although its ratio passes
    assert [tuple(r) for r in m] == [(0, 1), (2, 5)]
    assert list(d) == [6.0, 3.0]


def test_filter_no_pass_gives_flt_max():
    dist = np.array([[9.0, 10.0]], np.float32)
    m, d, md = M.filter_matches(dist, np.array([[0, 1]], np.int32))
    assert len(m) == 0 and md == np.finfo(np.float32).max


def test_too_few_train_raises():
    with pytest.raises(ValueError):
        M.knn2_int(synth.sift_like(4, 0), synth.sift_like(1, 1))


@pytest.mark.parametrize("name,expect", [("crazyhorse", [313, 507, 545, 386, 652, 389]),
                                         ("desktop", [871, 366, 142, 209])])
def test_golden_fixture_pins_oracle(golden, name, expect):
    """Fixtures were produced by cv2.batchDistance on the bundled datasets
    (tests/golden/make_golden.py); expected match counts are SURVEY.md section 4 item 3."""
    g = golden(name)
    n = int(g["n_img"])
    assert [len(g[f"match_{i}"]) for i in range(n - 1)] == expect
    pairs = range(n - 1) if name == "crazyhorse" else (2, 3)     # keep the CPU suite short
    for i in pairs:
        d, idx = M.knn2_int(g[f"desc_{i}"], g[f"desc_{i + 1}"])
        assert np.array_equal(idx, g[f"knn_idx_{i}"])
        assert np.array_equal(_bits(d), _bits(g[f"knn_dist_{i}"]))
        m, d0, md = M.filter_matches(d, idx)
        assert np.array_equal(m, g[f"match_{i}"]) and md == g[f"min_dist_{i}"]
    if name == "desktop":
        assert abs(float(g["min_dist_0"]) - 18.330) < 1e-3


def knn_sha(dist, idx) -> bytes:
    """The digest tests/golden/make_golden.py::knn_digest stores for large pairs."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(idx, np.int32).tobytes())
    h.update(np.ascontiguousarray(dist, np.float32).tobytes())
    return h.digest()


def test_dog_fixture_pins_oracle(golden):
    """dataset/dog (16 images; the largest bundled descriptor sets): the fixture holds the output of
    the reference's library call for all 15 consecutive pairs as match lists + a digest of every
    kNN row; the integer oracle reproduces the four cheapest pairs here (the GPU suite does all)."""
    g = golden("dog")
    sizes = [g[f"desc_{i}"].shape[0] for i in range(16)]
    assert sizes == [17176, 16794, 19092, 14916, 4596, 21408, 17623, 21531, 14034, 21813, 15312,
                     21436, 8644, 5112, 14509, 8688]
    for i in (3, 4, 12, 13):
        d, idx = M.knn2_int(g[f"desc_{i}"], g[f"desc_{i + 1}"])
        assert knn_sha(d, idx) == g[f"knn_sha_{i}"].tobytes()
        m, d0, md = M.filter_matches(d, idx)
        assert np.array_equal(m, g[f"match_{i}"]) and md == g[f"min_dist_{i}"]
        assert np.array_equal(_bits(d0), _bits(g[f"match_dist_{i}"]))


def test_all_pairs_fixture_pins_oracle(golden):
    """Exhaustive (i < j) schedule on dataset/crazyhorse: 21 pairs from the reference's library
    call; the integer oracle agrees on every one of them."""
    g = golden("crazyhorse")
    ap = golden("crazyhorse", "allpairs")
    assert int(ap["n_pairs"]) == 21
    for p in range(21):
        a, b = ap[f"pair_{p}"]
        d, idx = M.knn2_int(g[f"desc_{a}"], g[f"desc_{b}"])
        assert knn_sha(d, idx) == ap[f"knn_sha_{p}"].tobytes()
        m, _, md = M.filter_matches(d, idx)
        assert np.array_equal(m, ap[f"match_{p}"]) and md == ap[f"min_dist_{p}"]


# ---- property tests: integer restatement == the library call -----
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@st.composite
def descriptor_pair(draw, max_q=40, max_t=70):
    """(query, train) uint8 sets with planted duplicates (exact ties on d0 == d1), rows of 255s,
    zero rows and near-duplicates -- the cases the tie-break and the packed keys must survive."""
    nq = draw(st.integers(1, max_q))
    nt = draw(st.integers(2, max_t))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    kind = draw(st.sampled_from(["sift", "sparse", "tiny"]))
    if kind == "sift":
        q, t = synth.sift_like(nq, seed % 9973), synth.sift_like(nt, seed % 9973 + 1)
    elif kind == "sparse":
        q = (rng.random((nq, 128)) < 0.1).astype(np.uint8) * rng.integers(0, 120, (nq, 128)).astype(np.uint8)
        t = (rng.random((nt, 128)) < 0.1).astype(np.uint8) * rng.integers(0, 120, (nt, 128)).astype(np.uint8)
    else:
        q = rng.integers(0, 3, (nq, 128)).astype(np.uint8)
        t = rng.integers(0, 3, (nt, 128)).astype(np.uint8)
    for _ in range(draw(st.integers(0, 6))):          # duplicate train rows / copy query rows in
        a, b = rng.integers(0, nt, 2)
        t[a] = t[b]
    for _ in range(draw(st.integers(0, 4))):
        t[rng.integers(0, nt)] = q[rng.integers(0, nq)]
    if draw(st.booleans()):
        t[rng.integers(0, nt)] = 0
    if draw(st.booleans()):
        # a saturated row: 64 x 255 keeps |row|^2 = 4 161 600 > 2^21 out, 30 x 255 stays legal
        r = np.zeros(128, np.uint8)
        r[rng.choice(128, 30, replace=False)] = 255
        t[rng.integers(0, nt)] = r
        q[rng.integers(0, nq)] = r
    return q, t


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(descriptor_pair())
def test_property_int_oracle_equals_library(qt):
    q, t = qt
    d, i = M.knn2_int(q, t)
    dc, ic = M.knn2_cv(q, t)
    assert np.array_equal(i, ic) and np.array_equal(_bits(d), _bits(dc))
    # order = (distance, lower index): never a later equal-distance row in front
    assert ((d[:, 0] < d[:, 1]) | ((d[:, 0] == d[:, 1]) & (i[:, 0] < i[:, 1]))).all()

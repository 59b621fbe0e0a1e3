"""GPU parity on every bundled dataset (north_star: bit-exact kNN-ratio matches "on all bundled
datasets") and property tests of the kNN-2 kernel, through the C ABI."""
import numpy as np
import pytest

from oracle import matching as M
from ratio_sweep_model import undecidable_case
from test_oracle_matching import descriptor_pair, knn_sha

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _knn_arrays(k):
    idx = np.stack([k["trainIdx0"], k["trainIdx1"]], 1)
    dist = np.stack([k["distance0"], k["distance1"]], 1)
    return dist, idx


def _same_lists(a, b):
    return a.flat.tobytes() == b.flat.tobytes() and np.array_equal(a.offsets, b.offsets)


def test_dog_consecutive_pairs(ctx, golden):
    """dataset/dog: 16 images of 4.6k-21.8k SIFT descriptors, 15 consecutive pairs
    (match_features_for_all, NViewReconstuct.cpp:857-870): every kNN row (digest) and every match
    list equal what the reference's library call produced (tests/golden/make_golden.py --dog)."""
    g = golden("dog")
    n = int(g["n_img"])
    ctx.upload_descriptors([g[f"desc_{i}"] for i in range(n)])
    m, md, knn = ctx.match_pairs(M.consecutive_pairs(n), want_knn=True)
    for i in range(n - 1):
        assert knn_sha(*_knn_arrays(knn[i])) == g[f"knn_sha_{i}"].tobytes(), f"pair {i}"
        assert np.array_equal(m[i]["queryIdx"], g[f"match_{i}"][:, 0])
        assert np.array_equal(m[i]["trainIdx"], g[f"match_{i}"][:, 1])
        assert np.array_equal(_bits(m[i]["distance"]), _bits(g[f"match_dist_{i}"]))
        assert _bits(md[i]) == _bits(g[f"min_dist_{i}"])
    # match lists only (no raw kNN rows requested): the kernel stops tracking the second
    # neighbour of rows that cannot pass the ratio test any more -- same lists, same min_dist
    m2, md2, _ = ctx.match_pairs(M.consecutive_pairs(n))
    assert _same_lists(m, m2) and md.tobytes() == md2.tobytes()


@pytest.mark.parametrize("name", ["crazyhorse", "desktop"])
def test_real_datasets_all_pairs(ctx, golden, name):
    """The exhaustive i < j schedule (BASELINE config 3's) on real descriptor sets of very
    different sizes (desktop: 16717 ... 838) in ONE call."""
    g = golden(name)
    ap = golden(name, "allpairs")
    n = int(g["n_img"])
    ctx.upload_descriptors([g[f"desc_{i}"] for i in range(n)])
    pairs = [tuple(int(x) for x in ap[f"pair_{p}"]) for p in range(int(ap["n_pairs"]))]
    assert pairs == M.all_pairs(n)
    m, md, knn = ctx.match_pairs(pairs, want_knn=True)
    for p in range(len(pairs)):
        assert knn_sha(*_knn_arrays(knn[p])) == ap[f"knn_sha_{p}"].tobytes(), f"pair {pairs[p]}"
        assert np.array_equal(m[p]["queryIdx"], ap[f"match_{p}"][:, 0])
        assert np.array_equal(m[p]["trainIdx"], ap[f"match_{p}"][:, 1])
        assert np.array_equal(_bits(m[p]["distance"]), _bits(ap[f"match_dist_{p}"]))
        assert _bits(md[p]) == _bits(ap[f"min_dist_{p}"])
    for ratio in (0.6, 0.8, 0.95, 1.0, 0.3):
        a, amd, _ = ctx.match_pairs(pairs, ratio=ratio, want_knn=True)     # exact kNN rows
        b, bmd, _ = ctx.match_pairs(pairs, ratio=ratio)                    # match lists only
        assert _same_lists(a, b) and amd.tobytes() == bmd.tobytes(), ratio


# ---- property tests -------------------------------------------------------------------------
from hypothesis import HealthCheck, given, settings  # noqa: E402


@settings(max_examples=50, deadline=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(descriptor_pair(max_q=300, max_t=700))
def test_property_kernel_equals_oracle(ctx, qt):
    """Random sizes (not multiples of any tile), duplicate rows (ties on d0 == d1), zero rows,
    saturated 255 rows, sparse and tiny-valued descriptors."""
    q, t = qt
    ctx.upload_descriptors([q, t])
    m, md, knn = ctx.match_pairs([(0, 1)], want_knn=True)
    d, idx = M.knn2_int(q, t)
    dist, gi = _knn_arrays(knn[0])
    assert np.array_equal(gi, idx) and np.array_equal(_bits(dist), _bits(d))
    om, od, omd = M.filter_matches(d, idx)
    assert np.array_equal(m[0]["queryIdx"], om[:, 0]) and np.array_equal(m[0]["trainIdx"], om[:, 1])
    assert np.array_equal(_bits(m[0]["distance"]), _bits(od)) and _bits(md[0]) == _bits(omd)
    m2, md2, _ = ctx.match_pairs([(0, 1)])                   # match lists only
    assert _same_lists(m, m2) and md.tobytes() == md2.tobytes()


@settings(max_examples=20, deadline=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(descriptor_pair(max_q=200, max_t=2600))
def test_property_long_train_sets(ctx, qt):
    """Train sets longer than one packed-key window (1024 columns): window merge + joint bound."""
    q, t = qt
    ctx.upload_descriptors([q, t])
    _, _, knn = ctx.match_pairs([(0, 1)], want_knn=True)
    d, idx = M.knn2_int(q, t)
    dist, gi = _knn_arrays(knn[0])
    assert np.array_equal(gi, idx) and np.array_equal(_bits(dist), _bits(d))
    for ratio in (0.6, 0.9):
        a, amd, _ = ctx.match_pairs([(0, 1)], ratio=ratio, want_knn=True)
        b, bmd, _ = ctx.match_pairs([(0, 1)], ratio=ratio)
        assert _same_lists(a, b) and amd.tobytes() == bmd.tobytes()


def test_match_only_mode_on_the_headline_shape(ctx):
    """8192-row images with planted near-duplicates (BASELINE config 3's generator): the lists of
    the match-only sweep equal those of the exact sweep for every pair and several ratios."""
    from oracle import synth
    bank = synth.image_bank(5, 8192, seed0=31)
    ctx.upload_descriptors(bank)
    pairs = M.all_pairs(5)
    for ratio in (0.6, 0.75, 0.99):
        a, amd, _ = ctx.match_pairs(pairs, ratio=ratio, want_knn=True)
        b, bmd, _ = ctx.match_pairs(pairs, ratio=ratio)
        assert _same_lists(a, b) and amd.tobytes() == bmd.tobytes(), ratio
        assert len(a.flat) > 1000
    om, od, omd, _, _ = M.match_features(bank[0], bank[1], knn=M.knn2_cv)
    b, bmd, _ = ctx.match_pairs([(0, 1)])
    assert np.array_equal(b[0]["queryIdx"], om[:, 0]) and np.array_equal(b[0]["trainIdx"], om[:, 1])
    assert np.array_equal(_bits(b[0]["distance"]), _bits(od)) and _bits(bmd[0]) == _bits(omd)


def test_max_norm_rows_and_rejected_rows(ctx):
    """|row|^2 must stay below 2^21 (float sqrt injective on the distances): the largest legal
    rows match exactly, one more unit is SFM_E_RANGE -- an error, not a slow path."""
    import sfm_opencv_b200 as sfm
    from sfm_opencv_b200 import _capi
    rng = np.random.default_rng(8)
    t = np.zeros((300, 128), np.uint8)
    for r in range(300):
        cols = rng.choice(128, 32, replace=False)
        t[r, cols] = 255                                # 32 * 255^2 = 2 080 800 < 2^21 = 2 097 152
        t[r, rng.integers(0, 128)] = rng.integers(0, 100)
    q = t[::3].copy()
    ctx.upload_descriptors([q, t])
    m, md, knn = ctx.match_pairs([(0, 1)], want_knn=True)
    d, idx = M.knn2_int(q, t)
    dist, gi = _knn_arrays(knn[0])
    assert np.array_equal(gi, idx) and np.array_equal(_bits(dist), _bits(d))
    bad = t.copy()
    bad[5, :] = 0
    bad[5, :33] = 255                                   # 33 * 255^2 > 2^21
    with pytest.raises(sfm.SfmError) as e:
        ctx.upload_descriptors([q, bad])
    assert e.value.code == _capi.SFM_E_RANGE


def test_undecidable_rows_are_rechecked(ctx):
    """Rows the ratio-driven sweep cannot decide (tests/ratio_sweep_model.py: they would PASS with the
    second neighbour the sweep kept and FAIL with the true one) go through recheck_rows_kernel: the
    match-only call reports them and returns the exact lists."""
    q, t = undecidable_case(nq=300, nt=2000)
    q[7] = t[1500]                                         # one row with a true match
    ctx.upload_descriptors([q, t])
    b, bmd, _ = ctx.match_pairs([(0, 1)])
    # all but the true match, less the warp of the true match (its columns are looked at for all 32 rows)
    assert 300 - 32 <= ctx.last_rechecked_rows <= 299
    om, od, omd, _, _ = M.match_features(q, t)
    assert np.array_equal(b[0]["queryIdx"], om[:, 0]) and np.array_equal(b[0]["trainIdx"], om[:, 1])
    assert np.array_equal(_bits(b[0]["distance"]), _bits(od)) and _bits(bmd[0]) == _bits(omd)
    assert len(om) == 1
    a, amd, _ = ctx.match_pairs([(0, 1)], want_knn=True)   # the exact search never rechecks
    assert ctx.last_rechecked_rows == 0 and _same_lists(a, b)


def test_recheck_list_overflow_repeats_the_call_unpruned(ctx):
    """More undecidable rows than the recheck list holds (65 536, or an eighth of all query rows): the
    call is repeated with the plain match-only sweep; lists and min_dist are still exact."""
    q, t = undecidable_case(nq=17000, nt=1024)
    good = t.copy()
    good[300] = q[0]                                       # every query row has an exact copy here
    ctx.upload_descriptors([q, t, np.roll(t, 1, 0), np.roll(t, 2, 0), np.roll(t, 3, 0), good])
    pairs = [(0, j) for j in range(1, 6)]
    b, bmd, _ = ctx.match_pairs(pairs)
    assert ctx.last_rechecked_rows == 0                    # the sweep's verdicts were thrown away
    a, amd, _ = ctx.match_pairs(pairs, want_knn=True)
    assert _same_lists(a, b) and amd.tobytes() == bmd.tobytes()
    assert [len(m) for m in b] == [0, 0, 0, 0, 17000]
    om, od, omd, _, _ = M.match_features(q[:512], good)
    assert np.array_equal(b[4]["trainIdx"][:512], om[:, 1]) and _bits(bmd[4]) == _bits(omd)


def test_prune_mode_auto_small_calls_use_the_plain_sweep(monkeypatch):
    """Default SFM_PRUNE_MODE=2: a call below 2048 work items (256-row query blocks) runs the plain
    match-only sweep (nothing is rechecked), a larger one the ratio-driven sweep; the lists are the same
    as with the sweep forced off."""
    import sfm_opencv_b200 as sfm
    from oracle import synth
    q, t = undecidable_case(nq=300, nt=2000)
    bank = synth.image_bank(10, 12800, seed0=77)            # 45 pairs x 50 blocks = 2250 work items
    pairs = M.all_pairs(10)
    got = {}
    for mode in ("2", "0"):
        monkeypatch.setenv("SFM_PRUNE_MODE", mode)
        with sfm.Context(0) as c:
            c.upload_descriptors([q, t])
            small, smd, _ = c.match_pairs([(0, 1)])
            assert c.last_rechecked_rows == 0
            small = (small.flat.tobytes(), smd.tobytes())      # the lists live until the next call
            c.upload_descriptors(bank)
            big, bmd, _ = c.match_pairs(pairs)
            got[mode] = small + (big.flat.tobytes(), big.offsets.tobytes(), bmd.tobytes())
            assert len(big.flat) > 10000
    assert got["2"] == got["0"]
    om, od, omd, _, _ = M.match_features(bank[3], bank[4], knn=M.knn2_cv)
    monkeypatch.setenv("SFM_PRUNE_MODE", "2")
    with sfm.Context(0) as c:
        c.upload_descriptors(bank)
        b, bmd, _ = c.match_pairs(pairs)
        p = pairs.index((3, 4))
        assert np.array_equal(b[p]["trainIdx"], om[:, 1]) and _bits(bmd[p]) == _bits(omd)

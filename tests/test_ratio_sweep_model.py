"""CPU checks of the DECISION RULE of the ratio-driven match-only sweep (match_knn.cu, kPrune) through
its numpy model (tests/ratio_sweep_model.py): whatever the rule lets a row skip, the two filter passes
of match_features (NViewReconstuct.cpp:880-908) give the same list as on the exact kNN table
(oracle.matching).  The CUDA kernel itself is compared with the oracle in tests/test_gpu_*.py."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import matching as M
from oracle import synth
from ratio_sweep_model import ratio_sweep, undecidable_case


def _lists_equal(q, t, ratio, **kw):
    d2, idx, flagged, inserted = ratio_sweep(q, t, ratio=ratio, **kw)
    dist = np.sqrt(d2.astype(np.float32))
    got = M.filter_matches(dist, idx, ratio=ratio)
    want = M.match_features(q, t, ratio=ratio)[:3]
    ok = (np.array_equal(got[0], want[0]) and got[1].tobytes() == want[1].tobytes()
          and np.float32(got[2]).tobytes() == np.float32(want[2]).tobytes())
    return ok, flagged, inserted


@pytest.mark.parametrize("ratio", [0.3, 0.6, 0.8, 0.99, 1.0])
@pytest.mark.parametrize("warp_rows,partner", [(32, "fresh"), (1, "stale"), (1, "none"), (32, "mixed")])
@pytest.mark.parametrize("seed_group", [False, True])
def test_planted_bank_lists_equal(ratio, warp_rows, partner, seed_group):
    bank = synth.image_bank(2, 1500, seed0=5)
    ok, flagged, _ = _lists_equal(bank[0][:400], bank[1], ratio, warp_rows=warp_rows, partner=partner,
                                  seed_group=seed_group)
    assert ok
    assert flagged.mean() < 0.05          # the recheck is the exception, not the rule


def test_sweep_skips_most_groups():
    """The point of the rule: on unrelated images almost nothing is inserted after the cold tiles."""
    q, t = synth.sift_like(256, 1), synth.sift_like(4096, 2)
    ok, flagged, inserted = _lists_equal(q, t, 0.6, warp_rows=1)
    assert ok
    cold = 256 * 2 * 16                   # rows x cold tiles x groups per tile
    assert inserted - cold < 0.02 * (256 * 32 * 16 - cold)


def test_undecidable_rows_are_flagged_and_rechecked():
    q, t = undecidable_case()
    d, idx = M.knn2_int(q, t)
    assert (idx[:, 0] == 900).all() and (idx[:, 1] == 600).all()
    assert M.filter_matches(d, idx)[0].shape[0] == 0          # 120 > 0.6 * 160: every row fails
    ok, flagged, _ = _lists_equal(q, t, 0.6, warp_rows=1)
    assert ok and flagged.all()
    ok, flagged, _ = _lists_equal(q, t, 0.6, warp_rows=1, seed_group=True)
    assert ok and flagged.all()
    # without the recheck the rows would pass with a second neighbour that is too far
    d2, idx2, _, _ = ratio_sweep(q, t, warp_rows=1, recheck=False)
    assert M.filter_matches(np.sqrt(d2.astype(np.float32)), idx2)[0].shape[0] == len(q)


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10_000), nq=st.integers(1, 96), nt=st.integers(2, 700),
       ratio=st.sampled_from([0.5, 0.6, 0.75, 0.9]), dup=st.integers(0, 40), noise=st.integers(0, 12),
       warp_rows=st.sampled_from([1, 32]), partner=st.sampled_from(["fresh", "stale", "none", "mixed"]),
       cold=st.integers(0, 3))
def test_property_lists_equal(seed, nq, nt, ratio, dup, noise, warp_rows, partner, cold):
    """Random SIFT-like rows with planted noisy copies at several distances (near-threshold second
    neighbours included), exact duplicates and ragged sizes."""
    rng = np.random.default_rng(seed)
    q, t = synth.sift_like(nq, seed), synth.sift_like(nt, seed + 1)
    for _ in range(dup):
        r, c = rng.integers(0, nq), rng.integers(0, nt)
        nz = rng.integers(-noise, noise + 1, 128) * rng.integers(0, 2, 128)
        t[c] = np.clip(q[r].astype(np.int32) + nz * rng.integers(1, 6), 0, 255).astype(np.uint8)
    ok, _, _ = _lists_equal(q, t, ratio, warp_rows=warp_rows, partner=partner, cold_tiles=max(cold, 1),
                            seed_group=cold == 0, seed=seed)
    assert ok


@pytest.mark.parametrize("name,rows", [("crazyhorse", None), ("desktop", 768), ("dog", 768)])
def test_real_dataset_pairs_lists_equal(golden, name, rows):
    """The rule on real SIFT descriptors (bundled datasets, first consecutive pair): same lists as the
    exact table; the undecided rows are a fraction of the true matches."""
    z = golden(name)
    keys = sorted(k for k in z.files if k.startswith("desc"))
    q, t = M.as_u8(z[keys[0]]), M.as_u8(z[keys[1]])
    if rows is not None:
        q = q[:rows]
    for seed_group in (False, True):
        ok, flagged, _ = _lists_equal(q, t, 0.6, warp_rows=32, partner="mixed", seed_group=seed_group)
        assert ok
        assert flagged.mean() < 0.10        # a third to three quarters of the true matches (0.9 - 9 % of the rows)

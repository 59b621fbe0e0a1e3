"""GPU parity (through the C ABI) of the NORM_HAMMING2 path -- the configuration the live
reference runs (AKAZE + BFMatcher(NORM_HAMMING2), NViewReconstuct.cpp:797,876) -- against
the CPU oracle and the AKAZE fixture of the bundled desktop dataset.  Bit-exact.
Every test runs on both exact kernels: the CUDA-core XOR / POPC kernel (SFM_HAMMING_MODE=0) and
the tensor-core kernel over tetrahedron-coded s8 rows (SFM_HAMMING_MODE=2: also for calls too
small for the default rule to pick it)."""
import os

import numpy as np
import pytest

from oracle import matching as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["cuda_cores", "tensor_cores"])
def ctx(request):
    import sfm_opencv_b200 as sfm
    old = os.environ.get("SFM_HAMMING_MODE")
    os.environ["SFM_HAMMING_MODE"] = "0" if request.param == "cuda_cores" else "2"
    try:
        c = sfm.Context(0)
    finally:
        if old is None:
            del os.environ["SFM_HAMMING_MODE"]
        else:
            os.environ["SFM_HAMMING_MODE"] = old
    yield c
    c.close()


def _rand_bin(n, width, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, width), dtype=np.uint8)


def _check(out, d, idx):
    m, md, knn = out
    assert np.array_equal(knn["trainIdx0"], idx[:, 0]), "nearest index differs"
    assert np.array_equal(knn["trainIdx1"], idx[:, 1]), "second index differs"
    assert np.array_equal(knn["distance0"], d[:, 0]) and np.array_equal(knn["distance1"], d[:, 1])
    om, od, omd = M.filter_matches(d, idx)
    assert md == omd
    assert np.array_equal(m["queryIdx"], om[:, 0]) and np.array_equal(m["trainIdx"], om[:, 1])
    assert np.array_equal(m["distance"], od) and (m["imgIdx"] == 0).all()


def _run(ctx, bank, pairs):
    ctx.upload_descriptors(bank, norm="hamming2")
    m, md, knn = ctx.match_pairs(pairs, want_knn=True)
    for p, (a, b) in enumerate(pairs):
        d, idx = M.knn2_hamming2_int(bank[a], bank[b])
        _check((m[p], md[p], knn[p]), d, idx)
    return m


@pytest.mark.parametrize("nq,nt,width", [(1, 2, 61), (5, 64, 61), (129, 65, 61), (300, 1000, 61),
                                         (1000, 130, 32), (257, 4097, 64), (2000, 3000, 61),
                                         (3, 700, 1)])
def test_ragged_sizes(ctx, nq, nt, width):
    q, t = _rand_bin(nq, width, nq), _rand_bin(nt, width, 1000 + nt)
    k = min(nq, nt) // 3
    t[:k] = q[:k]                                     # distance-0 rows
    _run(ctx, [q, t], [(0, 1)])


def test_ties_and_duplicates(ctx):
    t = _rand_bin(900, 61, 7)
    t[10] = t[0]; t[20] = t[0]; t[64] = t[63]; t[600] = t[63]; t[899] = t[1]
    q = np.concatenate([t[:2], t[63:64], _rand_bin(70, 61, 8)])
    _run(ctx, [q, t], [(0, 1)])
    _, _, knn = ctx.match_pairs([(0, 1)], want_knn=True)
    assert list(knn[0]["trainIdx0"][:3]) == [0, 1, 63]
    assert list(knn[0]["trainIdx1"][:3]) == [10, 899, 64]


def test_low_entropy_descriptors_many_ties(ctx):
    # few distinct values: most rows tie on the distance and must resolve by index
    rng = np.random.default_rng(9)
    q = rng.integers(0, 2, (500, 61), dtype=np.uint8) * 0xFF
    t = rng.integers(0, 2, (1500, 61), dtype=np.uint8) * 0xFF
    q[:, 8:] = 0; t[:, 8:] = 0
    _run(ctx, [q, t], [(0, 1)])


def test_bank_of_images_all_pairs(ctx):
    sizes = [700, 130, 1100, 64, 513]
    bank = [_rand_bin(n, 61, 40 + i) for i, n in enumerate(sizes)]
    for j in range(1, len(bank)):
        k = min(len(bank[j]), len(bank[j - 1])) // 4
        noisy = bank[j - 1][k:2 * k].copy()
        noisy[:, ::7] ^= 0x11
        bank[j][:k] = noisy
    _run(ctx, bank, M.consecutive_pairs(len(bank)))
    _run(ctx, bank, M.all_pairs(len(bank)) + [(3, 0), (2, 2)])


def test_golden_akaze_desktop(ctx, golden):
    """Bundled dataset, live configuration: every consecutive pair must reproduce the
    reference library's kNN rows and the filtered match lists (2186/1063/230/553 matches,
    the counts behind the reference's bundled Viewer/structure.yml)."""
    g = golden("desktop", "akaze")
    n = int(g["n_img"])
    bank = [g[f"desc_{i}"] for i in range(n)]
    ctx.upload_descriptors(bank, norm="hamming2")
    pairs = M.consecutive_pairs(n)
    m, md, knn = ctx.match_pairs(pairs, want_knn=True)
    for p in range(n - 1):
        _check((m[p], md[p], knn[p]), g[f"knn_dist_{p}"], g[f"knn_idx_{p}"])
        assert np.array_equal(m[p]["trainIdx"], g[f"match_{p}"][:, 1])
    assert [len(x) for x in m] == [2186, 1063, 230, 553]


def test_reference_shaped_api(ctx, golden):
    import sfm_opencv_b200 as sfm
    g = golden("desktop", "akaze")
    one = sfm.match_features(ctx, g["desc_3"], g["desc_4"], norm="hamming2")
    assert np.array_equal(one["queryIdx"], g["match_3"][:, 0])
    assert np.array_equal(one["trainIdx"], g["match_3"][:, 1])


def test_errors(ctx):
    import sfm_opencv_b200 as sfm
    from sfm_opencv_b200 import _capi
    with pytest.raises(sfm.SfmError) as e:
        ctx.upload_descriptors([_rand_bin(10, 65, 1), _rand_bin(10, 65, 2)], norm="hamming2")
    assert e.value.code == _capi.SFM_E_DIM
    ctx.upload_descriptors([_rand_bin(10, 61, 1), _rand_bin(1, 61, 2)], norm="hamming2")
    with pytest.raises(sfm.SfmError) as e:
        ctx.match_pairs([(0, 1)])
    assert e.value.code == _capi.SFM_E_TOO_FEW_TRAIN
    # an L2 upload afterwards switches the context back
    from oracle import synth
    bank = synth.image_bank(2, 300, seed0=1)
    ctx.upload_descriptors(bank)
    m, _, _ = ctx.match_pairs([(0, 1)])
    om, _, _, _, _ = M.match_features(bank[0], bank[1])
    assert np.array_equal(m[0]["trainIdx"], om[:, 1])

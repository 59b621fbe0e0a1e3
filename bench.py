#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SfM hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (cv2)

Workload (BASELINE.json config 3): synthetic exhaustive all-pairs SIFT matching,
200 images x 8192 integer-valued 128-d descriptors = 19,900 pairs (i<j, query=i, train=j),
kNN k=2 + Lowe ratio + min-distance gate exactly as match_features()
(OpenCV_SFM/NViewReconstuct.cpp:873-913).  One "step" = one pass over all 19,900 pairs.
Pairs are sharded over ranks in contiguous cost-balanced blocks (no data-path collective on
results); total work is fixed by the config, so the scaling label is "strong".

The calls ask for match lists only (no raw kNN rows), so the library runs its ratio-driven sweep
(SFM_PRUNE_MODE=0 switches it off for an A/B run) -- lists, distances and min_dist are
those of the exact search, which `self_check` verifies every run.

Prints ONE JSON line (rank 0).  `value` = pairs/s with descriptors resident in HBM;
`e2e` = pairs/s through the reference-facing call (host CV_32F descriptor matrices in,
host DMatch lists out, copies inside the timed region).  With N > 1 ranks every image crosses
PCIe once: rank r uploads its 1/N slice of the images, and the packed u8 rows are all-gathered
over NVLink (NCCL) straight into every rank's bank.  The image list is cut into --stages regions
whose upload / exchange / commit are queued on the library's upload stream while ONE
sfm_match_pairs call matches the pairs in the order their images arrive.  --exchange push
(default): every rank pushes its rows into the peers' banks with the copy engines over NVLink and
raises flags in their mailboxes (sfm_peer_*, sfm_bank_push_range_async: no collective, no SMs);
--exchange nccl: an in-place NCCL all-gather per region on the upload stream.

Other workloads (their own metric; lines kept under profiles/):
    --workload pair65536   BASELINE config 4: one 65536 x 65536 pair, query rows sharded over the
                           ranks, min_dist reduced across ranks (MIN) between pass 1 and pass 2
    --workload geometry    BASELINE config 5: 4M points x V views, DLT triangulation + reprojection
                           residuals, points / observations sharded over the ranks, Huber cost summed
    --workload datasets    BASELINE configs 1-2: the bundled datasets (committed SIFT fixtures) in the
                           reference's pair schedule, end to end vs the reference's CPU library call
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INT8_DENSE_TOPS = 4500.0      # NVIDIA B200 datasheet, dense int8
HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = f"/tmp/sfm_bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
            self.f.close()
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm),
                       power_w_max=max(pw), reasons=sorted(reasons))
        return out


# ----------------------------------------------------------------------------- workload
def make_bank(n_img, n_desc):
    from oracle import synth            # input generator only (no reference arithmetic)
    return synth.image_bank(n_img, n_desc)


def all_pairs(n_img):
    return [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]


def cpu_match_pairs(bank_f32, pairs):
    """The reference's CPU path for a list of pairs: cv2 batchDistance (what
    BFMatcher::knnMatch runs) + the restated filter of NViewReconstuct.cpp:880-908."""
    from oracle import matching as M
    n = 0
    for (a, b) in pairs:
        dist, idx = M.knn2_cv(bank_f32[a], bank_f32[b])
        m, _, _ = M.filter_matches(dist, idx)
        n += len(m)
    return n


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation (OpenCV through cv2 4.13 of
    this image; the reference's C++ cannot be compiled here) on host cores."""
    if rank != 0:
        return
    import cv2
    n_img = min(args.images, 1 + args.ref_pairs_per_step)
    bank = [b.astype(np.float32) for b in make_bank(n_img, args.desc)]
    pairs = [(0, j) for j in range(1, n_img)][: args.ref_pairs_per_step]
    for _ in range(args.warmup):
        cpu_match_pairs(bank, pairs[:1])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_match_pairs(bank, pairs)
    dt = time.perf_counter() - t0
    done = args.steps * len(pairs)
    val = done / dt
    cores = cv2.getNumThreads()
    line = {
        "impl": "reference", "metric": "image pairs/s (8k SIFT/img, kNN k=2+ratio)",
        "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "reference",
                         "sample": f"{len(pairs)} pairs of {args.desc}x{args.desc} per step x "
                                   f"{args.steps} steps, cv2 {cv2.__version__} batchDistance(K=2,"
                                   f"NORM_L2)+filter, os.cpu_count={os.cpu_count()}"},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    n_pairs = args.images * (args.images - 1) // 2
    return {"workload": f"synthetic all-pairs SIFT matching: {args.images} images x {args.desc} "
                        f"x 128 u8-valued descriptors, {n_pairs} pairs, kNN k=2 + ratio 0.6 + "
                        f"5*max(min_dist,10) gate (BASELINE.json configs[2])",
            "pairs": n_pairs, "images": args.images, "desc_per_image": args.desc,
            "sharding": (f"contiguous pair blocks over {world} rank(s), no collective" if world == 1 else
                         f"{world} ranks; per arrival stage of the staged upload (image regions {args.stage_weights or args.stages}) "
                         "contiguous cost-balanced pair blocks; no collective on results"),
            "l2": f"descriptor bank {args.images * args.desc * 128 / 1e6:.0f} MB "
                  f"{'>' if args.images * args.desc * 128 > 126e6 else '<='} 126 MB L2; "
                  "no explicit flush"}


# ----------------------------------------------------------------------------- extras
L2_BYTES = 126e6


def _roof(bytes_, ms, hbm_gbs, peak_src, **kw):
    gbs = bytes_ / (ms * 1e-3) / 1e9
    out = dict({"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                "peak_source": peak_src}, **kw)
    if bytes_ < L2_BYTES:
        # the repeated launches of the timing loop re-use buffers that fit in the 126 MB L2: not an HBM number
        out.update(frac=None, note=f"working set {bytes_ / 1e6:.0f} MB fits in L2: repeated launches are L2-resident")
    return out


def extra_pair65536(ctx, n=65536, cpu=True):
    """BASELINE config 4 on one GPU: device-resident kernel time, end to end through host buffers,
    the reference's library call on a 4096-row query subset (parity + CPU time)."""
    from oracle import matching as M
    from oracle import synth
    q, t = synth.image_bank(2, n, seed0=1000)             # seeds 1000 / 1001, 20 % shared rows + noise
    ctx.upload_descriptors([q, t])
    ctx.match_pairs_resident([(0, 1)])
    best = min(ctx.match_pairs_resident([(0, 1)])[1] for _ in range(5))
    ops = 2.0 * n * n * 128
    out = {"knn_ms": best, "tops": ops / (best * 1e-3) / 1e12,
           "frac_of_4500": ops / (best * 1e-3) / 1e12 / INT8_DENSE_TOPS}
    hq = ctx.pinned_empty(q.shape, np.float32, "p65q"); hq[...] = q
    ht = ctx.pinned_empty(t.shape, np.float32, "p65t"); ht[...] = t
    for _ in range(2):
        ctx.upload_descriptors([hq, ht], overlap=True)
        ctx.match_pairs([(0, 1)], copy=False)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ctx.upload_descriptors([hq, ht], overlap=True)
        m, md, _ = ctx.match_pairs([(0, 1)], copy=False)
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    out["e2e"] = {"ms": e2e_ms, "tops": ops / (e2e_ms * 1e-3) / 1e12, "h2d_bytes": hq.nbytes + ht.nbytes,
                  "d2h_bytes": ctx.last_d2h_bytes, "matches": int(len(m[0])),
                  "api": "sfm_upload_descriptors_async(CV_32F) + sfm_match_pairs + sfm_fetch_matches"}
    if cpu:
        import cv2
        rows = np.sort(np.random.default_rng(0).choice(n, 4096, replace=False))
        tf = t.astype(np.float32)
        qf = q[rows].astype(np.float32)
        t0 = time.perf_counter()
        dist, idx = M.knn2_cv(qf, tf)
        dt = time.perf_counter() - t0
        _, _, knn = ctx.match_pairs([(0, 1)], want_knn=True)
        k = knn[0][rows]
        same = bool(np.array_equal(k["trainIdx0"], idx[:, 0]) and np.array_equal(k["trainIdx1"], idx[:, 1]) and
                    np.array_equal(k["distance0"].view(np.uint32), dist[:, 0].view(np.uint32)) and
                    np.array_equal(k["distance1"].view(np.uint32), dist[:, 1].view(np.uint32)))
        out["parity"] = {"rows": 4096, "bit_exact_vs_cv2": same}
        out["cpu_baseline"] = {"ms_per_pair": dt * (n / 4096) * 1e3, "cores": cv2.getNumThreads(), "kind": "reference",
                               "sample": f"4096 of {n} query rows x {n} train rows in {dt:.2f} s, "
                                         f"cv2 {cv2.__version__} batchDistance(K=2,NORM_L2), extrapolated x{n // 4096}"}
    return out


def extra_geometry(ctx, hbm_gbs, peak_src, fp64_peak, n=4_000_000, views=(2, 8), cpu=True):
    """BASELINE config 5 on one GPU: kernel rooflines (inputs resident), end to end through host
    buffers, the BA-loop form, and the reference's CPU calls on bounded samples."""
    import sfm_opencv_b200 as sfm
    from oracle import geometry as G
    from oracle import synth
    out = {}
    for V in views:
        sc = synth.scene(n, V)
        _, _, ms = ctx.triangulate_batch(sc["P"], sc["xy"], want_X4=True, want_xyz=False, iters=20)
        b = (8 * V + 16) * n
        tri = {"ms": ms, "points_per_s": n / (ms * 1e-3),
               "roofline": _roof(b, ms, hbm_gbs, peak_src, bytes_per_point=8 * V + 16)}
        if fp64_peak:
            # the kernel's real ceiling: 191 + 28 V fp64-pipe instructions per point (SASS census of
            # triangulate_kernel), each counted like the probe's DFMA (2 flop)
            fl = 2.0 * (191 + 28 * V) * n / (ms * 1e-3) / 1e12
            tri["fp64"] = {"achieved_tflops": fl, "peak_tflops": fp64_peak, "frac": fl / fp64_peak,
                           "note": "fp64-pipe bound; peak = in-run DFMA probe"}
        # end to end through pinned host buffers (a caller's reusable point pool)
        hxy = ctx.pinned_empty(sc["xy"].shape, np.float32, "g_xy"); hxy[...] = sc["xy"]
        hX4 = ctx.pinned_empty((4, n), np.float32, "g_X4")
        hxyz = ctx.pinned_empty((n, 3), np.float64, "g_xyz")
        ctx.triangulate_batch(sc["P"], hxy, out_X4=hX4, out_xyz=hxyz)
        t0 = time.perf_counter()
        for _ in range(3):
            X4, xyz = ctx.triangulate_batch(sc["P"], hxy, out_X4=hX4, out_xyz=hxyz)
        dt = (time.perf_counter() - t0) / 3
        tri["e2e"] = {"ms": dt * 1e3, "points_per_s": n / dt, "h2d_bytes": int(sc["xy"].nbytes + sc["P"].nbytes),
                      "d2h_bytes": int(X4.nbytes + xyz.nbytes),
                      "api": "sfm_triangulate_batch(pinned host xy -> pinned host X4 + xyz)"}
        cam, pt = synth.observations_camera_major(n, V)
        obs = sc["xy"].reshape(-1, 2)
        _, _, ms = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs, want_cost=False, iters=20)
        b = 32 * n * V + 24 * n
        res = {"ms": ms, "obs_per_s": n * V / (ms * 1e-3),
               "roofline": _roof(b, ms, hbm_gbs, peak_src, bytes_per_obs=32, bytes_per_point=24)}
        hX = ctx.pinned_empty(sc["X"].shape, np.float64, "g_X"); hX[...] = sc["X"]
        hcam = ctx.pinned_empty(cam.shape, np.int32, "g_cam"); hcam[...] = cam
        hpt = ctx.pinned_empty(pt.shape, np.int32, "g_pt"); hpt[...] = pt
        hobs = ctx.pinned_empty(obs.shape, np.float32, "g_obs"); hobs[...] = obs
        hres = ctx.pinned_empty((n * V, 2), np.float64, "g_res")
        ctx.reproject_residuals(sc["intr"], sc["ext"], hX, hcam, hpt, hobs, out_resid=hres)
        t0 = time.perf_counter()
        for _ in range(3):
            r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], hX, hcam, hpt, hobs, out_resid=hres)
        dt = (time.perf_counter() - t0) / 3
        res["e2e"] = {"ms": dt * 1e3, "obs_per_s": n * V / dt,
                      "h2d_bytes": int(sc["X"].nbytes + cam.nbytes + pt.nbytes + obs.nbytes),
                      "d2h_bytes": int(r.nbytes) + 8, "api": "sfm_reproject_residuals (all tables from pinned host memory every call)"}
        # the BA-loop form: tables resident, per evaluation only cameras + points down, cost (8 B) up
        pb = sfm.BAProblem(ctx, V, n, cam, pt, obs)
        pb.evaluate(sc["intr"], sc["ext"], hX, want_resid=False)
        t0 = time.perf_counter()
        for _ in range(5):
            _, _, c2 = pb.evaluate(sc["intr"], sc["ext"], hX, want_resid=False)
        dt = (time.perf_counter() - t0) / 5
        res["e2e_ba_loop"] = {"ms": dt * 1e3, "obs_per_s": n * V / dt, "kernel_ms": pb.kernel_ms,
                              "h2d_bytes": int(sc["X"].nbytes + sc["ext"].nbytes), "d2h_bytes": 8,
                              "cost_equals_one_shot": bool(c2 == cost),
                              "api": "sfm_ba_create once; sfm_ba_evaluate(ext, pts) -> Huber cost"}
        pb.close()
        if V == 2:
            _, _, ms = ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt, obs, want_resid=False, iters=10)
            b = (16 + 16 + 208) * n * V + 24 * n
            out["jacobians_4M_v2"] = {"ms": ms, "obs_per_s": n * V / (ms * 1e-3),
                                      "roofline": _roof(b, ms, hbm_gbs, peak_src, bytes_per_obs=240, bytes_per_point=24)}
        if cpu and V == 2:
            import cv2
            k = 400_000
            t0 = time.perf_counter()
            cv = G.triangulate_cv(sc["P"][0], sc["P"][1], sc["xy"][0, :k], sc["xy"][1, :k])
            ref = G.dehomogenize(cv)
            dt = time.perf_counter() - t0
            err = G.point_rel_err(xyz[:k], ref)
            tri["cpu_baseline"] = {"points_per_s": k / dt, "cores": 1, "kind": "reference",
                                   "sample": f"cv2 {cv2.__version__} triangulatePoints + de-homogenise on {k} of {n} points in {dt:.1f} s"}
            tri["parity"] = {"points": k, "max_rel_err_vs_cv2": float(err.max()), "tolerance": 1e-5}
            k = 2_000_000
            t0 = time.perf_counter()
            rr = G.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam[:k], pt[:k], obs[:k])
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"obs_per_s": k / dt, "cores": 1, "kind": "port",
                                   "sample": f"numpy fp64 restatement of ReprojectCost on {k} of {n * V} observations in {dt:.1f} s"}
            res["parity"] = {"observations": k, "max_abs_err_px": float(np.abs(r[:k] - rr).max())}
        out[f"triangulate_4M_v{V}"] = tri
        out[f"residuals_4M_v{V}"] = res
        del sc
    return out


def extras(ctx, hbm_gbs, peak_src, cpu=True):
    """Secondary lines (BASELINE.json configs 4 and 5, HAMMING2), rank 0 at N=1 only."""
    out = {}
    try:
        out["pair_65536"] = extra_pair65536(ctx, cpu=cpu)
    except Exception as e:                                   # pragma: no cover
        out["pair_65536"] = {"error": repr(e)}
    try:
        fp64_peak = ctx.probe_fp64_peak(4000)                # measured DFMA rate, TFLOP/s
        out["fp64_probe_tflops"] = fp64_peak
    except Exception as e:                                   # pragma: no cover
        fp64_peak = None
        out["fp64_probe"] = {"error": repr(e)}
    try:
        out.update(extra_geometry(ctx, hbm_gbs, peak_src, fp64_peak, cpu=cpu))
    except Exception as e:                                   # pragma: no cover
        out["geometry"] = {"error": repr(e)}
    try:
        # the live reference configuration: binary descriptors, NORM_HAMMING2 (AKAZE-sized: 61 B),
        # on both exact kernels (same context class, SFM_HAMMING_MODE picks the kernel)
        import sfm_opencv_b200 as sfm
        rng = np.random.default_rng(0)
        nb, nd = 16, 8192
        bank = [rng.integers(0, 256, (nd, 61), dtype=np.uint8) for _ in range(nb)]
        pairs = np.asarray(all_pairs(nb), np.int32)
        ham = {}
        for name, mode in (("tensor_cores", "1"), ("cuda_cores", "0")):
            os.environ["SFM_HAMMING_MODE"] = mode
            try:
                hc = sfm.Context(ctx.device)
            finally:
                del os.environ["SFM_HAMMING_MODE"]
            hc.upload_descriptors(bank, norm="hamming2")
            hc.match_pairs_resident(pairs)
            best = min(hc.match_pairs_resident(pairs)[1:] for _ in range(3))
            tot, _, _ = hc.match_pairs_resident(pairs)
            ham[name] = {"knn_ms": best[0], "total_ms": best[1], "image_pairs_per_s": len(pairs) / (best[1] * 1e-3),
                         "descriptor_pairs_per_s": len(pairs) * nd * nd / (best[0] * 1e-3), "matches": int(tot)}
            hc.close()
        # tensor pipe: 732 useful (768 stored) s8 dimensions per descriptor pair
        dp = ham["tensor_cores"]["descriptor_pairs_per_s"]
        ham["tensor_cores"]["roofline"] = {"bound": "tensor", "achieved": 2.0 * 732 * dp / 1e12, "peak": INT8_DENSE_TOPS,
                                           "unit": "TOP/s", "frac": 2.0 * 732 * dp / 1e12 / INT8_DENSE_TOPS,
                                           "frac_with_padding_dims": 2.0 * 768 * dp / 1e12 / INT8_DENSE_TOPS}
        ham["same_match_count"] = bool(ham["tensor_cores"]["matches"] == ham["cuda_cores"]["matches"])
        ham["note"] = ("tensor_cores: tetrahedron-coded s8 rows, tcgen05 kind::i8 (match_hamming_tc.cu); "
                       "cuda_cores: XOR / POPC kernel, 16 words per descriptor pair (match_hamming.cu)")
        out["hamming2_16x8192_allpairs"] = ham
    except Exception as e:                                   # pragma: no cover
        out["hamming2"] = {"error": repr(e)}
    return out


# ----------------------------------------------------------------------------- self-check
def self_check(ctx, sfm, local_rank, bank, my_pairs, rank, world, n_oracle=16):
    """Untimed: ties the step's output to (a) the unfiltered epilogue (SFM_KNN_MODE=0: same library,
    every group inserted) on ALL of this rank's pairs -- match lists byte for byte, kNN rows of every
    16th pair -- and (b) the CPU oracle (the reference's library call) on sampled pairs."""
    import hashlib
    out = {}
    m1, md1, _ = ctx.match_pairs(my_pairs)
    h1 = hashlib.sha256(m1.flat.tobytes() + m1.offsets.tobytes() + md1.tobytes()).hexdigest()
    sub = my_pairs[::16]
    _, _, k1 = ctx.match_pairs(sub, want_knn=True)
    hk1 = hashlib.sha256(b"".join(k.tobytes() for k in k1)).hexdigest()
    os.environ["SFM_KNN_MODE"] = "0"
    try:
        c0 = sfm.Context(local_rank)
    finally:
        del os.environ["SFM_KNN_MODE"]
    c0.upload_descriptors(bank)
    m0, md0, _ = c0.match_pairs(my_pairs)
    h0 = hashlib.sha256(m0.flat.tobytes() + m0.offsets.tobytes() + md0.tobytes()).hexdigest()
    _, _, k0 = c0.match_pairs(sub, want_knn=True)
    hk0 = hashlib.sha256(b"".join(k.tobytes() for k in k0)).hexdigest()
    c0.close()
    out["matches_sha256"] = h1
    out["filtered_equals_unfiltered_matches"] = bool(h1 == h0)
    out["filtered_equals_unfiltered_knn_rows"] = bool(hk1 == hk0)
    out["knn_rows_compared"] = int(sum(len(k) for k in k1))
    out["pairs_compared"] = len(my_pairs)
    if rank == 0 and n_oracle > 0:
        from oracle import matching as M
        rng = np.random.default_rng(123)
        pick = sorted(rng.choice(len(my_pairs), min(n_oracle, len(my_pairs)), replace=False).tolist())
        ok = True
        for p in pick:
            a, b = my_pairs[p]
            dist, idx = M.knn2_cv(bank[a].astype(np.float32), bank[b].astype(np.float32))
            om, od, omd = M.filter_matches(dist, idx)
            g = m1[p]
            ok &= bool(np.array_equal(g["queryIdx"], om[:, 0]) and np.array_equal(g["trainIdx"], om[:, 1]) and
                       np.array_equal(g["distance"].view(np.uint32), od.view(np.uint32)) and
                       np.float32(md1[p]).view(np.uint32) == np.float32(omd).view(np.uint32))
        out["oracle_pairs"] = len(pick)
        out["oracle_match_lists_equal"] = ok
    return out


# ----------------------------------------------------------------------------- main arm
class Ranks:
    """torch.distributed plumbing of one rank per GPU: barrier + device sync, max / sum / min
    over ranks (NCCL).  Nothing here is on the data path."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize()

    def _reduce(self, x, op, dtype=None):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=dtype or self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return t.item()

    def allmax(self, x):
        return float(self._reduce(float(x), self.dist.ReduceOp.MAX))

    def allsum(self, x):
        return float(self._reduce(float(x), self.dist.ReduceOp.SUM))

    def allmin_f32(self, x):
        """Exact MIN of a float32 over the ranks (min_dist of a row-sharded pair)."""
        return np.float32(self._reduce(float(np.float32(x)), self.dist.ReduceOp.MIN, self.torch.float32))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_b200(args, rank, local_rank, world):
    import sfm_opencv_b200 as sfm
    from sfm_opencv_b200.sharding import image_regions, shard_pairs_staged, staged_image_ranges
    rk = Ranks(rank, local_rank, world)
    torch, dist = rk.torch, rk.dist
    barrier, allmax, allsum = rk.barrier, rk.allmax, rk.allsum

    ctx = sfm.Context(local_rank)
    peaks, peak_src = measured_peaks()
    hbm_gbs = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))

    bank = make_bank(args.images, args.desc)                 # u8, every rank (replicated bank)
    pairs = all_pairs(args.images)
    n_desc = [len(b) for b in bank]
    pairs_np = np.asarray(pairs, np.int32)                   # the pair list as a caller's int array
    if world == 1:
        regions = None
        my_pairs = pairs_np
    else:
        # staged layout (see the e2e block below): image regions in arrival order, and pair shards
        # that give every rank its share of every arrival stage
        sw = [float(x) for x in args.stage_weights.split(",")] if args.stage_weights else None
        regions = staged_image_ranges(args.images, world, len(sw) if sw else max(1, args.stages), sw)
        mine_idx = shard_pairs_staged(pairs, n_desc, world, image_regions(args.images, regions))[rank]
        my_pairs = pairs_np[mine_idx]
    n_pairs = len(pairs)

    # ---- value: descriptors resident in HBM, result lists left on the device ------------
    ctx.upload_descriptors(bank)
    for _ in range(max(args.warmup, 3)):
        ctx.match_pairs_resident(my_pairs)
    sampler = ClockSampler(local_rank)
    l0 = ctx.launch_count
    barrier()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    knn_ms = 0.0
    matches = 0
    for _ in range(args.steps):
        tot, kms, _ = ctx.match_pairs_resident(my_pairs)
        knn_ms += kms
        matches = tot
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - l0
    ms_max = allmax(ms)
    knn_ms_max = allmax(knn_ms)
    launches = int(allsum(launches))
    matches = int(allsum(matches))
    value = n_pairs * args.steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (knn2_kernel), this rank's launches ------------
    my_ops = 2.0 * 128 * sum(n_desc[a] * n_desc[b] for a, b in my_pairs)
    tops = my_ops * args.steps / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else 0.0
    probe = ctx.probe_i8_peak(4000)                          # bare tcgen05 kind::i8 issue rate
    traffic = None                                           # ncu DRAM bytes per launch, if captured
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "knn2_traffic.json")))
        if (tj["images"], tj["desc_per_image"], tj["n_gpus"]) == (args.images, args.desc, world):
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass

    # ---- e2e: host CV_32F matrices in, host DMatch lists out, every step ----------------
    from sfm_opencv_b200.sharding import shard_range
    nvlink = 0
    if world == 1:
        host_f32 = []
        for k in range(args.images):                         # pinned, as a caller's cv::Mat pool
            a = ctx.pinned_empty(bank[k].shape, np.float32, f"desc{k}")
            a[...] = bank[k]
            host_f32.append(a)
        h2d = sum(a.nbytes for a in host_f32) + 8 * len(my_pairs)
        e2e_api = ("sfm_upload_descriptors_async(pinned CV_32F host matrices) + sfm_match_pairs + "
                   "sfm_fetch_matches (host DMatch lists); every step transfers the whole bank")

        def e2e_step():
            ctx.upload_descriptors(host_f32, overlap=True)   # H2D overlaps the matching kernels
            return ctx.match_pairs(my_pairs, copy=False)
    else:
        # every image crosses PCIe ONCE per step, and the exchange is overlapped with matching:
        # the image list is cut into `--stages` regions; per region this rank uploads its part
        # (1/N of the region), the packed u8 rows are all-gathered over NVLink in place on the
        # library's upload stream, the peers' parts are committed -- all queued without host
        # synchronisation -- and ONE sfm_match_pairs call visits the pairs in arrival order, each
        # launch waiting only for the images it reads (sfm_bank_*_async, include/sfm_b200.h)
        mine = [regions[k][rank] for k in range(len(regions))]          # (first image, count) per region
        host_f32 = []
        for first, count in mine:
            part = []
            for k in range(first, first + count):
                a = ctx.pinned_empty(bank[k].shape, np.float32, f"desc{k}")
                a[...] = bank[k]
                part.append(a)
            host_f32.append(part)
        h2d = sum(a.nbytes for part in host_f32 for a in part) + 8 * len(my_pairs)
        ctx.bank_layout(n_desc)
        spans = []                                           # per region: bank row range of every rank's part
        for row in regions:
            sp = []
            for first, count in row:
                if count > 0:
                    r0, _ = ctx.bank_image_rows(first)
                    rl, nl = ctx.bank_image_rows(first + count - 1)
                    sp.append((r0, rl + nl))
                else:
                    sp.append((0, 0))
            spans.append(sp)
        equal = all(len({b - a for a, b in sp}) == 1 and sp[0][1] > sp[0][0] for sp in spans)
        nvlink = sum(b - a for sp in spans for r, (a, b) in enumerate(sp) if r != rank) * 128
        push = args.exchange == "push"
        if push:
            # peer handles (CUDA IPC) once, any transport: here torch.distributed's object gather
            handles = [None] * world
            dist.all_gather_object(handles, ctx.peer_export())
            ctx.peer_connect(rank, handles)
            exch = ("sfm_bank_push_range_async: copy-engine pushes of the packed u8 rows into every peer's bank over "
                    "NVLink (CUDA IPC peer memory) + mailbox flags (stream memory operations), no collective, no SMs")
        else:
            up_stream = ctx.upload_stream_torch()
            exch = (f"NCCL {'all_gather_into_tensor (in place)' if equal else 'broadcast per rank'} of the packed u8 "
                    "rows over NVLink on the upload stream")
        e2e_api = (f"sfm_bank_layout_async + per region ({len(regions)} regions): sfm_bank_upload_range_async(this "
                   f"rank's 1/N of the region, pinned CV_32F) + {exch} + commit of the peers' parts; then ONE "
                   "sfm_match_pairs (pairs visited in arrival order, matching overlaps the later regions) + "
                   "sfm_fetch_matches")
        step_tag = [0]

        def e2e_step():
            ctx.bank_layout(n_desc, overlap=True)
            if push:
                step_tag[0] += 1
                tag = step_tag[0]
                ctx.bank_ready(tag)
                for k, row in enumerate(regions):
                    first, count = row[rank]
                    ctx.bank_upload_range(first, host_f32[k], overlap=True)
                    ctx.bank_push_range(first, count, k, tag)
                    for d in range(1, world):
                        r = (rank - d) % world               # the peer that pushes to this rank first
                        ctx.bank_pull_commit(r, row[r][0], 0, k, tag)      # wait for r's flag only ...
                    r_first, r_end = row[0][0], row[-1][0] + row[-1][1]
                    ctx.bank_commit(r_first, first - r_first, overlap=True)         # ... then commit the region's
                    ctx.bank_commit(first + count, r_end - first - count, overlap=True)   # two peer ranges at once
                return ctx.match_pairs(my_pairs, copy=False)
            bt = ctx.bank_as_torch()
            for k, row in enumerate(regions):
                first, count = row[rank]
                ctx.bank_upload_range(first, host_f32[k], overlap=True)
                sp = spans[k]
                with torch.cuda.stream(up_stream):
                    if equal:
                        dist.all_gather_into_tensor(bt[sp[0][0]:sp[-1][1]], bt[sp[rank][0]:sp[rank][1]])
                    else:
                        for r, (a, b) in enumerate(sp):
                            if b > a:
                                dist.broadcast(bt[a:b], src=r)
                r_first, r_end = row[0][0], row[-1][0] + row[-1][1]
                ctx.bank_commit(r_first, first - r_first, overlap=True)
                ctx.bank_commit(first + count, r_end - first - count, overlap=True)
            return ctx.match_pairs(my_pairs, copy=False)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        m, _, _ = e2e_step()
        d2h = ctx.last_d2h_bytes
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    if world > 1 and args.exchange == "push":
        ctx.peer_disconnect()                                # unmap the peers' banks before anyone frees one
        barrier()
    e2e_matches = int(allsum(len(m.flat)))
    e2e_s = allmax(e2e_s)
    e2e_val = n_pairs * args.steps / e2e_s
    h2d = int(allsum(h2d))
    d2h = int(allsum(d2h))
    nvlink = int(allsum(nvlink))

    # ---- self-check (untimed): the step's output against the unfiltered epilogue and the oracle
    check = None
    if not args.no_self_check:
        ctx.upload_descriptors(bank)
        check = self_check(ctx, sfm, local_rank, bank, my_pairs, rank, world)
        check["e2e_matches_equal_resident"] = bool(e2e_matches == matches)
        ok = all(v for k, v in check.items() if isinstance(v, bool))
        check["all_ranks_ok"] = bool(allsum(0 if ok else 1) == 0)

    if rank != 0:
        ctx.close()
        rk.close()
        return

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload -----------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import cv2
        bank_f32 = [bank[i].astype(np.float32) for i in range(min(args.images, 40))]
        cpu_match_pairs(bank_f32, [(0, 1)])
        t0 = time.perf_counter()
        done = 0
        while done < len(bank_f32) - 1 and time.perf_counter() - t0 < args.cpu_seconds:
            cpu_match_pairs(bank_f32, [(0, done + 1)])
            done += 1
        dt = time.perf_counter() - t0
        cpu = {"value": done / dt, "unit": "pairs/s", "cores": cv2.getNumThreads(),
               "kind": "reference",
               "sample": f"{done} pairs (image 0 vs 1..{done}) of {args.desc}x{args.desc} in "
                         f"{dt:.1f} s; cv2 {cv2.__version__} batchDistance(K=2,NORM_L2) = the "
                         f"reference's BFMatcher::knnMatch library call + restated filter; "
                         f"os.cpu_count={os.cpu_count()}"}

    line = {
        "metric": "image pairs/s (8k SIFT/img, kNN k=2+ratio)",
        "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args, world),
        "matches_per_step": matches,
        "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "nvlink_bytes_per_step": nvlink,
                "ms_per_step": e2e_s / args.steps * 1e3, "api": e2e_api},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "knn2_kernel", "achieved": tops,
                     "peak": INT8_DENSE_TOPS, "unit": "TOP/s", "frac": tops / INT8_DENSE_TOPS,
                     "traffic": traffic,
                     "traffic_note": "DRAM bytes of one knn2 launch from profiles/knn2_traffic.json "
                                     "(ncu --set full); null when the workload differs",
                     "peak_source": "B200 dense int8 datasheet (MEASURED_PEAKS.json has no int8 "
                                    "row); see measured_i8_probe_tops for the bare "
                                    "tcgen05.mma.kind::i8 rate measured in this run",
                     "measured_i8_probe_tops": probe, "frac_of_probe": tops / probe,
                     "ops_per_pair": 2.0 * args.desc * args.desc * 128,
                     "kernel_ms_per_step": knn_ms_max / args.steps,
                     "kernel_share_of_step": knn_ms_max / ms_max},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if check is not None:
        line["self_check"] = check
    if world == 1 and not args.no_extras:
        line["extra"] = extras(ctx, hbm_gbs, peak_src, cpu=not args.no_cpu_baseline)
    ctx.close()
    rk.close()
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- other workloads
def run_pair65536(args, rank, local_rank, world):
    """BASELINE config 4 over N GPUs: query rows of the one pair are sharded over
    the ranks (256-row aligned), the train image is replicated, min_dist (pass 1) is reduced with
    MIN across the ranks, pass 2 runs under the reduced value; the concatenated match lists equal
    the single-GPU list (checked: sha256 of the gathered list vs rank 0's unsharded run)."""
    import hashlib
    import sfm_opencv_b200 as sfm
    from oracle import synth
    from sfm_opencv_b200.sharding import shard_query_rows
    rk = Ranks(rank, local_rank, world)
    n = args.desc if args.desc != 8192 else 65536
    ctx = sfm.Context(local_rank)
    q, t = synth.image_bank(2, n, seed0=1000)             # seeds 1000 / 1001, 20 % shared rows + noise
    lo, hi = shard_query_rows(n, world)[rank]
    hq = ctx.pinned_empty(q.shape, np.float32, "q"); hq[...] = q
    ht = ctx.pinned_empty(t.shape, np.float32, "t"); ht[...] = t
    ops = 2.0 * n * n * 128

    def step(upload):
        if upload:
            ctx.upload_descriptors([hq, ht], overlap=False)
        md_local = ctx.match_rows_begin([(0, 1)], [lo], [hi - lo])[0]
        md = rk.allmin_f32(md_local)
        m, _ = ctx.match_rows_finish([md])
        return m[0], md

    ctx.upload_descriptors([q, t])
    for _ in range(max(args.warmup, 3)):
        step(False)
    rk.barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        m, md = step(False)
    ms = ctx.timer_stop()
    rk.barrier()
    ms = rk.allmax(ms) / args.steps
    for _ in range(2):
        step(True)
    rk.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m, md = step(True)
    e2e_ms = rk.allmax(time.perf_counter() - t0) / args.steps * 1e3
    # gathered list == unsharded list
    from sfm_opencv_b200.sharding import gather_match_lists
    parts = gather_match_lists([m.copy()], rank, rank + 1, world) if world > 1 else [m.copy()]
    if rank == 0:
        got = np.concatenate(parts)
        ctx.upload_descriptors([q, t])
        whole, wmd, _ = ctx.match_pairs([(0, 1)])
        same = bool(got.tobytes() == whole[0].tobytes() and np.float32(md).tobytes() == np.float32(wmd[0]).tobytes())
        line = {"metric": "one 65536 x 65536 SIFT pair (kNN k=2 + ratio): tera-ops/s, 2*Nq*Nt*128",
                "value": ops / (ms * 1e-3) / 1e12, "unit": "TOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"single pair {n} x {n} x 128 (BASELINE.json configs[3]), query rows sharded "
                                       f"over {world} rank(s), train image replicated, one float (min_dist) reduced with MIN"},
                "frac_of_4500_per_gpu": ops / (ms * 1e-3) / 1e12 / INT8_DENSE_TOPS / world,
                "e2e": {"value": ops / (e2e_ms * 1e-3) / 1e12, "unit": "TOP/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int((hq.nbytes + ht.nbytes) * world), "d2h_bytes_per_step": int(got.nbytes),
                        "api": "sfm_upload_descriptors(CV_32F, both images on every rank) + sfm_match_rows_begin + "
                               "MIN over ranks + sfm_match_rows_finish + sfm_fetch_matches"},
                "matches": int(len(got)), "sharded_list_equals_single_gpu": same,
                "matches_sha256": hashlib.sha256(got.tobytes()).hexdigest()}
        print(json.dumps(line), flush=True)
    ctx.close()
    rk.close()


def run_geometry(args, rank, local_rank, world):
    """BASELINE config 5 over N GPUs: contiguous point ranges (triangulation)
    and observation ranges (residuals) per rank, cameras replicated, no data-path collective; the
    Huber cost is the sum of the ranks' partial costs (one double, SUM)."""
    import sfm_opencv_b200 as sfm
    from oracle import synth
    from sfm_opencv_b200.sharding import shard_range
    rk = Ranks(rank, local_rank, world)
    ctx = sfm.Context(local_rank)
    peaks, peak_src = measured_peaks()
    hbm_gbs = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    n, V = args.points, args.views
    sc = synth.scene(n, V)
    s, e = shard_range(n, world)[rank]
    xy = np.ascontiguousarray(sc["xy"][:, s:e])
    cam, pt = synth.observations_camera_major(n, V)
    # this rank's observations: every camera's view of its point range, camera-major
    sel = np.concatenate([np.arange(v * n + s, v * n + e) for v in range(V)])
    cam_l, pt_l, obs_l = cam[sel], pt[sel], sc["xy"].reshape(-1, 2)[sel]
    it = 20
    _, _, tri_ms = ctx.triangulate_batch(sc["P"], xy, want_X4=True, want_xyz=False, iters=it)
    _, _, res_ms = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam_l, pt_l, obs_l, want_cost=False, iters=it)
    m = e - s
    hxy = ctx.pinned_empty(xy.shape, np.float32, "g_xy"); hxy[...] = xy
    hX4 = ctx.pinned_empty((4, m), np.float32, "g_X4")
    hxyz = ctx.pinned_empty((m, 3), np.float64, "g_xyz")
    hX = ctx.pinned_empty(sc["X"].shape, np.float64, "g_X"); hX[...] = sc["X"]
    hcam = ctx.pinned_empty(cam_l.shape, np.int32, "g_cam"); hcam[...] = cam_l
    hpt = ctx.pinned_empty(pt_l.shape, np.int32, "g_pt"); hpt[...] = pt_l
    hobs = ctx.pinned_empty(obs_l.shape, np.float32, "g_obs"); hobs[...] = obs_l
    hres = ctx.pinned_empty((m * V, 2), np.float64, "g_res")
    ctx.triangulate_batch(sc["P"], hxy, out_X4=hX4, out_xyz=hxyz)
    ctx.reproject_residuals(sc["intr"], sc["ext"], hX, hcam, hpt, hobs, out_resid=hres)
    rk.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        X4, xyz = ctx.triangulate_batch(sc["P"], hxy, out_X4=hX4, out_xyz=hxyz)
    tri_e2e = rk.allmax(time.perf_counter() - t0) / args.steps
    rk.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], hX, hcam, hpt, hobs, out_resid=hres)
    res_e2e = rk.allmax(time.perf_counter() - t0) / args.steps
    tri_ms, res_ms = rk.allmax(tri_ms), rk.allmax(res_ms)
    cost_sum = rk.allsum(cost)
    if rank == 0:
        _, c_all = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2),
                                           want_resid=False)
        line = {"metric": "points/s (batched DLT triangulation, 4M points x V views) at N GPUs",
                "value": n / (tri_ms * 1e-3), "unit": "points/s", "n_gpus": world, "steps": it, "warmup": 1,
                "ms_per_step": tri_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{n} points x {V} views (BASELINE.json configs[4]): contiguous point / "
                                       f"observation ranges over {world} rank(s), cameras replicated, no collective"},
                "roofline": _roof((8 * V + 16) * (e - s), tri_ms, hbm_gbs, peak_src, bytes_per_point=8 * V + 16, per="rank 0"),
                "e2e": {"value": n / tri_e2e, "unit": "points/s", "ms_per_step": tri_e2e * 1e3,
                        "h2d_bytes_per_step": int(sc["xy"].nbytes), "d2h_bytes_per_step": int(28 * n),
                        "api": "sfm_triangulate_batch(pinned host xy -> pinned host X4 + xyz) on every rank's point range"},
                "residuals": {"obs_per_s": n * V / (res_ms * 1e-3), "ms": res_ms,
                              "roofline": _roof(32 * (e - s) * V + 24 * (e - s), res_ms, hbm_gbs, peak_src, per="rank 0"),
                              "e2e_obs_per_s": n * V / res_e2e, "e2e_ms": res_e2e * 1e3,
                              "huber_cost_sum_over_ranks": cost_sum, "huber_cost_single_gpu": c_all,
                              "rel_diff": abs(cost_sum - c_all) / abs(c_all)}}
        print(json.dumps(line), flush=True)
    ctx.close()
    rk.close()


def run_datasets(args, rank, local_rank, world):
    """BASELINE configs 0-1 on the reference's bundled datasets (SIFT descriptors extracted once by
    tests/golden/make_golden.py with the image's cv2 and committed as fixtures: /root/reference does
    not exist on the GPU box).  Per dataset: the reference's own schedule -- consecutive pairs
    (match_features_for_all, NViewReconstuct.cpp:850-871) -- through the host API (pinned CV_32F in,
    DMatch lists out) against the reference's CPU library call on the same pairs, the match lists
    compared with the committed golden lists; desktop / crazyhorse also as exhaustive all-pairs."""
    if rank != 0:
        return
    import cv2
    import sfm_opencv_b200 as sfm
    from oracle import matching as M
    ctx = sfm.Context(local_rank)
    gdir = os.path.join(ROOT, "tests", "golden")
    out, tot_pairs, tot_gpu, tot_cpu = {}, 0, 0.0, 0.0
    for name in ("desktop", "crazyhorse", "dog"):
        g = np.load(os.path.join(gdir, f"{name}_sift.npz"))
        n = int(g["n_img"])
        bank = [g[f"desc_{i}"] for i in range(n)]
        host = []
        for i, b in enumerate(bank):
            h = ctx.pinned_empty(b.shape, np.float32, f"{name}{i}")
            h[...] = b
            host.append(h)
        for sched in ("consecutive", "allpairs"):
            if sched == "allpairs" and name == "dog":
                continue                                        # 120 pairs of ~17k x 17k: too long for the CPU leg
            pairs = M.consecutive_pairs(n) if sched == "consecutive" else M.all_pairs(n)
            pa = np.asarray(pairs, np.int32)
            for _ in range(2):
                ctx.upload_descriptors(host, overlap=True)
                ctx.match_pairs(pa, copy=False)
            reps = max(3, args.steps)
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.upload_descriptors(host, overlap=True)
                m, md, _ = ctx.match_pairs(pa, copy=False)
            gpu_s = (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            ref = []
            for a, b in pairs:
                dist, idx = M.knn2_cv(host[a], host[b])
                ref.append(M.filter_matches(dist, idx))
            cpu_s = time.perf_counter() - t0
            same = all(np.array_equal(m[p]["queryIdx"], ref[p][0][:, 0]) and np.array_equal(m[p]["trainIdx"], ref[p][0][:, 1]) and
                       np.array_equal(m[p]["distance"].view(np.uint32), ref[p][1].view(np.uint32)) and
                       np.float32(md[p]).view(np.uint32) == np.float32(ref[p][2]).view(np.uint32) for p in range(len(pairs)))
            golden_ok = None
            if sched == "consecutive":
                golden_ok = all(np.array_equal(m[p]["queryIdx"], g[f"match_{p}"][:, 0]) and
                                np.array_equal(m[p]["trainIdx"], g[f"match_{p}"][:, 1]) for p in range(len(pairs)))
            ops = 2.0 * 128 * sum(len(bank[a]) * len(bank[b]) for a, b in pairs)
            out[f"{name}_{sched}"] = {
                "images": n, "descriptors": [len(b) for b in bank], "pairs": len(pairs),
                "matches": [int(len(m[p])) for p in range(len(pairs))][:16],
                "gpu_e2e_ms": gpu_s * 1e3, "gpu_pairs_per_s": len(pairs) / gpu_s, "gpu_tops_e2e": ops / gpu_s / 1e12,
                "cpu_ms": cpu_s * 1e3, "cpu_pairs_per_s": len(pairs) / cpu_s, "speedup_e2e": cpu_s / gpu_s,
                "match_lists_equal_cv2": bool(same), "match_lists_equal_golden": golden_ok,
                "h2d_bytes": int(sum(h.nbytes for h in host)), "d2h_bytes": int(ctx.last_d2h_bytes)}
            if sched == "consecutive":
                tot_pairs += len(pairs); tot_gpu += gpu_s; tot_cpu += cpu_s
    line = {"metric": "image pairs/s on the reference's bundled datasets (consecutive pairs, SIFT kNN k=2 + ratio), end to end",
            "value": tot_pairs / tot_gpu, "unit": "pairs/s", "n_gpus": 1, "steps": max(3, args.steps), "warmup": 2,
            "ms_per_step": tot_gpu * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "bundled datasets (fixtures: SIFT descriptors of dataset/desktop, crazyhorse, dog)",
            "config": {"workload": "BASELINE.json configs[0-1]: matching of dataset/desktop (5 images), crazyhorse (7), dog (16) "
                                   "as the reference schedules it, host CV_32F descriptors in, host DMatch lists out"},
            "e2e": {"value": tot_pairs / tot_gpu, "unit": "pairs/s", "ms_per_step": tot_gpu * 1e3,
                    "api": "sfm_upload_descriptors_async(pinned CV_32F) + sfm_match_pairs + sfm_fetch_matches"},
            "cpu_baseline": {"value": tot_pairs / tot_cpu, "unit": "pairs/s", "cores": cv2.getNumThreads(), "kind": "reference",
                             "sample": f"all {tot_pairs} consecutive pairs, cv2 {cv2.__version__} batchDistance(K=2,NORM_L2) + restated filter"},
            "datasets": out}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=200)
    ap.add_argument("--desc", type=int, default=8192)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-pairs-per-step", type=int, default=8)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-self-check", action="store_true")
    ap.add_argument("--workload", default="allpairs", choices=["allpairs", "pair65536", "geometry", "datasets"])
    ap.add_argument("--points", type=int, default=4_000_000)
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="N > 1: how the packed rows travel GPU to GPU (copy-engine pushes + flags, or NCCL all-gather)")
    ap.add_argument("--stage-weights", default="1,2,3",
                    help="N > 1: relative sizes of the regions, e.g. 1,2 (overrides --stages; '' = equal regions)")
    ap.add_argument("--stages", type=int, default=2,
                    help="N > 1: regions of the image list whose upload + NVLink exchange is pipelined with matching")
    args = ap.parse_args()

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        raise SystemExit(subprocess.call(cmd + sys.argv[1:]))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "pair65536":
        run_pair65536(args, rank, local_rank, world)
    elif args.workload == "geometry":
        run_geometry(args, rank, local_rank, world)
    elif args.workload == "datasets":
        run_datasets(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

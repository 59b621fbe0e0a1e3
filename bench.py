#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SfM hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (cv2)

Workload (BASELINE.json config 3): synthetic exhaustive all-pairs SIFT matching,
200 images x 8192 integer-valued 128-d descriptors = 19,900 pairs (i<j, query=i, train=j),
kNN k=2 + Lowe ratio + min-distance gate exactly as match_features()
(OpenCV_SFM/NViewReconstuct.cpp:873-913).  One "step" = one pass over all 19,900 pairs.
Pairs are sharded over ranks in contiguous blocks (no data-path collective); total work is
fixed by the config, so the scaling label is "strong".

Prints ONE JSON line (rank 0).  `value` = pairs/s with descriptors resident in HBM;
`e2e` = pairs/s through the reference-facing call (host CV_32F descriptor matrices in,
host DMatch lists out, copies inside the timed region).
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
            "sharding": f"contiguous pair blocks over {world} rank(s), no collective",
            "l2": f"descriptor bank {args.images * args.desc * 128 / 1e6:.0f} MB "
                  f"{'>' if args.images * args.desc * 128 > 126e6 else '<='} 126 MB L2; "
                  "no explicit flush"}


# ----------------------------------------------------------------------------- extras
def extras(ctx, hbm_gbs, peak_src):
    """Secondary roofline lines (BASELINE.json configs 4 and 5), rank 0 at N=1 only."""
    from oracle import synth
    out = {}
    try:
        q = synth.sift_like(65536, 1000)
        t = synth.sift_like(65536, 1001)
        ctx.upload_descriptors([q, t])
        ctx.match_pairs_resident([(0, 1)])
        best = min(ctx.match_pairs_resident([(0, 1)])[1] for _ in range(5))
        ops = 2.0 * 65536 * 65536 * 128
        out["pair_65536"] = {"knn_ms": best, "tops": ops / (best * 1e-3) / 1e12,
                             "frac_of_4500": ops / (best * 1e-3) / 1e12 / INT8_DENSE_TOPS}
    except Exception as e:                                   # pragma: no cover
        out["pair_65536"] = {"error": str(e)}
    try:
        fp64_peak = ctx.probe_fp64_peak(4000)                # measured DFMA rate, TFLOP/s
        out["fp64_probe_tflops"] = fp64_peak
    except Exception as e:                                   # pragma: no cover
        fp64_peak = None
        out["fp64_probe"] = {"error": str(e)}
    try:
        n = 4_000_000
        for V in (2, 8):
            sc = synth.scene(n, V)
            _, _, ms = ctx.triangulate_batch(sc["P"], sc["xy"], want_X4=True, want_xyz=False,
                                             iters=20)
            b = (8 * V + 16) * n
            out[f"triangulate_4M_v{V}"] = {
                "ms": ms, "points_per_s": n / (ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": hbm_gbs,
                             "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / hbm_gbs,
                             "bytes_per_point": 8 * V + 16, "peak_source": peak_src}}
            if fp64_peak:
                # the kernel's real ceiling: ~(20 V + 170) fp64 FMA/MUL per point
                fl = 2.0 * (20 * V + 170) * n / (ms * 1e-3) / 1e12
                out[f"triangulate_4M_v{V}"]["fp64"] = {"achieved_tflops": fl, "peak_tflops": fp64_peak,
                                                       "frac": fl / fp64_peak,
                                                       "note": "fp64-pipe bound; peak = in-run DFMA probe"}
            cam, pt = synth.observations_camera_major(n, V)
            _, _, ms = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt,
                                               sc["xy"].reshape(-1, 2), want_cost=False,
                                               iters=20)
            b = 32 * n * V + 24 * n
            out[f"residuals_4M_v{V}"] = {
                "ms": ms, "obs_per_s": n * V / (ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": hbm_gbs,
                             "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / hbm_gbs,
                             "bytes_per_obs": 32, "bytes_per_point": 24,
                             "peak_source": peak_src}}
            if V == 2:
                # Jacobians of the same residual blocks: 16 B in + 16 B residual + 208 B of
                # derivatives per observation, 24 B per distinct point
                _, _, ms = ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt,
                                                   sc["xy"].reshape(-1, 2), want_resid=False, iters=10)
                b = (16 + 16 + 208) * n * V + 24 * n
                out["jacobians_4M_v2"] = {
                    "ms": ms, "obs_per_s": n * V / (ms * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": b / (ms * 1e-3) / 1e9, "peak": hbm_gbs,
                                 "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / hbm_gbs,
                                 "bytes_per_obs": 240, "bytes_per_point": 24, "peak_source": peak_src}}
            del sc
    except Exception as e:                                   # pragma: no cover
        out["geometry"] = {"error": str(e)}
    try:
        # the live reference configuration: binary descriptors, NORM_HAMMING2 (AKAZE-sized: 61 B)
        rng = np.random.default_rng(0)
        nb, nd = 16, 8192
        bank = [rng.integers(0, 256, (nd, 61), dtype=np.uint8) for _ in range(nb)]
        ctx.upload_descriptors(bank, norm="hamming2")
        pairs = all_pairs(nb)
        ctx.match_pairs_resident(pairs)
        best = min(ctx.match_pairs_resident(pairs)[1:] for _ in range(3))
        out["hamming2_16x8192_allpairs"] = {
            "knn_ms": best[0], "total_ms": best[1], "image_pairs_per_s": len(pairs) / (best[1] * 1e-3),
            "descriptor_pairs_per_s": len(pairs) * nd * nd / (best[0] * 1e-3),
            "note": "CUDA-core XOR/POPC kernel, 16 words per descriptor pair"}
    except Exception as e:                                   # pragma: no cover
        out["hamming2"] = {"error": str(e)}
    return out


# ----------------------------------------------------------------------------- main arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import sfm_opencv_b200 as sfm
    from sfm_opencv_b200.sharding import shard_pairs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = sfm.Context(local_rank)
    peaks, peak_src = measured_peaks()
    hbm_gbs = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))

    bank = make_bank(args.images, args.desc)                 # u8, every rank (replicated bank)
    pairs = all_pairs(args.images)
    n_desc = [len(b) for b in bank]
    lo, hi = shard_pairs(pairs, n_desc, world)[rank]
    my_pairs = pairs[lo:hi]
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
    ids = sorted({i for p in my_pairs for i in p})
    remap = {g: k for k, g in enumerate(ids)}
    loc_pairs = [(remap[a], remap[b]) for a, b in my_pairs]
    host_f32 = []
    for k, g in enumerate(ids):                              # pinned, as a caller's cv::Mat pool
        a = ctx.pinned_empty(bank[g].shape, np.float32, f"desc{k}")
        a[...] = bank[g]
        host_f32.append(a)
    h2d = sum(a.nbytes for a in host_f32) + 8 * len(loc_pairs)
    for _ in range(2):
        ctx.upload_descriptors(host_f32, overlap=True)
        ctx.match_pairs(loc_pairs, copy=False)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        ctx.upload_descriptors(host_f32, overlap=True)       # H2D overlaps the matching kernels
        m, _, _ = ctx.match_pairs(loc_pairs, copy=False)
        d2h = ctx.last_d2h_bytes
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = allmax(e2e_s)
    e2e_val = n_pairs * args.steps / e2e_s
    h2d = int(allsum(h2d))
    d2h = int(allsum(d2h))

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
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
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "sfm_upload_descriptors_async(pinned CV_32F host matrices) + sfm_match_pairs + "
                       "sfm_fetch_matches (host DMatch lists); every step transfers the whole bank"},
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
    if world == 1 and not args.no_extras:
        line["extra"] = extras(ctx, hbm_gbs, peak_src)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

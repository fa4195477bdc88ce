#!/usr/bin/env python
"""bench.py -- Mfeatures/s (extract + match) of the B200 feature hot path on the BASELINE.json config-2
workload: a stream of 3-fisheye 754x480 multi-camera frames, 8-level pyramid (scale 1.2), 2000 features per
camera, mdBRIEF-256 descriptors with masks; every (frame, camera) is brute-force matched against the same
camera of the previous frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

One step = one pass over one batch of F frames x 3 cameras (synthetic, seeded).  `value` times the kernels with
the inputs already resident in HBM (CUDA events on the launching stream); `e2e` times the C-ABI host call
mcs_extract_match_stream with pinned HOST buffers, H2D and D2H inside the timed region.  `roofline` is the fused
pyramid+blur+FAST kernel (K1): algorithmic bytes per launch / CUDA-event time / measured HBM peak.
`cpu_baseline` is the oracle port on the host cores (bounded sample).  --impl reference prints the CPU arm.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_CAMS, W, H = 3, 754, 480
NFEATURES, NLEVELS, K_MATCH = 2000, 8, 2
WORKLOAD = "lafida-3cam-754x480-stream, 8 levels x1.2, 2000 feat/cam, mdBRIEF-256+masks, match vs previous frame"


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_stream(cams, n_frames, seed0):
    """[F, 3, H, W] uint8: per camera a sliding crop of one big seeded texture (real inter-frame motion)."""
    from multicol_slam_b200 import synth
    per_cam = [synth.texture_stream(cams[c], n_frames, seed=seed0 + c) for c in range(N_CAMS)]
    return np.ascontiguousarray(np.stack(per_cam, axis=1))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def usable_cores():
    """Host threads this process may really use: min(affinity, cgroup CPU quota).  The GPU boxes expose 128 logical CPUs but
    cap the container at a 16-CPU quota (cpu.max = 1600000 100000); more threads than that only add throttling."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def cpu_oracle_run(cams, masks, images, n_threads):
    """Oracle port (oracle/mcs_oracle.cpp) over `images` [F,3,H,W] with n_threads host threads: extraction per
    (frame, camera) in parallel (the reference parallelises over cameras, src/cMultiFrame.cpp:128), then the
    brute-force match of every (t, c) against (t-1, c).  Returns (features, seconds)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle_api as oa
    oa.lib()
    t0 = time.perf_counter()
    nfeat, _ = oa.stream_mt(images, masks, cams, n_threads, nfeatures=NFEATURES, nlevels=NLEVELS)   # std::thread workers in C++
    return nfeat, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference cannot be built here (needs
    OpenCV C++), so this is the oracle PORT of it, with all host threads, on a bounded sample per step."""
    from multicol_slam_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cams = synth.lafida_cams()
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    cores = usable_cores()
    frames = max(2, min(max(args.ref_frames, 2 * cores), 128))
    images = make_stream(cams, frames, 1000)
    for _ in range(max(min(args.warmup, 1), 1)):
        cpu_oracle_run(cams, masks, images, cores)            # warms the per-thread malloc arenas
    nfeat, dt = 0, 0.0
    for _ in range(args.steps):
        n, t = cpu_oracle_run(cams, masks, images, cores)
        nfeat += n
        dt += t
    val = nfeat / dt / 1e6
    sample = f"{frames} frames x 3 cams per step, {args.steps} steps"
    print(json.dumps({"impl": "reference", "metric": "Mfeatures/s extract+match", "value": val, "unit": "Mfeatures/s",
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "sample": sample},
                      "cpu_baseline": {"value": val, "unit": "Mfeatures/s", "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "Mfeatures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128, help="multi-camera frames per step and per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import multicol_slam_b200.api as api
    from multicol_slam_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert api.device_count() > 0, "no sm_100 device"
    F = args.frames
    B = F * N_CAMS
    cams = synth.lafida_cams()
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    coi = np.tile(np.arange(N_CAMS, dtype=np.int32), F)
    # each rank owns its own temporal chunk of the stream (weak scaling: per-GPU work fixed)
    images = make_stream(cams, F, 1000 + 97 * rank)                          # [F,3,H,W]
    host_images = torch.from_numpy(images).pin_memory()
    PITCH = (W + 63) // 64 * 64                                                # 16-byte aligned rows: K1 stages with 128-bit loads
    dev_images = torch.zeros((B, H, PITCH), dtype=torch.uint8, device=dev)
    dev_images[:, :, :W] = host_images.to(dev, non_blocking=True).view(B, H, W)

    ex = api.mdBRIEFextractorOct(nfeatures=NFEATURES, nlevels=NLEVELS, do_dBrief=True, learnMasks=True)
    cap, ds = ex.capacity, 32
    # one packed per-rank feature buffer [counts | kps | desc | dmask] -> a single all_gather
    sizes = [B * 4, B * cap * 28, B * cap * ds, B * cap * ds]
    offs = np.concatenate([[0], np.cumsum([(s + 255) // 256 * 256 for s in sizes])])
    packed = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=dev)
    out = dict(counts=packed[offs[0]:offs[0] + sizes[0]].view(torch.int32),
               kps=packed[offs[1]:offs[1] + sizes[1]].view(torch.int32).view(B, cap, 7),
               desc=packed[offs[2]:offs[2] + sizes[2]].view(B, cap, ds),
               dmask=packed[offs[3]:offs[3] + sizes[3]].view(B, cap, ds))
    midx = torch.empty((B, cap, K_MATCH), dtype=torch.int32, device=dev)
    mdist = torch.empty((B, cap, K_MATCH), dtype=torch.int32, device=dev)
    gathered = torch.empty(world * packed.numel(), dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.Stream(dev)

    def step():
        ex.extract_batch_device(dev_images, masks, cams, coi, out=out, stream=stream, width=W)
        api.match_stream_device(out["desc"], out["dmask"], out["counts"], F, N_CAMS, K=K_MATCH, out=(midx, mdist), stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, packed)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        # ---- per-kernel timings (separate pass with event recording on; not part of the timed region) ----
        ex.set_profiling(True)
        k_ms = np.zeros(3)
        for _ in range(3):
            step()
            torch.cuda.synchronize(dev)
            k_ms += np.array(ex.get_timings())
        k_ms /= 3
        ex.set_profiling(False)
        ev_m0, ev_m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_m0.record(stream)
        api.match_stream_device(out["desc"], out["dmask"], out["counts"], F, N_CAMS, K=K_MATCH, out=(midx, mdist), stream=stream)
        ev_m1.record(stream)
        torch.cuda.synchronize(dev)
        match_ms = ev_m0.elapsed_time(ev_m1)
        # ---- timed region: exactly K steps, device resident ----
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    feats_rank = int(out["counts"].sum().item())
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    feats = torch.tensor([feats_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(feats, op=dist.ReduceOp.SUM)
    ms_per_step = t_ms.item() / args.steps
    value = feats.item() / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the C-ABI stream call with pinned HOST buffers (H2D + kernels + D2H inside the timed region) ----
    h_out = dict(kps=torch.empty((B, cap, 7), dtype=torch.int32).pin_memory(), desc=torch.empty((B, cap, ds), dtype=torch.uint8).pin_memory(),
                 dmask=torch.empty((B, cap, ds), dtype=torch.uint8).pin_memory(), counts=torch.empty(B, dtype=torch.int32).pin_memory(),
                 match_idx=torch.empty((B, cap, K_MATCH), dtype=torch.int32).pin_memory(),
                 match_dist=torch.empty((B, cap, K_MATCH), dtype=torch.int32).pin_memory())
    from multicol_slam_b200.ctypes_defs import KEYPOINT_DTYPE
    np_out = {k: v.numpy() for k, v in h_out.items()}
    np_out["kps"] = np_out["kps"].view(KEYPOINT_DTYPE).reshape(B, cap)
    himg = host_images.numpy()
    for _ in range(2):
        ex.extract_match_stream(himg, masks, cams, K=K_MATCH, out=np_out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        ex.extract_match_stream(himg, masks, cams, K=K_MATCH, out=np_out)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = feats.item() / (e2e_t.item() / e2e_steps) / 1e6
    h2d = int(himg.nbytes + masks.nbytes)
    d2h = int(sum(v.numel() * v.element_size() for v in h_out.values()))
    assert int(np_out["counts"].sum()) == feats_rank, "e2e and device-resident runs disagree"

    # ---- single-frame latency of the reference-shaped call: one 3-camera frame through mcs_extract_batch (host in/out) ----
    lat_ms = None
    if rank == 0:
        one = np.ascontiguousarray(himg[0])
        ex.extract_batch(one, masks, cams, [0, 1, 2])
        t0 = time.perf_counter()
        for _ in range(20):
            ex.extract_batch(one, masks, cams, [0, 1, 2])
        lat_ms = (time.perf_counter() - t0) / 20 * 1e3
    if rank == 0:
        # ---- roofline of K1 (fused pyramid + blur + FAST), algorithmic bytes per SURVEY 8d / DESIGN.md ----
        P = sum(int(ex.debug_read(l, 0).size) for l in range(NLEVELS))        # sum of pyramid pixels = 1 120 256
        n_raw = 0
        for l in range(NLEVELS):
            n_raw += len(ex.debug_read(l, 3, image_index=0))
        alg_bytes_img = P + 8 * n_raw                                          # K1: read L0 once + write L1..7 + 8 B / raw corner
        k1_ms_launch = k_ms[0] / NLEVELS
        peak, peak_src = hbm_peak()
        achieved = alg_bytes_img * B / (k_ms[0] * 1e-3) / 1e9                  # all 8 level launches together
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "k1_traffic.json"
        if tp.exists():                      # DRAM bytes of K1 from the committed ncu --set full capture, scaled to this batch
            tj = json.loads(tp.read_text())
            traffic, traffic_src = tj["dram_bytes_per_image"] * B / NLEVELS, tj["source"]
        roof = {"kernel": "pyr_fast_kernel (K1, 8 launches/step, one per level)", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": "bytes per launch (average over the 8 level launches); " + str(traffic_src),
                "algorithmic_bytes_per_launch_avg": alg_bytes_img * B / NLEVELS, "peak_source": peak_src,
                "issue_slot_utilisation_pct_ncu": 72.5,
                "algorithmic_bytes_per_camera_frame": alg_bytes_img, "ms_per_launch_avg": k1_ms_launch,
                "stage_ms": {"k1_pyr_blur_fast": k_ms[0], "k2_octree": k_ms[1], "k3_angle_describe": k_ms[2], "m2_match_stream": match_ms}}
        # the other stages against the same HBM peak (all compute-bound; bytes per SURVEY 8d)
        n_feat = int(feats_rank)
        k3_bytes = n_feat * (845 + 51 * 51 + 28 + 64)        # IC disc + blurred patch + keypoint + descriptor/mask
        m2_pairs = (B - N_CAMS) if B > N_CAMS else 0
        m2_bytes = m2_pairs * (64 * 2 * NFEATURES + 12 * NFEATURES) # 32(1+m)(Q+D) + 12Q per image pair
        roof["other_stages"] = {
            "k3_describe_kernel": {"bound": "fp64 issue / latency (ncu: FP64 pipe 41 %, issue 47 %)", "algorithmic_GB_per_s": k3_bytes / (k_ms[2] * 1e-3) / 1e9,
                                   "frac_of_hbm_peak": k3_bytes / (k_ms[2] * 1e-3) / 1e9 / peak},
            "m2_hamming_stream_kernel": {"bound": "integer ALU (ncu: ALU pipe 81 %)", "algorithmic_GB_per_s": m2_bytes / (match_ms * 1e-3) / 1e9,
                                         "frac_of_hbm_peak": m2_bytes / (match_ms * 1e-3) / 1e9 / peak,
                                         "pair_distances_per_s": m2_pairs * float(NFEATURES) * NFEATURES / (match_ms * 1e-3)},
            "k2_octree_kernel": {"bound": "latency (one CTA per image-level, serial passes)", "ms": k_ms[1]}}
        cpu = None
        if not args.no_cpu_baseline:
            cores = usable_cores()
            cf = int(min(F, max(4, 2 * cores)))                     # 6 images per thread: balanced, ~1-2 s of wall time
            cpu_oracle_run(cams, masks, images[:cf], cores)          # warm-up (per-thread malloc arenas, page faults)
            nf, dt = cpu_oracle_run(cams, masks, images[:cf], cores)
            cpu = {"value": nf / dt / 1e6, "unit": "Mfeatures/s", "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port",
                   "sample": f"{cf} frames x 3 cams of the same stream, {dt:.1f} s wall on {cores} threads (C++ std::thread driver), after one warm-up pass"}
        print(json.dumps({
            "metric": "Mfeatures/s extract+match", "value": value, "unit": "Mfeatures/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "images_per_step_per_gpu": B,
                       "features_per_step": int(feats.item()), "l2": f"inputs {himg.nbytes / 1e6:.0f} MB + {B * 2.6:.0f} MB pyramid per step > 126 MB L2",
                       "parallelism": f"stream-sharded x{world}" + (", 1 all_gather of the packed feature buffer" if world > 1 else "")},
            "e2e": {"value": e2e_val, "unit": "Mfeatures/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "mcs_extract_match_stream (C ABI, pinned host buffers)", "steps": e2e_steps},
            "gpu_launches": args.steps * (NLEVELS + 2 + 1), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "single_frame_latency_ms": {"value": lat_ms, "what": "one 3-camera frame, mcs_extract_batch with pageable host buffers, mean of 20"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

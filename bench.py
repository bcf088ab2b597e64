#!/usr/bin/env python
"""bench.py — frames/s of the per-volume fusion+raycast loop on synthetic KITTI-shaped streams.

    python bench.py --gpus N --steps K --warmup W            (own arm, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: CPU, rank 0 only)

A step is one frame through {AllocateSceneFromDepth, IntegrateIntoScene, CreateExpectedDepths,
CreateICPMaps (raycast), Decay(partial)} on one ITMScene volume (BASELINE.json configs[1]).
value  = frames/s with the frames already resident in HBM (whole job: all ranks' frames / max time)
e2e    = frames/s through b200_host_frame_submit/_wait (host buffers in, host image out): every frame's
         depth+RGB are copied from pinned host memory and its grey raycast image is copied back inside the
         timed region; the copies of neighbouring frames overlap the kernels (two staging slots)
roofline = IntegrateIntoScene: algorithmic bytes (8224 B per integrated block + w*h*8 B of images
         per launch) / mean launch duration from CUDA events around every launch of the timed region
cpu_baseline = the CPU oracle (oracle/tsdf_oracle.c, OpenMP at the reference's own pragma sites)
         timed on this box's host cores on a bounded sample of the same stream.
N > 1: one independent volume per rank (weak scaling; no data-path collective inside fusion); the
per-volume raycast image is gathered to rank 0 over NCCL every frame.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dynslam_b200 import synth  # noqa: E402

METRIC = "frames/sec TSDF fusion+raycast at KITTI 1242x375; Mvoxels/s integrated"
NUM_BLOCKS, NUM_BUCKETS, EXCESS = 0x60000, 0x100000, 0x80000   # ITMLibSettings.cpp:115, ITMLibDefines.h:42-53
DECAY = (1, 200)                                                # --max_decay_weight=1 --min_decay_age=200 (DynSLAMGUI.cpp:38-40)
BYTES_PER_BLOCK = 8192 + 32


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def pause(self, on):
        """SIGSTOP / SIGCONT nvidia-smi: its 20 ms polling was measured to stall the driver for 50-70 ms now and then, which a
        host-synchronous loop (the e2e phases) sees in full; the clocks are only needed for the device-timed region."""
        import signal
        if self.proc and self.proc.poll() is None:
            try:
                self.proc.send_signal(signal.SIGSTOP if on else signal.SIGCONT)
            except Exception:
                pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.pause(False)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def frame_gaps(marks):
    """Host-side spacing of consecutive pipelined submissions (ms): median / max / index of the max — exposes one-off stalls."""
    if len(marks) < 3:
        return None
    d = np.diff(np.asarray(marks)) * 1000.0
    return {"median": float(np.median(d)), "max": float(d.max()), "argmax": int(d.argmax()), "n": int(d.size)}


def gen_frames(seed, first, count, length_m):
    scene = synth.StreetScene(seed=seed, length_m=length_m)
    out = []
    for f in range(first, first + count):
        depth, rgb, M, proj = synth.kitti_frame(scene, f)
        out.append((depth, rgb, M, proj))
    return out


def gen_frames_parallel(seed, first, count, length_m, workers):
    if count <= 0:
        return []
    if workers <= 1 or count < 8:
        return gen_frames(seed, first, count, length_m)
    import multiprocessing as mp
    chunk = (count + workers - 1) // workers
    jobs = [(seed, first + i * chunk, min(chunk, count - i * chunk), length_m) for i in range(workers) if i * chunk < count]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.starmap(gen_frames, jobs)
    return [f for p in parts for f in p]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "integrate_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline of the own arm and the whole --impl reference arm)
# --------------------------------------------------------------------------------------------------
def cpu_run(seed, preroll, warmup, steps, length_m, omp=True):
    from tests import hostlib as H
    L = H.oracle()
    vol = H.HostVolume(NUM_BLOCKS, NUM_BUCKETS, EXCESS, synth.KITTI_W, synth.KITTI_H)
    frames = gen_frames_parallel(seed, 0, preroll + warmup + steps, length_m, min(8, os.cpu_count() or 1))
    times, vox = [], 0

    def one(fr):
        depth, rgb, M, proj = fr
        v = H.make_view(depth, rgb, M, proj)
        L.oracle_allocate_from_depth(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(v), 0, int(omp))
        L.oracle_integrate(vol.engine, C.byref(vol.scene), C.byref(vol.rs), C.byref(v), int(omp))
        n = L.oracle_integrated_blocks(vol.engine)
        cam = H.make_camera(M, proj)
        L.oracle_expected_depths(C.byref(vol.scene), C.byref(vol.rs), C.byref(cam))
        L.oracle_icp_maps(C.byref(vol.scene), C.byref(vol.rs), C.byref(v), H.vptr(vol.points), H.vptr(vol.normals), int(omp))
        L.oracle_decay(vol.engine, C.byref(vol.scene), C.byref(vol.rs), DECAY[0], DECAY[1], 0)
        return n

    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu}) if omp else [1]
    best = None
    for i, fr in enumerate(frames):
        if omp and preroll >= len(cands) + 1 and 1 <= i <= len(cands):
            L.oracle_set_threads(cands[i - 1])     # calibration during the pre-roll: one frame per thread count
        elif omp and i == len(cands) + 1 and best is not None:
            L.oracle_set_threads(best[1])
        t0 = time.perf_counter()
        n = one(fr)
        dt = time.perf_counter() - t0
        if omp and preroll >= len(cands) + 1 and 1 <= i <= len(cands) and (best is None or dt < best[0]):
            best = (dt, cands[i - 1])
        if i >= preroll + warmup:
            times.append(dt)
            vox += n * 512
    total = sum(times)
    return {"fps": len(times) / total if total > 0 else 0.0, "ms_per_step": 1000.0 * total / max(len(times), 1),
            "mvoxels_per_s": vox / total / 1e6 if total > 0 else 0.0, "cores": L.oracle_num_threads() if omp else 1,
            "frames": len(times), "preroll": preroll}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = min(args.steps, 16)
    r = cpu_run(6, args.ref_preroll, min(args.warmup, 3), steps, 200.0, omp=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "KITTI-odometry-06-shaped 1242x375 static-map fusion+raycast loop (configs[1])",
                   "voxel_m": 0.05, "mu_m": 0.75, "maxW": 50, "blocks": NUM_BLOCKS, "buckets": NUM_BUCKETS},
        "mvoxels_per_s": r["mvoxels_per_s"],
        "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['frames']} consecutive frames after a {r['preroll']}-frame CPU pre-roll of the same stream; "
                                   "oracle/tsdf_oracle.c with OpenMP at the reference's pragma sites (the reference's own _CPU "
                                   "hash engines are commented out, SURVEY finding 1)"},
        "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# the same stream through the REAL ITMLib objects (oracle/itm_harness.cpp): unmodified reference CUDA
# engines built for sm_100a vs. the B200 shim classes; pinned H2D of every frame inside the timing
# --------------------------------------------------------------------------------------------------
def run_itm_harness(frames, preroll, timed):
    import torch
    from tests import harnesslib as HL
    if not HL.available():
        return {"error": "oracle/_ref/libitmharness.so not built"}
    n = min(len(frames), preroll + timed)
    pinned = [(torch.from_numpy(f[0]).pin_memory(), torch.from_numpy(f[1]).pin_memory()) for f in frames[:n]]
    out = {}
    for name, impl in (("reference_cuda_build", HL.REFERENCE_CUDA), ("b200_itm_shim", HL.B200_SHIM)):
        hs = HL.Harness(impl, synth.KITTI_W, synth.KITTI_H, frames[0][3], numBlocks=NUM_BLOCKS)
        t0 = 0.0
        for i in range(n):
            if i == preroll:
                hs.sync()
                t0 = time.perf_counter()
            hs.process_frame(pinned[i][0].numpy(), pinned[i][1].numpy(), frames[i][2], decay=DECAY)
        hs.sync()
        dt = time.perf_counter() - t0
        c = hs.counters()
        hs.close()
        out[name] = {"value": (n - preroll) / dt, "unit": "frames/s", "ms_per_step": 1000.0 * dt / (n - preroll),
                     "visible_blocks": c["noVisibleBlocks"], "frames": n - preroll}
    out["what"] = ("per frame: pinned H2D of depth+RGB into the ITMView, then AllocateSceneFromDepth, IntegrateIntoScene, "
                   "CreateExpectedDepths, CreateICPMaps, Decay through the abstract ITMLib interfaces (synchronous, as DynSLAM "
                   f"calls them); {preroll}-frame pre-roll; reference = unmodified CUDA engines, nvcc sm_100a --use_fast_math")
    out["speedup_shim_vs_reference_cuda"] = out["b200_itm_shim"]["value"] / out["reference_cuda_build"]["value"]
    return out


# --------------------------------------------------------------------------------------------------
# roofline stress (BASELINE.json configs[4]): 4 mm voxels, mu 16 mm — the same street at
# 12.5x finer voxels makes the visible list two orders of magnitude longer, so IntegrateIntoScene runs for
# hundreds of microseconds and its bandwidth fraction can be read without launch effects.
# --------------------------------------------------------------------------------------------------
def run_hires(local_rank, frames_n):
    import torch
    from dynslam_b200 import engine as E
    dev = torch.device("cuda", local_rank)
    W, H_ = synth.KITTI_W, synth.KITTI_H
    nb, nbk, nex = 3000000, 0x400000, 0x100000
    frames = gen_frames_hires(5, frames_n)
    scene = E.Scene(E.SceneParams(voxelSize=0.004, mu=0.016, maxW=50), nb, nbk, nex, device=f"cuda:{local_rank}")
    eng = E.Engine(scene, (W, H_), decayRingItems=4 * nb)
    reco = E.SceneReconstructionEngine(eng)
    rs = E.VisualisationEngine(eng, scene).CreateRenderState((W, H_))
    reco.ResetScene(scene)
    views = [E.View(torch.from_numpy(f[0]).to(dev), torch.from_numpy(f[1]).to(dev), f[2], f[3]) for f in frames]
    torch.cuda.synchronize(dev)
    half = frames_n // 2
    for v in views[:half]:
        eng.process_frame_async(rs, v, None, None, decay=None, raycast=False)
    eng.sync(rs)
    b0 = eng.stats().totalIntegratedBlocks
    eng.set_timing(2)
    for v in views[half:]:
        eng.process_frame_async(rs, v, None, None, decay=None, raycast=False)
    eng.sync(rs)
    st = eng.stats()
    blocks, ms, n = st.totalIntegratedBlocks - b0, st.ring_ms_integrate, st.ring_count
    peak, _ = peaks()
    alg = blocks * BYTES_PER_BLOCK + n * W * H_ * 8
    out = {"workload": "4 mm voxels, mu 16 mm, depth clamp 8 m, same street (configs[4])", "frames": n,
           "visible_blocks": rs.noVisibleBlocks, "allocated_blocks": nb - 1 - scene.lastFreeBlockId,
           "mean_launch_us": 1000.0 * ms / max(n, 1), "achieved": alg / (ms / 1000.0) / 1e9 if ms > 0 else 0.0, "unit": "GB/s",
           "peak": peak, "mvoxels_per_s": blocks * 512 / (ms / 1000.0) / 1e6 if ms > 0 else 0.0}
    out["frac"] = out["achieved"] / peak
    eng.close()
    return out


# --------------------------------------------------------------------------------------------------
# view builder: ITMViewBuilder::UpdateView of one raw KITTI-shaped frame through the real
# ITMLib classes — the reference's ITMViewBuilder_CUDA (7 launches) vs ITMViewBuilder_B200 (one fused kernel)
# --------------------------------------------------------------------------------------------------
def run_view_builder(frames, iters=60):
    from tests import harnesslib as HL
    if not HL.available():
        return {"error": "oracle/_ref/libitmharness.so not built"}
    raw = np.round(frames[0][0] * 1000.0).astype(np.int16)
    out = {"what": "UpdateView(rgb, raw int16 depth, useBilateralFilter=true) at 1242x375: 'update_view_us' = the call as "
                   "ITMMainEngine makes it (two blocking H2D copies from ITMLib's pinned host images + conversion + 5 filter passes "
                   "+ copy), 'device_only_us' = the same stages on a device-resident raw image; wall clock, device synchronised",
           "alg_bytes": raw.size * 6}
    for name, impl in (("reference_cuda_build", HL.REFERENCE_CUDA), ("b200", HL.B200_SHIM)):
        vb = HL.ViewBuilderHarness(impl, synth.KITTI_W, synth.KITTI_H, frames[0][3])
        vb.update_view(raw, frames[0][1])
        vb.time_update_view(5); vb.time_device_only(5)
        out[name] = {"update_view_us": 1000.0 * vb.time_update_view(iters), "device_only_us": 1000.0 * vb.time_device_only(iters)}
        vb.close()
    out["speedup_device_only"] = out["reference_cuda_build"]["device_only_us"] / out["b200"]["device_only_us"]
    return out


# --------------------------------------------------------------------------------------------------
# instance frame splitting + compositing: HBM-bound byte work, so each gets a roofline line.
# 7 cars (configs[2]) at 1242x375; L2 flushed before every timed launch; CUDA events on the engine's stream.
# --------------------------------------------------------------------------------------------------
def run_frames_ops(local_rank, iters=30, ncars=7):
    import torch
    from dynslam_b200 import engine as E
    dev = torch.device("cuda", local_rank)
    W, H_ = synth.KITTI_W, synth.KITTI_H
    stream = torch.cuda.Stream(device=dev)
    peak, _ = peaks()
    with torch.cuda.stream(stream):
        eng = E.Engine(E.Scene(E.SceneParams(), 2048, 0x800, 0x400, device=f"cuda:{local_rank}"), (W, H_), stream=stream.cuda_stream)
        fr = E.InstanceFrames(eng)
        rng = np.random.default_rng(7)
        rgb0 = torch.from_numpy(rng.integers(0, 256, (H_, W, 4), dtype=np.uint8)).to(dev)
        depth0 = torch.from_numpy(rng.uniform(0.5, 20.0, (H_, W)).astype(np.float32)).to(dev)
        ops, keep = [], []
        for i in range(ncars):           # car-sized boxes (~150x90 px) spread over the frame, elliptical silhouettes
            bw, bh = 150 + 10 * i, 90 + 4 * i
            x0, y0 = 40 + i * 160, 120 + (i % 3) * 40
            yy, xx = np.mgrid[0:bh, 0:bw]
            m = ((((xx - bw / 2) / (bw / 2)) ** 2 + ((yy - bh / 2) / (bh / 2)) ** 2) <= 1.0).astype(np.uint8)
            t = torch.from_numpy(m).to(dev)
            mask = E.make_mask((x0, y0, x0 + bw - 1, y0 + bh - 1), t)
            drgb = torch.zeros((H_, W, 4), dtype=torch.uint8, device=dev)
            ddep = torch.zeros((H_, W), dtype=torch.float32, device=dev)
            keep.append((t, mask, drgb, ddep))
            ops.append((E.EXTRACT, mask, mask, drgb, ddep))
        flush = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
        rgb, depth = rgb0.clone(), depth0.clone()

        def timed(fn, reset, kernel):
            # Kernel time from the library's launch trace (an event pair right around the launch, b200_set_timing(3)); the
            # L2 flush is enqueued just before and nothing synchronises in between, so the launch never waits for the host.
            eng.set_timing(3)
            for it in range(iters + 3):
                reset()
                flush.add_(1)
                fn()
            stream.synchronize()
            d = [b - a for name, a, b in eng.trace() if name == kernel][3:]
            eng.set_timing(0)
            return sum(d) / len(d)

        def reset_split():
            rgb.copy_(rgb0); depth.copy_(depth0)
        us_split = timed(lambda: fr.ProcessSilhouettes(rgb, depth, ops, sync=False, wait_inputs=False), reset_split, "k_process_silhouettes")
        # algorithmic bytes: the frame is read once (8 B/px); every instance frame is written in full (8 B/px each, the
        # reference's two memsets + copies); the blanked pixels of the main frame are written back
        blanked = int((depth == 0).sum().item())
        bytes_split = W * H_ * 8 + ncars * W * H_ * 8 + blanked * 8
        layers = [(k[2], k[3], E.MATPLOTLIB2_PALETTE[i % 10]) for i, k in enumerate(keep)]
        out_c, out_d = rgb0.clone(), depth0.clone()

        def reset_cmp():
            out_c.copy_(rgb0); out_d.copy_(depth0)
        us_cmp = timed(lambda: fr.CompositeInstances(out_c, out_d, layers, dim_factor=0.10, tint_strength=1.0, wait_inputs=False), reset_cmp,
                       "k_composite_layers")
        # every layer's depth is read (4 B/px); its colour only where it wins; background read + written (16 B/px)
        bytes_cmp = W * H_ * 16 + ncars * W * H_ * 4
        eng.close()
    return {"instance_split": {"us": us_split, "ops": ncars, "alg_bytes": bytes_split, "achieved": bytes_split / us_split / 1e3,
                               "peak": peak, "unit": "GB/s", "frac": bytes_split / us_split / 1e3 / peak,
                               "what": "b200_process_silhouettes_async: 7 detections cut out of a 1242x375 frame into 7 instance frames, one launch"},
            "composite": {"us": us_cmp, "layers": ncars, "alg_bytes": bytes_cmp, "achieved": bytes_cmp / us_cmp / 1e3, "peak": peak,
                          "unit": "GB/s", "frac": bytes_cmp / us_cmp / 1e3 / peak,
                          "what": "b200_composite_instances (synchronous call): background dim + 7 layers z-composited, one launch"}}


def gen_frames_hires(seed, count):
    scene = synth.StreetScene(seed=seed, length_m=60.0)
    return [synth.kitti_frame(scene, f, zmax=8.0) for f in range(count)]


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
def run_own(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from dynslam_b200 import abi, engine as E

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H_ = synth.KITTI_W, synth.KITTI_H
    K, Wm = args.steps, args.warmup
    n_e2e = args.e2e_steps
    n_raw = args.e2e_raw_steps if world == 1 else 0
    total_frames = args.preroll + Wm + K + 3 + n_e2e + n_raw
    length_m = total_frames * 0.8 + 60.0
    seed = 6 + rank
    t_gen = time.perf_counter()
    frames = gen_frames_parallel(seed, 0, total_frames, length_m, max(1, min(16, (os.cpu_count() or 2) // max(world, 1))))
    log(f"[rank {rank}] generated {len(frames)} frames in {time.perf_counter() - t_gen:.1f}s")

    # Host buffers of the e2e phases are pinned NOW, long before they are used: pinning several hundred MB was followed, some
    # milliseconds later, by a one-off 50-70 ms host stall (seen as a single gap between two pipelined submissions, with the
    # clock sampler paused and the GC off), which a 50 ms e2e phase cannot absorb.
    e2e_first = args.preroll + Wm + K + 3
    h_depth = [torch.from_numpy(frames[e2e_first + i][0]).pin_memory() for i in range(n_e2e)]
    h_rgb = [torch.from_numpy(frames[e2e_first + i][1]).pin_memory() for i in range(n_e2e)]
    h_raw = [torch.from_numpy(np.round(frames[e2e_first + n_e2e + i][0] * 1000.0).astype(np.int16)).pin_memory() for i in range(n_raw)]
    h_rgb2 = [torch.from_numpy(frames[e2e_first + n_e2e + i][1]).pin_memory() for i in range(n_raw)]
    h_out = [torch.zeros(H_ * W * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        scene = E.Scene(E.SceneParams(), NUM_BLOCKS, NUM_BUCKETS, EXCESS, device=f"cuda:{local_rank}")
        eng = E.Engine(scene, (W, H_), stream=stream.cuda_stream)
        reco = E.SceneReconstructionEngine(eng)
        vis = E.VisualisationEngine(eng, scene)
        rs = vis.CreateRenderState((W, H_))
        reco.ResetScene(scene)
        points = torch.zeros(H_ * W * 4, dtype=torch.float32, device=dev)
        normals = torch.zeros(H_ * W * 4, dtype=torch.float32, device=dev)
        gather_buf = [torch.zeros(H_ * W * 4, dtype=torch.uint8, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
        flush_buf = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev) if args.flush_l2 else None   # 2x the 126 MB L2

        def dev_view(fr):
            depth, rgb, M, proj = fr
            d = torch.from_numpy(depth).to(dev, non_blocking=False)
            c = torch.from_numpy(rgb).to(dev, non_blocking=False)
            return E.View(d, c, M, proj)

        def step(view):
            eng.process_frame_async(rs, view, points, normals, decay=DECAY)
            if world > 1:
                dist.gather(rs.raycastImage, gather_buf, dst=0)

        diag = os.environ.get("B200_BENCH_DIAG", "")
        sampler = ClockSampler(local_rank)
        if "nosampler" not in diag:
            sampler.start()
        if "nogc" in diag:
            import gc
            gc.disable()
        # ---- pre-roll: build the map to steady state (untimed set-up) ----
        idx = 0
        for _ in range(args.preroll):
            step(dev_view(frames[idx])); idx += 1
        eng.sync(rs)
        views = [dev_view(frames[idx + i]) for i in range(Wm + K + 3)]
        idx += Wm + K + 3
        for i in range(Wm):
            step(views[i])
        eng.sync(rs)
        launches0 = eng.stats().launches
        blocks0 = eng.stats().totalIntegratedBlocks
        eng.set_timing(2)
        clk_first = len(sampler.rows)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)] if args.flush_l2 else []
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(K):
            if args.flush_l2:       # evict L2 between timed iterations; the flush itself is timed and subtracted
                flush_ev[i][0].record(stream)
                flush_buf.add_(1)
                flush_ev[i][1].record(stream)
            if args.profile_step == i:   # ncu --profile-from-start off: exactly this frame's kernels are captured
                torch.cuda.synchronize(dev)
                torch.cuda.cudart().cudaProfilerStart()
            step(views[Wm + i])
            if args.profile_step == i:
                torch.cuda.synchronize(dev)
                torch.cuda.cudart().cudaProfilerStop()
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        flush_ms = sum(a.elapsed_time(b) for a, b in flush_ev)
        gpu_ms = ev0.elapsed_time(ev1) - flush_ms
        wall_ms -= flush_ms
        eng.sync(rs)
        st = eng.stats()
        time.sleep(0.05)
        sampler.pause(True)
        clk_timed_end = len(sampler.rows)      # the sampler process stays alive (paused) through the e2e phases: terminating nvidia-smi stalls
                                               # the driver for tens of milliseconds, which used to land in the e2e timing
        launches = st.launches - launches0
        blocks = st.totalIntegratedBlocks - blocks0
        int_ms, int_n = st.ring_ms_integrate, st.ring_count
        n_vis = rs.noVisibleBlocks
        # sanity of the timed work: the last timed frame's raycast must have hit the surface on a large part of the image
        # (a broken expected-depth image makes every ray exit at once and the frame look fast)
        rays_hit = int((rs.raycastResult.view(-1, 4)[:, 3] > 0).sum().item())
        if rays_hit < 0.3 * W * H_ or n_vis < 1000:
            raise RuntimeError(f"bench sanity check failed: {rays_hit} of {W * H_} rays hit the surface, {n_vis} visible blocks")
        used_blocks = NUM_BLOCKS - 1 - scene.lastFreeBlockId
        decayed = reco.GetDecayedBlockCount()
        # per-stage breakdown of a few extra frames (per-frame sync; not part of the timed region)
        eng.set_timing(1)
        stage = np.zeros(6)
        for i in range(3):
            step(views[Wm + K + i])
            eng.sync(rs)
            s = eng.stats()
            stage += np.array([s.ms_allocate, s.ms_integrate, s.ms_expected, s.ms_raycast, s.ms_decay, s.ms_total])
        stage /= 3.0
        eng.set_timing(0)

        # ---- e2e: host buffers -> H2D -> frame -> D2H image, every step ----
        assert idx == e2e_first, (idx, e2e_first)
        ev = E.View(torch.zeros((H_, W), dtype=torch.float32, device=dev), torch.zeros((H_, W, 4), dtype=torch.uint8, device=dev),
                    frames[idx][2], frames[idx][3])
        e2e_warm = min(3, n_e2e // 2)
        e2e_marks = []
        t_e2e = 0.0
        for i in range(n_e2e):
            if i == e2e_warm:
                eng.host_frame_wait(0); eng.host_frame_wait(1)
                torch.cuda.synchronize(dev)
                if world > 1:
                    dist.barrier()
                t_e2e = time.perf_counter()
            slot = i & 1
            eng.host_frame_wait(slot)          # frame i-2 (same staging slot) has delivered its image
            e2e_marks.append(time.perf_counter())
            ev.set_pose(frames[idx + i][2])
            # public API: host depth+RGB in, grey raycast image out; copies of neighbouring frames overlap the kernels
            eng.host_frame_submit(rs, ev, h_depth[i], h_rgb[i], points, normals, decay=DECAY, h_out=h_out[slot], slot=slot)
            if world > 1:
                dist.gather(rs.raycastImage, gather_buf, dst=0)
        eng.host_frame_wait(0); eng.host_frame_wait(1)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t_e2e = time.perf_counter() - t_e2e
        e2e_frames = n_e2e - e2e_warm

        # ---- e2e from RAW sensor frames: int16 depth + RGB in, UpdateView (conversion + 5-pass bilateral filter) on the
        # device, fused frame, grey image out (what DynSLAM does per frame from InfiniTamDriver::UpdateView onwards) ----
        e2e_raw = None
        if n_raw > 3:
            calib = E.make_view_calib()
            base = idx + n_e2e
            t_raw = 0.0
            raw_marks = []
            for i in range(n_raw):
                if i == 3:
                    eng.host_frame_wait(0); eng.host_frame_wait(1)
                    torch.cuda.synchronize(dev)
                    t_raw = time.perf_counter()
                slot = i & 1
                eng.host_frame_wait(slot)
                raw_marks.append(time.perf_counter())
                ev.set_pose(frames[base + i][2])
                eng.host_frame_submit_raw(rs, ev, h_raw[i], h_rgb2[i], calib, points, normals, decay=DECAY, h_out=h_out[slot], slot=slot)
            eng.host_frame_wait(0); eng.host_frame_wait(1)
            torch.cuda.synchronize(dev)
            t_raw = time.perf_counter() - t_raw
            e2e_raw = {"frame_ms": frame_gaps(raw_marks[3:]), "value": (n_raw - 3) / t_raw, "unit": "frames/s", "h2d_bytes_per_step": W * H_ * 6, "d2h_bytes_per_step": W * H_ * 4,
                       "steps": n_raw - 3, "what": "raw int16 depth + RGB in -> UpdateView with bilateral filter -> fused frame -> image out"}

    sampler.stop()
    clk_all = len(sampler.rows)
    sampler.rows = sampler.rows[max(clk_first - 1, 0):clk_timed_end] or sampler.rows   # samples taken during the timed region
    clocks = sampler.summary()
    clocks["samples_since_start"] = clk_all

    # ---- reduce over ranks (max time) ----
    ms = max(gpu_ms, 0.0)
    if world > 1:
        t = torch.tensor([ms, wall_ms, t_e2e, float(blocks), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, wall_ms, t_e2e = float(tmax[0]), float(tmax[1]), float(tmax[2])
        blocks_all, launches_all = float(tsum[3]), float(tsum[4])
    else:
        blocks_all, launches_all = float(blocks), float(launches)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fps = world * K / (ms / 1000.0)
    mvox = blocks_all * 512 / (ms / 1000.0) / 1e6
    peak, peak_src = peaks()
    alg_bytes = blocks * BYTES_PER_BLOCK + int_n * W * H_ * 8
    achieved = alg_bytes / (int_ms / 1000.0) / 1e9 if int_ms > 0 else 0.0
    footprint_mb = (n_vis * 4096 * 2 + (NUM_BUCKETS + EXCESS) * 21 + W * H_ * (8 + 8 + 16 + 4 + 32)) / 1e6

    cpu = None
    if args.cpu_steps > 0:
        try:
            c = cpu_run(6, args.cpu_preroll, 1, args.cpu_steps, 120.0, omp=True)
            cpu = {"value": c["fps"], "unit": "frames/s", "cores": c["cores"], "kind": "port",
                   "sample": f"{c['frames']} frames after a {c['preroll']}-frame CPU pre-roll of the same stream (smaller map than "
                             f"the GPU's {args.preroll}-frame pre-roll, which favours the CPU); {c['mvoxels_per_s']:.1f} Mvoxels/s",
                   "ms_per_step": c["ms_per_step"]}
        except Exception as ex:  # the baseline is reported, never required for the GPU numbers
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    itm = None
    if world == 1 and args.harness_frames > 0:
        try:
            itm = run_itm_harness(frames, args.harness_preroll, args.harness_frames)
        except Exception as ex:
            itm = {"error": str(ex)}

    hires = None
    if world == 1 and args.hires_frames > 0:
        try:
            hires = run_hires(local_rank, args.hires_frames)
        except Exception as ex:
            hires = {"error": str(ex)}

    vbuild = None
    if world == 1 and args.harness_frames > 0:
        try:
            vbuild = run_view_builder(frames)
        except Exception as ex:
            vbuild = {"error": str(ex)}

    frames_ops = None
    if world == 1 and args.harness_frames > 0:
        try:
            frames_ops = run_frames_ops(local_rank)
        except Exception as ex:
            frames_ops = {"error": str(ex)}

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "KITTI-odometry-06-shaped 1242x375 static-map fusion+raycast loop (configs[1]); one volume per GPU",
                   "voxel_m": 0.05, "mu_m": 0.75, "maxW": 50, "blocks": NUM_BLOCKS, "buckets": NUM_BUCKETS, "excess": EXCESS,
                   "decay": {"maxWeight": DECAY[0], "minAge": DECAY[1]}, "preroll_frames": args.preroll,
                   "l2": (f"explicit flush between timed steps (256 MB read-modify-write, timed with CUDA events and subtracted, "
                          f"{flush_ms / max(K, 1) * 1000:.0f} us each); per-step footprint ~{footprint_mb:.0f} MB") if args.flush_l2 else
                         f"no flush (--no-flush-l2): per-step footprint ~{footprint_mb:.0f} MB, consecutive frames reuse L2",
                   "integrate_impl": os.environ.get("B200_INTEGRATE_IMPL", "v3"),
                   "parallelism": f"{world} independent volume(s), NCCL gather of raycast images to rank 0" if world > 1 else "1 volume"},
        "mvoxels_per_s": mvox, "rays_hit": rays_hit, "visible_blocks": n_vis, "allocated_blocks": used_blocks, "decayed_blocks": int(decayed),
        "wall_ms_per_step": wall_ms / K,
        "stage_ms": {"allocate": stage[0], "integrate": stage[1], "expected_depths": stage[2], "raycast_icp": stage[3],
                     "decay": stage[4], "total": stage[5]},
        "roofline": {"kernel": "k_integrate_" + os.environ.get("B200_INTEGRATE_IMPL", "v3"), "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": ncu_traffic(),
                     "peak_source": peak_src, "launches_timed": int_n, "mean_launch_us": 1000.0 * int_ms / max(int_n, 1),
                     "alg_bytes_per_launch": alg_bytes / max(int_n, 1)},
        "roofline_hires": hires,
        "cpu_baseline": cpu,
        "itmlib_harness": itm,
        "view_builder": vbuild,
        "frames_ops": frames_ops,
        "e2e_raw": e2e_raw,
        "e2e": {"value": world * e2e_frames / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": W * H_ * 8,
                "d2h_bytes_per_step": W * H_ * 4, "steps": e2e_frames, "frame_ms": frame_gaps(e2e_marks[e2e_warm:])},
        "gpu_launches": int(launches_all),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--preroll", type=int, default=230, help="untimed frames that build the map (> decay minAge)")
    ap.add_argument("--e2e-steps", type=int, default=203)
    ap.add_argument("--e2e-raw-steps", type=int, default=103, help="frames of the raw-sensor-frame e2e variant (1 GPU only)")
    ap.add_argument("--flush-l2", dest="flush_l2", action="store_true", default=True)
    ap.add_argument("--no-flush-l2", dest="flush_l2", action="store_false")
    ap.add_argument("--profile-step", type=int, default=-1,
                    help="bracket this timed step with cudaProfilerStart/Stop (for ncu --profile-from-start off; not a bench run)")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--harness-frames", type=int, default=100, help="frames timed through the real ITMLib objects (0 = skip)")
    ap.add_argument("--harness-preroll", type=int, default=60)
    ap.add_argument("--hires-frames", type=int, default=24, help="frames of the 4 mm roofline-stress stream (0 = skip)")
    ap.add_argument("--cpu-preroll", type=int, default=12)
    ap.add_argument("--ref-preroll", type=int, default=12)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

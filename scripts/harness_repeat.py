"""Fresh-process timing of the synchronous ITMLib loop (oracle/_ref/libitmharness.so): the UNMODIFIED reference CUDA engines
(sm_100a, --use_fast_math) and the B200 shim behind the same abstract interfaces, same frames.
No torch in this process: nothing but the harness has touched the CUDA allocator (the reference number moved
4.6x between a run inside bench.py's process and a stand-alone one). Prints ONE JSON object:
  {impl: {"fps": [r0, r1, r2], "median_fps", "spread", "stages_us": {...}}, "speedup_median"}.
usage: python scripts/harness_repeat.py [frames=160] [preroll=60] [repeats=3]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynslam_b200 import synth  # noqa: E402
from tests import harnesslib as HL  # noqa: E402

NUM_BLOCKS = 0x60000
DECAY = (1, 200)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    preroll = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    decay_age = int(sys.argv[4]) if len(sys.argv) > 4 else DECAY[1]
    import bench
    frames = bench.gen_frames_parallel(6, 0, n, n * 0.8 + 60.0, min(16, os.cpu_count() or 1))
    L = HL.lib()
    for f in frames:
        L.harness_pin(f[0].ctypes.data, f[0].nbytes)
        L.harness_pin(f[1].ctypes.data, f[1].nbytes)
    out = {"frames_timed": n - preroll, "preroll": preroll, "repeats": repeats, "decay": [DECAY[0], decay_age]}
    names = ["h2d", "allocate", "integrate", "expected_depths", "icp_maps", "decay"]
    for name, impl in (("reference_cuda_build", HL.REFERENCE_CUDA), ("b200_itm_shim", HL.B200_SHIM)):
        fps, stages = [], None
        for r in range(repeats + 1):            # the last pass is the per-stage one (synchronises after every call)
            hs = HL.Harness(impl, synth.KITTI_W, synth.KITTI_H, frames[0][3], numBlocks=NUM_BLOCKS)
            st = (C.c_double * 6)()
            t0 = 0.0
            for i in range(n):
                if i == preroll:
                    hs.sync()
                    t0 = time.perf_counter()
                if r == repeats and i >= preroll:
                    hs.process_frame_timed(frames[i][0], frames[i][1], frames[i][2], st, decay=(DECAY[0], decay_age))
                else:
                    hs.process_frame(frames[i][0], frames[i][1], frames[i][2], decay=(DECAY[0], decay_age))
            hs.sync()
            dt = time.perf_counter() - t0
            vis = hs.counters()["noVisibleBlocks"]
            hs.close()
            if r < repeats:
                fps.append((n - preroll) / dt)
            else:
                stages = {k: st[j] / (n - preroll) for j, k in enumerate(names)}
        med = float(np.median(fps))
        out[name] = {"fps": fps, "median_fps": med, "spread": (max(fps) - min(fps)) / med, "stages_us": stages, "visible_blocks": vis}
    out["speedup_median"] = out["b200_itm_shim"]["median_fps"] / out["reference_cuda_build"]["median_fps"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Generates tests/golden/kitti_small.json: per-frame digests of every piece of state of the golden
sequence (tests/parity.py GOLDEN_CFG). Run in the build container, where /root/reference is mounted:
every frame's integration result is recomputed block by block with the REFERENCE's own
ComputeUpdatedVoxelInfo (oracle/_ref/libitmref.so) and asserted identical before the digests are
written, and tests/test_oracle_vs_ref.py pins marking, visibility, projection, raycast, shading and
ICP the same way. The committed file then travels to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import hostlib as H  # noqa: E402
from tests import parity as P  # noqa: E402

if __name__ == "__main__":
    assert H.ref_available(), "oracle/_ref/libitmref.so missing: run oracle/build_ref.sh where /root/reference exists"
    cfg = P.Cfg(**P.GOLDEN_CFG)
    digests = P.oracle_sequence_digests(cfg, ref_check=True)
    out = {"cfg": {k: (list(v) if isinstance(v, tuple) else v) for k, v in P.GOLDEN_CFG.items()},
           "certified_by": "oracle/_ref/libitmref.so (reference DeviceAgnostic functions, g++ -O2 -ffp-contract=off)",
           "frames": digests}
    path = os.path.join(ROOT, "tests", "golden", "kitti_small.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, "frames:", len(digests), "visible blocks last frame:", digests[-1]["counters"][2])

"""GPU: b200_evaluate_depth (csrc/evalc.cu) against the CPU oracle (oracle/eval_oracle.c, pinned to the reference's
Evaluation::EvaluateDepth by tests/test_eval_oracle.py): every counter of every callback must be equal, at the KITTI frame size and
LIDAR density, with and without the static / dynamic association image; the negative-disparity exception maps to RuntimeError."""
import numpy as np
import pytest
import torch

from dynslam_b200 import abi, engine as E
from tests import evallib as V

pytestmark = pytest.mark.gpu
CBS = E.Evaluation.default_callbacks()


def make_eval(eng, w, h, rigt, min_depth=0.5, max_depth=30.0):
    v, pl, pr, b = rigt
    return E.Evaluation(eng, v, pl, pr, b, w, h, min_depth, max_depth)


@pytest.mark.parametrize("seed,n,with_assoc", [(11, 120000, False), (12, 120000, True), (13, 257, True), (14, 0, False)])
def test_evaluate_depth_equals_oracle(seed, n, with_assoc):
    w, h = 1242, 375
    p, rigt = V.params(w, h)
    pts = V.lidar_cloud(max(n, 1), seed, rigt, w, h)[:n]
    rendered, inp = V.depth_images(w, h, seed + 100)
    assoc = None
    if with_assoc:
        assoc = (np.random.default_rng(seed).uniform(size=(h, w)) * 3).astype(np.uint8)
        assoc[:, : w // 3] = abi.EVAL_STATIC
    rc, st_o, dy_o, summ = V.run_oracle(p, pts, rendered, inp, CBS, assoc, with_dynamic=with_assoc)
    assert rc == 0
    dev = torch.device("cuda:0")
    scene = E.Scene(E.SceneParams(), 4096, 0x4000, 0x1000, device="cuda:0")
    eng = E.Engine(scene, (w, h))
    ev = make_eval(eng, w, h, rigt)
    d_pts = torch.from_numpy(pts.reshape(-1, 4)).to(dev) if n else torch.zeros((0, 4), dtype=torch.float32, device=dev)
    st, dy, summary = ev.EvaluateDepth(d_pts, torch.from_numpy(rendered).to(dev), torch.from_numpy(inp).to(dev), CBS,
                                       association=torch.from_numpy(assoc).to(dev) if with_assoc else None, with_dynamic=with_assoc)
    assert st == st_o and dy == dy_o
    assert summary["valid_lidar_points"] == summ.valid_lidar_points and summary["epi_errors"] == summ.epi_errors
    assert summary["skipped_lidar_points"] == summ.skipped_lidar_points
    if n >= 100000:
        assert st[0]["measurement_count"] > 20000
    # a second call on the same engine starts from zero
    st2, _, _ = ev.EvaluateDepth(d_pts, torch.from_numpy(rendered).to(dev), torch.from_numpy(inp).to(dev), CBS,
                                 association=torch.from_numpy(assoc).to(dev) if with_assoc else None, with_dynamic=with_assoc)
    assert st2 == st_o


def test_negative_disparity_raises_like_the_reference():
    w, h = 310, 94
    p, rigt = V.params(w, h)
    v, pl, pr, b = rigt
    pr = pr.copy(); pr[0, 3] = pl[0, 3] + 50.0
    dev = torch.device("cuda:0")
    scene = E.Scene(E.SceneParams(), 4096, 0x4000, 0x1000, device="cuda:0")
    eng = E.Engine(scene, (w, h))
    ev = E.Evaluation(eng, v, pl, pr, b, w, h, 0.5, 30.0)
    pts = V.lidar_cloud(500, 3, rigt, w, h)
    rendered, inp = V.depth_images(w, h, 4)
    with pytest.raises(RuntimeError, match="Negative disparity"):
        ev.EvaluateDepth(torch.from_numpy(pts).to(dev), torch.from_numpy(rendered).to(dev), torch.from_numpy(inp).to(dev))

"""Experimental pixel-major RIC producer (DSU_RIC_PIXEL_MAJOR=1 -> conv_ric_persist_kernel<4>, ric_producer.cuh ric_produce_px):
it must reproduce the default producer bit for bit (same arithmetic, same order), then the per-layer table of both.
NOT validated on hardware yet (written after the round-1 GPU budget was spent) - run this first.

    python tools/ric_px_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import synth  # noqa: E402
from drawingspinup_b200.pipeline import DEFAULT_ARGS  # noqa: E402

dev = torch.device("cuda:0")
m = dsu.GeneratorJ_RIC(precision="fp16", **DEFAULT_ARGS)
m.load_state_dict(synth.to_torch_state_dict(synth.make_state_dict(1, out_gain=0.25)))
m = m.to(dev).eval()
rng = np.random.default_rng(0)
for (b, h, w) in [(1, 4, 4), (2, 8, 12), (5, 20, 36), (1, 132, 68), (3, 64, 48), (2, 512, 512)]:
    x = torch.from_numpy(rng.uniform(-1, 1, (b, 6, h, w)).astype(np.float32)).to(dev)
    ys = []
    for mode in ("0", "1"):
        os.environ["DSU_RIC_PIXEL_MAJOR"] = mode       # read per launch by the planner
        with torch.no_grad():
            ys.append(m(x).clone())
    torch.cuda.synchronize()
    same = torch.equal(ys[0], ys[1])
    print("shape %-14s pixel-major == default: %s (max|d| %.2e)" % ((b, h, w), same, (ys[0] - ys[1]).abs().max().item()), flush=True)
c, p, _ = synth.make_frames(16, 512, 512, seed=3)
cd, pd = torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)
for mode in ("0", "1"):
    os.environ["DSU_RIC_PIXEL_MAJOR"] = mode
    with torch.no_grad():
        for _ in range(3):
            m.forward_frames(cd, pd, None)
        rows = m.profile_layers(16, 512, 512, reps=5)
    pick = [(n, round(ms, 3)) for n, ms, _ in rows if n in ("conv1", "resnets.0.conv_0", "upconv2", "upconv1", "conv_11", "conv_11_a.3")]
    print("DSU_RIC_PIXEL_MAJOR=%s" % mode, pick, "stage-1 total %.3f ms" % sum(ms for _, ms, _ in rows), flush=True)

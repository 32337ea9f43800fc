"""Experimental fused stage-1 conv0 (DSU_RIC_FIRST=1 -> conv_ric_first.cu) against the default ric_expand + 1x1 path:
conv0 activations and final output on ragged shapes (same fp32 blend and fp16 rounding, so conv0 should agree to
accumulation order), then the per-layer table.  NOT validated on hardware yet - run this first.

    python tools/ric_first_check.py
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
    out = {}
    for mode in ("0", "1"):
        os.environ["DSU_RIC_FIRST"] = mode             # read per forward by the planner
        with torch.no_grad():
            y = m(x).cpu()
        out[mode] = (y, m.debug_buffer(0, 0, (b, h, w, 40)).float()[..., :32])
    d0 = (out["0"][1] - out["1"][1]).abs().max().item()
    dy = (out["0"][0] - out["1"][0]).abs().max().item()
    print("shape %-14s conv0 |default - fused| %.2e  output %.2e  %s" % ((b, h, w), d0, dy, "OK" if d0 <= 2 ** -7 and dy < 5e-3 else "FAIL"), flush=True)
c, p, _ = synth.make_frames(16, 512, 512, seed=3)
cd, pd = torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)
for mode in ("0", "1"):
    os.environ["DSU_RIC_FIRST"] = mode
    with torch.no_grad():
        for _ in range(3):
            m.forward_frames(cd, pd, None)
        rows = m.profile_layers(16, 512, 512, reps=5)
    print("DSU_RIC_FIRST=%s" % mode, [(n, round(ms, 3)) for n, ms, _ in rows[:3]], "stage-1 total %.3f ms" % sum(ms for _, ms, _ in rows), flush=True)

"""Experimental sub-pixel up-convolutions (DSU_SUBPIXEL=1, engine.cu add_up / conv_halo_persist_kernel<true>) on the GPU:
parity of stage 2 against the oracle on ragged shapes in both precisions, then the per-layer table next to the default
plan (B=16, 512x512).  NOT validated on hardware yet (written after the round-1 GPU budget was spent) - run this first.

    python tools/subpixel_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import synth  # noqa: E402
from drawingspinup_b200.pipeline import DEFAULT_ARGS  # noqa: E402
from oracle import reference_port as rp  # noqa: E402

dev = torch.device("cuda:0")
sd = synth.to_torch_state_dict(synth.make_state_dict(2, out_gain=0.25))


def model(prec, sub):
    os.environ["DSU_SUBPIXEL"] = "1" if sub else "0"      # read by dsu_create (plan time)
    m = dsu.GeneratorJ(precision=prec, **DEFAULT_ARGS)
    m.load_state_dict(sd)
    return m.to(dev).eval()


rng = np.random.default_rng(0)
for prec, tol in (("fp16x3", 1e-3), ("fp16", 2.5e-2)):
    m = model(prec, True)
    for (b, h, w) in [(1, 4, 4), (2, 8, 12), (5, 20, 36), (1, 132, 68), (2, 64, 48)]:
        x = torch.from_numpy(rng.uniform(-1, 1, (b, 6, h, w)).astype(np.float32))
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
            ref = rp.generator_j_forward(sd, x)
        err = (y - ref).abs().max().item()
        print("sub-pixel %-6s shape %-14s max|err| %.2e %s" % (prec, (b, h, w), err, "OK" if err < tol else "FAIL"), flush=True)

c, p, e = synth.make_frames(16, 512, 512, seed=3)
cd, pd, ed = (torch.from_numpy(t).to(dev) for t in (c, p, e))
for sub in (False, True):
    m = model("fp16", sub)
    with torch.no_grad():
        for _ in range(3):
            out = m.forward_frames(cd, pd, ed)
        rows = m.profile_layers(16, 512, 512, reps=5)
    up = [(n, round(ms, 3)) for n, ms, _ in rows if n.startswith("upconv")]
    print("DSU_SUBPIXEL=%d" % sub, up, "stage-2 total %.3f ms" % sum(ms for _, ms, _ in rows), flush=True)
    if sub:
        print("uint8 frames equal to the default plan within 1 LSB:", int((out.int() - base.int()).abs().max()) <= 1)
    else:
        base = out

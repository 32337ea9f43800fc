"""conv_first.cu vs the tap-mode kernel on GeneratorJ.conv0 (development aid, run under gpurun):
activations of both paths on ragged shapes, then conv0 timing of a 16 x 512 x 512 batch over the knobs
DSU_FIRST (0 = tap mode), DSU_FIRST_KS (issuers), DSU_FIRST_SETS (accumulator sets)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import synth  # noqa: E402
from drawingspinup_b200.pipeline import DEFAULT_ARGS  # noqa: E402

dev = torch.device("cuda:0")
sd = synth.to_torch_state_dict(synth.make_state_dict(2, out_gain=0.25))
m = dsu.GeneratorJ(precision="fp16", **DEFAULT_ARGS)
m.load_state_dict(sd)
m = m.to(dev).eval()
rng = np.random.default_rng(0)
for (b, h, w) in [(1, 4, 4), (2, 8, 12), (1, 12, 4), (5, 20, 36), (1, 132, 68), (33, 16, 16), (2, 64, 48), (3, 512, 512)]:
    x = torch.from_numpy(rng.uniform(-1, 1, (b, 6, h, w)).astype(np.float32)).to(dev)
    out = {}
    for mode in ("0", "1"):
        os.environ["DSU_FIRST"] = mode
        with torch.no_grad():
            y = m(x)
        torch.cuda.synchronize()
        out[mode] = (y.cpu(), m.debug_buffer(0, 0, (b, h, w, 40)).float()[..., :32])
    dy = (out["0"][0] - out["1"][0]).abs().max().item()
    ds = (out["0"][1] - out["1"][1]).abs().max().item()
    print("shape %-14s conv0 |tap - first| %.2e  output %.2e  %s" % ((b, h, w), ds, dy, "OK" if ds < 2e-2 and dy < 5e-3 else "FAIL"), flush=True)
c, p, e = synth.make_frames(16, 512, 512, seed=3)
cd, pd, ed = (torch.from_numpy(t).to(dev) for t in (c, p, e))
for mode, ks, sets in (("0", "2", "4"), ("1", "1", "2"), ("1", "1", "4"), ("1", "2", "2"), ("1", "2", "4"), ("1", "4", "2"), ("1", "4", "4")):
    os.environ.update(DSU_FIRST=mode, DSU_FIRST_KS=ks, DSU_FIRST_SETS=sets)
    with torch.no_grad():
        for _ in range(3):
            m.forward_frames(cd, pd, ed)
        rows = m.profile_layers(16, 512, 512, reps=5)
    print("DSU_FIRST=%s KS=%s SETS=%s" % (mode, ks, sets), [(n, round(ms, 3)) for n, ms, _ in rows[:2]], "total %.3f ms" % sum(ms for _, ms, _ in rows))

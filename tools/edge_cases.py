"""Edge-case sweep on the GPU (development aid): tiny frames, ragged sizes, large batch, 1024x1024."""
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
models = {}
for stage, cls in ((1, dsu.GeneratorJ_RIC), (2, dsu.GeneratorJ)):
    sd = synth.to_torch_state_dict(synth.make_state_dict(stage, out_gain=0.25))
    for prec in ("fp16x3", "fp16"):
        m = cls(precision=prec, **DEFAULT_ARGS)
        m.load_state_dict(sd)
        models[(stage, prec)] = (m.to(dev).eval(), sd)

rng = np.random.default_rng(0)
for (b, h, w) in [(1, 4, 4), (2, 8, 12), (1, 12, 4), (5, 20, 36), (1, 132, 68), (33, 16, 16)]:
    x = torch.from_numpy(rng.uniform(-1, 1, (b, 6, h, w)).astype(np.float32))
    for stage in (1, 2):
        m, sd = models[(stage, "fp16x3")]
        with torch.no_grad():
            y = m(x.to(dev)).cpu()
            ref = rp.generator_j_ric_forward(sd, x, use_torchvision=True) if stage == 1 else rp.generator_j_forward(sd, x)
        err = (y - ref).abs().max().item()
        print("shape %-14s stage %d fp16x3 max|err| %.2e %s" % ((b, h, w), stage, err, "OK" if err < 1e-3 else "FAIL"), flush=True)
# 1024x1024 (BASELINE configs[4]): finite, deterministic, batch-invariant
c, p, e = synth.make_frames(2, 1024, 1024, seed=3)
m, _ = models[(1, "fp16")]
with torch.no_grad():
    cd, pd = torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)
    o2, y2 = m.forward_frames(cd, pd, None, return_float=True)
    o1, y1 = m.forward_frames(cd[1:2], pd[1:2], None, return_float=True)
torch.cuda.synchronize()
print("1024x1024 stage1 fp16: finite", bool(torch.isfinite(y2).all()), "batch-invariant", bool(torch.equal(y2[1:2], y1)),
      "alpha exact", bool(torch.equal(o2[..., 3].cpu(), torch.from_numpy(c[..., 3]))))

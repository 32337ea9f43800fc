"""First-contact check of the tensor-memory RIC kernel (conv_ric_tm.cu): stage 1 against the CPU oracle on small and ragged
shapes, both precisions, then the per-layer table at the benchmark shape.   python tools/tm_check.py [quick]"""
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
sd = synth.to_torch_state_dict(synth.make_state_dict(1, out_gain=0.25))
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
shapes = [(1, 8, 16), (1, 16, 32), (2, 20, 36), (1, 132, 68), (3, 64, 48)]
for prec in ("fp16", "fp16x3"):
    m = dsu.GeneratorJ_RIC(precision=prec, **DEFAULT_ARGS)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    for (b, h, w) in shapes:
        color, pos, edge = synth.make_frames(b, h, w, seed=h + w)
        x = torch.from_numpy(np.stack([rp.frame_to_tensor(color[i], pos[i])[0] for i in range(b)]))
        try:
            with torch.no_grad():
                y = m(x.to(dev)).cpu()
        except Exception as exc:
            print("FAILED", prec, (b, h, w), str(exc).splitlines()[0][:200])
            print("watchdog (warp, tag, a, b, block):", m.watchdog_records(), flush=True)
            sys.exit(3)
        with torch.no_grad():
            ref = rp.generator_j_ric_forward(sd, x, use_torchvision=True)
        err = (y - ref).abs().max().item()
        print("%-6s shape %-14s max|dy| %.3e %s" % (prec, (b, h, w), err, "OK" if err < (1e-3 if prec == "fp16x3" else 2.5e-2) else "BAD"), flush=True)
    if quick:
        continue
    c, p, _ = synth.make_frames(16, 512, 512, seed=3)
    cd, pd = torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)
    with torch.no_grad():
        for _ in range(3):
            m.forward_frames(cd, pd, None)
        rows = m.profile_layers(16, 512, 512, reps=5)
    print("stage 1 %s: %.3f ms per 16 frames" % (prec, sum(ms for _, ms, _ in rows)))
    for n, ms, fl in rows:
        print("   %-22s %7.3f ms %7.0f TF" % (n, ms, fl / ms / 1e9 if ms > 0 else 0), flush=True)

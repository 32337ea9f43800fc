"""One forward of one stage at B=16, 512x512 (for `ncu -k regex:... -s N -c 1` captures).
    python tools/ncu_target.py <stage> [precision] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import synth  # noqa: E402
from drawingspinup_b200.pipeline import DEFAULT_ARGS  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cls = dsu.GeneratorJ if stage == 2 else dsu.GeneratorJ_RIC
m = cls(precision=prec, **DEFAULT_ARGS)
m.load_state_dict(synth.to_torch_state_dict(synth.make_state_dict(stage, out_gain=0.25)))
m = m.to("cuda:0").eval()
c, p, e = synth.make_frames(16, 512, 512, seed=1)
c, p, e = torch.from_numpy(c).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(e).cuda()
with torch.no_grad():
    for _ in range(reps):
        out = m.forward_frames(c, p, e if stage == 2 else None)
torch.cuda.synchronize()
print("done", out.shape)

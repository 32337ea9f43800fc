import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu
from drawingspinup_b200 import synth
from drawingspinup_b200.pipeline import DEFAULT_ARGS
from oracle import reference_port as rp
dev = torch.device("cuda:0")
sd = synth.to_torch_state_dict(synth.make_state_dict(1, out_gain=0.25))
for (b, h, w) in [(1, 64, 48), (1, 256, 256)]:
    color, pos, edge = synth.make_frames(b, h, w, seed=7)
    x = torch.from_numpy(np.stack([rp.frame_to_tensor(color[i], pos[i])[0] for i in range(b)]))
    taps = {}
    with torch.no_grad():
        rp.generator_j_ric_forward(sd, x, use_torchvision=True, taps=taps)
    for ni in (6, 1):
        m = dsu.GeneratorJ_RIC(precision="fp16", **DEFAULT_ARGS); m.load_state_dict(sd); m = m.to(dev).eval()
        m.set_knob("tm_ni", ni, device=dev)
        with torch.no_grad():
            y = m(x.to(dev)); torch.cuda.synchronize()
        sk0 = m.debug_buffer(0, 0, (b, h, w, 40)).float()
        got = sk0[..., :32].permute(0, 3, 1, 2)
        ref = taps["conv0"]
        err = (got - ref).abs()
        print("shape", (b, h, w), "ni", ni, "conv0 |got| max %.3f mean %.4f nonzero frac %.3f; err max %.3f; frac of pixels with err>0.05: %.3f" % (
            got.abs().max().item(), got.abs().mean().item(), (got != 0).float().mean().item(), err.max().item(), (err.amax(1) > 0.05).float().mean().item()))
        bad = (err.amax(1)[0] > 0.05)
        ys, xs = torch.nonzero(bad, as_tuple=True)
        if len(ys):
            print("   bad rows range", ys.min().item(), ys.max().item(), "cols", xs.min().item(), xs.max().item(), "count", len(ys))
            # pattern by (y%8, x%16)
            pat = torch.zeros(8, 16)
            for yy, xx in zip(ys.tolist(), xs.tolist()): pat[yy % 8, xx % 16] += 1
            print("   bad count by tile-row (y%8):", pat.sum(1).int().tolist())
        o1 = m.debug_buffer(2, 0, (b, h // 2, w // 2, 64)).float().permute(0, 3, 1, 2)
        e1 = (o1 - taps["conv1"]).abs()
        print("   conv1 err max %.3f; by channel half: %.3f %.3f; nonzero frac %.3f; bad-pixel frac %.3f" % (
            e1.max().item(), e1[:, :32].max().item(), e1[:, 32:].max().item(), (o1 != 0).float().mean().item(), (e1.amax(1) > 0.05).float().mean().item()))
        badm = e1.amax(1)[0] > 0.05
        ys, xs = torch.nonzero(badm, as_tuple=True)
        if len(ys):
            pat = torch.zeros(8, 16)
            for yy, xx in zip(ys.tolist(), xs.tolist()): pat[yy % 8, xx % 16] += 1
            print("   conv1 bad by x%16:", pat.sum(0).int().tolist(), " by y%8:", pat.sum(1).int().tolist())
            ch_bad = (e1[0] > 0.05).float().mean((1, 2))
            print("   conv1 bad frac per channel (first 8, 32..39):", [round(v, 2) for v in ch_bad[:8].tolist()], [round(v, 2) for v in ch_bad[32:40].tolist()])
        p0 = m.debug_buffer(1, 0, (b, h // 2, w // 2, 32)).float().permute(0, 3, 1, 2)
        import torch.nn.functional as F
        p0_ref = F.max_pool2d(taps["conv0"], 2, 2)
        print("   P0 (maxpool) err max %.4f" % (p0 - p0_ref).abs().max().item())

"""Layer-by-layer GPU diagnostic: engine activations vs the CPU oracle (run under gpurun).

    python tools/gpu_check.py [H W B]

Prints, per stage and precision, the max-abs error of every observable activation buffer and of
the final output against the fp32 oracle (oracle/reference_port.py).  Development aid; the
asserted versions of these checks live in tests/test_gpu_parity.py.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import synth  # noqa: E402
from oracle import reference_port as rp  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 48
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ARGS = dict(use_bias=False, tanh=True, append_smoothers=True, resnet_blocks=7,
            filters=[32, 64, 128, 128, 128, 64], input_channels=6)
BUF = dict(SK0=0, P0=1, O1=2, P1=3, O2=4, T=5, U=6, V2=7, V1=8, C11=9, S0=10)


def nhwc(m, buf, b, h, w, c, exact):
    if exact and isinstance(m, dsu.GeneratorJ_RIC):      # stage 1 keeps fp32 activations in split-fp16 mode
        return m.debug_buffer(BUF[buf], 0, (b, h, w, c), torch.float32).permute(0, 3, 1, 2)
    t = m.debug_buffer(BUF[buf], 0, (b, h, w, c)).float()
    if exact:
        t = t + m.debug_buffer(BUF[buf], 1, (b, h, w, c)).float()
    return t.permute(0, 3, 1, 2)


def main():
    dev = torch.device("cuda:0")
    print("device:", torch.cuda.get_device_name(0), "| frames", (B, H, W))
    color, pos, edge = synth.make_frames(B, H, W, seed=7)
    stages = ((2, dsu.GeneratorJ), (1, dsu.GeneratorJ_RIC))
    if os.environ.get("CHECK_STAGE"):
        stages = tuple(s for s in stages if s[0] == int(os.environ["CHECK_STAGE"]))
    for stage, cls in stages:
        sd = synth.to_torch_state_dict(synth.make_state_dict(stage, seed=1234, out_gain=0.25))
        x = torch.from_numpy(np.stack([rp.frame_to_tensor(color[i], pos[i], edge[i] if stage == 2 else None)[0]
                                       for i in range(B)]))
        taps = {}
        with torch.no_grad():
            t0 = time.time()
            if stage == 2:
                y_ref = rp.generator_j_forward(sd, x, taps=taps)
            else:
                y_ref = rp.generator_j_ric_forward(sd, x, use_torchvision=True, taps=taps)
            t_cpu = time.time() - t0
        for prec in ("fp16x3", "fp16"):
            m = cls(precision=prec, **ARGS)
            m.load_state_dict(sd)
            m = m.to(dev).eval()
            with torch.no_grad():
                y = m(x.to(dev))
                torch.cuda.synchronize()
                y = y.cpu()
            ex = prec == "fp16x3"
            f = ARGS["filters"]
            rows = [("conv0", nhwc(m, "SK0", B, H, W, f[0] + 8, ex)[:, :f[0]], taps["conv0"]),
                    ("x(in)", nhwc(m, "SK0", B, H, W, f[0] + 8, ex)[:, f[0]:f[0] + 6], x),
                    ("conv1", nhwc(m, "O1", B, H // 2, W // 2, f[1], ex), taps["conv1"]),
                    ("conv2", nhwc(m, "O2", B, H // 4, W // 4, f[2], ex), taps["conv2"]),
                    ("res_out", m.debug_buffer(100, 0, (B, H // 4, W // 4, f[2]), torch.float32).permute(0, 3, 1, 2), taps["res6"]),
                    ("upconv2", nhwc(m, "V2", B, H // 2, W // 2, f[4], ex), taps["upconv2"]),
                    ("upconv1", nhwc(m, "V1", B, H, W, f[4], ex), taps["upconv1"]),
                    ("conv_11", nhwc(m, "C11", B, H, W, f[5], ex), taps["conv_11"]),
                    ("output", y, y_ref)]
            print("stage %d  precision %-6s (oracle CPU %.2fs)" % (stage, prec, t_cpu))
            for name, got, want in rows:
                err = (got - want).abs()
                print("   %-8s max|err| %.3e  mean|err| %.3e  (ref max|v| %.3f)%s" % (
                    name, err.max().item(), err.mean().item(), want.abs().max().item(),
                    "  NaN!" if torch.isnan(got).any() else ""))
            del m
    # fused uint8 frame path vs oracle composite
    for stage, cls in ((1, dsu.GeneratorJ_RIC), (2, dsu.GeneratorJ)):
        sd = synth.to_torch_state_dict(synth.make_state_dict(stage, seed=1234, out_gain=0.25))
        m = cls(precision="fp16x3", **ARGS)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        e = torch.from_numpy(edge).to(dev) if stage == 2 else None
        with torch.no_grad():
            out, yf = m.forward_frames(torch.from_numpy(color).to(dev), torch.from_numpy(pos).to(dev), e, return_float=True)
        torch.cuda.synchronize()
        want = np.stack([rp.compose_rgba(yf[i].cpu().numpy(), rp.frame_to_tensor(color[i], pos[i])[1]) for i in range(B)])
        print("stage %d fused u8: composite == to_image_space(own fp32 y)+alpha: %s" % (stage, np.array_equal(out.cpu().numpy(), want)))


if __name__ == "__main__":
    main()

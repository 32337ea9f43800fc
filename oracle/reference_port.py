"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference hot path.

Restates, function by function, what ``/root/reference/3_style_translator`` computes
for per-frame stylization inference.  Written from the reference's behaviour, not
copied from it; every function cites the file:line it follows.  fp32 torch /
numpy arithmetic on whatever device the input tensors live on (the CPU in the tests and the
CPU baseline; ``cuda`` only for bench.py's reference-on-GPU comparison arm, where torch's own
cuDNN / torchvision CUDA kernels play the reference), no dependency on ``/root/reference`` at run time.

Pinning status: the reference ships NO golden vectors / KATs for this path
("parity unpinned" by the reference's own tests).  This port
is instead pinned against the reference *itself*, imported live in the build
container by ``oracle/make_golden.py``, which commits the resulting fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the port against
them everywhere (no ``/root/reference`` needed).

Third-party arithmetic on the path: ``torchvision.ops.deform_conv2d``
(reference pins torchvision==0.15.1+cu118, README.md:29; not vendored).  Its
published sampling rule is restated in :func:`deform_conv3x3_port`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by models.py:31-33 via norm_layer(num_features=...)


# ----------------------------------------------------------------------------
# a3: generate_coordinates  (training/models.py:551-604)
# ----------------------------------------------------------------------------
def ric_offsets(height: int, width: int) -> torch.Tensor:
    """Rotation-invariant-coordinate offset field, ``[18, h, w]`` fp32.

    Follows models.py:551-598 op for op (same torch ops, same dtype, same order, so the
    result is bit-identical on the same host) but without the batch expand (:600) and
    without the unconditional ``.cuda()`` (:602).  Channel 2k is the row offset and
    2k+1 the column offset of raster tap k; the centre tap (k=4) stays zero (:567-568).
    """
    dims = torch.zeros(3)
    dims[1] = height
    dims[2] = width
    centre_r = torch.sub(torch.div(dims[1], 2.0), 0.5)
    centre_c = torch.sub(torch.div(dims[2], 2.0), 0.5)
    rr, cc = torch.meshgrid(torch.arange(0, dims[1]), torch.arange(0, dims[2]), indexing="ij")
    d_r = torch.sub(rr, centre_r)
    d_c = torch.sub(cc, centre_c)
    two_pi = torch.mul(torch.Tensor([math.pi]), 2.0)
    theta = torch.atan2(d_c, d_r) % two_pi[0]
    theta = torch.round(10000.0 * theta) / 10000.0
    step = torch.div(two_pi[0], 8.0)
    out = torch.zeros(height, width, 18)
    for rot in range(8):
        tap = rot if rot < 4 else rot + 1          # raster tap index, centre skipped
        i, j = divmod(tap, 3)
        ang = torch.add(theta, torch.mul(step, float(rot)))
        out[:, :, 2 * tap] = torch.add(torch.cos(ang), float(1 - i))
        out[:, :, 2 * tap + 1] = torch.add(torch.sin(ang), float(1 - j))
    return out.permute(2, 0, 1).contiguous()


# ----------------------------------------------------------------------------
# T: torchvision.ops.deform_conv2d restated
# ----------------------------------------------------------------------------
def bilinear_taps(offsets: torch.Tensor, height: int, width: int):
    """Per (tap, pixel) bilinear stencil of a 3x3, pad-1, stride-1 deformable conv.

    Returns ``(idx[9,4,h,w] int64 (clamped flat indices), wgt[9,4,h,w] fp32)`` such that
    sample(tap, y, x) = sum_c wgt[tap,c,y,x] * img.flatten()[idx[tap,c,y,x]].
    Rule (torchvision deform_conv2d kernel, bilinear_interpolate): the whole sample is 0
    when py<=-1, py>=H, px<=-1 or px>=W; each corner outside [0,H-1]x[0,W-1]
    contributes 0; weights are (1-lh)(1-lw), (1-lh)lw, lh(1-lw), lh*lw in fp32.
    """
    ys = torch.arange(height, dtype=torch.float32).view(height, 1).expand(height, width)
    xs = torch.arange(width, dtype=torch.float32).view(1, width).expand(height, width)
    idx = torch.zeros(9, 4, height, width, dtype=torch.int64)
    wgt = torch.zeros(9, 4, height, width, dtype=torch.float32)
    for tap in range(9):
        i, j = divmod(tap, 3)
        py = (ys - 1.0 + float(i)) + offsets[2 * tap]
        px = (xs - 1.0 + float(j)) + offsets[2 * tap + 1]
        inside = ~((py <= -1) | (py >= height) | (px <= -1) | (px >= width))
        h_lo = torch.floor(py)
        w_lo = torch.floor(px)
        lh = py - h_lo
        lw = px - w_lo
        hh = 1.0 - lh
        hw = 1.0 - lw
        h_lo = h_lo.long()
        w_lo = w_lo.long()
        h_hi = h_lo + 1
        w_hi = w_lo + 1
        corners = ((h_lo, w_lo, hh * hw), (h_lo, w_hi, hh * lw), (h_hi, w_lo, lh * hw), (h_hi, w_hi, lh * lw))
        for c, (hc, wc, wv) in enumerate(corners):
            ok = inside & (hc >= 0) & (hc <= height - 1) & (wc >= 0) & (wc <= width - 1)
            wgt[tap, c] = torch.where(ok, wv, torch.zeros_like(wv))
            idx[tap, c] = hc.clamp(0, height - 1) * width + wc.clamp(0, width - 1)
    return idx, wgt


def deform_conv3x3_port(x: torch.Tensor, offsets: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """``torchvision.ops.deform_conv2d(x, offset, weight, padding=(1,1))`` for a fixed,
    batch-independent offset field ``offsets[18,h,w]`` (call sites models.py:302-351).
    Pure torch gather + matmul; no bias, no mask, stride 1, dilation 1, groups 1."""
    b, c, h, w = x.shape
    idx, wgt = bilinear_taps(offsets.cpu(), h, w)
    idx, wgt = idx.to(x.device), wgt.to(x.device)
    flat = x.reshape(b, c, h * w)
    out = torch.zeros(b, weight.shape[0], h, w, dtype=x.dtype, device=x.device)
    for tap in range(9):
        i, j = divmod(tap, 3)
        samp = torch.zeros(b, c, h * w, dtype=x.dtype, device=x.device)
        for cn in range(4):
            samp = samp + flat[:, :, idx[tap, cn].reshape(-1)] * wgt[tap, cn].reshape(1, 1, -1)
        out = out + torch.einsum("oc,bcp->bop", weight[:, :, i, j], samp).reshape(b, -1, h, w)
    return out


def _deform(x, offsets, weight, use_torchvision):
    if use_torchvision:
        import torchvision
        off = offsets.unsqueeze(0).expand(x.shape[0], -1, -1, -1)
        return torchvision.ops.deform_conv2d(input=x, offset=off, weight=weight, padding=(1, 1))
    return deform_conv3x3_port(x, offsets, weight)


# ----------------------------------------------------------------------------
# building blocks shared by both generators
# ----------------------------------------------------------------------------
def _bn(x, sd, prefix, cfg=None):
    """Eval-mode BatchNorm2d = per-channel affine from running stats (scripts call
    generator.eval(): test_stage1.py:48, test_stage2.py:55).  A ``norm_layer`` module without state is either absent
    (norm_layer=None) or nn.InstanceNorm2d with its defaults (models.py:34-35: affine=False, track_running_stats=False,
    eps=1e-5 -> per-(frame, channel) statistics over H x W, biased variance, in eval mode too)."""
    if prefix + ".weight" not in sd:
        if cfg is not None and cfg.get("norm") == "instance_norm":
            return F.instance_norm(x, eps=BN_EPS)
        return x
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(v + BN_EPS)
    return (x - m.view(1, -1, 1, 1)) * (inv * g).view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _bias(sd, key):
    return sd.get(key, None)


def default_config(stage: int) -> Dict:
    """Generator hyper-parameters of configs/config_stage{1,2}.yaml:5-10 plus the
    +1 (mask) +2 (pos) input channels added by test_stage*.py:33-39."""
    return dict(filters=(32, 64, 128, 128, 128, 64), resnet_blocks=7, tanh=True,
                append_smoothers=True, use_bias=False, input_channels=6, stage=stage)


# ----------------------------------------------------------------------------
# a5: GeneratorJ.forward  (training/models.py:113-129) - stage 2, plain convs
# ----------------------------------------------------------------------------
def generator_j_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: Optional[Dict] = None,
                        taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    cfg = cfg or default_config(2)
    rec = (lambda k, v: taps.__setitem__(k, v.clone())) if taps is not None else (lambda k, v: None)
    o0 = F.leaky_relu(_bn(F.conv2d(x, sd["conv0.conv.weight"], _bias(sd, "conv0.conv.bias"), 1, 3),
                          sd, "conv0.normalization", cfg), 0.2)
    rec("conv0", o0)
    o1 = F.leaky_relu(_bn(F.conv2d(o0, sd["conv1.conv.weight"], _bias(sd, "conv1.conv.bias"), 2, 1),
                          sd, "conv1.normalization", cfg), 0.2)
    rec("conv1", o1)
    o2 = F.leaky_relu(_bn(F.conv2d(o1, sd["conv2.conv.weight"], _bias(sd, "conv2.conv.bias"), 2, 1),
                          sd, "conv2.normalization", cfg), 0.2)
    rec("conv2", o2)
    out = o2
    for i in range(cfg["resnet_blocks"]):
        p = "resnets.%d." % i
        t = F.conv2d(F.relu(out), sd[p + "conv_0.weight"], _bias(sd, p + "conv_0.bias"), 1, 1)
        t = F.relu(_bn(t, sd, p + "normalization", cfg))
        out = F.conv2d(t, sd[p + "conv_1.weight"], _bias(sd, p + "conv_1.bias"), 1, 1) + out
        rec("res%d" % i, out)
    t = F.interpolate(torch.cat((out, o2), 1), scale_factor=2, mode="nearest")
    out = F.relu(_bn(F.conv2d(t, sd["upconv2.1.weight"], None, 1, 1), sd, "upconv2.2", cfg))
    rec("upconv2", out)
    t = F.interpolate(torch.cat((out, o1), 1), scale_factor=2, mode="nearest")
    out = F.relu(_bn(F.conv2d(t, sd["upconv1.1.weight"], None, 1, 1), sd, "upconv1.2", cfg))
    rec("upconv1", out)
    out = F.relu(F.conv2d(torch.cat((out, o0, x), 1), sd["conv_11.0.weight"], _bias(sd, "conv_11.0.bias"), 1, 3))
    rec("conv_11", out)
    if cfg["append_smoothers"]:
        t = F.relu(F.conv2d(out, sd["conv_11_a.0.weight"], _bias(sd, "conv_11_a.0.bias"), 1, 1))
        t = _bn(t, sd, "conv_11_a.2")
        out = F.relu(F.conv2d(t, sd["conv_11_a.3.weight"], _bias(sd, "conv_11_a.3.bias"), 1, 1))
        rec("conv_11_a", out)
    w12 = "conv_12.0" if cfg["tanh"] else "conv_12"
    out = F.conv2d(out, sd[w12 + ".weight"], sd[w12 + ".bias"])
    return torch.tanh(out) if cfg["tanh"] else out


# ----------------------------------------------------------------------------
# a2: GeneratorJ_RIC.forward  (training/models.py:293-356) - stage 1, RIC deformable convs
# ----------------------------------------------------------------------------
def generator_j_ric_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, cfg: Optional[Dict] = None,
                            use_torchvision: bool = False,
                            taps: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """Note the reference quirks reproduced here: conv biases are never applied by the
    deformable calls (only ``.weight`` is passed, models.py:302-351); the first smoother
    conv is dead code because the second one reads ``output`` not ``tmp`` (:348-352)."""
    cfg = cfg or default_config(1)
    rec = (lambda k, v: taps.__setitem__(k, v.clone())) if taps is not None else (lambda k, v: None)
    h, w = x.shape[2], x.shape[3]
    # the field is computed on the CPU (bit-identical on every host) and moved to x's device: the reference does the same
    # (CPU torch ops, then .cuda(), models.py:551-602); tests run this port on the CPU, bench.py's reference-on-GPU arm on cuda
    c0 = ric_offsets(h, w).to(x.device)
    c1 = ric_offsets(int(h / 2), int(w / 2)).to(x.device)
    c2 = ric_offsets(int(h / 4), int(w / 4)).to(x.device)
    dc = lambda a, off, key: _deform(a, off, sd[key], use_torchvision)
    o0 = F.leaky_relu(_bn(dc(x, c0, "conv0.conv.weight"), sd, "conv0.normalization", cfg), 0.2)
    rec("conv0", o0)
    o1 = F.leaky_relu(_bn(dc(F.max_pool2d(o0, 2, 2), c1, "conv1.conv.weight"), sd, "conv1.normalization", cfg), 0.2)
    rec("conv1", o1)
    o2 = F.leaky_relu(_bn(dc(F.max_pool2d(o1, 2, 2), c2, "conv2.conv.weight"), sd, "conv2.normalization", cfg), 0.2)
    rec("conv2", o2)
    out = o2
    for i in range(cfg["resnet_blocks"]):
        p = "resnets.%d." % i
        t = F.relu(_bn(dc(F.relu(out), c2, p + "conv_0.weight"), sd, p + "normalization", cfg))
        out = dc(t, c2, p + "conv_1.weight") + out
        rec("res%d" % i, out)
    t = F.interpolate(torch.cat((out, o2), 1), scale_factor=2, mode="nearest")
    out = F.relu(_bn(dc(t, c1, "upconv2.1.weight"), sd, "upconv2.2", cfg))
    rec("upconv2", out)
    t = F.interpolate(torch.cat((out, o1), 1), scale_factor=2, mode="nearest")
    out = F.relu(_bn(dc(t, c0, "upconv1.1.weight"), sd, "upconv1.2", cfg))
    rec("upconv1", out)
    out = F.relu(dc(torch.cat((out, o0, x), 1), c0, "conv_11.0.weight"))
    rec("conv_11", out)
    if cfg["append_smoothers"]:
        out = F.relu(dc(out, c0, "conv_11_a.3.weight"))   # conv_11_a.0/.2 are dead (models.py:348-352)
        rec("conv_11_a", out)
    w12 = "conv_12.0" if cfg["tanh"] else "conv_12"
    out = F.conv2d(out, sd[w12 + ".weight"], sd[w12 + ".bias"])
    return torch.tanh(out) if cfg["tanh"] else out


# ----------------------------------------------------------------------------
# a10-a12: uint8 steps (custom_transforms.py:7-35, data.py:23-47, test_stage1.py:68-70)
# ----------------------------------------------------------------------------
def to_image_space(x: np.ndarray) -> np.ndarray:
    """fp32 [-1,1] -> uint8, clip then (x+1)/2*255 then TRUNCATE (custom_transforms.py:7-8)."""
    x = np.asarray(x, dtype=np.float32)
    return ((np.clip(x, -1, 1) + 1) / 2 * 255).astype(np.uint8)


def overlap_edge_on_img(edge: np.ndarray, rgba: np.ndarray) -> np.ndarray:
    """Stage-2 edge burn-in: where edge<255, RGB<-0 and A<-255 (custom_transforms.py:30-35)."""
    out = np.array(rgba, dtype=np.uint8, copy=True)
    hit = np.asarray(edge) < 255
    out[hit, 0:3] = 0
    out[hit, 3] = 255
    return out


def frame_to_tensor(color_rgba: np.ndarray, pos_rgba: np.ndarray, edge: Optional[np.ndarray] = None):
    """DatasetFullImages.__getitem__ (data.py:23-47) on already-decoded uint8 arrays.

    colour ``[H,W,4]``, pos ``[H,W,4]``, optional edge ``[H,W]``.  Returns
    ``(pre[6,H,W] fp32, pre_mask[1,H,W] fp32)``: RGB -> /255 -> (v-0.5)/0.5
    (ToTensor + Normalize, custom_transforms.py:18-22; alpha dropped, no premultiply :11-15);
    mask = alpha of the colour image taken BEFORE the edge burn-in (data.py:28) -> /255;
    pos RGB same transform, channels 0:2 only (data.py:40).
    """
    color_rgba = np.asarray(color_rgba, dtype=np.uint8)
    pos_rgba = np.asarray(pos_rgba, dtype=np.uint8)
    mask = color_rgba[..., 3].astype(np.float32) / np.float32(255)
    if edge is not None:
        color_rgba = overlap_edge_on_img(edge, color_rgba)
    norm = lambda u8: ((u8.astype(np.float32) / np.float32(255)) - np.float32(0.5)) / np.float32(0.5)
    rgb = norm(color_rgba[..., 0:3]).transpose(2, 0, 1)
    pos = norm(pos_rgba[..., 0:2]).transpose(2, 0, 1)
    pre = np.concatenate([rgb, mask[None], pos], 0).astype(np.float32)
    return pre, mask[None].astype(np.float32)


def compose_rgba(net_out: np.ndarray, pre_mask: np.ndarray) -> np.ndarray:
    """Result image of test_stage1.py:68-70 / test_stage2.py:75-78: RGB = to_image_space(out),
    A = (mask*255) truncated; returns uint8 ``[H,W,4]``."""
    img = to_image_space(net_out).transpose(1, 2, 0)
    alpha = (np.asarray(pre_mask, dtype=np.float32).transpose(1, 2, 0) * 255).astype(np.uint8)
    return np.concatenate((img, alpha), 2)


# ----------------------------------------------------------------------------
# f1 ("next" row): pos2edge (run_render.py:31-57)
# ----------------------------------------------------------------------------
def _sobel_reflect101(img64: np.ndarray):
    """cv2.Sobel(ksize=3, CV_64F) with the default BORDER_REFLECT_101, separable form."""
    p = np.pad(img64, 1, mode="reflect")
    gx = (p[:-2, 2:] - p[:-2, :-2]) + 2.0 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    gy = (p[2:, :-2] - p[:-2, :-2]) + 2.0 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    return gx, gy


def pos2edge(pos_rgba: np.ndarray) -> np.ndarray:
    """Edge map from a pos RGBA image (run_render.py:31-57): /255 in fp32, background
    (alpha<1) -> 2, per-channel Sobel-3 in float64, max gradient magnitude > 0.3 -> 255.
    (The caller stores 255-edge, run_render.py:117-120.)  Channel order is irrelevant
    because the three channels are treated identically and max-reduced."""
    f = np.asarray(pos_rgba, dtype=np.uint8).astype(np.float32) / np.float32(255.0)
    bg = f[..., 3] < 1
    mags = []
    for ch in range(3):
        c = f[..., ch].copy()
        c[bg] = 2
        gx, gy = _sobel_reflect101(c.astype(np.float64))
        mags.append(np.sqrt(np.square(gx) + np.square(gy)))
    e = np.maximum(np.maximum(mags[0], mags[1]), mags[2])
    return ((e > 0.3) * 255).astype(np.uint8)

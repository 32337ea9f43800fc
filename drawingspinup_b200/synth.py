"""Seeded synthetic frame stacks and generator weights.

No frames, meshes or checkpoints ship with the reference, and there is no network, so
benchmarks and parity tests run on synthetic data with the on-disk shapes of
``<uid>/mesh/blender_render/<motion>/{color,pos,edge}/NNNN.png`` (decoded) and on a
seeded ``state_dict`` with the 89-key layout of ``training/models.py``.
Pure numpy; deterministic for a given seed on every host.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Sequence, Tuple

import numpy as np

DEFAULT_FILTERS = (32, 64, 128, 128, 128, 64)


def _lowfreq(rng, n, h, w, ch, grid=8):
    """[n,h,w,ch] smooth noise in [0,1]: grid x grid control points, bilinear upsampling.  (Frame by frame: the same fp32
    operations in the same order as the whole-array form, ~5x faster because the temporaries stay in cache - the synthetic
    clips of the multi-GPU bench configs are hundreds of frames per rank.)"""
    ctl = rng.random((n, grid + 1, grid + 1, ch), dtype=np.float32)
    ys = np.linspace(0, grid, h, endpoint=False, dtype=np.float32)
    xs = np.linspace(0, grid, w, endpoint=False, dtype=np.float32)
    y0 = ys.astype(np.int64); x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None, None]; fx = (xs - x0)[None, :, None]
    out = np.empty((n, h, w, ch), np.float32)
    for i in range(n):
        r0 = ctl[i][y0]; r1 = ctl[i][y0 + 1]
        out[i] = r0[:, x0] * (1 - fy) * (1 - fx) + r0[:, x0 + 1] * (1 - fy) * fx + r1[:, x0] * fy * (1 - fx) + r1[:, x0 + 1] * fy * fx
    return out


def make_frames(n_frames: int, height: int = 512, width: int = 512, seed: int = 1234
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Synthetic decoded animation frames.

    Returns ``(color[n,H,W,4] u8, pos[n,H,W,4] u8, edge[n,H,W] u8)``:
    colour RGBA with a filled character blob (union of moving ellipses, 1-px antialiased
    rim, RGB = smooth noise inside and 0 outside); pos RGBA with R,G = warped (x,y)
    ramps inside the blob and the same alpha; edge L = 255 with 0 on the outline and a
    few interior strokes (the stored ``255 - edge`` of run_render.py:117-120).
    """
    rng = np.random.default_rng(seed)
    n, h, w = n_frames, height, width
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    n_ell = 6
    cx = rng.uniform(0.3, 0.7, n_ell) * w; cy = rng.uniform(0.25, 0.75, n_ell) * h
    ax = rng.uniform(0.10, 0.22, n_ell) * w; ay = rng.uniform(0.10, 0.22, n_ell) * h
    ph = rng.uniform(0, 2 * np.pi, (n_ell, 2)); amp = rng.uniform(0.01, 0.04, (n_ell, 2)) * min(h, w)
    alpha = np.zeros((n, h, w), np.float32)
    for f in range(n):
        t = 2 * np.pi * f / max(n, 16)
        soft = np.zeros((h, w), np.float32)
        for e in range(n_ell):
            ex = cx[e] + amp[e, 0] * np.sin(t + ph[e, 0]); ey = cy[e] + amp[e, 1] * np.cos(t + ph[e, 1])
            q = ((xx - ex) / ax[e]) ** 2 + ((yy - ey) / ay[e]) ** 2
            soft = np.maximum(soft, np.clip((1.0 - q) * (0.5 * min(ax[e], ay[e])) + 0.5, 0.0, 1.0))
        alpha[f] = soft
    a8 = np.rint(alpha * 255).astype(np.uint8)
    inside = a8 > 0
    rgb = (_lowfreq(rng, n, h, w, 3) * 255).astype(np.uint8)
    rgb[~inside] = 0
    color = np.concatenate([rgb, a8[..., None]], -1)

    warp = (_lowfreq(rng, n, h, w, 2, grid=4) - 0.5) * 0.15
    px = np.clip(xx[None] / w + warp[..., 0], 0, 1); py = np.clip(yy[None] / h + warp[..., 1], 0, 1)
    pos = np.zeros((n, h, w, 4), np.uint8)
    pos[..., 0] = (px * 255).astype(np.uint8); pos[..., 1] = (py * 255).astype(np.uint8)
    pos[..., 2] = (_lowfreq(rng, n, h, w, 1, grid=4)[..., 0] * 255).astype(np.uint8)
    pos[~inside] = 0
    pos[..., 3] = a8

    rim = (a8 > 0) & (a8 < 255)
    grow = rim.copy()
    grow[:, 1:, :] |= rim[:, :-1, :]; grow[:, :-1, :] |= rim[:, 1:, :]
    grow[:, :, 1:] |= rim[:, :, :-1]; grow[:, :, :-1] |= rim[:, :, 1:]
    strokes = np.zeros((n, h, w), bool)
    for s in range(4):
        col = int(rng.uniform(0.3, 0.7) * w); r0 = int(rng.uniform(0.2, 0.5) * h); r1 = r0 + int(0.2 * h)
        strokes[:, r0:r1, col:col + 2] = True
    edge = np.full((n, h, w), 255, np.uint8)
    edge[grow | (strokes & inside)] = 0
    return color, pos, edge


def _conv_w(rng, cout, cin, k):
    bound = np.sqrt(6.0 / (cin * k * k))       # He-uniform keeps activations O(1) through the ReLU stacks
    return rng.uniform(-bound, bound, (cout, cin, k, k)).astype(np.float32)


def _bn(rng, sd, prefix, c):
    sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + ".bias"] = rng.normal(0, 0.1, c).astype(np.float32)
    sd[prefix + ".running_mean"] = rng.normal(0, 0.1, c).astype(np.float32)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(1000, dtype=np.int64)


def make_state_dict(stage: int, seed: int = 1234, filters: Sequence[int] = DEFAULT_FILTERS,
                    resnet_blocks: int = 7, input_channels: int = 6, tanh: bool = True,
                    append_smoothers: bool = True, use_bias: bool = False,
                    out_gain: float = 1.0, norm: str = "batch_norm") -> "OrderedDict[str, np.ndarray]":
    """Seeded weights with the reference key order / shapes (models.py:24-111 stage 2,
    :200-291 stage 1).  BN running stats are non-trivial so that
    folding bugs cannot hide; ``conv_12`` is scaled by ``out_gain`` so the tanh output
    uses its full range like a trained network.  ``norm`` = 'instance_norm' / None: the twelve ``norm_layer`` modules have no
    state (nn.InstanceNorm2d defaults, models.py:34-35); ``conv_11_a.2`` is a hard-coded BatchNorm2d either way (models.py:98)."""
    rng = np.random.default_rng(seed * 7919 + stage)
    _bn_fixed = _bn
    bn_l = (lambda *a: None) if norm != "batch_norm" else _bn_fixed
    f = list(filters)
    k0 = 3 if stage == 1 else 7
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(key, cout, cin, k, bias=use_bias):
        sd[key + ".weight"] = _conv_w(rng, cout, cin, k)
        if bias:
            sd[key + ".bias"] = rng.normal(0, 0.05, cout).astype(np.float32)

    conv("conv0.conv", f[0], input_channels, k0); bn_l(rng, sd, "conv0.normalization", f[0])
    conv("conv1.conv", f[1], f[0], 3); bn_l(rng, sd, "conv1.normalization", f[1])
    conv("conv2.conv", f[2], f[1], 3); bn_l(rng, sd, "conv2.normalization", f[2])
    for i in range(resnet_blocks):
        p = "resnets.%d." % i
        conv(p + "conv_0", f[2], f[2], 3); bn_l(rng, sd, p + "normalization", f[2])
        conv(p + "conv_1", f[2], f[2], 3)
        sd[p + "conv_1.weight"] *= np.float32(0.5)     # keep the residual stream from blowing up
    conv("upconv2.1", f[4], f[3] + f[2], 3, bias=False); bn_l(rng, sd, "upconv2.2", f[4])
    conv("upconv1.1", f[4], f[4] + f[1], 3, bias=False); bn_l(rng, sd, "upconv1.2", f[4])
    conv("conv_11.0", f[5], f[0] + f[4] + input_channels, k0)
    if append_smoothers:
        conv("conv_11_a.0", f[5], f[5], 3); _bn_fixed(rng, sd, "conv_11_a.2", f[5])
        conv("conv_11_a.3", f[5], f[5], 3)
    k12 = "conv_12.0" if tanh else "conv_12"
    sd[k12 + ".weight"] = (_conv_w(rng, 3, f[5], 1) * np.float32(out_gain)).astype(np.float32)
    sd[k12 + ".bias"] = rng.normal(0, 0.1, 3).astype(np.float32)
    return sd


def to_torch_state_dict(sd_np: Dict[str, np.ndarray]):
    import torch
    return OrderedDict((k, torch.from_numpy(np.array(v, copy=True, order='C'))) for k, v in sd_np.items())


def write_character_tree(root_dir: str, uid: str, actions: Dict[str, int], height: int, width: int, seed: int = 1234,
                         state_dicts=None, frames=None) -> Dict[str, Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Materialise a synthetic character in the reference's on-disk layout (README.md "dataset" tree):
    ``<root>/<uid>/mesh/blender_render/<action>/{color,pos,edge}/NNNN.png`` for ``actions = {name: n_frames}`` and,
    when ``state_dicts = (sd_stage1, sd_stage2)`` (torch tensors) is given, the two ``model_99999.pth`` checkpoints.
    ``frames[action] = (color, pos, edge)`` overrides the generated stacks.  Returns the stacks per action."""
    import os
    from PIL import Image
    out = {}
    base = os.path.join(root_dir, uid, "mesh")
    for ai, (action, n) in enumerate(sorted(actions.items())):
        color, pos, edge = frames[action] if frames and action in frames else make_frames(n, height, width, seed=seed + ai)
        for sub in ("color", "pos", "edge"):
            os.makedirs(os.path.join(base, "blender_render", action, sub), exist_ok=True)
        for i in range(color.shape[0]):
            name = "%04d.png" % i
            Image.fromarray(color[i]).save(os.path.join(base, "blender_render", action, "color", name))
            Image.fromarray(pos[i]).save(os.path.join(base, "blender_render", action, "pos", name))
            Image.fromarray(edge[i]).save(os.path.join(base, "blender_render", action, "edge", name))
        out[action] = (color, pos, edge)
    if state_dicts is not None:
        import torch
        for log, sd in zip(("logs_stage1_mask_pos", "logs_stage2_mask_pos_edge"), state_dicts):
            os.makedirs(os.path.join(base, log), exist_ok=True)
            torch.save(sd, os.path.join(base, log, "model_99999.pth"))
    return out

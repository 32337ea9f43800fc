"""CPU: the oracle port (oracle/reference_port.py) against the golden vectors minted from the live
reference (oracle/make_golden.py) and against the known-answer tests of SURVEY.md 8c."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import DEFAULT_ARGS, VARIANT_ARGS
from drawingspinup_b200 import synth
from oracle import reference_port as rp


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("stage", [1, 2])
def test_generator_matches_reference_golden(golden_dir, stage):
    g = _load(golden_dir, "generator_stage%d.npz" % stage)
    sd = synth.to_torch_state_dict(synth.make_state_dict(stage, seed=int(g["seed"]), out_gain=float(g["out_gain"])))
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = rp.generator_j_ric_forward(sd, x) if stage == 1 else rp.generator_j_forward(sd, x)
    assert y.shape == g["y"].shape
    assert np.abs(y.numpy() - g["y"]).max() < 1e-4      # fp32 summation-order noise only


@pytest.mark.parametrize("stage", [1, 2])
def test_generator_variant_config_matches_reference_golden(golden_dir, stage):
    g = _load(golden_dir, "generator_variant_stage%d.npz" % stage)
    a = VARIANT_ARGS
    sd = synth.to_torch_state_dict(synth.make_state_dict(stage, seed=77, filters=a["filters"], resnet_blocks=a["resnet_blocks"],
                                                         input_channels=a["input_channels"], tanh=a["tanh"],
                                                         append_smoothers=a["append_smoothers"], use_bias=a["use_bias"], out_gain=0.25))
    cfg = dict(rp.default_config(stage), **{k: a[k] for k in ("resnet_blocks", "tanh", "append_smoothers", "use_bias")})
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = rp.generator_j_ric_forward(sd, x, cfg) if stage == 1 else rp.generator_j_forward(sd, x, cfg)
    assert np.abs(y.numpy() - g["y"]).max() < 2e-4


@pytest.mark.parametrize("norm,stage", [("instance_norm", 1), ("instance_norm", 2), (None, 2)])
def test_generator_norm_layer_variants_match_reference_golden(golden_dir, norm, stage):
    """norm_layer='instance_norm' (nn.InstanceNorm2d defaults: per-(frame, channel) statistics, no state) and norm_layer=None
    (models.py:29-35), recorded from the live reference.  (norm_layer=None is not runnable for stage 1 in the reference
    itself: its forward indexes self.conv0[2], models.py:303.)"""
    g = _load(golden_dir, "generator_%s_stage%d.npz" % (norm or "no_norm", stage))
    sd = synth.to_torch_state_dict(synth.make_state_dict(stage, seed=91, resnet_blocks=2, out_gain=0.25, norm=norm or "none"))
    assert not any("normalization" in k or k.startswith("upconv2.2") for k in sd) and "conv_11_a.2.weight" in sd
    cfg = dict(rp.default_config(stage), resnet_blocks=2, norm=norm)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = rp.generator_j_ric_forward(sd, x, cfg) if stage == 1 else rp.generator_j_forward(sd, x, cfg)
    assert np.abs(y.numpy() - g["y"]).max() < 2e-4


def test_ric_port_equals_torchvision_path():
    sd = synth.to_torch_state_dict(synth.make_state_dict(1, seed=3, out_gain=0.25))
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((1, 6, 16, 24)).astype(np.float32))
    with torch.no_grad():
        a = rp.generator_j_ric_forward(sd, x, use_torchvision=False)
        b = rp.generator_j_ric_forward(sd, x, use_torchvision=True)
    assert (a - b).abs().max().item() < 1e-4


def test_ric_coordinates_bit_exact(golden_dir):
    g = _load(golden_dir, "ric_coords_24x20.npz")
    assert np.array_equal(rp.ric_offsets(24, 20).numpy(), g["coords"])


def test_ric_offsets_are_unit_circle_samples():
    off = rp.ric_offsets(12, 16).numpy()
    assert np.all(off[8:10] == 0)                       # centre tap untouched (models.py:567-568)
    for tap in (0, 1, 2, 3, 5, 6, 7, 8):
        i, j = divmod(tap, 3)
        dy = off[2 * tap] + (i - 1)
        dx = off[2 * tap + 1] + (j - 1)
        assert np.allclose(dy * dy + dx * dx, 1.0, atol=1e-5)


def test_zero_offset_deform_equals_conv2d():            # KAT (i)
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((2, 5, 9, 11)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((7, 5, 3, 3)).astype(np.float32))
    out = rp.deform_conv3x3_port(x, torch.zeros(18, 9, 11), w)
    assert (out - F.conv2d(x, w, padding=1)).abs().max().item() < 1e-4


def test_deform_border_rule():                          # KAT (ii): half-pixel outside halves, a full pixel outside zeroes
    x = torch.ones(1, 1, 4, 4)
    w = torch.zeros(1, 1, 3, 3)
    w[0, 0, 1, 1] = 1.0                                 # only the centre tap
    off = torch.zeros(18, 4, 4)
    off[8] = -0.5                                       # centre tap row offset
    out = rp.deform_conv3x3_port(x, off, w)
    assert torch.allclose(out[0, 0, 0], torch.full((4,), 0.5)) and torch.allclose(out[0, 0, 1:], torch.ones(3, 4))
    off[8] = -1.0
    out = rp.deform_conv3x3_port(x, off, w)
    assert torch.all(out[0, 0, 0] == 0) and torch.allclose(out[0, 0, 1:], torch.ones(3, 4))


def test_stage1_dead_smoother_branch_is_ignored():      # KAT (iii), models.py:348-352
    sd = synth.to_torch_state_dict(synth.make_state_dict(1, seed=5, out_gain=0.25))
    x = torch.from_numpy(np.random.default_rng(2).standard_normal((1, 6, 16, 16)).astype(np.float32))
    with torch.no_grad():
        a = rp.generator_j_ric_forward(sd, x)
        sd["conv_11_a.0.weight"] = torch.randn_like(sd["conv_11_a.0.weight"])
        sd["conv_11_a.2.running_mean"] = torch.randn_like(sd["conv_11_a.2.running_mean"])
        b = rp.generator_j_ric_forward(sd, x)
    assert torch.equal(a, b)


def test_to_image_space_known_answers(golden_dir):      # KAT (v)
    g = _load(golden_dir, "to_image_space.npz")
    assert list(rp.to_image_space(g["kat_in"])) == [0, 0, 63, 127, 127, 191, 254, 255, 255]
    assert np.array_equal(rp.to_image_space(g["kat_in"]), g["kat_out"])
    assert np.array_equal(rp.to_image_space(g["rnd_in"]), g["rnd_out"])


def test_alpha_round_trip_all_256():                    # KAT (iv)
    a = np.arange(256, dtype=np.uint8)
    m = a.astype(np.float32) / np.float32(255)
    assert np.array_equal((m * 255).astype(np.uint8), a)


def test_dataset_transform_matches_reference(golden_dir):
    g = _load(golden_dir, "dataset_transform.npz")
    for i in range(g["color"].shape[0]):
        p1, m1 = rp.frame_to_tensor(g["color"][i], g["pos"][i])
        p2, _ = rp.frame_to_tensor(g["color"][i], g["pos"][i], g["edge"][i])
        assert np.array_equal(p1, g["pre_stage1"][i]) and np.array_equal(p2, g["pre_stage2"][i])
        assert np.array_equal(m1, g["pre_mask"][i])


def test_overlap_edge_semantics():
    rgba = np.full((3, 3, 4), 7, np.uint8)
    edge = np.full((3, 3), 255, np.uint8)
    edge[1, 1] = 0
    edge[0, 2] = 254
    out = rp.overlap_edge_on_img(edge, rgba)
    assert list(out[1, 1]) == [0, 0, 0, 255] and list(out[0, 2]) == [0, 0, 0, 255] and list(out[0, 0]) == [7, 7, 7, 7]


def test_pos2edge_matches_reference(golden_dir):
    g = _load(golden_dir, "pos2edge.npz")
    for i in range(g["pos"].shape[0]):
        assert np.array_equal(rp.pos2edge(g["pos"][i]), g["edges"][i])


def test_batch_invariance_of_port():                    # KAT (vi)
    sd = synth.to_torch_state_dict(synth.make_state_dict(2, seed=9, out_gain=0.25))
    x = torch.from_numpy(np.random.default_rng(4).standard_normal((3, 6, 16, 16)).astype(np.float32))
    with torch.no_grad():
        full = rp.generator_j_forward(sd, x)
        one = rp.generator_j_forward(sd, x[1:2])
    assert (full[1:2] - one).abs().max().item() < 1e-5

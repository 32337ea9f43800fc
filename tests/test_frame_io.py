"""On-disk frame format and the per-character driver (drawingspinup_b200/frame_io.py) - SURVEY.md 8f rank 2, host side.

CPU tests cover the folder / PNG / GIF logic against the reference's conventions (data.py:18-47, test_stage1.py:51-71,
test_stage2.py:59-79, gif_writer.py:13-30) with a stand-in pipeline; the GPU test runs the real engine over a
synthetic character tree and compares the written PNGs with the oracle chain."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from drawingspinup_b200 import frame_io, synth
from drawingspinup_b200.pipeline import DEFAULT_ARGS
from oracle import reference_port as rp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class _StandInPipeline:
    """Same call contract as StylizationPipeline.run_host; 'stage 1' inverts RGB, 'stage 2' burns the edge map in."""

    def run_host(self, color, pos, edge, out, keep_stage1=False):
        mid = color.clone()
        mid[..., :3] = 255 - mid[..., :3]
        res = mid.clone()
        res[..., :3][edge < 255] = 0
        out.copy_(res)
        return mid if keep_stage1 else out


def test_tree_listing_and_stack_loading(tmp_path):
    stacks = synth.write_character_tree(str(tmp_path), "u1", {"walk": 3, "rest_pose": 2}, 24, 32, seed=5)
    data_root = tmp_path / "u1" / "mesh" / "blender_render"
    os.makedirs(data_root / ".hidden")
    (data_root / "notes.txt").write_text("x")
    assert frame_io.list_actions(str(data_root)) == ["rest_pose", "walk"]
    assert frame_io.list_frames(str(data_root / "walk")) == ["0000.png", "0001.png", "0002.png"]
    fs = frame_io.FrameSet.load(str(data_root / "walk"), pin=False)
    color, pos, edge = stacks["walk"]
    assert len(fs) == 3 and fs.color.dtype == torch.uint8
    assert np.array_equal(fs.color.numpy(), color) and np.array_equal(fs.pos.numpy(), pos) and np.array_equal(fs.edge.numpy(), edge)
    sub = frame_io.FrameSet.load(str(data_root / "walk"), names=["0001.png", "0002.png"], pin=False, workers=1)
    assert np.array_equal(sub.color.numpy(), color[1:]) and sub.names == ["0001.png", "0002.png"]
    no_edge = frame_io.FrameSet.load(str(data_root / "walk"), need_edge=False, pin=False)
    assert no_edge.edge is None


def test_png_tree_reproduces_reference_dataset_tensors(tmp_path):
    """The golden holds what the live reference's DatasetFullImages (data.py:23-47) returned for these PNG frames:
    PNG tree -> FrameSet -> oracle transform must reproduce it bit for bit."""
    g = np.load(os.path.join(GOLDEN, "dataset_transform.npz"))
    synth.write_character_tree(str(tmp_path), "u", {"a": 2}, 32, 40, frames={"a": (g["color"], g["pos"], g["edge"])})
    fs = frame_io.FrameSet.load(str(tmp_path / "u" / "mesh" / "blender_render" / "a"), pin=False)
    for i in range(2):
        x1, m1 = rp.frame_to_tensor(fs.color[i].numpy(), fs.pos[i].numpy(), None)
        x2, _ = rp.frame_to_tensor(fs.color[i].numpy(), fs.pos[i].numpy(), fs.edge[i].numpy())
        assert np.array_equal(x1, g["pre_stage1"][i]) and np.array_equal(x2, g["pre_stage2"][i]) and np.array_equal(m1, g["pre_mask"][i])


def test_rejects_frames_the_path_does_not_define(tmp_path):
    synth.write_character_tree(str(tmp_path), "u", {"a": 2}, 16, 16)
    adir = tmp_path / "u" / "mesh" / "blender_render" / "a"
    Image.fromarray(np.zeros((16, 16, 3), np.uint8)).save(adir / "color" / "0001.png")          # RGB instead of RGBA
    with pytest.raises(ValueError, match="expected a RGBA PNG"):
        frame_io.FrameSet.load(str(adir), pin=False)
    Image.fromarray(np.zeros((16, 20, 4), np.uint8)).save(adir / "color" / "0001.png")          # other size
    with pytest.raises(ValueError, match="frame size differs"):
        frame_io.FrameSet.load(str(adir), pin=False)


def test_save_frames_and_gif(tmp_path):
    rng = np.random.default_rng(0)
    rgba = torch.from_numpy(rng.integers(0, 256, (3, 12, 10, 4), dtype=np.uint8))
    names = ["0000.png", "0001.png", "0002.png"]
    frame_io.save_frames(str(tmp_path / "rgba"), names, rgba)
    frame_io.save_frames(str(tmp_path / "rgb"), names, rgba, save_alpha=False)
    for i, n in enumerate(names):
        with Image.open(tmp_path / "rgba" / n) as im:
            assert im.mode == "RGBA" and np.array_equal(np.asarray(im), rgba[i].numpy())
        with Image.open(tmp_path / "rgb" / n) as im:
            assert im.mode == "RGB" and np.array_equal(np.asarray(im), rgba[i].numpy()[..., :3])
    with pytest.raises(ValueError):
        frame_io.save_frames(str(tmp_path / "bad"), names[:2], rgba)
    assert frame_io.write_gif(str(tmp_path / "rgba"), str(tmp_path / "gif" / "clip.gif")) == 3
    with Image.open(tmp_path / "gif" / "clip.gif") as im:                                         # gif_writer.py:30
        assert im.n_frames == 3 and im.info["duration"] == 30 and im.info["loop"] == 0
        im.seek(1)
        assert im.disposal_method == 2


def test_stylize_character_folder_logic(tmp_path):
    sd = ({"w": torch.zeros(1)}, {"w": torch.ones(1)})
    stacks = synth.write_character_tree(str(tmp_path), "u", {"jump": 5, "rest_pose": 1}, 16, 24, seed=3, state_dicts=sd)
    seen = {}

    def factory(sd1, sd2):
        seen["sd"] = (float(sd1["w"]), float(sd2["w"]))
        return _StandInPipeline()

    rep = frame_io.stylize_character(str(tmp_path), "u", pipeline_factory=factory, gif=True, workers=2)
    assert seen["sd"] == (0.0, 1.0)                                   # model_99999.pth of stage 1 and stage 2, in that order
    assert rep.frames == 6 and rep.actions == {"jump": 5, "rest_pose": 1}
    root = tmp_path / "u" / "mesh"
    color, _, edge = stacks["jump"]
    for i in range(5):
        with Image.open(root / "blender_render" / "jump" / frame_io.STAGE1_RES / ("%04d.png" % i)) as im:
            want = color[i].copy(); want[..., :3] = 255 - want[..., :3]
            assert np.array_equal(np.asarray(im), want)
        with Image.open(root / "blender_render" / "jump" / frame_io.STAGE2_RES / ("%04d.png" % i)) as im:
            want2 = want.copy(); want2[..., :3][edge[i] < 255] = 0
            assert np.array_equal(np.asarray(im), want2)
    assert frame_io.STAGE1_RES == "res_stage1_mask_pos" and frame_io.STAGE2_RES == "res_stage2_mask_pos_edge"
    assert os.path.isfile(root / "gif" / "jump_res_stage2_mask_pos_edge.gif")
    assert not os.path.exists(root / "gif" / "rest_pose_res_stage2_mask_pos_edge.gif")          # gif_writer.py:15


def test_stylize_character_rank_sharding(tmp_path):
    synth.write_character_tree(str(tmp_path), "u", {"jump": 5}, 16, 16, seed=9, state_dicts=({}, {}))
    for rank in range(2):
        rep = frame_io.stylize_character(str(tmp_path), "u", pipeline=_StandInPipeline(), rank=rank, world=2,
                                         keep_stage1=False, save_alpha=False)
        assert rep.frames == (3 if rank == 0 else 2)
    out = tmp_path / "u" / "mesh" / "blender_render" / "jump" / frame_io.STAGE2_RES
    assert sorted(os.listdir(out)) == ["%04d.png" % i for i in range(5)]
    with Image.open(out / "0004.png") as im:
        assert im.mode == "RGB"
    assert not os.path.exists(tmp_path / "u" / "mesh" / "blender_render" / "jump" / frame_io.STAGE1_RES)


@pytest.mark.gpu
def test_stylize_character_matches_oracle_chain(tmp_path):
    """Real engine over a PNG tree: every written stage-1 / stage-2 frame equals the oracle chain
    (test_stage1.py -> test_stage2.py) up to the 1-LSB uint8 boundary effect of the fp16x3 forward."""
    sd1 = synth.to_torch_state_dict(synth.make_state_dict(1, seed=21, out_gain=0.25))
    sd2 = synth.to_torch_state_dict(synth.make_state_dict(2, seed=22, out_gain=0.25))
    stacks = synth.write_character_tree(str(tmp_path), "u", {"walk": 3}, 32, 48, seed=4, state_dicts=(sd1, sd2))
    rep = frame_io.stylize_character(str(tmp_path), "u", device="cuda:0", precision="fp16x3", batch=2)
    assert rep.frames == 3
    color, pos, edge = stacks["walk"]
    adir = tmp_path / "u" / "mesh" / "blender_render" / "walk"
    cfg1, cfg2 = rp.default_config(1), rp.default_config(2)
    for i in range(3):
        x1, mask = rp.frame_to_tensor(color[i], pos[i], None)
        with torch.no_grad():
            y1 = rp.generator_j_ric_forward(sd1, torch.from_numpy(x1)[None], cfg1, use_torchvision=True)[0].numpy()
        want1 = rp.compose_rgba(y1, mask)
        with Image.open(adir / frame_io.STAGE1_RES / ("%04d.png" % i)) as im:
            got1 = np.asarray(im)
        assert np.array_equal(got1[..., 3], color[i][..., 3])
        assert np.abs(got1.astype(int) - want1.astype(int)).max() <= 1
        x2, mask2 = rp.frame_to_tensor(got1, pos[i], edge[i])          # stage 2 consumes the frame stage 1 wrote
        with torch.no_grad():
            y2 = rp.generator_j_forward(sd2, torch.from_numpy(x2)[None], cfg2)[0].numpy()
        want2 = rp.compose_rgba(y2, mask2)
        with Image.open(adir / frame_io.STAGE2_RES / ("%04d.png" % i)) as im:
            got2 = np.asarray(im)
        assert np.abs(got2.astype(int) - want2.astype(int)).max() <= 1
        assert (got2 != want2).mean() < 0.02


def test_missing_edge_maps_and_empty_clips(tmp_path):
    synth.write_character_tree(str(tmp_path), "u", {"a": 2}, 16, 16, state_dicts=({}, {}))
    data_root = tmp_path / "u" / "mesh" / "blender_render"
    os.makedirs(data_root / "empty" / "color")                      # a clip without frames is skipped
    for f in os.listdir(data_root / "a" / "edge"):
        os.remove(data_root / "a" / "edge" / f)
    os.rmdir(data_root / "a" / "edge")
    assert len(frame_io.FrameSet.load(str(data_root / "empty"), pin=False)) == 0
    with pytest.raises(FileNotFoundError, match="edge"):             # stage 2 cannot run without run_render.py's edge maps
        frame_io.stylize_character(str(tmp_path), "u", pipeline=_StandInPipeline())

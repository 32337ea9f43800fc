"""Raw frame-stack format (drawingspinup_b200/frame_stack.py) - SURVEY.md 8f rank 2: the on-disk format either side of the
per-frame path without a codec.  CPU tests: the stacks hold exactly the pixels of the reference's PNG tree
(training/data.py:18-33), sharded ranks read / write disjoint ranges of the same files, and the per-character driver gives
the same result from stacks as from PNGs."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from drawingspinup_b200 import frame_io, frame_stack, synth


class _StandInPipeline:
    derive_edge = False

    def run_host(self, color, pos, edge, out, keep_stage1=False):
        mid = color.clone()
        mid[..., :3] = 255 - mid[..., :3]
        res = mid.clone()
        res[..., :3][edge < 255] = 0
        out.copy_(res)
        return mid if keep_stage1 else out


def _tree(tmp_path, frames=5, h=24, w=32):
    sd = ({"w": torch.zeros(1)}, {"w": torch.ones(1)})
    stacks = synth.write_character_tree(str(tmp_path), "u", {"walk": frames, "rest_pose": 2}, h, w, seed=9, state_dicts=sd)
    return stacks, tmp_path / "u" / "mesh" / "blender_render"


def test_pack_is_bit_exact_and_memory_mappable(tmp_path):
    stacks, root = _tree(tmp_path)
    adir = str(root / "walk")
    assert not frame_stack.has_stack(adir)
    assert frame_stack.pack_action(adir, workers=2) == 5 and frame_stack.has_stack(adir)
    color, pos, edge = stacks["walk"]
    mm = frame_stack.open_layer(adir, "color")
    assert isinstance(mm, np.memmap) and mm.shape == color.shape and np.array_equal(mm, color)
    names, c, p, e = frame_stack.load_range(adir, 1, 4, pin=False)
    assert names == ["0001.png", "0002.png", "0003.png"]
    assert np.array_equal(c.numpy(), color[1:4]) and np.array_equal(p.numpy(), pos[1:4]) and np.array_equal(e.numpy(), edge[1:4])
    n0, c0, _, _ = frame_stack.load_range(adir, 2, 2, pin=False)            # empty shard
    assert n0 == [] and c0.shape == (0, 24, 32, 4)
    with pytest.raises(ValueError):
        frame_stack.load_range(adir, 3, 9, pin=False)
    with pytest.raises(ValueError):
        frame_stack.write_stack(adir, ["a.png"], color[:1, :, :, :3], pos[:1])          # RGB is not a defined input


def test_sharded_ranks_fill_one_output_stack(tmp_path):
    _, root = _tree(tmp_path, frames=7)
    adir = str(root / "walk")
    rng = np.random.default_rng(1)
    full = rng.integers(0, 256, (7, 24, 32, 4), dtype=np.uint8)
    for lo, hi in ((4, 7), (0, 2), (2, 4)):                                  # any order, disjoint ranges
        frame_stack.save_range(adir, "res_stage2_mask_pos_edge", torch.from_numpy(full[lo:hi]), lo, 7)
    got = np.load(os.path.join(frame_stack.stack_dir(adir), "res_stage2_mask_pos_edge.npy"))
    assert np.array_equal(got, full)
    with pytest.raises(ValueError):
        frame_stack.save_range(adir, "res_stage2_mask_pos_edge", full[:3], 5, 7)
    # a stale file of another size is replaced, not silently reused
    frame_stack.save_range(adir, "res_stage2_mask_pos_edge", full[:2], 0, 2)
    assert np.load(os.path.join(frame_stack.stack_dir(adir), "res_stage2_mask_pos_edge.npy")).shape[0] == 2


def test_driver_gives_the_same_result_from_stacks_and_from_pngs(tmp_path):
    stacks, root = _tree(tmp_path)
    factory = lambda sd1, sd2: _StandInPipeline()
    rep_png = frame_io.stylize_character(str(tmp_path), "u", pipeline_factory=factory, workers=2)
    want = {a: np.stack([np.asarray(Image.open(root / a / frame_io.STAGE2_RES / n)) for n in frame_io.list_frames(str(root / a))])
            for a in ("walk", "rest_pose")}
    for a in ("walk", "rest_pose"):
        frame_stack.pack_action(str(root / a))
    # two ranks over the stacks, no PNG written
    for a in ("walk", "rest_pose"):
        for f in os.listdir(root / a / frame_io.STAGE2_RES):
            os.remove(root / a / frame_io.STAGE2_RES / f)
    reps = [frame_io.stylize_character(str(tmp_path), "u", pipeline_factory=factory, rank=r, world=2) for r in range(2)]
    assert sum(r.frames for r in reps) == rep_png.frames == 7
    for a in ("walk", "rest_pose"):
        assert not os.listdir(root / a / frame_io.STAGE2_RES)
        got = np.load(os.path.join(frame_stack.stack_dir(str(root / a)), frame_io.STAGE2_RES + ".npy"))
        assert np.array_equal(got, want[a])
        mid = np.load(os.path.join(frame_stack.stack_dir(str(root / a)), frame_io.STAGE1_RES + ".npy"))
        assert np.array_equal(mid[..., :3], 255 - stacks[a][0][..., :3])
    # and back to the reference's PNG layout
    assert frame_stack.unpack_action(str(root / "walk"), frame_io.STAGE2_RES, workers=2) == 5
    back = np.stack([np.asarray(Image.open(root / "walk" / frame_io.STAGE2_RES / n)) for n in frame_io.list_frames(str(root / "walk"))])
    assert np.array_equal(back, want["walk"])
    # the command line: --pack / --unpack
    assert frame_io.main(["--root", str(tmp_path), "--uid", "u", "--pack"]) == 0
    assert frame_io.main(["--root", str(tmp_path), "--uid", "u", "--unpack"]) == 0


def test_stack_without_edge_needs_a_pipeline_that_derives_it(tmp_path):
    stacks, root = _tree(tmp_path)
    adir = str(root / "walk")
    color, pos, _ = stacks["walk"]
    frame_stack.write_stack(adir, frame_io.list_frames(adir), color, pos, None)
    with pytest.raises(FileNotFoundError):
        frame_io.stylize_character(str(tmp_path), "u", pipeline_factory=lambda a, b: _StandInPipeline(), stack=True)

    class Deriving(_StandInPipeline):
        derive_edge = True

        def run_host(self, color, pos, edge, out, keep_stage1=False):
            assert edge is None
            out.copy_(color)
            return color.clone() if keep_stage1 else out

    frame_stack.write_stack(str(root / "rest_pose"), frame_io.list_frames(str(root / "rest_pose")), *stacks["rest_pose"][:2])
    rep = frame_io.stylize_character(str(tmp_path), "u", pipeline_factory=lambda a, b: Deriving(), stack=True)
    assert rep.frames == 7


@pytest.mark.gpu
def test_engine_over_stacks_equals_engine_over_pngs(tmp_path):
    """Real engine (deterministic mode: bitwise reproducible) over the same clip as a PNG tree and as raw frame stacks:
    identical uint8 results; and without edge.npy the stage-2 ingest derives the edges from pos (run_render.py:31-57) -
    identical to feeding the edge maps pos2edge would have written."""
    from drawingspinup_b200.pipeline import StylizationPipeline
    from oracle import reference_port as rp
    sd1 = synth.to_torch_state_dict(synth.make_state_dict(1, seed=21, out_gain=0.25))
    sd2 = synth.to_torch_state_dict(synth.make_state_dict(2, seed=22, out_gain=0.25))
    stacks = synth.write_character_tree(str(tmp_path), "u", {"walk": 5}, 32, 48, seed=4, state_dicts=(sd1, sd2))
    color, pos, _ = stacks["walk"]
    adir = str(tmp_path / "u" / "mesh" / "blender_render" / "walk")
    edge = np.stack([255 - rp.pos2edge(pos[i]) for i in range(5)]).astype(np.uint8)        # what run_render.py:117-120 writes to edge/NNNN.png
    for i, n in enumerate(frame_io.list_frames(adir)):
        Image.fromarray(edge[i]).save(os.path.join(adir, "edge", n))
    pipe = StylizationPipeline(sd1, sd2, "cuda:0", batch=2, deterministic=True)
    frame_io.stylize_character(str(tmp_path), "u", pipeline=pipe)
    want = np.stack([np.asarray(Image.open(os.path.join(adir, frame_io.STAGE2_RES, n))) for n in frame_io.list_frames(adir)])
    frame_stack.pack_action(adir)
    for r in range(2):
        frame_io.stylize_character(str(tmp_path), "u", pipeline=pipe, rank=r, world=2, stack=True)
    got = np.load(os.path.join(frame_stack.stack_dir(adir), frame_io.STAGE2_RES + ".npy"))
    assert np.array_equal(got, want)
    os.remove(os.path.join(frame_stack.stack_dir(adir), "edge.npy"))
    pipe_d = StylizationPipeline(sd1, sd2, "cuda:0", batch=2, deterministic=True, derive_edge=True)
    frame_io.stylize_character(str(tmp_path), "u", pipeline=pipe_d, stack=True)
    got_d = np.load(os.path.join(frame_stack.stack_dir(adir), frame_io.STAGE2_RES + ".npy"))
    assert np.array_equal(got_d, want)

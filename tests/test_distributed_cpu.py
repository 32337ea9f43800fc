"""CPU, world_size 2, gloo: the only collective of the path (weight broadcast at load) and the
frame sharding that replaces any per-frame communication."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drawingspinup_b200 import synth
from drawingspinup_b200.pipeline import broadcast_state_dict, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = synth.to_torch_state_dict(synth.make_state_dict(2, seed=21)) if rank == 0 else None
        got = broadcast_state_dict(sd, src=0, device="cpu")
        want = synth.to_torch_state_dict(synth.make_state_dict(2, seed=21))
        same = list(got.keys()) == list(want.keys()) and all(
            torch.equal(got[k], want[k]) and got[k].dtype == want[k].dtype and got[k].shape == want[k].shape for k in want)
        # every rank "processes" its shard (here: a checksum per frame); gathering the shards must
        # reproduce the single-process result frame for frame
        color, _, _ = synth.make_frames(n_frames, 16, 16, seed=2)
        lo, hi = shard_range(n_frames, rank, world)
        local = torch.tensor([int(color[i].astype(np.int64).sum()) for i in range(lo, hi)], dtype=torch.int64)
        sizes = [shard_range(n_frames, r, world) for r in range(world)]
        bufs = [torch.zeros(b - a, dtype=torch.int64) for a, b in sizes]
        padded = [torch.zeros(max(b - a for a, b in sizes), dtype=torch.int64) for _ in range(world)]
        mine = torch.zeros(padded[0].numel(), dtype=torch.int64)
        mine[:local.numel()] = local
        dist.all_gather(padded, mine)
        merged = torch.cat([p[:b.numel()] for p, b in zip(padded, bufs)])
        full = torch.tensor([int(color[i].astype(np.int64).sum()) for i in range(n_frames)], dtype=torch.int64)
        ret[rank] = bool(same) and torch.equal(merged, full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8])
def test_broadcast_and_shard_world2(n_frames):
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_frames, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_broadcast_without_process_group_is_identity():
    sd = synth.to_torch_state_dict(synth.make_state_dict(1, seed=2))
    out = broadcast_state_dict(sd)
    assert list(out.keys()) == list(sd.keys()) and all(out[k] is sd[k] for k in sd)


def _char_worker(rank, world, port, ret):
    """BASELINE configs[3] on CPU: 8 characters (distinct checkpoints) -> ranks, nothing broadcast; every rank reports which
    characters it owns, a checksum of each one's weights and its frame count; all-reduced totals must tile the job."""
    from drawingspinup_b200.pipeline import assign_work
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port)
    try:
        share = assign_work(8 * 16, 8, rank, world)
        owned = torch.zeros(8, dtype=torch.int64)
        sums = torch.zeros(8, dtype=torch.float64)
        for c, nf in share.items():
            sd = synth.make_state_dict(1, seed=1234 + c, resnet_blocks=1)         # what bench.py's _weights(1234 + c) loads
            owned[c] += nf
            sums[c] = float(sum(np.asarray(v, dtype=np.float64).sum() for v in sd.values()))
        dist.all_reduce(owned)
        dist.all_reduce(sums)
        ret[rank] = (sorted(share), owned.tolist(), [round(v, 6) for v in sums.tolist()])
    finally:
        dist.destroy_process_group()


def test_character_to_rank_assignment_world2():
    from drawingspinup_b200.pipeline import assign_work
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_char_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    assert r0[0] == [0, 2, 4, 6] and r1[0] == [1, 3, 5, 7]
    assert r0[1] == r1[1] == [16] * 8                                   # every character's frames counted exactly once
    assert r0[2] == r1[2] and len(set(r0[2])) == 8                      # eight distinct checkpoints, each built by one rank
    # shapes of the job the planner must handle
    for world in (1, 2, 3, 8, 16):
        shares = [assign_work(1024, 8, r, world) for r in range(world)]
        assert sum(sum(s.values()) for s in shares) == 1024 and sorted(c for s in shares for c in s) == list(range(8))
    assert [assign_work(10, 1, r, 4) for r in range(4)] == [{0: 3}, {0: 3}, {0: 2}, {0: 2}]
    with pytest.raises(ValueError):
        assign_work(8, 2, 2, 2)

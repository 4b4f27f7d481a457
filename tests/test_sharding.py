"""Host logic of the multi-GPU path on CPU: partitioning, and a world_size-2 gloo run of the ESTM clip pipeline with
a stand-in model whose prepare/fuse mimic the hidden-state protocol (memory FIFO, stale pose)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from estdepth_b200 import sharding


def test_partition_is_contiguous_balanced_and_complete():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            ranges = [sharding.partition(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.clip_steps(20, 3, 1, 0) == (0, 18)           # 20-frame ESTM clip -> 18 forwards (BASELINE cfg3)
    assert sharding.partition(32, 8, 3) == (12, 16)              # cfg4: 32 sequences, 4 per rank
    with pytest.raises(ValueError):
        sharding.partition(4, 2, 2)


class ToyModel(object):
    """prepare: per-step feature; fuse: mixes in the memory exactly like the real protocol (order-sensitive)."""
    shape = (1, 16, 2, 3, 4)
    in_flight = 0
    max_in_flight = 0

    def prepare(self, imgs, poses, K):
        ToyModel.in_flight += 1
        ToyModel.max_in_flight = max(ToyModel.max_in_flight, ToyModel.in_flight)
        return {"x": imgs.sum() * torch.ones(self.shape), "pose": poses[:, 1]}

    def fuse(self, prep, pre_costs, pre_poses):
        ToyModel.in_flight -= 1
        v = prep["x"].clone()
        if pre_costs is not None:
            for i, (k, p) in enumerate(zip(pre_costs["values"], pre_poses)):
                v = v + 0.5 ** (i + 1) * k + p.sum()
        pose = pre_poses[-1] if pre_poses else prep["pose"]          # quirk Q4
        return {("depth", 0, 2): v.mean().reshape(1)}, {"keys": [v * 2], "values": [v]}, [pose]


def _frames(s):
    g = torch.Generator().manual_seed(s)
    return torch.rand(1, 3, 3, 4, 4, generator=g), torch.rand(1, 3, 4, 4, generator=g), torch.eye(3).unsqueeze(0)


def _sequential(n_frames):
    m, mem, out = ToyModel(), [], []
    for s in range(n_frames - 2):
        pre = sharding._flatten_memory(mem)
        o, c, p = m.fuse(m.prepare(*_frames(s)), pre[0], pre[1])
        mem.append((c, p))
        if len(mem) > 2:
            mem.pop(0)
        out.append(o[("depth", 0, 2)])
    return torch.cat(out) if out else torch.zeros(0)


def _worker(rank, world, port, n_frames, q, max_ahead=4):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipe = sharding.EstmClipPipeline(ToyModel(), window=3, memory_size=2, max_ahead=max_ahead)
    (start, stop), results = pipe.run(n_frames, _frames, ToyModel.shape, torch.device("cpu"))
    assert ToyModel.max_in_flight <= max_ahead, (ToyModel.max_in_flight, max_ahead)      # bounded look-ahead
    local = torch.cat([r[("depth", 0, 2)] for r in results]).reshape(-1, 1) if results else torch.zeros(0, 1)
    gathered = sharding.gather_maps(local)
    if rank == 0:
        q.put(torch.cat(gathered).reshape(-1).tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,max_ahead", [(9, 4), (9, 1), (4, 4), (3, 4), (2, 4)])
def test_clip_pipeline_world2_equals_sequential(n_frames, max_ahead):
    """9 frames: 7 steps cut 4 + 3 (two states travel); 4 frames: 1 + 1 (one state travels); 3 frames: rank 1 owns nothing and
    nothing travels; 2 frames: no step at all -- neither rank may wait for the other.  max_ahead = 1: prepare and fuse
    strictly alternate (bounded look-ahead)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q, max_ahead)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _sequential(n_frames)
    assert torch.allclose(torch.tensor(got), want, rtol=0, atol=0)


def _dp_worker(rank, world, port, n_seq, q):
    """BASELINE cfg4 in miniature: independent sequences partitioned over the ranks, depth maps all_gather'ed (ragged)."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, stop = sharding.partition(n_seq, world, rank)
    local = torch.stack([_frames(s)[0].sum(dim=(0, 1, 2)) for s in range(start, stop)]) if stop > start else torch.zeros(0, 4, 4)
    gathered = sharding.gather_maps(local)
    counts = [b - a for a, b in (sharding.partition(n_seq, world, r) for r in range(world))]
    known = sharding.gather_maps(local, counts=counts)               # sizes known from the partition: one collective, no host read
    assert all(torch.equal(a, b) for a, b in zip(gathered, known))
    if rank == 0:
        q.put(torch.cat(gathered).tolist())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_seq", [5, 1])
def test_data_parallel_sequences_world2_gather_in_order(n_seq):
    """No data-path collective: each rank owns a contiguous block of sequences; one ragged all_gather at the end returns
    the maps in sequence order (a rank may own nothing: n_seq = 1)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, n_seq, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.tensor(q.get(timeout=120))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.stack([_frames(s)[0].sum(dim=(0, 1, 2)) for s in range(n_seq)])
    assert got.shape == want.shape and torch.equal(got, want)

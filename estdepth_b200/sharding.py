"""Multi-GPU sharding of the inference hot path: one process per GPU, torch.distributed.

The reference has no multi-GPU inference (its eval drivers compute ``num_gpus`` and never use it,
eval_hybrid.py:68-69); what the path offers is:

* independent units -- sequences (scenes) never interact: ``partition`` assigns contiguous ranges of sequences
  to ranks, there is NO data-path collective, and ``gather_maps`` collects the depth maps at the end (one
  all_gather; 14.7 MB/rank for BASELINE config 4).
* one real dependency -- inside a scene, step k+1 needs the hidden state of step k (hybrid_depth_decoder.py:292).
  ``EstmClipPipeline`` splits ONE long ESTM sequence into contiguous clips, one per rank: every rank runs the
  memory-independent ~89 % of each of its steps (``DepthNetHybrid.prepare``) immediately, receives its
  predecessor's last ``memory_size`` hidden states with point-to-point ``recv`` (2 x 78.6 MB + a pose per state at
  480x640/D=64), runs the sequential fusion tail (``DepthNetHybrid.fuse``) and ``send``s its own last states on.
  Results are bit-identical to the single-process loop of eval_hybrid_seq.py:169-193.

Backends: "nccl" on GPUs (NVLink/NVSwitch), "gloo" in the CPU tests of the host logic.
"""
import torch
import torch.distributed as dist


def partition(n_items, world, rank):
    """Contiguous, balanced [start, stop) range of ``n_items`` for ``rank`` (first ``n_items % world`` ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def clip_steps(n_frames, window, world, rank):
    """ESTM steps (one per new frame once ``window`` frames are buffered, eval_hybrid_seq.py:169-171) owned by ``rank``:
    step s consumes frames [s, s+window).  Returns the [start, stop) range of steps."""
    return partition(max(0, n_frames - window + 1), world, rank)


def gather_maps(local, world=None):
    """all_gather of a [n_local, ...] tensor of depth maps -> list over ranks (ragged first dim allowed)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size() if world is None else world
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    nmax = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((nmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded)
    return [b[:int(c.item())] for b, c in zip(bufs, counts)]


class StateExchange(object):
    """Point-to-point hand-off of hidden states (lists of ({"keys":[k],"values":[v]}, [pose])) between neighbour ranks."""

    def __init__(self, shape, device, dtype=torch.float32):
        self.shape = tuple(shape)          # [B,16,D,H,W]
        self.device = device
        self.dtype = dtype

    def send(self, memory, dst):
        """memory: list (oldest first) of (costs dict, [pose]).  Sends the count, then key/value/pose of each entry."""
        n = torch.tensor([len(memory)], device=self.device, dtype=torch.int64)
        dist.send(n, dst)
        for costs, poses in memory:
            dist.send(costs["keys"][0].contiguous(), dst)
            dist.send(costs["values"][0].contiguous(), dst)
            dist.send(poses[0].to(self.dtype).contiguous(), dst)

    def recv(self, src):
        n = torch.zeros(1, device=self.device, dtype=torch.int64)
        dist.recv(n, src)
        memory = []
        for _ in range(int(n.item())):
            k = torch.empty(self.shape, device=self.device, dtype=self.dtype)
            v = torch.empty(self.shape, device=self.device, dtype=self.dtype)
            p = torch.empty((self.shape[0], 4, 4), device=self.device, dtype=self.dtype)
            dist.recv(k, src)
            dist.recv(v, src)
            dist.recv(p, src)
            memory.append(({"keys": [k], "values": [v]}, [p]))
        return memory


def _flatten_memory(memory):
    """lw2batch's flattening (eval_hybrid_seq.py:102-116)."""
    if not memory:
        return None, None
    return ({"keys": [c["keys"][0] for c, _ in memory], "values": [c["values"][0] for c, _ in memory]},
            [p[0] for _, p in memory])


class EstmClipPipeline(object):
    """One long ESTM sequence split into contiguous clips over the ranks (exact: same results as one process).

    ``model`` needs ``prepare(imgs, poses, K)`` and ``fuse(prep, pre_costs, pre_cam_poses)``; ``frames(s)`` returns
    the (imgs [1,window,3,H,W], cam_poses [1,window,4,4], cam_intr [1,3,3]) tensors of step ``s`` on the device.
    """

    def __init__(self, model, window=3, memory_size=2):
        self.model = model
        self.window = window
        self.memory_size = memory_size

    def run(self, n_frames, frames, state_shape, device):
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        start, stop = clip_steps(n_frames, self.window, world, rank)
        # 1) memory-independent work of every local step, concurrently on all ranks
        prepared = [self.model.prepare(*frames(s)) for s in range(start, stop)]
        # 2) wait for the predecessor's memory, then the sequential fusion tail
        xchg = StateExchange(state_shape, device)
        memory = xchg.recv(rank - 1) if (rank > 0 and start > 0) else []
        results = []
        for prep in prepared:
            pre_costs, pre_poses = _flatten_memory(memory)
            outputs, costs, poses = self.model.fuse(prep, pre_costs, pre_poses)
            memory.append((costs, poses))
            if len(memory) > self.memory_size:
                memory.pop(0)
            results.append(outputs)
        if rank + 1 < world:
            xchg.send(memory, rank + 1)
        return (start, stop), results

"""Multi-GPU sharding of the inference hot path: one process per GPU, torch.distributed.

The reference has no multi-GPU inference (its eval drivers compute ``num_gpus`` and never use it,
eval_hybrid.py:68-69); what the path offers is:

* independent units -- sequences (scenes) never interact: ``partition`` assigns contiguous ranges of sequences
  to ranks, there is NO data-path collective, and ``gather_maps`` collects the depth maps at the end (one
  all_gather; 14.7 MB/rank for BASELINE config 4).
* one real dependency -- inside a scene, step k+1 needs the hidden state of step k (hybrid_depth_decoder.py:292).
  ``EstmClipPipeline`` splits ONE long ESTM sequence into contiguous clips, one per rank: every rank runs the
  memory-independent ~89 % of each of its steps (``DepthNetHybrid.prepare``) immediately, receives its
  predecessor's last ``memory_size`` hidden states as ONE point-to-point message (2 x 78.6 MB + a pose per state at
  480x640/D=64; the ``irecv`` is posted before the local work starts), runs the sequential fusion tail
  (``DepthNetHybrid.fuse``) and ``isend``s its own last states on.
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


def gather_maps(local, world=None, counts=None):
    """all_gather of a [n_local, ...] tensor of depth maps -> list over ranks (ragged first dim allowed).

    ``counts`` (optional): the per-rank first dims when the caller already knows them (``partition`` gives them for free);
    the size exchange and its host read are skipped then, and the gather is ONE collective with no host synchronisation."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    world = dist.get_world_size() if world is None else world
    if counts is None:
        n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        got = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got, n)
        counts = [int(c.item()) for c in got]
    nmax = max(counts)
    if local.shape[0] == nmax:
        padded = local.contiguous()
    else:
        padded = torch.zeros((nmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        padded[:local.shape[0]] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, padded)
    return [out[r * nmax:r * nmax + counts[r]] for r in range(world)]


class StateExchange(object):
    """Point-to-point hand-off of the hidden-state memory (list, oldest first, of ({"keys":[k],"values":[v]}, [pose]))
    between neighbour ranks as ONE message.

    Both sides know how many entries travel -- ``count`` = min(memory_size, steps that precede the receiver's first step) --
    so there is no header and no host read: every entry's key, value and pose are packed into one flat buffer
    (2 x 78.6 MB + 64 B per entry at 480x640 / D=64), sent with a single ``isend`` and received with a single ``irecv``
    that the receiver posts BEFORE it starts its memory-independent work.  With NCCL the wait is stream-ordered (the host
    never blocks); with gloo (CPU tests) it blocks the calling thread."""

    def __init__(self, shape, device, count, dtype=torch.float32):
        self.shape = tuple(shape)          # [B,16,D,H,W]
        self.device = device
        self.dtype = dtype
        self.count = int(count)
        self.vol = 1
        for d in self.shape:
            self.vol *= int(d)
        self.pose = self.shape[0] * 16
        self.entry = 2 * self.vol + self.pose
        self._keep = None                  # buffers of an in-flight isend

    def _buffer(self):
        return torch.empty(self.count * self.entry, device=self.device, dtype=self.dtype)

    def isend(self, memory, dst):
        """memory: the sender's final memory; its last ``count`` entries travel.  Returns the request (or None)."""
        if self.count == 0:
            return None
        assert len(memory) >= self.count, "sender holds %d states, receiver expects %d" % (len(memory), self.count)
        buf = self._buffer()
        for i, (costs, poses) in enumerate(memory[-self.count:]):
            e = buf[i * self.entry:(i + 1) * self.entry]
            e[:self.vol].view(self.shape).copy_(costs["keys"][0])
            e[self.vol:2 * self.vol].view(self.shape).copy_(costs["values"][0])
            e[2 * self.vol:].view(self.shape[0], 4, 4).copy_(poses[0].to(device=self.device, dtype=self.dtype))
        req = dist.isend(buf, dst)
        self._keep = (buf, req)
        return req

    def irecv(self, src):
        """Posts the receive; returns (request, buffer) for ``unpack`` (or (None, None))."""
        if self.count == 0:
            return None, None
        buf = self._buffer()
        return dist.irecv(buf, src), buf

    def unpack(self, buf):
        """Views into the received buffer (no copies): list of (costs dict, [pose]), oldest first."""
        memory = []
        for i in range(self.count):
            e = buf[i * self.entry:(i + 1) * self.entry]
            k = e[:self.vol].view(self.shape)
            v = e[self.vol:2 * self.vol].view(self.shape)
            p = e[2 * self.vol:].view(self.shape[0], 4, 4)
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

    ``max_ahead`` bounds how many prepared steps (each pins its key / value volumes, 157 MB at 480x640 / D=64, plus the
    context maps and the initial outputs) may exist before their fuse: the memory-independent part of at most that many
    steps is issued ahead of the fusion cursor, however long the clip is.
    """

    def __init__(self, model, window=3, memory_size=2, max_ahead=4):
        self.model = model
        self.window = window
        self.memory_size = memory_size
        self.max_ahead = max(1, int(max_ahead))
        self.stats = {}

    def run(self, n_frames, frames, state_shape, device, on_result=None):
        """-> ((start, stop), [outputs of the local steps]).  ``on_result(step, outputs)``, when given, consumes each step's
        outputs as they are produced (nothing is accumulated then: results is empty)."""
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
        start, stop = clip_steps(n_frames, self.window, world, rank)
        # who exchanges: a rank receives iff it has steps and a predecessor; it sends iff its successor has steps.  Both sides
        # evaluate the SAME partition, so a send always meets its receive (empty ranges only occur at the tail).
        receives = rank > 0 and stop > start
        nxt = clip_steps(n_frames, self.window, world, rank + 1) if rank + 1 < world else (0, 0)
        sends = nxt[1] > nxt[0]
        recv_x = StateExchange(state_shape, device, min(self.memory_size, start)) if receives else None
        send_x = StateExchange(state_shape, device, min(self.memory_size, nxt[0])) if sends else None
        req, buf = recv_x.irecv(rank - 1) if receives else (None, None)          # posted before any local work

        cuda = torch.device(device).type == "cuda"
        prepared = []
        cursor = [start]

        def top_up():
            # the memory-independent ~89 % of the next steps, at most max_ahead of them ahead of the fusion cursor
            while cursor[0] < stop and len(prepared) < self.max_ahead:
                prepared.append(self.model.prepare(*frames(cursor[0])))
                cursor[0] += 1

        top_up()
        memory = []
        if req is not None:
            if cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            req.wait()                                   # NCCL: stream-ordered, the host does not block
            if cuda:
                e1.record()
                self.stats["recv_wait_events"] = (e0, e1)
            memory = recv_x.unpack(buf)
        results = []
        for s in range(start, stop):
            top_up()
            prep = prepared.pop(0)
            pre_costs, pre_poses = _flatten_memory(memory)
            outputs, costs, poses = self.model.fuse(prep, pre_costs, pre_poses)
            memory.append((costs, poses))
            if len(memory) > self.memory_size:
                memory.pop(0)
            if on_result is not None:
                on_result(s, outputs)
            else:
                results.append(outputs)
        if sends:
            sreq = send_x.isend(memory, rank + 1)
            if sreq is not None:
                sreq.wait()
        if hasattr(self.model, "check") and cuda:
            self.model.check(device)                     # end of this rank's clip: blocking fp16-range check
        return (start, stop), results

    def recv_wait_ms(self):
        """Time the stream spent between the end of the prepared-ahead work and the arrival of the predecessor's memory
        (CUDA ranks only; synchronises)."""
        ev = self.stats.get("recv_wait_events")
        if ev is None:
            return 0.0
        ev[1].synchronize()
        return float(ev[0].elapsed_time(ev[1]))

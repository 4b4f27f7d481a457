"""Multi-GPU check (not collected by pytest; run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/run_clip_pipeline_nccl.py

One ESTM sequence (9 frames -> 7 steps, 128x160, D=32, R18) is split into contiguous clips over the ranks with
`sharding.EstmClipPipeline` (NCCL send/recv of the hidden state); the gathered depth maps must be BIT-IDENTICAL to
the single-process sliding-window loop of eval_hybrid_seq.py:169-193 run on rank 0."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from estdepth_b200 import DepthNetHybrid, sharding, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = False          # identical algorithm choice on every rank
model = DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
N_FRAMES, H, W = 9, 128, 160


def frames(s):
    imgs, poses, K, _ = synth.synth_inputs(3, H, W, seed=0, start=s)
    return imgs.to(dev), poses.to(dev), K.to(dev)


with torch.no_grad():
    pipe = sharding.EstmClipPipeline(model, window=3, memory_size=2)
    (start, stop), results = pipe.run(N_FRAMES, frames, (1, 16, 32, H // 4, W // 4), dev)
    local_maps = torch.cat([r[("depth", 0, 2)] for r in results]) if results else torch.zeros(0, 1, H, W, device=dev)
    gathered = torch.cat(sharding.gather_maps(local_maps))
    if rank == 0:
        mem, ref = [], []
        for s in range(N_FRAMES - 2):
            pre = sharding._flatten_memory(mem)
            out, costs, poses = model(*frames(s), None, pre[0], pre[1], mode="val")
            mem.append((costs, poses))
            if len(mem) > 2:
                mem.pop(0)
            ref.append(out[("depth", 0, 2)])
        ref = torch.cat(ref)
        diff = (ref - gathered).abs().max().item()
        print("clip pipeline over %d GPUs: %d steps, ranks own %s..., max |pipeline - sequential| = %.3e"
              % (world, ref.shape[0], (start, stop), diff))
        assert gathered.shape == ref.shape and diff == 0.0, diff
        print("OK")
dist.barrier()
dist.destroy_process_group()

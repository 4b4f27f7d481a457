"""Does torch.linalg.inv_ex on CUDA give the same bits (a) batched vs one matrix at a time, (b) as torch.inverse on the CPU?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from estdepth_b200 import synth
g = torch.Generator().manual_seed(0)
poses = synth.camera_track(40)                       # [40,4,4] cam->world
rel = torch.stack([poses[(i * 7) % 40] @ torch.inverse(poses[i]) for i in range(40)])
K = synth.intrinsics(480, 640).clone(); K[:2] *= 0.25
for name, m in (("poses", poses), ("rel", rel), ("K", K.unsqueeze(0).repeat(4, 1, 1) * torch.tensor([1.0, 1.1, 0.9, 1.3]).view(4, 1, 1))):
    cpu = torch.stack([torch.inverse(x.unsqueeze(0))[0] for x in m])
    d = m.cuda()
    single = torch.stack([torch.linalg.inv_ex(x.unsqueeze(0))[0][0] for x in d]).cpu()
    inv_single = torch.stack([torch.inverse(x.unsqueeze(0))[0] for x in d]).cpu()
    batched = torch.linalg.inv_ex(d)[0].cpu()
    print("%-6s n=%d: inv_ex single == torch.inverse single (cuda): %s | batched == single (cuda): %s | cuda single == cpu: %s (max diff %.1e)"
          % (name, m.shape[0], torch.equal(single, inv_single), torch.equal(batched, single), torch.equal(single, cpu), float((single - cpu).abs().max())))
from estdepth_b200 import ops
one = torch.linalg.inv_ex(poses[7:8].cuda())[0].cpu()
print("padded batch of one == single: %s" % torch.equal(ops._inv(poses[7:8].cuda()).cpu(), one))
three = torch.stack([torch.linalg.inv_ex(x.unsqueeze(0))[0][0] for x in rel[:3].cuda()]).cpu()
print("padded batch of three == singles: %s" % torch.equal(ops._inv(rel[:3].cuda()).cpu(), three))
mm_cpu = poses[3:4] @ torch.inverse(poses[5:6])
mm_gpu = (poses[3:4].cuda() @ torch.linalg.inv_ex(poses[5:6].cuda())[0]).cpu()
print("pose_j @ inverse(pose_i): cuda == cpu: %s (max diff %.1e)" % (torch.equal(mm_cpu, mm_gpu), float((mm_cpu - mm_gpu).abs().max())))

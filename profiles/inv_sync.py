"""Does torch.linalg.inv_ex / small matmul on CUDA block the host (hidden synchronisation)?"""
import time, torch
dev = torch.device("cuda:0")
a = torch.randn(8192, 8192, device=dev)
m = torch.eye(4, device=dev).repeat(5, 1, 1) + 0.01 * torch.randn(5, 4, 4, device=dev)
for _ in range(3):
    torch.linalg.inv_ex(m); (m[:1] @ m[1:2])
torch.cuda.synchronize()
def probe(name, fn):
    torch.cuda.synchronize()
    for _ in range(10):
        a @ a                      # ~50 ms of queued GPU work
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-28s host time for 10 calls behind a busy GPU: %7.2f ms (GPU drained after %7.2f ms)" % (name, (t1 - t0) * 1e3, (t2 - t0) * 1e3))
probe("linalg.inv_ex [5,4,4]", lambda: torch.linalg.inv_ex(m))
probe("linalg.inv_ex [1,4,4]", lambda: torch.linalg.inv_ex(m[:1]))
probe("linalg.inv_ex [1,3,3]", lambda: torch.linalg.inv_ex(m[:1, :3, :3]))
probe("matmul [1,4,4]@[1,4,4]", lambda: m[:1] @ m[1:2])
probe("torch.inverse [1,4,4]", lambda: torch.inverse(m[:1]))
side = torch.cuda.Stream()
def on_side():
    with torch.cuda.stream(side):
        torch.linalg.inv_ex(m)
probe("inv_ex on a side stream", on_side)

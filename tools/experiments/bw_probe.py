"""HBM ceilings on this box: pure write (memset), pure read (sum), copy; all on 8 GiB buffers, CUDA events."""
import torch
dev = torch.device("cuda")
n = 1 << 32   # 4 Gi halves = 8 GiB
a = torch.empty(n, dtype=torch.float16, device=dev)
b = torch.empty(n, dtype=torch.float16, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_());  print(f"memset  8 GiB: {ms:.2f} ms  write {a.numel()*2/ms/1e6:.0f} GB/s")
ms = t(lambda: a.fill_(1.5)); print(f"fill    8 GiB: {ms:.2f} ms  write {a.numel()*2/ms/1e6:.0f} GB/s")
av = a.view(torch.float32)
ms = t(lambda: av.sum());  print(f"sum     8 GiB: {ms:.2f} ms  read  {a.numel()*2/ms/1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy    8 GiB: {ms:.2f} ms  r+w   {2*a.numel()*2/ms/1e6:.0f} GB/s")

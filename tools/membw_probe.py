"""Write-only / read-only / copy bandwidth of the B200 memory system as seen by plain kernels (torch fill_, sum, copy_):
the ceilings for the write-bound layers (1x1 expand convs) whose epilogues run at ~5 B/cycle/SM."""
import torch

dev = torch.device("cuda")
for mb in (64, 256, 2048):
    n = mb * (1 << 20) // 4
    a = torch.empty(n, device=dev)
    b = torch.empty(n, device=dev)

    def t(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3
    tw = t(lambda: a.fill_(1.0))
    tr = t(lambda: a.sum())
    tc = t(lambda: b.copy_(a))
    print("%5d MB: write-only %.0f GB/s, read-only %.0f GB/s, copy (r+w) %.0f GB/s" %
          (mb, mb * 1.048576e-3 / tw, mb * 1.048576e-3 / tr, 2 * mb * 1.048576e-3 / tc), flush=True)

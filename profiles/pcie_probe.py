"""Host <-> device copy rates of the box with pinned memory (what bounds the bench's end-to-end leg)."""
import torch

dev = torch.device("cuda", 0)
for mb in (0.5, 1, 2, 16, 64):
    n = int(mb * (1 << 20))
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 200 if mb <= 2 else 40
        h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
        d2 = torch.empty(n, dtype=torch.uint8, device=dev)
        a.record()
        for _ in range(reps):
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d.copy_(h, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps
        print(f"{mb:5.1f} MB {name:5s}: {us:8.1f} us per copy, {n / us / 1e3:6.1f} GB/s per direction")

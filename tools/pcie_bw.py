"""Pinned host <-> device copy bandwidth on this box (context for bench.py's e2e figure)."""
import torch
dev = torch.device("cuda:0")
n = 165 * (1 << 20)
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {n / 1e6:.0f} MB in {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"both directions at once: {ms:.3f} ms per {n/1e6:.0f} MB each way = {n / ms / 1e6:.1f} GB/s per direction")

import torch, time
n = 1<<28  # 2 GiB of f64? no: 256Mi doubles = 2 GiB
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, k=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/k
gb = n*8/1e9
print("H2D GB/s", gb/t(lambda: d.copy_(h, non_blocking=True)))
print("D2H GB/s", gb/t(lambda: h.copy_(d, non_blocking=True)))
h2 = torch.empty(n, dtype=torch.float64).pin_memory(); d2 = torch.empty(n, dtype=torch.float64, device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("both: each GB/s", gb/t(both))
import subprocess; print(subprocess.run("nvidia-smi -q | grep -A6 'GPU Link Info' | head -12; nproc; free -g | head -2", shell=True, capture_output=True, text=True).stdout)

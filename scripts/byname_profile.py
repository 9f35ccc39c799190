#!/usr/bin/env python
"""Where the by-name drop-in path (scripts/byname_path.py) spends its time: torch.profiler over 5 iterations, kernels by GPU time,
and the GPU-busy share of the wall time (is it launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import importlib.util
spec = importlib.util.spec_from_file_location("byname_path_mod", os.path.join(os.path.dirname(os.path.abspath(__file__)), "byname_path.py"))
src = open(spec.origin).read().split("l_by, l_fused =")[0]          # definitions only, no timing
ns_ = {"__name__": "byname_defs", "__file__": spec.origin}
exec(compile(src, spec.origin, "exec"), ns_)
by_name = ns_["by_name"]
for _ in range(3):
    by_name()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    by_name()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        by_name()
    torch.cuda.synchronize()
ka = prof.key_averages()
gpu_total = sum(e.self_device_time_total for e in ka) / 5 / 1e3
print(f"wall {wall * 1e3:.2f} ms per fwd+bwd; GPU kernel time {gpu_total:.2f} ms per fwd+bwd")
print(ka.table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=70))

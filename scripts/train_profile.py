#!/usr/bin/env python
"""Where the training step's GPU time goes (torch profiler, kernel level), config 2 on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from snde_b200.trainer import GeoNetTrainer, TrainConfig
dev = torch.device("cuda:0")
B, S, H, W = 32, 2, 128, 416
g = torch.Generator().manual_seed(1)
frames = torch.randint(0, 256, (B, 3 * (1 + S), H, W), generator=g, dtype=torch.uint8).to(dev)
K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1).to(dev)
tr = GeoNetTrainer(TrainConfig(), device=dev, ddp=False)
for _ in range(5):
    tr.step(frames[:, :3], frames[:, 3:], K)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        tr.step(frames[:, :3], frames[:, 3:], K)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

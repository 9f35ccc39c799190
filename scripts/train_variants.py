#!/usr/bin/env python
"""Training-step throughput under cuDNN settings that do not change the arithmetic type (fp32 storage, cuDNN's default TF32
convolution math): cudnn.benchmark (algorithm autotuning), channels_last memory format, fused Adam."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snde_b200.trainer import GeoNetTrainer, TrainConfig
dev = torch.device("cuda:0")
B, S, H, W = 32, 2, 128, 416
g = torch.Generator().manual_seed(1)
frames = torch.randint(0, 256, (B, 3 * (1 + S), H, W), generator=g, dtype=torch.uint8).to(dev)
K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1).to(dev)
for bench, cl, fused in itertools.product((False, True), (False, True), (False, True)):
    torch.backends.cudnn.benchmark = bench
    tr = GeoNetTrainer(TrainConfig(), device=dev, ddp=False)
    if cl:
        tr.nets.to(memory_format=torch.channels_last)
    if fused:
        tr.optimizer = torch.optim.Adam(tr.nets.parameters(), lr=2e-4, betas=(0.9, 0.999), fused=True)
    for _ in range(6):
        out = tr.step(frames[:, :3], frames[:, 3:], K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(15):
        out = tr.step(frames[:, :3], frames[:, 3:], K)
    e1.record(); torch.cuda.synchronize()
    print(f"cudnn.benchmark={bench} channels_last={cl} fused_adam={fused}: {e0.elapsed_time(e1) / 15:.2f} ms/step  loss {out['loss_total'].item():.4f}", flush=True)
    del tr
    torch.cuda.empty_cache()

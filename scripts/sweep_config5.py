#!/usr/bin/env python
"""BASELINE config 5: stand-alone warp + photometric loss sweep, single scale, S = 1, both directions, fwd+bwd,
128x416 .. 1024x2048, batch 1 .. 64, reference ('zeros') vs border padding.  Prints a markdown table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snde_b200 import ops
from snde_b200.synthetic import make_inputs

dev = torch.device("cuda:0")
print("| H x W | B | padding | kernel instance | us / call | G warped-px/s | algorithmic GB/s (87 B/px at S=1: 96 B) |")
print("|---|---|---|---|---|---|---|")
for H, W in ((128, 416), (256, 832), (512, 1024), (1024, 2048)):
    for B in (1, 8, 64):
        inp = make_inputs(B=B, S=1, H=H, W=W, num_scales=1, seed=B, smooth=False).to(dev)
        for pad in ("zeros", "border"):
            plan = ops.FusedLossPlan(inp.tgt_pyr, inp.src_pyr, inp.depth_pyr, inp.poses, inp.K_ms, want_grad=True, padding_mode=pad)
            for _ in range(3):
                plan.run()
            torch.cuda.synchronize()
            n = 20 if H * W * B < 5e7 else 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                plan.run()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            px = 2 * B * H * W
            bytes_ = 60 * px + 36 * (2 * B * H * W)      # S = 1: as many smoothness pixels as warped pixels
            print(f"| {H}x{W} | {B} | {pad} | {'lean <GRAD,FAST>' if pad == 'zeros' else 'general <EXTRA>'} | {us:.1f} | {px / us / 1e3:.2f} | {bytes_ / us / 1e3:.0f} |")
        del inp, plan
        torch.cuda.empty_cache()

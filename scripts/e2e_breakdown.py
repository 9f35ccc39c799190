#!/usr/bin/env python
"""Where the end-to-end step time goes: H2D alone, compute alone (public API, device inputs), D2H alone, host-side call cost."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from snde_b200.loss_functions import geonet_loss_from_frames
from snde_b200.synthetic import make_inputs
cfg = dict(B=32, S=2, H=128, W=416, num_scales=4)
dev = torch.device("cuda:0")
host = make_inputs(seed=7, **cfg)
pin = lambda t: t.contiguous().pin_memory()
to_u8 = lambda x: ((x + 1) * 127.5).round().clamp(0, 255).to(torch.uint8)
h_tgt = pin(to_u8(host.tgt_pyr[0])); h_src = pin(to_u8(host.src_pyr[0]).view(2, 32, 3, 128, 416).transpose(0, 1).reshape(32, 6, 128, 416))
h_dep = [pin(t) for t in host.depth_pyr]; h_poses, h_K = pin(host.poses), pin(host.K)
d_tgt, d_src = h_tgt.to(dev), h_src.to(dev); d_dep = [t.to(dev) for t in h_dep]; d_poses, d_K = h_poses.to(dev), h_K.to(dev)
outs = [torch.empty_like(t).pin_memory() for t in h_dep]
def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n
def h2d():
    d_tgt.copy_(h_tgt, non_blocking=True); d_src.copy_(h_src, non_blocking=True)
    for a, b in zip(d_dep, h_dep): a.copy_(b, non_blocking=True)
    d_poses.copy_(h_poses, non_blocking=True); d_K.copy_(h_K, non_blocking=True)
grads = []
def compute():
    dep = [t.detach().requires_grad_(True) for t in d_dep]; poses = d_poses.detach().requires_grad_(True)
    total, detail = geonet_loss_from_frames(d_tgt, d_src, dep, poses, d_K); total.backward()
    grads[:] = [t.grad for t in dep]
def d2h():
    for o, g in zip(outs, grads): o.copy_(g, non_blocking=True)
for name, fn in (("H2D 42.5 MB", h2d), ("compute (public API, device inputs)", compute), ("D2H 27 MB", d2h)):
    gpu_ms, host_ms = timed(fn)
    print(f"{name:40s} gpu {gpu_ms:.3f} ms/step   host issue {host_ms:.3f} ms/step")

if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile()
    for _ in range(5): compute()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(50): compute()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)

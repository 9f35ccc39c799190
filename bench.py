#!/usr/bin/env python
"""bench.py -- warped-pixel fwd+bwd per second of the fused rigid-warp + photometric + smoothness loss.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config1|nyu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one fused fwd+bwd of the whole loss (all scales, both directions, smoothness) over one batch of
synthetic KITTI-shaped frames (BASELINE.json configs[1]: 128x416, B=32 per GPU, S=2, 4 scales).
  value  : warped pixels / s with inputs resident in HBM (CUDA-graph replay of the C-ABI call, 4 rotating
           input sets = 540 MB > L2, CUDA events, max over ranks).
  e2e    : same metric through the public autograd API (snde_b200.loss_functions.geonet_loss_from_frames + backward) with
           HOST buffers as the reference's trainer has them: uint8 frames + intrinsics from the DataLoader, depth pyramid +
           poses from the networks.  Pinned H2D copy of every input, on-device normalisation + pyramid (one kernel per
           frame set), fused loss, D2H read of loss + all gradients, all inside the timed region (copies and compute of
           consecutive steps overlap on three streams).
  roofline : algorithmic bytes (60 B per warped pixel + 36 B per smoothness pixel) / step time.
  cpu_baseline : oracle/torch_port.py (the reference's ATen op chain restated) on this box's host cores.
--impl reference times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "config2": dict(B=32, S=2, H=128, W=416, num_scales=4, intrinsics="kitti"),
    "config1": dict(B=4, S=2, H=128, W=416, num_scales=4, intrinsics="kitti"),
    "nyu": dict(B=8, S=2, H=480, W=640, num_scales=4, intrinsics="nyu"),                 # BASELINE config 3 shape
    "hires": dict(B=8, S=4, H=376, W=1248, num_scales=4, intrinsics="kitti"),             # config 4: B=64 sharded over 8 GPUs
}
METRIC = "warped_pixel_fwd_bwd_per_sec"
UNIT = "warped-px/s"


def algorithmic_bytes(B, S, H, W, num_scales, **_):
    P = sum(int(H / 2 ** s) * int(W / 2 ** s) for s in range(num_scales))
    warped = 2 * B * S * P
    smooth = B * (1 + S) * P
    return warped, 60 * warped + 36 * smooth


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(cfg, iters=3, batch=None):
    import torch
    from oracle import torch_port
    from snde_b200.synthetic import make_inputs
    c = dict(cfg)
    if batch:
        c["B"] = batch
    torch.set_num_threads(os.cpu_count() or 1)
    inp = make_inputs(seed=0, **c)
    sec, _ = torch_port.timed_fwd_bwd(inp, iters=iters, warmup=1)
    return inp.warped_pixels() / sec, torch.get_num_threads(), f"B={c['B']} S={c['S']} {c['H']}x{c['W']} x{c['num_scales']} scales, 1 warm-up + {iters} timed fwd+bwd, median"


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU implementation of the path (ATen op chain, oracle/torch_port.py) on the
    host cores.  /root/reference is Python-only and does not travel to the GPU box, hence kind='port'."""
    if rank != 0:
        return
    import torch
    from oracle import torch_port
    from snde_b200.synthetic import make_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    # bounded sample of the same workload per step, sized so that (steps + warmup) steps take ~2 minutes at most
    per_sample_s = 0.07 * (cfg["H"] * cfg["W"]) / (128 * 416)
    c = dict(cfg); c["B"] = int(max(1, min(cfg["B"], 8, 120.0 / ((args.steps + args.warmup) * per_sample_s))))
    inp = make_inputs(seed=0, **c)
    depth = [d.clone().requires_grad_(True) for d in inp.depth_pyr]
    poses = inp.poses.clone().requires_grad_(True)
    to_u8 = lambda x: ((x + 1) * 127.5).round().clamp(0, 255).to(torch.uint8)
    tgt_u8, src_u8 = to_u8(inp.tgt_pyr[0]), to_u8(inp.src_pyr[0])          # what the DataLoader hands over
    ns = c["num_scales"]
    def step():
        for t in depth + [poses]:
            t.grad = None
        # iter_data_preparation (baseline.py:761-792) + loss path, as in one trainer iteration
        tgt_pyr, src_pyr = torch_port.frames_to_pyramid(tgt_u8, ns), torch_port.frames_to_pyramid(src_u8, ns)
        total, _, _ = torch_port.loss_path(tgt_pyr, src_pyr, depth, poses, inp.K_ms)
        total.backward()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = inp.warped_pixels() * args.steps / dt
    sample = f"B={c['B']} per step (bounded sample of B={cfg['B']}), S={c['S']} {c['H']}x{c['W']} x{c['num_scales']} scales"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: fused-loss fwd+bwd, " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the end-to-end leg (its keys are then null)")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from snde_b200 import ops
    from snde_b200.loss_functions import geonet_loss_from_frames
    from snde_b200.synthetic import make_inputs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- device-resident: 4 rotating input sets (4 x 135 MB > 126 MB L2), one captured graph each ----
    NSETS = 4
    plans = []
    for i in range(NSETS):
        d = make_inputs(seed=100 * rank + i, **cfg).to(dev)
        plans.append(ops.FusedLossPlan(d.tgt_pyr, d.src_pyr, d.depth_pyr, d.poses, d.K_ms, want_grad=True))
    warped_px, algo_bytes = algorithmic_bytes(**cfg)
    stream = torch.cuda.Stream(dev)
    graphs = []
    with torch.cuda.stream(stream):
        for p in plans:
            p.run()                               # first launch outside capture (sets the smem attribute)
        stream.synchronize()
        if not args.no_graph:
            for p in plans:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    p.run()
                graphs.append(g)
    def step(i):
        if graphs:
            graphs[i % NSETS].replay()
        else:
            plans[i % NSETS].run()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()                        # samples from here to the end of the timed region (+ probe steps below)
        t_load = time.perf_counter()
        for i in range(args.warmup):
            step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        # nvidia-smi samples every 100 ms: keep the same load running (untimed) until >= 1.5 s have been observed
        while rank == 0 and time.perf_counter() - t_load < 1.5:
            for i in range(200):
                step(i)
            torch.cuda.synchronize(dev)
        clk = clocks.stop() if rank == 0 else None
    loss_check = float(plans[0].loss_out[0].item())
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * warped_px / (ms_per_step * 1e-3)

    # ---- end to end through the public API with host buffers ----
    # host inputs as the reference's DataLoader delivers them (models/baseline.py:746-770): uint8 frames, tgt (B,3,H,W) and
    # src (B,3*S,H,W), intrinsics (B,3,3); plus the network outputs the loss consumes (depth pyramid, poses)
    host = make_inputs(seed=7 + rank, **cfg)
    pin = lambda t: t.contiguous().pin_memory()
    B_, S_ = cfg["B"], cfg["S"]
    to_u8 = lambda x: ((x + 1) * 127.5).round().clamp(0, 255).to(torch.uint8)
    h_tgt = pin(to_u8(host.tgt_pyr[0]))
    h_src = pin(to_u8(host.src_pyr[0]).view(S_, B_, 3, cfg["H"], cfg["W"]).transpose(0, 1).reshape(B_, 3 * S_, cfg["H"], cfg["W"]))
    h_dep = [pin(t) for t in host.depth_pyr]
    h_poses, h_K = pin(host.poses), pin(host.K)
    h2d = sum(t.numel() * t.element_size() for t in [h_tgt, h_src] + h_dep + [h_poses, h_K])
    out_host = [torch.empty_like(t).pin_memory() for t in h_dep] + [torch.empty_like(host.poses).pin_memory(), torch.empty(3 + 3 * cfg["num_scales"]).pin_memory()]
    d2h = sum(t.numel() * 4 for t in out_host)

    # Three device input sets and three streams: the H2D copy of step i+1, the fused loss of step i and the D2H read of
    # step i-1 overlap (every byte still moves every step, inside the timed region).
    NBUF = int(os.environ.get("SNDE_E2E_NBUF", "2"))
    dev_in = [dict(tgt=torch.empty_like(h_tgt, device=dev), src=torch.empty_like(h_src, device=dev),
                   dep=[torch.empty_like(t, device=dev) for t in h_dep], poses=torch.empty_like(h_poses, device=dev),
                   K=torch.empty_like(h_K, device=dev)) for _ in range(NBUF)]
    s_h2d, s_run, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]      # inputs of set j are on the device
    ev_free = [torch.cuda.Event() for _ in range(NBUF)]    # set j has been consumed
    ev_res = [torch.cuda.Event() for _ in range(NBUF)]     # results of set j are in their staging buffers
    ev_read = [torch.cuda.Event() for _ in range(NBUF)]    # ... and have been read back
    # results leave through persistent device staging buffers: the autograd-allocated gradient tensors are released in
    # stream order on the compute stream (no cross-stream allocator traffic), the D2H copy runs on its own stream
    stage = [[torch.empty_like(t, device=dev) for t in out_host] for _ in range(NBUF)]
    step_no = [0]

    def e2e_step():
        j = step_no[0] % NBUF
        step_no[0] += 1
        d = dev_in[j]
        with torch.cuda.stream(s_h2d):
            s_h2d.wait_event(ev_free[j])
            d["tgt"].copy_(h_tgt, non_blocking=True)
            d["src"].copy_(h_src, non_blocking=True)
            for a, b in zip(d["dep"], h_dep):
                a.copy_(b, non_blocking=True)
            d["poses"].copy_(h_poses, non_blocking=True)
            d["K"].copy_(h_K, non_blocking=True)
            ev_in[j].record(s_h2d)
        with torch.cuda.stream(s_run):
            s_run.wait_event(ev_in[j])
            dep = [t.detach().requires_grad_(True) for t in d["dep"]]
            poses = d["poses"].detach().requires_grad_(True)
            total, detail = geonet_loss_from_frames(d["tgt"], d["src"], dep, poses, d["K"])
            total.backward()
            s_run.wait_event(ev_read[j])                       # the staging set of two steps ago has been read back
            torch._foreach_copy_(stage[j], [t.grad for t in dep] + [poses.grad, detail])
            ev_free[j].record(s_run)
            ev_res[j].record(s_run)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_res[j])
            for o, g in zip(out_host, stage[j]):
                o.copy_(g, non_blocking=True)
            ev_read[j].record(s_d2h)

    e2e_steps = 0 if args.no_e2e else max(5, min(args.steps, 100))
    for _ in range(3 if e2e_steps else 0):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    for s_ in (s_h2d, s_run, s_d2h):
        s_.wait_event(e0)
    for _ in range(e2e_steps):
        e2e_step()
    for s_ in (s_h2d, s_run, s_d2h):
        cur.wait_stream(s_)
    e1.record(cur)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * warped_px * e2e_steps / (e2e_ms * 1e-3) if e2e_steps else None
    e2e_loss = float(out_host[-1][0]) if e2e_steps else None

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the strip kernel, one launch, from the committed ncu capture
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: fused rigid-warp+SSIM/L1+smoothness loss fwd+bwd, B={cfg['B']} per GPU, S={cfg['S']}, "
                                   f"{cfg['H']}x{cfg['W']}, {cfg['num_scales']} scales", "warped_px_per_step_per_gpu": warped_px,
                       "l2": f"{NSETS} rotating input sets ({NSETS} x {plans[0].num_tiles} tiles, > L2)", "cuda_graph": bool(graphs),
                       "parallelism": f"dp{world} (batch sharded, no data-path collective)", "loss_check": loss_check},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps if e2e_steps else None, "loss_check": e2e_loss},
            "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_step": algo_bytes, "traffic_source": traffic_src,
                         "note": "duration = whole step (prep + strip kernel + finalize); the strip kernel is 98% of it "
                                 "(profiles/). Bytes per SURVEY 8(d): 60 B/warped px + 36 B/smooth px, as for separate fwd and bwd "
                                 "passes; the single-pass kernel moves less and is bound by instruction issue and dependent-issue latency (DESIGN.md section 5)."},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_baseline(cfg, iters=3, batch=min(cfg["B"], 8))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

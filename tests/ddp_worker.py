"""torchrun worker of tests/test_gpu_trainer.py::test_ddp_nccl_two_gpus_...: every rank trains one step on its shard under DDP
(NCCL); rank 0 then repeats the step alone on the whole batch and writes both results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def frames(B, S, H, W, seed=3):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3 * (1 + S), H // 8, W // 8, generator=g)
    fr = (torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False).clamp(0, 1) * 255).round().to(torch.uint8)
    K = torch.tensor([[0.58 * W, 0, 0.5 * W], [0, 1.92 * H, 0.5 * H], [0, 0, 1]]).repeat(B, 1, 1)
    return fr[:, :3].contiguous(), fr[:, 3:].contiguous(), K


def main():
    from snde_b200.sharding import shard_range
    from snde_b200.trainer import GeoNetTrainer, TrainConfig
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dist.init_process_group("nccl", device_id=dev)
    B, S, H, W = 8, 2, 128, 416
    tgt, src, K = (t.to(dev) for t in frames(B, S, H, W))
    tr = GeoNetTrainer(TrainConfig(seed=5), device=dev)
    lo, hi = shard_range(B, rank, world)
    out = tr.step(tgt[lo:hi], src[lo:hi], K[lo:hi])
    loss = out["loss_total"].clone()
    dist.all_reduce(loss)
    loss /= world
    l2 = np.array([float(p.grad.double().norm()) for p in tr.nets.parameters()])
    after = np.array([float(p.detach().double().abs().sum()) for p in tr.nets.parameters()])
    nbytes = tr.gradient_bytes()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        single = GeoNetTrainer(TrainConfig(seed=5), device=dev, ddp=False)
        o = single.step(tgt, src, K)
        np.savez(sys.argv[1], loss_ddp=loss.item(), loss_single=o["loss_total"].item(), l2_ddp=l2,
                 l2_single=np.array([float(p.grad.double().norm()) for p in single.nets.parameters()]), after_ddp=after,
                 after_single=np.array([float(p.detach().double().abs().sum()) for p in single.nets.parameters()]), bytes=nbytes)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Stand-alone NCCL timings at the message sizes of the data-parallel step (run under torchrun):
all-reduce of the GRU gradient buckets (fp32 / bf16), all-gather of the encoder-MLP factors.  One JSON line per case
(rank 0), CUDA-event timed, max over ranks."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def timed(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    cases = []
    for mb, dt in ((12.6, torch.float32), (25.2, torch.float32), (50.3, torch.float32), (76.6, torch.float32),
                   (38.3, torch.bfloat16), (190.0, torch.float32)):
        n = int(mb * 1e6 / torch.empty((), dtype=dt).element_size())
        x = torch.zeros(n, device=dev, dtype=dt)
        ms = timed(lambda: dist.all_reduce(x))
        cases.append({"op": "all_reduce", "dtype": str(dt), "MB": mb, "ms": ms, "algbw_GB/s": mb / ms})
        del x
    for mb_rank in (1.57, 4.7, 9.4):
        n = int(mb_rank * 1e6 / 2)
        i = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        o = torch.zeros(n * world, device=dev, dtype=torch.bfloat16)
        ms = timed(lambda: dist.all_gather_into_tensor(o, i))
        cases.append({"op": "all_gather", "MB_per_rank": mb_rank, "ms": ms, "recv_GB/s": mb_rank * (world - 1) / ms})
    n = int(1.57e6 / 2)
    ins = [torch.zeros(n, device=dev, dtype=torch.bfloat16) for _ in range(6)]
    outs = [torch.zeros(n * world, device=dev, dtype=torch.bfloat16) for _ in range(6)]

    def grouped():
        with dist._coalescing_manager(device=dev, async_ops=False):
            for o, i in zip(outs, ins):
                dist.all_gather_into_tensor(o, i)
    cases.append({"op": "all_gather x6 grouped", "MB_per_rank": 1.57, "ms": timed(grouped)})
    if rank == 0:
        for c in cases:
            c["world"] = world
            print(json.dumps(c), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Scratch: what transport does NCCL pick on this box, and what does the accumulator-sized all_reduce cost?"""
import os, time
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if local == 0:
    print("peer access 0->1:", torch.cuda.can_device_access_peer(0, 1), flush=True)
for dtype, n in ((torch.int64, 9283304), (torch.int32, 9283304), (torch.float32, 2 * 9283304)):
    x = torch.ones(n, dtype=dtype, device="cuda")
    for _ in range(3):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.all_reduce(x); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    if local == 0:
        print(dtype, n * x.element_size() / 1e6, "MB: all_reduce ms", [round(t, 2) for t in ts], flush=True)
dist.destroy_process_group()

"""Probe (2+ GPUs, torchrun): does torch symmetric memory rendezvous work on this box, and do raw peer pointers
read each other's data?  Prints one line per rank."""
import os
import sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(4096, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peers = [hdl.get_buffer(r, (4096,), torch.float32) for r in range(world)]
    vals = [float(p[0]) for p in peers]
    hdl.barrier(channel=0)
    print(f"rank {rank}: symmetric memory ok, peer values {vals}, ptrs {[hex(x) for x in hdl.buffer_ptrs]}", flush=True)
except Exception as exc:                                            # noqa: BLE001
    print(f"rank {rank}: symmetric memory FAILED: {type(exc).__name__}: {exc}", flush=True)
dist.barrier()
dist.destroy_process_group()

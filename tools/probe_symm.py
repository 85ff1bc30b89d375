"""Probe: torch symmetric memory across processes (peer pointers for hand-written NVLink kernels)."""
import os, sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(4096, dtype=torch.float32, device=dev)
h = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs],
      "multicast", hex(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else None, "signal_pad_size", h.signal_pad_size, flush=True)
t.fill_(float(rank + 1))
dist.barrier(); torch.cuda.synchronize()
peer = h.get_buffer((rank + 1) % world, (4096,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
dist.barrier()
dist.destroy_process_group()

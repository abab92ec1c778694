import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((1 << 20,), dtype=torch.uint8, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "backend", symm_mem.get_backend(dev) if hasattr(symm_mem, "get_backend") else None, "multicast support", hdl.has_multicast_support if hasattr(hdl, "has_multicast_support") else None,
      "mc ptr", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "size", hdl.buffer_size, flush=True)
t.fill_(rank + 1)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.uint8)
print(rank, "peer value", int(peer[0].item()), flush=True)
dist.barrier(); dist.destroy_process_group()

"""2-rank probe: does torch's symmetric memory (CUDA VMM + NVSwitch multicast) rendezvous on this box?"""
import os, sys, traceback
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
name = dist.group.WORLD.group_name
for attempt in ("plain", "enable_group"):
    try:
        if attempt == "enable_group":
            symm_mem.enable_symm_mem_for_group(name)
        t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
        t.fill_(rank + 1)
        hdl = symm_mem.rendezvous(t, name)
        print(rank, attempt, "OK backend", symm_mem.get_backend(dev), "multicast", hdl.has_multicast_support(),
              "mc_ptr", hex(hdl.multicast_ptr), "bufs", [hex(p) for p in hdl.buffer_ptrs], "pads",
              [hex(p) for p in hdl.signal_pad_ptrs], "pad size", hdl.signal_pad_size, "buffer_size", hdl.buffer_size,
              "rank", hdl.rank, "world", hdl.world_size, flush=True)
        hdl.barrier()
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", name)
        torch.cuda.synchronize()
        print(rank, "all-reduce result", float(t[0]), flush=True)
        break
    except Exception:
        print(rank, attempt, "FAILED", flush=True)
        traceback.print_exc()
dist.barrier(); dist.destroy_process_group()

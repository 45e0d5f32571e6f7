"""N-rank probe of the fused in-switch all-reduce (sbr_adam_step_mc with apply_adam = 0): sums == closed form?"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import sibrar_b200  # noqa
from sibrar_b200 import ops
from sibrar_b200.parallel import DataParallelTrainer
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
total = int(sys.argv[1]) if len(sys.argv) > 1 else 688512
fake = types.SimpleNamespace(world=world)
mc = DataParallelTrainer._setup_multicast(fake, total, dev)
for k in ("grads_hdl", "sums_hdl", "flags_hdl"):
    h = mc[k]
    print(rank, k, "offset", getattr(h, "offset", None), "buffer_size", h.buffer_size, "local ptr",
          hex(h.buffer_ptrs[rank]), "tensor ptr", hex(mc[k[:-4]].data_ptr()), "mc", hex(h.multicast_ptr), flush=True)
g = mc["grads"]
idx = torch.arange(g.numel(), device=dev, dtype=torch.float32)
p = torch.zeros(total, device=dev)
plan = ops.AdamPlan([dict(param=p, grad=g[:total], exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))], dev)
step = torch.ones(1, dtype=torch.int64, device=dev)
for it in range(3):
    g.copy_((idx % 1000) * (rank + 1) + it)
    torch.cuda.synchronize(); dist.barrier()
    plan.step_mc(mc["comm"], mc["grid"], 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, step, 1.0, apply_adam=False)
    torch.cuda.synchronize()
    want = (idx % 1000) * (world * (world + 1) // 2) + it * world
    bad = (mc["sums"] != want).nonzero().reshape(-1)
    print(rank, "iter", it, "wrong entries", bad.numel(), "of", g.numel(),
          ("first", int(bad[0]), "last", int(bad[-1]), "got", float(mc["sums"][bad[0]]), "want", float(want[bad[0]])) if bad.numel() else "",
          flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
a.record()
for it in range(50):
    plan.step_mc(mc["comm"], mc["grid"], 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, step, 1.0, apply_adam=True)
b.record(); torch.cuda.synchronize()
print(rank, f"fused all-reduce + adam over {total} floats: {a.elapsed_time(b) / 50 * 1e3:.1f} us / call", flush=True)
dist.barrier(); dist.destroy_process_group()

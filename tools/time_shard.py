"""Timing experiments for the shard exchange (torchrun, N ranks): whole-scan time and per-kernel times of the sharded PCI scan."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kubevirt-gpu-device-plugin_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
import kvgpu
from oracle import oracle as O
import bench as B

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
text = B.load_pciids(); ids = O.nv_ids(text)
n = int(os.environ.get("EXP_N", 1_000_000)); gbits = B.GROUP_BITS_FOR(n * world)
ctx = kvgpu.Context(local)
ext = torch.cuda.ExternalStream(ctx.stream, device=local)
pad = ctx.text_pad(len(text))
h_text = np.full(pad + 16, 10, dtype=np.uint8); h_text[:len(text)] = np.frombuffer(text, dtype=np.uint8)
d_text = torch.from_numpy(h_text).cuda()
d_recs = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.dev_gen_pci(d_recs.data_ptr(), rank * n, n, ids, gbits)
ctx.dev_pciids_parse(d_text.data_ptr(), len(text), pad + 16, 1)
def bcast(b, src):
    if world == 1: return b
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == src: t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
    dist.broadcast(t, src); return bytes(t.cpu().numpy().tobytes())
def allgather(b):
    if world == 1: return [b]
    t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t); return [bytes(o.cpu().numpy().tobytes()) for o in outs]
sh = kvgpu.ShardedScan(ctx, rank, world, bcast, allgather, n + 1)
def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()
for _ in range(5):
    sh.scan_device_shard(d_recs.data_ptr(), n)
sync()
evs = []
for _ in range(30):
    ctx.dev_flush_l2()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext); sh.scan_device_shard(d_recs.data_ptr(), n); e1.record(ext); evs.append((e0, e1))
sync()
ms = sorted(a.elapsed_time(b) for a, b in evs)
per = {}
for _ in range(5):
    ctx.dev_flush_l2(); ctx.set_kernel_timing(True)
    sh.scan_device_shard(d_recs.data_ptr(), n)
    for name, t in ctx.kernel_times(): per.setdefault(name, []).append(t)
ctx.set_kernel_timing(False)
sync()
print("rank", rank, "scan_ms median %.4f min %.4f" % (ms[len(ms) // 2], ms[0]),
      {k: round(1e3 * sum(v) / 5, 1) for k, v in per.items()}, flush=True)
del d_recs, d_text
torch.cuda.synchronize()
sh.close(); ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import kvgpu, util
from oracle import oracle as O
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29755")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1"); os.environ.setdefault("LOCAL_RANK", "0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200003
mode = sys.argv[2] if len(sys.argv) > 2 else "p2p"
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
text = util.pciids_text(); ids = O.nv_ids(text)
ctx = kvgpu.Context(0); ctx.pciids_load(text)
def bcast(b, src): return b
def allgather(b): return [b]
sh = kvgpu.ShardedScan(ctx, 0, 1, bcast, allgather if mode == "p2p" else None, n + 1)
print("mode", sh.mode)
buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
ctx.dev_gen_pci(buf.data_ptr(), 0, n, ids, 17)
ctx.dev_scan_pci(buf.data_ptr(), n)
ref = ctx.dev_scan_pci_fetch()
sh.scan_device_shard(buf.data_ptr(), n)
res = sh.fetch()
print("local", len(res.local), "ref surv", len(ref.survivors), "equal", np.array_equal(res.local, ref.survivors))
print("dev members", len(res.dev.survivors), "keys", len(res.dev.dev_keys), "ref keys", len(ref.dev_keys))
print("dev members == surv", np.array_equal(res.dev.survivors, ref.survivors))
print("grp members == surv", np.array_equal(res.grp.survivors, ref.survivors))
print("dev_keys eq", np.array_equal(res.dev.dev_keys, ref.dev_keys), "off eq", np.array_equal(res.dev.dev_off, ref.dev_off), "perm eq", np.array_equal(res.dev.dev_perm, ref.dev_perm), "name eq", np.array_equal(res.dev.dev_name_slot, ref.dev_name_slot))
print("grp_keys eq", np.array_equal(res.grp.grp_keys, ref.grp_keys), "off eq", np.array_equal(res.grp.grp_off, ref.grp_off), "perm eq", np.array_equal(res.grp.grp_perm, ref.grp_perm))
if not np.array_equal(res.dev.dev_keys, ref.dev_keys):
    print(res.dev.dev_keys[:10], ref.dev_keys[:10], res.dev.dev_off[:5], ref.dev_off[:5])
if not np.array_equal(res.grp.grp_keys, ref.grp_keys):
    print(len(res.grp.grp_keys), len(ref.grp_keys), res.grp.grp_keys[:10], ref.grp_keys[:10])
a = kvgpu.canonical_dump(kvgpu.pci_maps_from_result(ref)); b = kvgpu.canonical_dump(kvgpu.merge_parts([kvgpu.pci_maps_from_shard(res)]))
print("dump equal", a == b)
if a != b:
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            print("first diff line", i, x[:100], y[:100]); break
    print(len(la), len(lb))
d = res.dev.survivors; s = ref.survivors
bad = np.nonzero(d.view(np.uint32).reshape(-1, 4) != s.view(np.uint32).reshape(-1, 4))[0]
bad = np.unique(bad)
print("bad records", len(bad), "first", bad[:20], "tiles", np.unique(bad // 2048)[:40], "within-tile", np.unique(bad % 2048)[:10], np.unique(bad % 2048)[-10:])
print("zero records", int((d.view(np.uint32).reshape(-1, 4).sum(axis=1) == 0).sum()))
g = res.grp.survivors
badg = np.unique(np.nonzero(g.view(np.uint32).reshape(-1, 4) != s.view(np.uint32).reshape(-1, 4))[0])
print("grp bad", len(badg), np.unique(badg // 2048)[:40])

"""Worker for tests/test_sharding_gloo.py: one rank of a world_size-N gloo job on CPU.

Checks the sharding theory of BASELINE.json config 4 without a GPU: range shards classified
independently and concatenated in rank order by the allgatherv ARE the unsharded result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))

import numpy as np
import torch
import torch.distributed as dist

import kvgpu
from oracle import oracle as O


def alive(r):
    return (r["vendor"] == 0x10de) & ((r["flags"] & 15) == 0) & ((r["driver"] == 1) | (r["driver"] == 2))


def main():
    n = int(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = np.arange(0x1b00, 0x1c00, dtype=np.uint16)
    lo, hi = kvgpu.shard_range(n, rank, world)
    shard = O.gen_pci(lo, hi - lo, ids, 12)          # counter-based: a shard is generated in place
    local = shard[alive(shard)]                        # per-rank classification (checker code here)
    surv = np.zeros(len(local), dtype=kvgpu.PCI_SURV)
    surv["addr"], surv["iommu_group"], surv["device"] = local["addr"], local["iommu_group"], local["device"]
    surv["numa"] = np.where((local["flags"] & 16) | (local["numa"] < 0), 0, local["numa"])
    t = torch.from_numpy(surv.view(np.uint8).copy())
    out, counts = kvgpu.allgatherv_torch(t)
    got = np.frombuffer(out.numpy().tobytes(), dtype=kvgpu.PCI_SURV)
    # the unsharded truth, computed by every rank
    full = O.gen_pci(0, n, ids, 12)
    keep = full[alive(full)]
    assert sum(counts) == len(keep) * 16, (counts, len(keep))
    assert np.array_equal(got["addr"], keep["addr"]), "rank-ordered concatenation is not Walk order"
    assert np.array_equal(got["iommu_group"], keep["iommu_group"])
    assert np.array_equal(got["numa"], np.where((keep["flags"] & 16) | (keep["numa"] < 0), 0, keep["numa"]))
    # byte broadcast used for the 128-byte NCCL unique id
    blob = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        blob.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(blob, 0)
    assert blob.tolist() == list(range(128))
    dist.barrier()
    if rank == 0:
        print("gloo-ok world=%d n=%d survivors=%d counts=%s" % (world, n, len(keep), counts))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

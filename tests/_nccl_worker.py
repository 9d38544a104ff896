"""Worker for the multi-GPU parity test: rank r scans its shard on GPU r; the gathered result on
every rank must equal the oracle on the unsharded snapshot."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import kvgpu
import util
from oracle import oracle as O


def main():
    n = int(sys.argv[1])
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    text = util.pciids_text()
    ids = O.nv_ids(text)
    ctx = kvgpu.Context(local_rank)
    ctx.pciids_load(text)

    def bcast(b, src):
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == src:
            t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
        dist.broadcast(t, src)
        return bytes(t.cpu().numpy().tobytes())

    def allgather(b):
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [bytes(o.cpu().numpy().tobytes()) for o in outs]

    mode = sys.argv[2] if len(sys.argv) > 2 else "p2p"
    sh = kvgpu.ShardedScan(ctx, rank, world, bcast, allgather if mode == "p2p" else None,
                           (n + world - 1) // world + 1)
    if mode == "p2p" and sh.mode != "p2p" and rank == 0:
        print("note: peer-memory path unavailable, fell back to NCCL")
    lo, hi = kvgpu.shard_range(n, rank, world)
    buf = torch.empty(max(hi - lo, 1) * 16, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.dev_gen_pci(buf.data_ptr(), lo, hi - lo, ids, 17)
    trace = os.environ.get("KVG_WORKER_TRACE") == "1"
    for rep in range(4):   # > 2: exercises window reuse and the consumed-acks
        if trace:
            print("rank %d rep %d scan" % (rank, rep), file=sys.stderr, flush=True)
        sh.scan_device_shard(buf.data_ptr(), hi - lo)
        res = sh.fetch()
        if trace:
            print("rank %d rep %d fetched S=%d" % (rank, rep, len(res.survivors)), file=sys.stderr, flush=True)
        part = kvgpu.pci_maps_from_result(res)
        # the bucketing is partitioned by key: rank r owns the keys with key % world == r
        assert all(int(k, 16) % world == rank for k in part.deviceMap), "foreign device key"
        assert all(int(k) % world == rank for k in part.iommuMap), "foreign iommu group"
        parts = [None] * world
        dist.all_gather_object(parts, (part.deviceMap, part.iommuMap, part.deviceNames))
        merged = kvgpu.Maps(bdfToIommuMap=part.bdfToIommuMap)   # survivor list is replicated
        for dm, im, nm in parts:                                # disjoint key sets: plain union
            assert not (set(dm) & set(merged.deviceMap)) and not (set(im) & set(merged.iommuMap))
            merged.deviceMap.update(dm)
            merged.iommuMap.update(im)
            merged.deviceNames.update(nm)
        got = kvgpu.canonical_dump(merged)
        m = O.Maps()
        m.create_iommu_device_map_flat(O.gen_pci(0, n, ids, 17))
        want = m.dump(text)
        assert got == want, "rank %d rep %d: sharded dump differs (%s vs %s)" % (
            rank, rep, hashlib.sha256(got).hexdigest()[:12], hashlib.sha256(want).hexdigest()[:12])
    dist.barrier()
    if rank == 0:
        print("nccl-ok world=%d n=%d sha=%s exchange=%s" % (world, n, hashlib.sha256(got).hexdigest()[:16], sh.mode))
    sh.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    # a failing rank must take the job down instead of hanging in collective teardown: dump where
    # every thread is if the run stalls, and leave without running destructors on an exception
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(int(os.environ.get("KVG_WORKER_STALL_S", "150")), exit=True)
    try:
        main()
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)

"""Multi-GPU sharding of the scan (BASELINE.json config 4): one process per GPU.

Records are range-partitioned in Walk order, rank r owns [r*N/P, (r+1)*N/P).  Every rank classifies its
shard; its survivors stay local (rank order == Walk order: the host concatenates them for bdfToIommuMap).
ONE exchange step sends every survivor to the OWNER of its key (key % P), once per group-by map, so each
rank ends up with all members of the keys it owns — constant volume per GPU whatever P is.  Transport:
stores into the owners' peer windows over NVLink (CUDA IPC), or — fallback — one NCCL allgatherv of the
survivor lists followed by a local select.  The pci.ids table is parsed by every rank itself.

`ShardedScan` needs only byte collectives (broadcast for the 128-byte NCCL unique id, all-gather for the
64-byte IPC handles), so the same class is driven by torch.distributed (bench.py) or any other launcher.
"""
from __future__ import annotations

from .context import Context, MdevShardResult, PciResult, PciShardResult
from .plugin import Maps, mdev_maps_from_result, pci_maps_from_result


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; shards tile [0, n) exactly, sizes differ by <= 1."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world


def concat_in_rank_order(parts: list) -> list:
    """The host-side statement of what the allgatherv does: shard outputs concatenated in rank
    order are the global Walk-order list (used by the gloo CPU tests)."""
    out = []
    for p in parts:
        out.extend(p)
    return out


def allgatherv_torch(local, dist_mod=None):
    """Backend-agnostic statement of the exchange step (same algorithm as the NCCL path in
    kvg_dev_scan_pci_sharded): all-gather the per-rank counts, then one broadcast per root into the
    rank-ordered output.  `local` is a 1-D torch tensor; works on gloo (CPU tests) and nccl."""
    import torch
    import torch.distributed as dist
    dist = dist_mod or dist
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    out = torch.empty(sum(counts), dtype=local.dtype, device=local.device)
    off = 0
    for root in range(world):
        seg = out[off:off + counts[root]]
        if root == rank:
            seg.copy_(local)
        if counts[root]:
            dist.broadcast(seg, root)
        off += counts[root]
    return out, counts


def pci_maps_from_shard(sh: PciShardResult) -> Maps:
    """This rank's part of the three PCI maps: deviceMap / iommuMap for the keys it owns (all members),
    bdfToIommuMap for its own shard's survivors."""
    m = pci_maps_from_result(sh.dev)
    part = Maps(deviceMap=m.deviceMap, deviceNames=m.deviceNames)
    part.iommuMap = pci_maps_from_result(sh.grp).iommuMap
    e = sh.dev.grp_perm[:0]
    local = PciResult(sh.n_records, sh.local, sh.dev.dev_keys[:0], sh.dev.grp_off[:1], e, e, sh.dev.grp_keys[:0],
                      sh.dev.grp_off[:1], e, b"")
    part.bdfToIommuMap = pci_maps_from_result(local).bdfToIommuMap
    return part


def mdev_maps_from_shard(sh: MdevShardResult) -> Maps:
    """This rank's part of vGpuMap (type labels it owns) and gpuVgpuMap (parents it owns)."""
    a = mdev_maps_from_result(sh.by_type)
    b = mdev_maps_from_result(sh.by_parent)
    return Maps(vGpuMap=a.vGpuMap, gpuVgpuMap=b.gpuVgpuMap, deviceNames=a.deviceNames)


def merge_parts(parts: list) -> Maps:
    """Union of the ranks' parts (rank order): key sets are disjoint, bdfToIommuMap concatenates in rank
    order == Walk order."""
    out = Maps()
    for p in parts:
        for name in ("deviceMap", "iommuMap", "vGpuMap", "gpuVgpuMap"):
            mine, theirs = getattr(out, name), getattr(p, name)
            clash = set(mine) & set(theirs)
            if clash:
                raise ValueError("%s: key owned by two ranks: %r" % (name, sorted(clash)[:3]))
            mine.update(theirs)
        out.bdfToIommuMap.update(p.bdfToIommuMap)
        out.deviceNames.update(p.deviceNames)
    return out


class ShardedScan:
    def __init__(self, ctx: Context, rank: int, world: int, broadcast_bytes, allgather_bytes=None,
                 p2p_cap: int = 0):
        """broadcast_bytes(b: bytes | None, src=0) -> bytes : collective byte broadcast (NCCL unique id).
        allgather_bytes(b: bytes) -> list[bytes] (rank order) + p2p_cap (PCI records per shard; an mdev
        record takes two): also set up the peer windows (CUDA IPC over NVLink); on any failure the NCCL
        path is used."""
        self.ctx, self.rank, self.world = ctx, rank, world
        self.mode = "nccl"
        uid = ctx.comm_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 0)
        ctx.comm_init(rank, world, uid)
        if allgather_bytes is not None and p2p_cap > 0:
            try:
                mine = ctx.comm_p2p_export(rank, world, p2p_cap)
                ok = b"\1"
            except Exception:
                mine, ok = b"\0" * 64, b"\0"
            blobs = allgather_bytes(mine + ok)            # every rank learns whether all exported
            if all(b[64:65] == b"\1" for b in blobs):
                try:
                    ctx.comm_p2p_import(b"".join(b[:64] for b in blobs))
                    good = b"\1"
                except Exception:
                    good = b"\0"
                if all(g == b"\1" for g in allgather_bytes(good)):
                    ctx.comm_p2p_enable(True)
                    self.mode = "p2p"

    def scan_device_shard(self, d_recs: int, n_local: int):
        self.ctx.dev_scan_pci_sharded(d_recs, n_local)

    def fetch(self) -> PciShardResult:
        return self.ctx.dev_scan_pci_shard_fetch()

    def scan_device_mdev_shard(self, d_recs: int, n_local: int, raw_types: list):
        self.ctx.dev_scan_mdev_sharded(d_recs, n_local, raw_types)

    def fetch_mdev(self) -> MdevShardResult:
        return self.ctx.dev_scan_mdev_shard_fetch()

    def close(self):
        self.ctx.comm_destroy()

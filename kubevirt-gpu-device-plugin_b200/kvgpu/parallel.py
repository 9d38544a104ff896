"""Multi-GPU sharding of the scan (BASELINE.json config 4): one process per GPU.

Records are range-partitioned in Walk order, rank r owns [r*N/P, (r+1)*N/P).  Every rank
classifies its shard; ONE exchange step — an allgatherv of the 16-byte survivor records over
NCCL/NVLink, rank order == Walk order — rebuilds the global survivor list on every rank, which
then runs the (replicated) bucketing.  The pci.ids table is parsed by every rank itself.

`ShardedScan` needs only a byte-broadcast callable to distribute the 128-byte NCCL unique id, so
the same class is driven by torch.distributed (bench.py) or by any other launcher.
"""
from __future__ import annotations

from .context import Context, PciResult


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


class ShardedScan:
    def __init__(self, ctx: Context, rank: int, world: int, broadcast_bytes, allgather_bytes=None,
                 p2p_cap: int = 0):
        """broadcast_bytes(b: bytes | None, src=0) -> bytes : collective byte broadcast (NCCL unique id).
        allgather_bytes(b: bytes) -> list[bytes] (rank order) + p2p_cap (records per shard): also set
        up the peer-memory exchange (CUDA IPC over NVLink); on any failure the NCCL path is used."""
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

    def fetch(self) -> PciResult:
        return self.ctx.dev_scan_pci_fetch()

    def close(self):
        self.ctx.comm_destroy()

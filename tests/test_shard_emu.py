"""The exchange step of the sharded scan (csrc/kvg_shard.cuh: k_shard_send or the fused k_classify_send, then
k_shard_gather — which also publishes) executed on the CPU from its real kernel source: P emulated ranks with their own windows and
control blocks, several back-to-back steps (window parities, acks), peer-window mode and the NCCL local mode,
16-byte (PCI) and 32-byte (mdev) records.  Every rank must end up with exactly the records whose key it owns
(key % P == rank), in Walk order (source-rank order, then position), for both orderings."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import conftest  # noqa: F401

sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "emu"))
import build as emu_build  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_shard())
    L.emu_exchange.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def keys_of(recs, units):
    w = recs.reshape(-1, units * 4)
    if units == 1:
        return w[:, 2] & 0xffff, w[:, 1]
    return w[:, 5] & 0xffff, w[:, 4]


@pytest.mark.parametrize("units", [1, 2])
@pytest.mark.parametrize("P,local_mode", [(1, 0), (2, 0), (3, 0), (8, 0), (16, 0), (3, 1)])
def test_exchange_by_owner(emu, units, P, local_mode):
    rng = np.random.default_rng(100 * P + units + local_mode)
    sizes = [int(x) for x in rng.integers(0, 5000, P)]
    sizes[0] = 4500 if P > 1 else 2300
    if P > 2:
        sizes[1] = 0                                   # an empty shard
    shards = []
    for r in range(P):
        recs = rng.integers(0, 1 << 32, (sizes[r], units * 4), dtype=np.uint64).astype(np.uint32)
        recs[:, 0] = (r << 24) | np.arange(sizes[r])   # a Walk-order tag
        shards.append(np.ascontiguousarray(recs))
    if local_mode:                                     # NCCL: every rank works on the gathered list
        gathered = np.ascontiguousarray(np.concatenate(shards))
        lists = [gathered] * P
        n = np.full(P, len(gathered), dtype=np.uint32)
        cap = max(sizes) + 1
        while P * cap < len(gathered) + 1:
            cap += 1
    else:
        lists = shards
        n = np.array(sizes, dtype=np.uint32)
        cap = max(sizes) + 1
    ptrs = (C.c_void_p * P)(*[a.ctypes.data if len(a) else None for a in lists])
    owned_cap = P * cap
    o0 = np.zeros((P, owned_cap, units * 4), dtype=np.uint32)
    o1 = np.zeros((P, owned_cap, units * 4), dtype=np.uint32)
    n_own = np.zeros(2 * P, dtype=np.uint32)
    mx = np.zeros(2 * P, dtype=np.uint32)
    rc = emu.emu_exchange(units, ptrs, n.ctypes.data, P, 4, cap, local_mode, o0.ctypes.data, o1.ctypes.data,
                          n_own.ctypes.data, mx.ctypes.data)
    assert rc == 0, rc
    allrecs = np.concatenate(shards) if P else shards[0]
    k0, k1 = keys_of(allrecs, units)
    for r in range(P):
        for o, (keys, got) in enumerate(((k0, o0), (k1, o1))):
            want = allrecs[keys % P == r]
            cnt = int(n_own[2 * r + o])
            assert cnt == len(want), (r, o, cnt, len(want))
            assert np.array_equal(got[r, :cnt], want), (r, o)
            assert int(mx[2 * r + o]) == (int(keys[keys % P == r].max()) if len(want) else 0)


@pytest.mark.parametrize("P", [1, 2, 5, 8, 16])
def test_classify_and_send_in_one_kernel(emu, P):
    """k_classify_send: raw PCI records in; every rank's dense survivor list (Walk order) and both owned lists
    out — the owned lists must equal those of the separate classify + exchange path (the oracle's drop rules
    applied with numpy, then key % P)."""
    from oracle import oracle as O
    import util
    emu.emu_classify_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ids = O.nv_ids(util.pciids_text())
    rng = np.random.default_rng(7 * P)
    sizes = [int(x) for x in rng.integers(0, 5000, P)]
    sizes[0] = 4097
    if P > 2:
        sizes[1] = 0
    nv_index = np.full(65536, 0xffffffff, dtype=np.uint32)
    nv_index[ids] = np.arange(len(ids), dtype=np.uint32) * 3 + 5
    shards, first = [], 0
    for r in range(P):
        shards.append(np.ascontiguousarray(O.gen_pci(first, sizes[r], ids, 9)))
        first += sizes[r]
    cap = max(sizes) + 1
    ptrs = (C.c_void_p * P)(*[a.ctypes.data if len(a) else None for a in shards])
    n = np.array(sizes, dtype=np.uint32)
    surv = np.zeros((P, cap, 4), dtype=np.uint32)
    n_surv = np.zeros(P, dtype=np.uint32)
    owned_cap = P * cap
    o0 = np.zeros((P, owned_cap, 4), dtype=np.uint32)
    o1 = np.zeros((P, owned_cap, 4), dtype=np.uint32)
    n_own = np.zeros(2 * P, dtype=np.uint32)
    mx = np.zeros(2 * P, dtype=np.uint32)
    rc = emu.emu_classify_exchange(ptrs, n.ctypes.data, P, 3, cap, nv_index.ctypes.data, surv.ctypes.data, n_surv.ctypes.data,
                                   o0.ctypes.data, o1.ctypes.data, n_own.ctypes.data, mx.ctypes.data)
    assert rc == 0, rc
    drop = 1 | 2 | 4 | 8
    want_all = []
    for r in range(P):
        recs = shards[r]
        alive = (recs["vendor"] == 0x10de) & ((recs["flags"] & drop) == 0) & ((recs["driver"] == 1) | (recs["driver"] == 2))
        a = recs[alive]
        numa = np.where(((a["flags"] & 16) != 0) | (a["numa"] < 0), 0, a["numa"]).astype(np.uint32)
        w = np.zeros((len(a), 4), dtype=np.uint32)
        w[:, 0] = a["addr"]
        w[:, 1] = a["iommu_group"]
        w[:, 2] = a["device"].astype(np.uint32) | (numa << 16)
        w[:, 3] = nv_index[a["device"]]
        assert int(n_surv[r]) == len(w), (r, int(n_surv[r]), len(w))
        assert np.array_equal(surv[r, :len(w)], w), r
        want_all.append(w)
    allrecs = np.concatenate(want_all)
    k0, k1 = keys_of(allrecs, 1)
    for r in range(P):
        for o, (keys, got) in enumerate(((k0, o0), (k1, o1))):
            want = allrecs[keys % P == r]
            cnt = int(n_own[2 * r + o])
            assert cnt == len(want), (r, o, cnt, len(want))
            assert np.array_equal(got[r, :cnt], want), (r, o)


def test_owner_without_division_is_exact():
    """shard_owner (kvg_shard.cuh): hi(key * floor(2^32 / P)) is the quotient or one less for every 32-bit key,
    so one conditional subtract makes key % P exact.  Checked on the corners and a random sample for every P."""
    rng = np.random.default_rng(3)
    keys = np.concatenate([np.arange(0, 70, dtype=np.uint64), np.uint64(2 ** 32) - np.arange(1, 70, dtype=np.uint64),
                           rng.integers(0, 2 ** 32, 200_000, dtype=np.uint64),
                           (np.arange(1, 17, dtype=np.uint64)[:, None] * np.arange(0, 2 ** 32, 2 ** 27, dtype=np.uint64)[None, :]).ravel() % (2 ** 32)])
    for P in range(1, 17):
        m = np.uint64(0xffffffff if P == 1 else (2 ** 32) // P)
        qd = (keys * m) >> np.uint64(32)
        r = keys - qd * np.uint64(P)
        assert np.all(r < 2 * P)
        r = np.where(r >= P, r - np.uint64(P), r)
        assert np.array_equal(r, keys % np.uint64(P)), P

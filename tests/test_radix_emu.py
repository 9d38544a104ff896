"""K4, the stable orderings of csrc/kvg_order.cuh (k_order_tilescan, k_order_scatter<8|11, 8>, k_order_final),
executed on the CPU from their real kernel source under the warp emulator of tools/emu/ and checked against a
stable numpy sort: device-chosen digit widths, multi-tile inputs, ragged last tiles, duplicate-heavy and
full-width keys; both forms of the final step (one fused launch / count - offsets - emit)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import conftest  # noqa: F401

sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "emu"))
import build as emu_build  # noqa: E402

PAIR = np.dtype([("key", "<u4"), ("idx", "<u4")])


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_radix())
    L.emu_radix_sort.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    return L


def device_sort(emu, keys, key_bits_max, max_bits, variant):
    pairs = np.zeros(len(keys) + 1, dtype=PAIR)
    pairs["key"][:len(keys)] = keys
    pairs["idx"][:len(keys)] = np.arange(len(keys), dtype=np.uint32)
    np_ = emu.emu_radix_sort(pairs.ctypes.data, len(keys), key_bits_max, max_bits, variant)
    assert np_ >= 1
    return pairs[:len(keys)], np_


def expected(keys):
    order = np.argsort(keys, kind="stable")
    return keys[order], order.astype(np.uint32)


CASES = [  # (n, key bits in the data, key_bits_max, max_bits)
    (1, 1, 32, 11), (31, 5, 16, 11), (2049, 11, 32, 11), (3000, 12, 32, 11), (40_000, 17, 32, 11),
    (5000, 19, 32, 11), (2500, 23, 32, 11), (2200, 32, 32, 11), (2300, 16, 16, 11), (2600, 24, 32, 8),
]


def want_passes(bits, kmax, mb):
    return -(-max(1, min(bits, kmax)) // mb)


def test_device_radix_sort_is_a_stable_sort(emu):
    rng = np.random.default_rng(17)
    for n, bits, kmax, mb in CASES:
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
        keys[rng.integers(0, n)] = (1 << bits) - 1          # the widest key is present: the plan sees it
        if n > 100:
            keys[rng.integers(0, n, n // 3)] = keys[0]       # a heavy bucket: long runs of equal keys
        want_k, want_i = expected(keys)
        for layout in (0, 1):   # histogram layout: digit-major (row tile scan) / tile-major (column tile scan)
            got, npass = device_sort(emu, keys, kmax, mb, layout)
            assert npass == want_passes(bits, kmax, mb), (n, bits, mb)
            assert np.array_equal(got["key"], want_k) and np.array_equal(got["idx"], want_i), (n, bits, kmax, mb, layout)


def test_skewed_input(emu):
    rng = np.random.default_rng(5)
    keys = np.concatenate([np.full(3000, 7, np.uint32), rng.integers(0, 1 << 19, 1500, dtype=np.uint64).astype(np.uint32),
                           np.zeros(700, np.uint32)])
    rng.shuffle(keys)
    want_k, want_i = expected(keys)
    for layout in (0, 1):
        a, _ = device_sort(emu, keys, 32, 11, layout)
        assert np.array_equal(a["key"], want_k) and np.array_equal(a["idx"], want_i), layout


def test_whole_ordering_permutation_segments_and_bucket_names(emu):
    """radix passes + the final step (fused k_order_final, or k_order_heads<false> -> k_tile_offsets ->
    k_order_heads<true>): the final permutation, the distinct keys with their segment offsets, and (device-id
    ordering) each bucket's joined name slot."""
    emu.emu_ordering.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(23)
    for n, bits, kmax in ((1, 3, 16), (2048, 6, 16), (4500, 16, 16), (3000, 19, 32), (2500, 25, 32), (0, 1, 32)):
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
        if n:
            keys[0] = (1 << bits) - 1
        surv = np.zeros((max(n, 1), 4), dtype=np.uint32)
        surv[:n, 3] = (keys * 7 + 1) & 0xffff          # name slot: a function of the key, like the real join
        pairs = np.zeros(n + 1, dtype=PAIR)
        pairs["key"][:n], pairs["idx"][:n] = keys, np.arange(n, dtype=np.uint32)
        perm = np.zeros(n + 1, np.uint32)
        seg_key, seg_off, seg_name = np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32)
        order = np.argsort(keys, kind="stable")
        uniq, first = np.unique(keys[order], return_index=True)
        # forms of the final step: 0 = count / offsets / emit, 1 = one chained-scan launch
        for form in (0, 1):
            p2 = pairs.copy()
            perm[:] = 0xdeadbeef
            seg_key[:] = seg_off[:] = seg_name[:] = 0xdeadbeef
            n_seg = emu.emu_ordering(p2.ctypes.data, n, surv.ctypes.data, kmax, 11, perm.ctypes.data, seg_key.ctypes.data,
                                     seg_off.ctypes.data, seg_name.ctypes.data, form, None, 0)
            assert n_seg == len(uniq), (n_seg, form)
            assert np.array_equal(perm[:n], order.astype(np.uint32))
            assert np.array_equal(seg_key[:n_seg], uniq) and np.array_equal(seg_off[:n_seg], first.astype(np.uint32))
            assert int(seg_off[n_seg]) == n
            assert np.array_equal(seg_name[:n_seg], (uniq * 7 + 1) & 0xffff)


def test_deferred_name_join_in_the_final_kernel(emu):
    """The parse may still be running when the scan classifies (side stream): the records then carry no name slot
    and the final kernel of an ordering joins it while it writes the permutation — by the ordering's key (device-id
    ordering: join 1) or by the device id read from the record (any other ordering of such records: join 2).  The
    segment heads take their name straight from the table."""
    emu.emu_ordering.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(41)
    table = rng.integers(0, 1 << 20, 65536, dtype=np.uint64).astype(np.uint32)
    for n, mode, form in ((4500, 1, 1), (4500, 1, 0), (3000, 2, 1), (2049, 2, 0), (1, 1, 1)):
        dev = rng.integers(0, 700, n, dtype=np.uint64).astype(np.uint32)       # device ids
        surv = np.zeros((n, 4), dtype=np.uint32)
        surv[:, 2] = dev | (rng.integers(0, 4, n, dtype=np.uint64).astype(np.uint32) << 16)   # device | numa << 16
        surv[:, 3] = 0xffffffff                                               # no name yet
        keys = dev if mode == 1 else rng.integers(0, 1 << 19, n, dtype=np.uint64).astype(np.uint32)
        pairs = np.zeros(n + 1, dtype=PAIR)
        pairs["key"][:n], pairs["idx"][:n] = keys, np.arange(n, dtype=np.uint32)
        perm = np.zeros(n + 1, np.uint32)
        seg_key, seg_off, seg_name = np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32)
        n_seg = emu.emu_ordering(pairs.ctypes.data, n, surv.ctypes.data, 16 if mode == 1 else 32, 11, perm.ctypes.data,
                                 seg_key.ctypes.data, seg_off.ctypes.data, seg_name.ctypes.data, form, table.ctypes.data, mode)
        order = np.argsort(keys, kind="stable")
        uniq = np.unique(keys)
        assert n_seg == len(uniq)
        assert np.array_equal(perm[:n], order.astype(np.uint32))
        assert np.array_equal(surv[:, 3], table[dev]), (n, mode, form)        # every record joined
        if mode == 1:
            assert np.array_equal(seg_name[:n_seg], table[uniq])

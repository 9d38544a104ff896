"""The radix family of csrc/kvg_scan.cuh (k_radix_hist, k_radix_tilescan, k_radix_scatter<8|11>) and the
experimental k_radix_tilescan_warp (csrc/kvg_radix_exp.cuh), executed on the CPU from their real kernel
source under the warp emulator of tools/emu/ and checked against a stable numpy sort: device-chosen
digit widths, multi-tile inputs, ragged last tiles, duplicate-heavy and full-width keys."""
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
    (1, 1, 32, 11), (31, 5, 16, 11), (2047, 9, 16, 11), (2048, 10, 32, 11), (2049, 11, 32, 11),
    (3000, 12, 32, 11), (5000, 19, 32, 11), (4100, 22, 32, 11), (2500, 23, 32, 11), (2200, 32, 32, 11),
    (2300, 16, 16, 11), (4500, 19, 32, 8), (2100, 8, 16, 8), (2600, 24, 32, 8),
]


@pytest.mark.parametrize("variant", [0, 1], ids=["tilescan", "tilescan_warp"])
def test_device_radix_sort_is_a_stable_sort(emu, variant):
    rng = np.random.default_rng(17 + variant)
    for n, bits, kmax, mb in CASES:
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
        keys[rng.integers(0, n)] = (1 << bits) - 1          # the widest key is present: the plan sees it
        if n > 100:
            keys[rng.integers(0, n, n // 3)] = keys[0]       # a heavy bucket: long runs of equal keys
        got, npass = device_sort(emu, keys, kmax, mb, variant)
        want_k, want_i = expected(keys)
        assert npass == -(-min(bits, kmax) // mb), (n, bits, mb)
        assert np.array_equal(got["key"], want_k) and np.array_equal(got["idx"], want_i), (n, bits, kmax, mb)


def test_both_tile_scans_agree_on_skewed_input(emu):
    rng = np.random.default_rng(5)
    keys = np.concatenate([np.full(3000, 7, np.uint32), rng.integers(0, 1 << 19, 1500, dtype=np.uint64).astype(np.uint32),
                           np.zeros(700, np.uint32)])
    rng.shuffle(keys)
    a, _ = device_sort(emu, keys, 32, 11, 0)
    b, _ = device_sort(emu, keys, 32, 11, 1)
    assert np.array_equal(a, b)
    want_k, want_i = expected(keys)
    assert np.array_equal(a["key"], want_k) and np.array_equal(a["idx"], want_i)

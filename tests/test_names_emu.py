"""getDeviceName (device_plugin.go:371-438) end to end from KERNEL SOURCE on the CPU: the barrier-free
parse, k_nv_index, k_pciids_sanitise_lines (the lane-parallel name transform with its Unicode fall-back),
k_section_lines / k_lookup_general / k_sanitise_matches (prefix semantics for arbitrary keys) — sequenced
like libkvgpu.so does and compared with the oracle on the reference's Ginkgo fixture, the shipped
pci.ids and the grammar fuzz.  Runs under the warp emulator of tools/emu/."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import conftest  # noqa: F401
import util
from oracle import oracle as O

sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "emu"))
import build as emu_build  # noqa: E402
from test_gpu_parity import _random_pciids  # noqa: E402
from test_parse_v2_emu import pad  # noqa: E402

NAME_CAP = 4096


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_names())
    L.emu_get_device_names.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                       C.c_uint32, C.c_void_p]
    return L


def device_names(emu, text, keys, cap_log2=12):
    buf = pad(text)
    blob = b"".join(keys) + b"\0"
    off = np.zeros(len(keys) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    kb = np.frombuffer(blob, dtype=np.uint8).copy()
    out = np.zeros(len(keys) * NAME_CAP, dtype=np.uint8)
    ln = np.zeros(len(keys), dtype=np.uint32)
    rc = emu.emu_get_device_names(buf.ctypes.data, len(text), cap_log2, kb.ctypes.data, off.ctypes.data, len(keys),
                                  out.ctypes.data, NAME_CAP, ln.ctypes.data)
    assert rc == 0
    return [bytes(out[i * NAME_CAP:i * NAME_CAP + int(ln[i])]).decode("latin-1") for i in range(len(keys))]


def check(emu, text, keys, cap_log2=12):
    if not text:
        return
    keys = [k if isinstance(k, bytes) else k.encode("utf-8") for k in keys]
    got = device_names(emu, text, keys, cap_log2)
    for k, g in zip(keys, got):
        assert g == O.get_device_name(text, k), (k, text[:100])


def test_ginkgo_kats_from_kernel_source(emu):
    G = util.ginkgo()["get_device_name"]
    text = G["fixture"].encode()
    for kat in G["kats"]:
        if kat["missing_file"]:
            continue
        assert device_names(emu, text, [kat["key"].encode()])[0] == kat["want"], kat["title"]


def test_shipped_pciids_names_from_kernel_source(emu):
    text = util.pciids_text()
    names = util.pciids_names()["names"]
    keys = sorted(names)[::9] + ["1b38", "2901", "2330", "05be", "ffff", "0000"]
    got = device_names(emu, text, [k.encode() for k in keys], cap_log2=15)
    for k, g in zip(keys, got):
        assert g == names.get(k, O.get_device_name(text, k.encode())), k
    general = ["", "1", "1b", "1b3", "1b38 ", "1b38  GP102GL", "\t1043", "1B38", "2901  ", "x", "10de",
               "0008  NV1 [STG2000X-B Series]", "0008  NV1 [STG2000X-B Series]x", "#", "\n", "1b38\n", "ffffff", "é", "1b3\r"]
    check(emu, text, general, cap_log2=15)


def test_grammar_fuzz_names_from_kernel_source(emu):
    rng = np.random.default_rng(20250711)
    keys = ["%04x" % i for i in range(0, 40, 3)] + ["", "0", "00", "000", "0001 ", "\t", "001\r", "0001\r"]
    for it in range(25):
        check(emu, _random_pciids(rng, int(rng.integers(1, 300))), keys)

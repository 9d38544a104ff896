"""getDeviceName (device_plugin.go:371-438) end to end from KERNEL SOURCE on the CPU: K1 (per-warp TMA
ring, span summaries, resolve + finalize), k_pciids_names (the lane-parallel name transform with its Unicode fall-back),
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
from test_parse_k1_emu import pad  # noqa: E402

NAME_CAP = 4096


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_names())
    L.emu_get_device_names.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                       C.c_void_p]
    return L


PARSERS = pytest.mark.parametrize("parser", [3], ids=["K1"])


def device_names(emu, text, keys, cap_log2=12, parser=1, want_info=False):
    buf = pad(text)
    blob = b"".join(keys) + b"\0"
    off = np.zeros(len(keys) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    kb = np.frombuffer(blob, dtype=np.uint8).copy()
    out = np.zeros(len(keys) * NAME_CAP, dtype=np.uint8)
    ln = np.zeros(len(keys), dtype=np.uint32)
    info = np.zeros(8, dtype=np.uint32)
    rc = emu.emu_get_device_names(buf.ctypes.data, len(text), kb.ctypes.data, off.ctypes.data, len(keys),
                                  out.ctypes.data, NAME_CAP, ln.ctypes.data, info.ctypes.data, None, None, 0, None)
    assert rc == 0
    names = [bytes(out[i * NAME_CAP:i * NAME_CAP + int(ln[i])]).decode("latin-1") for i in range(len(keys))]
    return (names, info) if want_info else names


def check(emu, text, keys, cap_log2=12, parser=1):
    if not text:
        return
    keys = [k if isinstance(k, bytes) else k.encode("utf-8") for k in keys]
    got = device_names(emu, text, keys, cap_log2, parser)
    for k, g in zip(keys, got):
        assert g == O.get_device_name(text, k), (k, text[:100])


@PARSERS
def test_ginkgo_kats_from_kernel_source(emu, parser):
    G = util.ginkgo()["get_device_name"]
    text = G["fixture"].encode()
    for kat in G["kats"]:
        if kat["missing_file"]:
            continue
        assert device_names(emu, text, [kat["key"].encode()], parser=parser)[0] == kat["want"], kat["title"]


@PARSERS
def test_shipped_pciids_names_from_kernel_source(emu, parser):
    text = util.pciids_text()
    names = util.pciids_names()["names"]
    keys = sorted(names)[::9] + ["1b38", "2901", "2330", "05be", "ffff", "0000"]
    got, info = device_names(emu, text, [k.encode() for k in keys], cap_log2=15, parser=parser, want_info=True)
    assert int(info[0]) == text.index(b"\n10de  NVIDIA") + 1 and int(info[2]) == 1931 and int(info[3]) == text.count(b"\n")
    for k, g in zip(keys, got):
        assert g == names.get(k, O.get_device_name(text, k.encode())), k
    general = ["", "1", "1b", "1b3", "1b38 ", "1b38  GP102GL", "\t1043", "1B38", "2901  ", "x", "10de",
               "0008  NV1 [STG2000X-B Series]", "0008  NV1 [STG2000X-B Series]x", "#", "\n", "1b38\n", "ffffff", "é", "1b3\r"]
    check(emu, text, general, cap_log2=15, parser=parser)


@PARSERS
def test_grammar_fuzz_names_from_kernel_source(emu, parser):
    rng = np.random.default_rng(20250711)
    keys = ["%04x" % i for i in range(0, 40, 3)] + ["", "0", "00", "000", "0001 ", "\t", "001\r", "0001\r"]
    for it in range(25):
        check(emu, _random_pciids(rng, int(rng.integers(1, 300))), keys, parser=parser)

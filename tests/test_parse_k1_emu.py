"""csrc/kvg_parse_k1.cuh — K1, the pci.ids parse (scan with the per-warp TMA ring -> resolve + finalize ->
names, self-cleaning: no clearing launch) — executed on the CPU from its REAL kernel source under the warp emulator (tools/emu/) and
compared with the oracle and with the Python model of the span decomposition (tools/span_model.py:
same spans, same ownership rule, same section / scanner-limit arithmetic).  Output buffers are
poisoned where the kernels write them whole, and the persistent scan kernel runs
with several grid sizes so that the two-stage ring is exercised over many iterations per warp."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import conftest  # noqa: F401
import util
from oracle import oracle as O

sys.path.insert(0, os.path.join(conftest.ROOT, "tools"))
sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "emu"))
import build as emu_build  # noqa: E402
import span_model as M  # noqa: E402
from test_gpu_parity import _random_pciids  # noqa: E402

NONE = 0xFFFFFFFF


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(emu_build.build_names())
    L.emu_parse_k1.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    return L


def pad(text: bytes) -> np.ndarray:
    """kvg_text_pad: '\\n' up to a multiple of the 8 KiB tile, plus the 16-byte halo."""
    n = (len(text) + 8191) // 8192 * 8192 + 16
    buf = np.full(max(n, 16), 0x0A, dtype=np.uint8)
    buf[:len(text)] = np.frombuffer(text, dtype=np.uint8)
    return buf


def run(emu, text: bytes, n_files: int = 1, cap_log2: int = 0, scan_ctas: int = 0):
    buf = pad(text)
    stride = len(buf)
    images = np.tile(buf, n_files)
    info = np.zeros((n_files, 8), dtype=np.uint32)
    tables = np.zeros((n_files, 65536), dtype=np.uint32)
    assert emu.emu_parse_k1(images.ctypes.data, stride, len(text), n_files, scan_ctas, info.ctypes.data,
                            tables.ctypes.data) == 0
    return info, tables


def emu_name(emu, text, info, table, cap_log2, key):
    """What kvg_name_lookup's table path returns: the id's slot, the section check of k_pciids_names, the line."""
    off = int(table[int(key, 16)])
    v_off, sec_end = int(info[0]), int(info[1])
    if off == NONE or v_off == NONE or not (v_off < off < sec_end):
        return ""
    end = text.find(b"\n", off)
    line = text[off:end if end >= 0 else len(text)]
    return O.get_device_name(b"10de\n" + line + b"\n", key.encode())


def check(emu, text, keys, cap_log2=15, scan_ctas=0):
    if len(text) == 0:
        return            # the product never launches the parse on an empty file (kvg_pciids_load, len == 0)
    info, tables = run(emu, text, 1, cap_log2, scan_ctas)
    model = M.parse(text)
    v_off, sec_end, n_entries, n_lines, limit, overflow = (int(x) for x in info[0][:6])
    assert overflow == 0
    assert (v_off, sec_end, limit, n_lines) == (model["v_off"], model["sec_end"], model["limit"], model["n_lines"])
    for dev, off in model["table"].items():       # every (device -> first line) pair, straight from the table
        assert int(tables[0][dev]) == off
    assert int((tables[0] != NONE).sum()) == len(model["table"])   # nothing else was recorded
    for k in keys:
        assert emu_name(emu, text, info[0], tables[0], cap_log2, k) == O.get_device_name(text, k.encode()), k


def test_k1_on_shipped_pciids(emu):
    text = util.pciids_text()
    names = util.pciids_names()["names"]
    rng = np.random.default_rng(5)
    keys = list(names)[::7] + ["%04x" % int(k) for k in rng.integers(0, 65536, 60)] + ["2330", "2901", "1b38", "ffff"]
    check(emu, text, keys, cap_log2=13)
    check(emu, text, keys[:40], cap_log2=13, scan_ctas=7)     # 28 warps x ~14 spans each: the ring wraps many times


def test_k1_two_images_are_independent(emu):
    text = util.pciids_text()[:200_000] + b"10de  NVIDIA tail\n\t2901  GB100 [B200]\n"
    info, tables = run(emu, text, 2, scan_ctas=5)
    assert (info[0] == info[1]).all() and (tables[0] == tables[1]).all()
    assert int((tables[1] != NONE).sum()) == len(M.parse(text)["table"])


def test_k1_on_grammar_fuzz(emu):
    rng = np.random.default_rng(20250711)
    keys = ["%04x" % i for i in range(0, 40)]
    for it in range(40):
        check(emu, _random_pciids(rng, int(rng.integers(1, 400))), keys, cap_log2=12, scan_ctas=it % 3)
    for it in range(3):
        check(emu, _random_pciids(rng, int(rng.integers(3000, 6000))), keys, cap_log2=14, scan_ctas=1 + it)


def test_k1_span_boundaries_and_scanner_limit(emu):
    S = M.SPAN
    base = b"8086  Intel\n\t1234  wrong vendor\n"
    for delta in list(range(-8, 9)) + [S - 8, S, S + 5]:
        pad_len = S - len(base) + delta - 2
        text = base + b"#" + b"c" * pad_len + b"\n" + b"10de  NVIDIA\n\t1234  Edge [case]\n" + \
            b"#" + b"d" * (S - 40) + b"\n\t5678  second tile\n10df  next\n\t9999  other\n"
        check(emu, text, ("1234", "5678", "9999", "abcd"), cap_log2=10, scan_ctas=delta % 2)
    many = b"10de  NVIDIA\n" + b"".join(b"\t%04x  dev %d\n" % (i, i) for i in range(0, 9000)) + b"1000 x\n\t0001  y\n"
    check(emu, many, ("0000", "0100", "1fff", "2327", "2328", "0001"), cap_log2=15, scan_ctas=2)
    tail = b"10de  NVIDIA\n\t1234  name\n"
    for n in (65535, 65536, 70000):
        check(emu, b"x" * n + b"\n" + tail, ("1234",), cap_log2=10)
    check(emu, b"10de\n\t1234  name\n\t" + b"y" * 65536, ("1234",), cap_log2=10)
    check(emu, b"10de\n\t" + b"y" * 65536 + b"\n\t1234  name\n", ("1234",), cap_log2=10)
    for text in (b"\n", b"10de", b"10de\n", b"\t1234  orphan\n10de\n", b"10de\r\n\t1234  crlf\r\n",
                 b"10de  a\n\t1234  first\n\t1234  second\n", b"10de\n\n\t1234  after blank\n",
                 b"10de\n# c\n\t1234  after comment\n", b"10de  x\n10de  dup\n\t1234  under dup\n",
                 b"\t1234  line zero is a device line\n", b"10de  no newline at all\t1234"):
        check(emu, text, ("1234", "0000"), cap_log2=10)
    # the densest span: every line a 10de device line in front of any header of its span
    dense = b"10de\n" + b"".join(b"\t%04x\n" % (i & 0xffff) for i in range(3000))
    check(emu, dense, ("0000", "0abc"), cap_log2=13)
    # a 10de header whose device lines start exactly on a span boundary, and one ending a span
    for k in (S - 13, S - 12, S - 1, S, S + 1):
        text = b"#" + b"z" * (k - 2) + b"\n" + b"10de  NVIDIA\n\t1111  first\n" + b"#" + b"q" * 5000 + b"\n\t2222  resolved\n"
        check(emu, text, ("1111", "2222"), cap_log2=10)

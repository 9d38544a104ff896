"""The decomposition of the barrier-free parse K1 (tools/span_model.py: what a span decides alone, what it
defers, how the deferred part resolves, section bounds from span summaries) against the oracle: the shipped
pci.ids, the grammar fuzz and the span-boundary / scanner-limit cases the GPU parity tests use.  CPU-only.  The
model is what tests/test_parse_k1_emu.py compares the kernel's device-id table with, so it is pinned here."""
import os
import sys

import numpy as np

import conftest  # noqa: F401
import util
from oracle import oracle as O

sys.path.insert(0, os.path.join(conftest.ROOT, "tools"))
import span_model as M  # noqa: E402

from test_gpu_parity import _random_pciids  # noqa: E402  (a pure-Python generator)


def model_name(text, parsed, key):
    line = M.lookup_line(text, parsed, key)
    if line is None:
        return ""
    # the name depends only on the matched line: let the oracle transform it in a minimal file
    return O.get_device_name(b"10de\n" + line + b"\n", key.encode())


def check(text, keys):
    parsed = M.parse(text)
    assert parsed["n_lines"] == text.count(b"\n")
    for k in keys:
        assert model_name(text, parsed, k) == O.get_device_name(text, k.encode()), (k, text[:120])
    return parsed


def test_model_on_shipped_pciids():
    text = util.pciids_text()
    names = util.pciids_names()["names"]
    rng = np.random.default_rng(3)
    keys = list(names) + ["%04x" % int(k) for k in rng.integers(0, 65536, 200)] + ["2330", "ffff", "0000", "10de"]
    parsed = check(text, keys)
    assert parsed["v_off"] == text.index(b"\n10de  NVIDIA") + 1
    assert len(parsed["table"]) == 1931          # only lines under vendor 10de reach the table


def test_model_on_grammar_fuzz():
    rng = np.random.default_rng(20250711)
    keys = ["%04x" % i for i in range(0, 40)]
    for it in range(120):
        check(_random_pciids(rng, int(rng.integers(1, 400))), keys)
    for it in range(6):                            # many spans: context carried across span edges
        check(_random_pciids(rng, int(rng.integers(3000, 9000))), keys)


def test_model_span_boundaries_and_scanner_limit():
    S = M.SPAN
    base = b"8086  Intel\n\t1234  wrong vendor\n"
    for delta in list(range(-8, 9)) + [S - 8, S, S + 5]:
        pad_len = S - len(base) + delta - 2
        text = base + b"#" + b"c" * pad_len + b"\n" + b"10de  NVIDIA\n\t1234  Edge [case]\n" + \
            b"#" + b"d" * (S - 40) + b"\n\t5678  second tile\n10df  next\n\t9999  other\n"
        check(text, ("1234", "5678", "9999", "abcd"))
    many = b"10de  NVIDIA\n" + b"".join(b"\t%04x  dev %d\n" % (i, i) for i in range(0, 9000)) + b"1000 x\n\t0001  y\n"
    check(many, ("0000", "0100", "1fff", "2327", "2328", "0001"))
    tail = b"10de  NVIDIA\n\t1234  name\n"
    for n in (65535, 65536, 70000):
        check(b"x" * n + b"\n" + tail, ("1234",))
    check(b"10de\n\t1234  name\n\t" + b"y" * 65536, ("1234",))
    check(b"10de\n\t" + b"y" * 65536 + b"\n\t1234  name\n", ("1234",))
    for text in (b"", b"\n", b"10de", b"10de\n", b"\t1234  orphan\n10de\n", b"10de\r\n\t1234  crlf\r\n",
                 b"10de  a\n\t1234  first\n\t1234  second\n", b"10de\n\n\t1234  after blank\n",
                 b"10de\n# c\n\t1234  after comment\n", b"10de  x\n10de  dup\n\t1234  under dup\n"):
        check(text, ("1234", "0000"))

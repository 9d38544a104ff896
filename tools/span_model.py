#!/usr/bin/env python3
"""Executable model of the barrier-free pci.ids parse K1 (k_pciids_scan + k_pciids_resolve_finalize).

The CUDA kernels of csrc/kvg_parse_k1.cuh follow this decomposition step by step; the model exists
so that the DECOMPOSITION (what a 4 KiB span decides alone, what it defers, how the deferred part is
resolved, how the section bounds come out of the span summaries) can be checked against the oracle
on a CPU-only box (tests/test_span_model.py), and so that the emulated kernel's device-id table has
something to be compared with.  It is test/tool code, never on the product path.

  span        4096 text bytes owned by one warp; it owns the lines that START in (a, a + 4096]
              (line start = newline position + 1; the line at offset 0 belongs to span 0)
  scan        per span, independent of every other span:
                n_newlines, first / last newline offset, first header-type line offset,
                v_off candidates (header lines whose first four bytes are "10de"),
                state    = last header of the span: None | (valid, vendor)
                direct   = device lines ("\\t" + 4 lower hex) AFTER a header of the span whose vendor is 10de
                pending  = device lines BEFORE the span's first header (their vendor context is
                           whatever the previous spans say)
  resolve     per span with pending lines: walk back to the nearest span that has a header; if that
              header is a valid 10de, the pending lines are inserted
  finalize    exactly k_pciids_finalize: scanner limit from the newline summaries, section end from
              the first-header summaries (reference: device_plugin.go:371-438, bufio.Scanner)
"""
from __future__ import annotations

SPAN = 4096
SCAN_TOKEN_MAX = 65536
NONE = 0xFFFFFFFF
HEX = b"0123456789abcdef"


def parse_hex4(b: bytes):
    """(valid, value) of the first four bytes as LOWER-case hex."""
    if len(b) < 4 or any(c not in HEX for c in b[:4]):
        return False, 0
    return True, int(b[:4], 16)


def scan_span(text: bytes, s: int):
    n = len(text)
    a = s * SPAN
    out = dict(n_nl=0, first_nl=NONE, last_nl=0, first_hdr=NONE, v_off=NONE, state=None, direct=[], pending=[])
    starts = []
    if s == 0 and n > 0:
        starts.append(0)
    for p in range(a, min(a + SPAN, n)):
        if text[p] == 0x0A:
            out["n_nl"] += 1
            if out["first_nl"] == NONE:
                out["first_nl"] = p
            out["last_nl"] = p
            if p != n - 1:                       # a newline that is the last byte starts no line
                starts.append(p + 1)
    ctx = None                                    # None: no header seen in this span yet
    for p in starts:
        b0 = text[p:p + 1]
        if b0 == b"\t":
            ok, dev = parse_hex4(text[p + 1:p + 5])
            if not ok:
                continue
            if ctx is None:
                out["pending"].append((p, dev))
            elif ctx == (True, 0x10de):
                out["direct"].append((p, dev))
        elif b0 != b"#":                          # header-type line (also an empty line)
            ok, ven = parse_hex4(text[p:p + 4])
            ctx = (ok, ven)
            if out["first_hdr"] == NONE:
                out["first_hdr"] = p
            if ok and ven == 0x10de and out["v_off"] == NONE:
                out["v_off"] = p
    out["state"] = ctx
    return out


def parse(text: bytes):
    """-> dict(table={dev: first line offset}, v_off, sec_end, limit, n_lines)"""
    n = len(text)
    n_spans = (n + SPAN - 1) // SPAN
    spans = [scan_span(text, s) for s in range(n_spans)]
    table = {}

    def insert(p, dev):
        if dev not in table or p < table[dev]:   # first line wins (atomicMin on the offset)
            table[dev] = p

    for sp in spans:
        for p, dev in sp["direct"]:
            insert(p, dev)
    for s, sp in enumerate(spans):                # resolve
        if not sp["pending"]:
            continue
        ctx = None
        for u in range(s - 1, -1, -1):
            if spans[u]["state"] is not None:
                ctx = spans[u]["state"]
                break
        if ctx == (True, 0x10de):
            for p, dev in sp["pending"]:
                insert(p, dev)
    # ---- finalize (k_pciids_finalize)
    v_off = min([sp["v_off"] for sp in spans] + [NONE])
    limit = n
    for t in range(n_spans + 1):
        if t == n_spans:
            fn = n
        else:
            fn = spans[t]["first_nl"]
            if fn == NONE:
                continue
        line_start = 0
        for u in range(t - 1, -1, -1):
            if spans[u]["first_nl"] != NONE:
                line_start = spans[u]["last_nl"] + 1
                break
        if line_start < n and fn - line_start >= SCAN_TOKEN_MAX:
            limit = min(limit, line_start)
    if v_off == NONE or v_off >= limit:
        return dict(table=table, v_off=NONE, sec_end=NONE, limit=limit, n_lines=sum(sp["n_nl"] for sp in spans))
    tv = 0 if v_off == 0 else (v_off - 1) // SPAN
    end = n
    tile_end = min(n, (tv + 1) * SPAN + 1)
    for p in range(v_off + 1, tile_end):
        if text[p - 1] == 0x0A and text[p:p + 1] not in (b"\t", b"#"):
            end = min(end, p)
    if end == n:
        for t in range(tv + 1, n_spans):
            if spans[t]["first_hdr"] != NONE:
                end = spans[t]["first_hdr"]
                break
    return dict(table=table, v_off=v_off, sec_end=min(end, limit), limit=limit,
                n_lines=sum(sp["n_nl"] for sp in spans))


def lookup_line(text: bytes, parsed, key: str):
    """The line a 4-lower-hex key resolves to (bytes without the newline), or None."""
    if len(key) != 4 or any(c not in "0123456789abcdef" for c in key):
        raise ValueError("the hash path only serves canonical keys")
    off = parsed["table"].get(int(key, 16))
    if off is None or parsed["v_off"] == NONE or not (parsed["v_off"] < off < parsed["sec_end"]):
        return None
    end = text.find(b"\n", off)
    return text[off:end if end >= 0 else len(text)]

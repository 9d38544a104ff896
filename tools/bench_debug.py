#!/usr/bin/env python3
"""Decomposed classify timings + a plain device copy of the same bytes, to locate the ceiling."""
import ctypes as C, gzip, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
import numpy as np, torch
import kvgpu
from oracle import oracle as O
text = gzip.open(os.path.join(ROOT, "tests", "golden", "pci.ids.gz"), "rb").read()
ids = O.nv_ids(text)
ctx = kvgpu.Context(0); ctx.pciids_load(text); lib = kvgpu.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda"); dst = torch.empty_like(buf); torch.cuda.synchronize()
ctx.dev_gen_pci(buf.data_ptr(), 0, n, ids, 23); ctx.dev_scan_pci(buf.data_ptr(), n); S = ctx.dev_scan_pci_count()[0]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): dst.copy_(buf)
torch.cuda.synchronize(); e0.record()
for _ in range(5): dst.copy_(buf)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
print("torch copy %d MB: %.1f us = %.0f GB/s (r+w)" % (n * 16 >> 20, t * 1e3, 2 * n * 16 / t / 1e6))
for rows in (4, 8, 16):
    for mode in (0, 1, 2):
        ms = C.c_float(); best = 1e9
        for _ in range(4):
            rc = lib.kvg_dev_debug_classify(ctx.handle, buf.data_ptr(), n, mode, rows, C.byref(ms)); assert rc == 0
            best = min(best, ms.value)
        nbytes = 16 * n + (16 * S if mode else 0)
        print("rows=%2d mode=%d: %.1f us = %.0f GB/s" % (rows, mode, best * 1e3, nbytes / best / 1e6))

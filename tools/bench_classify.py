#!/usr/bin/env python3
"""Kernel-only timing of the classify/compact kernel and the whole scan at a given size."""
import gzip, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
import numpy as np, torch
import kvgpu
from oracle import oracle as O
text = gzip.open(os.path.join(ROOT, "tests", "golden", "pci.ids.gz"), "rb").read()
ids = O.nv_ids(text)
ctx = kvgpu.Context(0); ctx.pciids_load(text)
for n in [int(a) for a in sys.argv[1:]] or [1 << 20, 1 << 24]:
    buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
    ctx.dev_gen_pci(buf.data_ptr(), 0, n, ids, max(1, int(np.ceil(np.log2(n // 2)))))
    for _ in range(3): ctx.dev_scan_pci(buf.data_ptr(), n)
    acc = {}
    for _ in range(5):
        ctx.dev_flush_l2(); ctx.set_kernel_timing(True); ctx.dev_scan_pci(buf.data_ptr(), n)
        for k, v in ctx.kernel_times(): acc.setdefault(k, []).append(v)
    ctx.set_kernel_timing(False)
    S = ctx.dev_scan_pci_count()[0]
    reps = 5
    per = {k: sum(v) / reps for k, v in acc.items()}
    c = per["classify_compact"]
    print("%s n=%d S=%d classify %.1f us = %.0f GB/s (%.1f%% of 6567)  whole scan %.1f us  %s" % (
        os.environ.get("KVG_CLASSIFY", "default"), n, S, c * 1e3, (16 * n + 16 * S) / c / 1e6,
        100 * (16 * n + 16 * S) / c / 1e6 / 6567.4, sum(per.values()) * 1e3,
        {k: round(v * 1e3, 1) for k, v in per.items()}))
    del buf

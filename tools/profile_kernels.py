#!/usr/bin/env python3
"""Short driver for ncu: the two headline kernels on inputs larger than L2.
   ncu --set full -k regex:'k_compact|k_pciids_parse' ... python tools/profile_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
import gzip

import numpy as np
import torch

import kvgpu
from oracle import oracle as O

text = gzip.open(os.path.join(ROOT, "tests", "golden", "pci.ids.gz"), "rb").read()
ids = O.nv_ids(text)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = kvgpu.Context(0)
ctx.pciids_load(text)
buf = torch.empty(n * 16, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.dev_gen_pci(buf.data_ptr(), 0, n, ids, 23)
for _ in range(reps):
    ctx.dev_scan_pci(buf.data_ptr(), n)
print("scan", ctx.dev_scan_pci_count())
pad = ctx.text_pad(len(text))
stride = pad + 16
host = np.full(stride * nf, 10, dtype=np.uint8)
for f in range(nf):
    host[f * stride:f * stride + len(text)] = np.frombuffer(text, dtype=np.uint8)
dev = torch.from_numpy(host).cuda()
torch.cuda.synchronize()
c2 = kvgpu.Context(0)
for _ in range(reps):
    c2.dev_pciids_parse(dev.data_ptr(), len(text), stride, nf)
print("parse", c2.pciids_info())

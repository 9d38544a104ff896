"""Host-clock phase trace of the e2e step (kvg_pciids_load + kvg_scan_pci with pinned buffers).
KVG_TRACE=1 python tools/trace_e2e.py [n_records]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kubevirt-gpu-device-plugin_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import kvgpu
import util
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
text = util.pciids_text()
ids = O.nv_ids(text)
ctx = kvgpu.Context(0)
lib = kvgpu.load()
p_text = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).pin_memory()
p_recs = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
p_recs.numpy()[:] = np.frombuffer(O.gen_pci(0, n, ids, 20).tobytes(), dtype=np.uint8)
for i in range(8):
    t0 = time.perf_counter()
    rc = lib.kvg_pciids_load(ctx.handle, p_text.data_ptr(), len(text))
    t1 = time.perf_counter()
    res = C.POINTER(kvgpu._lib.PciResultC)()
    rc2 = lib.kvg_scan_pci(ctx.handle, p_recs.data_ptr(), n, C.byref(res))
    t2 = time.perf_counter()
    lib.kvg_result_free(res)
    t3 = time.perf_counter()
    assert rc == 0 and rc2 == 0
    print("step %d: load %.0f us  scan %.0f us  free %.0f us" % (i, (t1 - t0) * 1e6, (t2 - t1) * 1e6, (t3 - t2) * 1e6),
          file=sys.stderr, flush=True)
ctx.close()

"""createIommuDeviceMap (device_plugin.go:187-247) END TO END FROM KERNEL SOURCE on the CPU: the parse and
name table (tools/emu names), the classification pipeline, both stable orderings with their segment
heads — composed like kvg_dev_scan_pci composes them — rebuilt into the five maps by the product's own
host code (kvgpu.pci_maps_from_result) and compared with the oracle through the SAME canonical dump the
GPU parity tests use."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import conftest  # noqa: F401
import kvgpu
import util
from oracle import oracle as O

sys.path.insert(0, os.path.join(conftest.ROOT, "tools", "emu"))
import build as emu_build  # noqa: E402
from test_parse_k1_emu import pad  # noqa: E402

PAIR = np.dtype([("key", "<u4"), ("idx", "<u4")])


@pytest.fixture(scope="module")
def libs():
    names = C.CDLL(emu_build.build_names())
    names.emu_get_device_names.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_void_p]
    cls = C.CDLL(emu_build.build_classify())
    cls.emu_classify_pci.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rdx = C.CDLL(emu_build.build_radix())
    rdx.emu_ordering.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return names, cls, rdx


def name_table(names, text, parser):
    buf = pad(text)
    nv_index = np.zeros(65536, dtype=np.uint32)
    pool = np.zeros(len(text) + 64, dtype=np.uint8)
    pool_len, info = C.c_uint32(), np.zeros(8, dtype=np.uint32)
    key = np.zeros(1, dtype=np.uint8)
    off = np.zeros(1, dtype=np.uint32)
    rc = names.emu_get_device_names(buf.ctypes.data, len(text), key.ctypes.data, off.ctypes.data, 0, None, 64, None,
                                    info.ctypes.data, nv_index.ctypes.data, pool.ctypes.data, len(pool), C.byref(pool_len))
    assert rc == 0
    return nv_index, bytes(pool[:pool_len.value])


def ordering(rdx, surv, field, key_bits):
    n = len(surv)
    pairs = np.zeros(n + 1, dtype=PAIR)
    pairs["key"][:n], pairs["idx"][:n] = surv[field], np.arange(n, dtype=np.uint32)
    perm = np.zeros(n + 1, np.uint32)
    seg_key, seg_off, seg_name = np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32), np.zeros(n + 2, np.uint32)
    raw = np.ascontiguousarray(surv).view(np.uint32).reshape(-1, 4) if n else np.zeros((1, 4), np.uint32)
    k = rdx.emu_ordering(pairs.ctypes.data, n, raw.ctypes.data, key_bits, 11, perm.ctypes.data, seg_key.ctypes.data,
                         seg_off.ctypes.data, seg_name.ctypes.data, 1, None, 0)
    assert k >= 0
    return seg_key[:k].copy(), seg_off[:k + 1].copy(), perm[:n].copy(), seg_name[:k].copy()


@pytest.mark.parametrize("parser,variant", [(3, 1), (3, 0)], ids=["K1+oneshot", "K1+ragged"])
def test_create_iommu_device_map_from_kernel_source(libs, parser, variant):
    names, cls, rdx = libs
    text = util.pciids_text()
    nv_index, pool = name_table(names, text, parser)
    ids = O.nv_ids(text)
    for n, gbits in ((0, 0), (700, 0), (2600, 11)):
        recs = O.gen_pci(3, n, ids, gbits)
        surv = np.zeros(n + 1, dtype=kvgpu.PCI_SURV)
        ctrl = np.zeros(3, dtype=np.uint32)
        buf = np.ascontiguousarray(recs) if n else np.zeros(1, dtype=kvgpu.PCI_REC)
        assert cls.emu_classify_pci(buf.ctypes.data, n, nv_index.ctypes.data, variant, surv.ctypes.data, ctrl.ctypes.data) == 0
        surv = surv[:int(ctrl[0])].copy()
        dk, doff, dperm, dname = ordering(rdx, surv, "device", 16)
        gk, goff, gperm, _ = ordering(rdx, surv, "iommu_group", 32)
        res = kvgpu.PciResult(n_records=n, survivors=surv, dev_keys=dk.astype(np.uint16), dev_off=doff, dev_perm=dperm,
                              dev_name_slot=dname, grp_keys=gk, grp_off=goff, grp_perm=gperm, name_pool=pool)
        got = kvgpu.canonical_dump(kvgpu.pci_maps_from_result(res))
        m = O.Maps()
        m.create_iommu_device_map_flat(recs)
        assert got == m.dump(text), (n, gbits)

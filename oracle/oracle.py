"""ctypes loader for the CPU ORACLE (oracle/kvg_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkvgoracle.so")

# wire-format dtypes (include/kvgpu.h)
PCI_REC = np.dtype(
    [("addr", "<u4"), ("vendor", "<u2"), ("device", "<u2"), ("iommu_group", "<u4"),
     ("driver", "u1"), ("flags", "u1"), ("numa", "<i2")], align=False)
PCI_SURV = np.dtype(
    [("addr", "<u4"), ("iommu_group", "<u4"), ("device", "<u2"), ("numa", "<u2"),
     ("name_slot", "<u4")], align=False)
MDEV_REC = np.dtype(
    [("uuid", "u1", (16,)), ("parent", "<u4"), ("type_idx", "<u2"), ("flags", "u1"),
     ("pad0", "u1"), ("parent_numa", "<i2"), ("pad1", "u1", (6,))], align=False)
MDEV_SURV = np.dtype(
    [("uuid", "u1", (16,)), ("parent", "<u4"), ("type_key", "<u2"), ("numa", "<u2"),
     ("src", "<u4"), ("pad", "<u4")], align=False)
assert PCI_REC.itemsize == 16 and PCI_SURV.itemsize == 16
assert MDEV_REC.itemsize == 32 and MDEV_SURV.itemsize == 32


class TypeDict(C.Structure):
    _fields_ = [("n_types", C.c_uint32), ("off", C.POINTER(C.c_uint32)),
                ("bytes", C.POINTER(C.c_uint8))]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "kvg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "kvg_oracle.h")),
            os.path.getmtime(os.path.join(_HERE, "..", "include", "kvgpu.h"))):
        subprocess.run(["make", "-C", _HERE, "-B", "libkvgoracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, sz = C.c_void_p, C.c_size_t
        L.kvo_get_device_name.restype = C.c_long
        L.kvo_get_device_name.argtypes = [u8p, sz, u8p, sz, u8p, sz]
        L.kvo_get_device_name_file.restype = C.c_long
        L.kvo_get_device_name_file.argtypes = [C.c_char_p, u8p, sz, u8p, sz]
        L.kvo_nv_ids.restype = sz
        L.kvo_nv_ids.argtypes = [u8p, sz, u8p, sz]
        L.kvo_trim_space.restype = None
        L.kvo_trim_space.argtypes = [u8p, sz, C.POINTER(sz), C.POINTER(sz)]
        L.kvo_to_upper.restype = sz
        L.kvo_to_upper.argtypes = [u8p, sz, u8p]
        for fn in ("kvo_read_id_from_file", "kvo_read_link", "kvo_read_vgpu_id_from_file"):
            getattr(L, fn).restype = C.c_long
            getattr(L, fn).argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, sz]
        L.kvo_read_gpu_id_for_vgpu.restype = C.c_long
        L.kvo_read_gpu_id_for_vgpu.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, sz]
        L.kvo_read_numa_node.restype = C.c_int
        L.kvo_read_numa_node.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int64)]
        L.kvo_is_supported_vfio_driver.restype = C.c_int
        L.kvo_is_supported_vfio_driver.argtypes = [C.c_char_p]
        L.kvo_maps_new.restype = C.c_void_p
        L.kvo_maps_free.argtypes = [C.c_void_p]
        L.kvo_create_iommu_device_map_tree.argtypes = [C.c_void_p, C.c_char_p]
        L.kvo_create_vgpu_id_map_tree.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.kvo_create_iommu_device_map_flat.argtypes = [C.c_void_p, u8p, sz]
        L.kvo_create_vgpu_id_map_flat.argtypes = [C.c_void_p, u8p, sz, C.POINTER(TypeDict)]
        L.kvo_dump.argtypes = [C.c_void_p, u8p, sz, C.POINTER(C.c_void_p), C.POINTER(sz)]
        L.kvo_free.argtypes = [C.c_void_p]
        L.kvo_sha256.argtypes = [u8p, sz, u8p]
        L.kvo_maps_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 5
        L.kvo_mix.restype = C.c_uint64
        L.kvo_mix.argtypes = [C.c_uint64]
        L.kvo_gen_pci.restype = None
        L.kvo_gen_pci.argtypes = [u8p, C.c_uint64, sz, u8p, C.c_uint32, C.c_uint32]
        L.kvo_gen_mdev.restype = None
        L.kvo_gen_mdev.argtypes = [u8p, C.c_uint64, sz]
        L.kvo_gen_type_name.restype = sz
        L.kvo_gen_type_name.argtypes = [C.c_uint32, C.c_char_p, sz]
        L.kvo_format_bdf.argtypes = [C.c_uint32, C.c_char_p]
        L.kvo_format_uuid.argtypes = [u8p, C.c_char_p]
        for fn in ("kvo_bench_faithful", "kvo_bench_threads"):
            getattr(L, fn).restype = C.c_double
        L.kvo_bench_faithful.argtypes = [u8p, sz, u8p, sz, C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64)]
        L.kvo_bench_threads.argtypes = [u8p, sz, u8p, sz, C.c_int, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _buf(b: bytes):
    return C.cast(C.c_char_p(b), C.c_void_p), len(b)


def get_device_name(text: bytes, key: bytes | str) -> str:
    """getDeviceName (device_plugin.go:371-422) on an in-memory pci.ids."""
    if isinstance(key, str):
        key = key.encode()
    out = C.create_string_buffer(3 * 70000)
    n = lib().kvo_get_device_name(C.cast(C.c_char_p(text), C.c_void_p), len(text),
                                  C.cast(C.c_char_p(key), C.c_void_p), len(key),
                                  C.cast(out, C.c_void_p), len(out))
    assert n >= 0
    return out.raw[:n].decode("latin-1")


def get_device_name_file(path: str, key: bytes | str) -> str:
    if isinstance(key, str):
        key = key.encode()
    out = C.create_string_buffer(3 * 70000)
    n = lib().kvo_get_device_name_file(path.encode(), C.cast(C.c_char_p(key), C.c_void_p),
                                       len(key), C.cast(out, C.c_void_p), len(out))
    assert n >= 0
    return out.raw[:n].decode("latin-1")


def nv_ids(text: bytes) -> np.ndarray:
    out = np.zeros(65536, dtype=np.uint16)
    n = lib().kvo_nv_ids(C.cast(C.c_char_p(text), C.c_void_p), len(text), out.ctypes.data, 65536)
    return out[:min(n, 65536)].copy()


def trim_space(s: bytes) -> bytes:
    a, b = C.c_size_t(), C.c_size_t()
    lib().kvo_trim_space(C.cast(C.c_char_p(s), C.c_void_p), len(s), C.byref(a), C.byref(b))
    return s[a.value:b.value]


def to_upper(s: bytes) -> bytes:
    out = C.create_string_buffer(3 * len(s) + 1)
    n = lib().kvo_to_upper(C.cast(C.c_char_p(s), C.c_void_p), len(s), C.cast(out, C.c_void_p))
    return out.raw[:n]


def _reader3(fn, base, addr, prop):
    out = C.create_string_buffer(4096)
    rc = getattr(lib(), fn)(base.encode(), addr.encode(), prop.encode(), out, len(out))
    if rc == -2:
        raise IndexError("Go panic: slice bounds out of range")
    if rc < 0:
        return "", True
    return out.raw[:rc].decode("latin-1"), False


def read_id_from_file(base, addr, prop):
    return _reader3("kvo_read_id_from_file", base, addr, prop)


def read_link(base, addr, link):
    return _reader3("kvo_read_link", base, addr, link)


def read_vgpu_id_from_file(base, addr, prop):
    return _reader3("kvo_read_vgpu_id_from_file", base, addr, prop)


def read_gpu_id_for_vgpu(base, addr):
    out = C.create_string_buffer(4096)
    rc = lib().kvo_read_gpu_id_for_vgpu(base.encode(), addr.encode(), out, len(out))
    if rc == -2:
        raise IndexError("Go panic: index out of range")
    if rc < 0:
        return "", True
    return out.raw[:rc].decode("latin-1"), False


def read_numa_node(base, addr):
    v = C.c_int64()
    rc = lib().kvo_read_numa_node(base.encode(), addr.encode(), C.byref(v))
    return v.value, rc != 0


def is_supported_vfio_driver(name: str) -> bool:
    return bool(lib().kvo_is_supported_vfio_driver(name.encode()))


def make_type_dict(raw_types: list[bytes]):
    """-> (TypeDict, keepalive)"""
    off = np.zeros(len(raw_types) + 1, dtype=np.uint32)
    for i, t in enumerate(raw_types):
        off[i + 1] = off[i] + len(t)
    blob = np.frombuffer(b"".join(raw_types) or b"\0", dtype=np.uint8).copy()
    td = TypeDict(len(raw_types), off.ctypes.data_as(C.POINTER(C.c_uint32)),
                  blob.ctypes.data_as(C.POINTER(C.c_uint8)))
    return td, (off, blob)


class Maps:
    """The reference's five package-level maps (device_plugin.go:55-68)."""

    def __init__(self):
        self._h = lib().kvo_maps_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kvo_maps_free(self._h)
            self._h = None

    def create_iommu_device_map_tree(self, base_path: str) -> int:
        return lib().kvo_create_iommu_device_map_tree(self._h, base_path.encode())

    def create_vgpu_id_map_tree(self, vgpu_base: str, pci_base: str) -> int:
        return lib().kvo_create_vgpu_id_map_tree(self._h, vgpu_base.encode(), pci_base.encode())

    def create_iommu_device_map_flat(self, recs: np.ndarray) -> int:
        assert recs.dtype == PCI_REC and recs.flags.c_contiguous
        return lib().kvo_create_iommu_device_map_flat(self._h, recs.ctypes.data, len(recs))

    def create_vgpu_id_map_flat(self, recs: np.ndarray, raw_types: list[bytes]) -> int:
        assert recs.dtype == MDEV_REC and recs.flags.c_contiguous
        td, keep = make_type_dict(raw_types)
        rc = lib().kvo_create_vgpu_id_map_flat(self._h, recs.ctypes.data, len(recs), C.byref(td))
        del keep
        return rc

    def dump(self, pciids: bytes | None) -> bytes:
        out, n = C.c_void_p(), C.c_size_t()
        if pciids is None:
            lib().kvo_dump(self._h, None, 0, C.byref(out), C.byref(n))
        else:
            lib().kvo_dump(self._h, C.cast(C.c_char_p(pciids), C.c_void_p), len(pciids),
                           C.byref(out), C.byref(n))
        data = C.string_at(out.value, n.value)
        lib().kvo_free(out)
        return data

    def counts(self) -> dict:
        v = [C.c_uint64() for _ in range(5)]
        lib().kvo_maps_counts(self._h, *[C.byref(x) for x in v])
        return dict(zip(("dev_keys", "groups", "bdfs", "types", "parents"),
                        [x.value for x in v]))


def sha256(data: bytes) -> str:
    out = C.create_string_buffer(32)
    lib().kvo_sha256(C.cast(C.c_char_p(data), C.c_void_p), len(data), C.cast(out, C.c_void_p))
    return out.raw.hex()


def gen_pci(first: int, n: int, ids: np.ndarray, group_bits: int = 0) -> np.ndarray:
    out = np.zeros(n, dtype=PCI_REC)
    ids = np.ascontiguousarray(ids, dtype=np.uint16)
    lib().kvo_gen_pci(out.ctypes.data, first, n, ids.ctypes.data, len(ids), group_bits)
    return out


def gen_mdev(first: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=MDEV_REC)
    lib().kvo_gen_mdev(out.ctypes.data, first, n)
    return out


def gen_type_names(n: int = 256) -> list[bytes]:
    res = []
    for k in range(n):
        b = C.create_string_buffer(64)
        ln = lib().kvo_gen_type_name(k, b, 64)
        res.append(b.raw[:ln])
    return res


def format_bdf(packed: int) -> str:
    b = C.create_string_buffer(16)
    lib().kvo_format_bdf(packed, b)
    return b.value.decode()


def bench_faithful(recs: np.ndarray, pciids: bytes):
    s, h = C.c_uint64(), C.c_uint64()
    t = lib().kvo_bench_faithful(recs.ctypes.data, len(recs),
                                 C.cast(C.c_char_p(pciids), C.c_void_p), len(pciids),
                                 C.byref(s), C.byref(h))
    return t, s.value, h.value


def bench_threads(recs: np.ndarray, pciids: bytes, threads: int):
    s, h = C.c_uint64(), C.c_uint64()
    t = lib().kvo_bench_threads(recs.ctypes.data, len(recs),
                                C.cast(C.c_char_p(pciids), C.c_void_p), len(pciids), threads,
                                C.byref(s), C.byref(h))
    return t, s.value, h.value


def py_sha256(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()

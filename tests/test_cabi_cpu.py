"""CPU: the C-ABI library loads, exports every declared symbol and fails loudly without a GPU;
host-side logic (snapshotter, map rebuild, dump, sharding) on CPU."""
import os

import numpy as np
import pytest

import conftest
import util


def test_library_exports_every_declared_symbol():
    import kvgpu
    lib = kvgpu.load()
    syms = kvgpu.declared_symbols()
    assert len(syms) >= 30
    assert [s for s in syms if not hasattr(lib, s)] == []
    assert lib.kvg_abi_version() == 1
    tile = lib.kvg_text_pad(1) - 16           # text is padded to whole TMA tiles + a 16-byte halo
    assert tile % 16 == 0 and tile >= 4096
    assert lib.kvg_text_pad(tile) == tile + 16 and lib.kvg_text_pad(tile + 1) == 2 * tile + 16


@pytest.mark.skipif(conftest.HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    import kvgpu
    with pytest.raises(kvgpu.KvgError) as e:
        kvgpu.Context(0)
    assert e.value.rc == -2  # KVG_ECUDA
    with pytest.raises(kvgpu.KvgError):
        kvgpu.DiscoveryScan("/nonexistent", "/nonexistent", "/nonexistent")


def test_product_never_imports_oracle():
    pkg = os.path.join(conftest.PKG)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".go")):
                src = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "kvg_oracle" not in src or f.endswith((".cuh", ".cu")) and "twin" in src, f


def test_wire_format_sizes():
    import kvgpu
    assert kvgpu.PCI_REC.itemsize == 16 and kvgpu.PCI_SURV.itemsize == 16
    assert kvgpu.MDEV_REC.itemsize == 32 and kvgpu.MDEV_SURV.itemsize == 32
    from oracle import oracle as O
    assert O.PCI_REC == kvgpu.PCI_REC and O.MDEV_REC == kvgpu.MDEV_REC


def test_snapshot_pci_tree_config1(tmp_path):
    import kvgpu
    from oracle import oracle as O
    base = util.make_pci_tree(str(tmp_path), util.c1_tree_entries())
    snap = kvgpu.snapshot_pci_tree(base)
    assert snap.packed_addr and snap.group_names is None
    assert snap.names == sorted(util.c1_tree_entries())
    by = {n: r for n, r in zip(snap.names, snap.recs)}
    r = by["0000:04:00.0"]
    assert (r["vendor"], r["device"], r["driver"], r["flags"], r["iommu_group"], r["numa"]) == (
        0x10de, 0x1b38, 1, 0, 40, 0)
    assert by["0000:04:00.1"]["numa"] == -1            # clamp happens on the GPU
    assert by["0000:01:00.0"]["vendor"] == 0x8086       # only vendor is read for non-NVIDIA
    assert by["0000:08:00.0"]["driver"] == 3            # "nvidia": unsupported
    assert by["0000:09:00.0"]["flags"] == 2             # no driver link
    assert by["0000:87:00.0"]["addr"] == kvgpu.parse_bdf("0000:87:00.0")
    # the snapshot, run through the ORACLE's flat path, equals the oracle's tree walk
    m1, m2 = O.Maps(), O.Maps()
    m1.create_iommu_device_map_tree(base)
    m2.create_iommu_device_map_flat(snap.recs)
    assert m1.dump(None) == m2.dump(None)


def test_snapshot_index_mode_and_interned_groups(tmp_path):
    import kvgpu
    G = util.ginkgo()["create_iommu_device_map"]
    base = util.make_pci_tree(str(tmp_path), G["entries"])
    snap = kvgpu.snapshot_pci_tree(base)
    assert not snap.packed_addr and snap.group_names == ["io_1", "io_2", "io_3"]
    assert snap.names == ["1", "2", "3", "4", "5", "6"]
    assert list(snap.recs["flags"]) == [16, 16, 16 | 8, 2, 4, 1]   # numa_node absent everywhere
    assert list(snap.recs["addr"]) == [0, 1, 2, 3, 4, 5]


def test_snapshot_mdev_tree(tmp_path):
    import kvgpu
    spec = util.ginkgo()["create_vgpu_id_map"]
    mdev, pci = util.make_mdev_tree(str(tmp_path), {spec["parent_dir"]: spec["parent_numa_content"]},
                                    spec["entries"])
    snap = kvgpu.snapshot_mdev_tree(mdev, pci)
    assert snap.names == ["1", "2", "3", "4", "5"]
    assert snap.raw_types == [b"vGPUId", b"vGPUId1"] and snap.parent_names == ["GpuId"]
    assert list(snap.recs["flags"]) == [0, 0, 0, 1, 1]  # a real tree cannot make only the link fail
    assert list(snap.recs["parent_numa"][:3]) == [2, 2, 2]


def test_reference_panic_is_surfaced(tmp_path):
    import kvgpu
    ent = {"0000:01:00.0": dict(vendor="10de", device="1b38", driver="vfio-pci", iommu_group="1")}
    base = util.make_pci_tree(str(tmp_path), ent)
    with open(os.path.join(base, "0000:01:00.0", "device"), "w") as f:
        f.write("0")
    with pytest.raises(kvgpu.ReferencePanic):
        kvgpu.snapshot_pci_tree(base)


def test_maps_and_dump_from_flat_results_cpu():
    """pci_maps_from_result + canonical_dump on a hand-built flat result (no GPU involved)."""
    import kvgpu
    surv = np.zeros(3, dtype=kvgpu.PCI_SURV)
    surv["addr"] = [kvgpu.parse_bdf(b) for b in ("0000:04:00.0", "0000:04:00.1", "0000:05:00.0")]
    surv["iommu_group"] = [40, 40, 9]
    surv["device"] = [0x1b38, 0x10f0, 0x1b38]
    surv["numa"] = [0, 0, 1]
    pool = b"\x00" * 4 + bytes([3, 0]) + b"P40"
    res = kvgpu.PciResult(3, surv, np.array([0x10f0, 0x1b38], np.uint16), np.array([0, 1, 3], np.uint32),
                          np.array([1, 0, 2], np.uint32), np.array([0xFFFFFFFF, 4], np.uint32),
                          np.array([9, 40], np.uint32), np.array([0, 1, 3], np.uint32),
                          np.array([2, 0, 1], np.uint32), pool)
    m = kvgpu.pci_maps_from_result(res)
    assert kvgpu.canonical_dump(m) == (
        b"D 10f0 - nvidia.com/10f0 1\n  0000:04:00.1 0\n"
        b"D 1b38 P40 nvidia.com/P40 2\n  0000:04:00.0 0\n  0000:05:00.0 1\n"
        b"I 40 2\n  0000:04:00.0 0\n  0000:04:00.1 0\nI 9 1\n  0000:05:00.0 1\n"
        b"B 0000:04:00.0 40\nB 0000:04:00.1 40\nB 0000:05:00.0 9\n")


def test_shard_ranges_tile_exactly():
    import kvgpu
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [kvgpu.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


# ---- native (C++) host layer: libkvghost.so / kvg-discover ------------------------------------
def _host_lib():
    import ctypes as C
    path = os.path.join(conftest.PKG, "libkvghost.so")
    L = C.CDLL(path)
    L.kvgh_snapshot_pci.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.kvgh_free.argtypes = [C.c_void_p]
    L.kvgh_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    return L


def _native_snapshot(base):
    import ctypes as C
    import kvgpu
    L = _host_lib()
    recs, n, names, nl, groups, gl = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
    rc = L.kvgh_snapshot_pci(base.encode(), C.byref(recs), C.byref(n), C.byref(names), C.byref(nl),
                             C.byref(groups), C.byref(gl), None, None)
    if rc != 0:
        return rc, None, None, None
    arr = np.frombuffer(C.string_at(recs, n.value * 16), dtype=kvgpu.PCI_REC).copy()
    nm = C.string_at(names, nl.value).split(b"\0")[:-1]
    gr = C.string_at(groups, gl.value).split(b"\0")[:-1]
    for p in (recs, names, groups):
        L.kvgh_free(p)
    return 0, arr, [x.decode() for x in nm], [x.decode() for x in gr]


def test_native_snapshot_equals_python_snapshot(tmp_path):
    """The C++ snapshotter (index mode) and the Python one see the same sysfs facts."""
    import kvgpu
    ent = util.c1_tree_entries()
    ent.update(util.ginkgo()["create_iommu_device_map"]["entries"])
    base = util.make_pci_tree(str(tmp_path), ent)
    rc, recs, names, groups = _native_snapshot(base)
    assert rc == 0
    py = kvgpu.snapshot_pci_tree(base)
    assert names == py.names and list(recs["addr"]) == list(range(len(names)))
    for f in ("vendor", "device", "driver", "flags", "numa"):
        assert np.array_equal(recs[f], py.recs[f]), f
    py_groups = py.group_names if py.group_names is not None else None
    nat = [groups[g] if (fl & 4) == 0 and drv in (1, 2) and v == 0x10de and (fl & 3) == 0 else None
           for g, fl, drv, v in zip(recs["iommu_group"], recs["flags"], recs["driver"], recs["vendor"])]
    pyg = [(py_groups[g] if py_groups is not None else str(g)) if x is not None else None
           for g, x in zip(py.recs["iommu_group"], nat)]
    assert nat == pyg


def test_native_snapshot_reports_reference_panic(tmp_path):
    ent = {"0000:01:00.0": dict(vendor="10de", device="1b38", driver="vfio-pci", iommu_group="1")}
    base = util.make_pci_tree(str(tmp_path), ent)
    with open(os.path.join(base, "0000:01:00.0", "device"), "w") as f:
        f.write("0")
    rc, *_ = _native_snapshot(base)
    assert rc == -100  # KVGH_EPANIC


@pytest.mark.skipif(conftest.HAS_GPU, reason="checks the no-GPU failure mode")
def test_native_host_has_no_cpu_fallback():
    import ctypes as C
    import subprocess
    L = _host_lib()
    h = C.c_void_p()
    assert L.kvgh_create(b"/x", b"/y", b"/z", 0, C.byref(h)) == -2  # KVG_ECUDA
    r = subprocess.run([os.path.join(conftest.PKG, "kvg-discover")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_radix_plan_covers_every_key_width():
    """The device-side digit plan (kvg_scan.cuh: radix_plan), evaluated on the host: passes are
    contiguous, no wider than the limit, as few as possible, and cover exactly the key's bits."""
    import ctypes as C
    import kvgpu
    lib = kvgpu.load()
    for max_bits in (8, 11):
        for key_bits_max in (16, 32):
            for kb in range(1, key_bits_max + 1):
                for max_key in {1 << (kb - 1), (1 << kb) - 1}:
                    np_, sh, bt = C.c_uint32(), (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
                    assert lib.kvg_debug_radix_plan(max_key, key_bits_max, max_bits, C.byref(np_), sh, bt) == 0
                    n = np_.value
                    assert n == -(-kb // max_bits) and 1 <= n <= 4
                    assert sh[0] == 0 and all(0 < bt[p] <= max_bits for p in range(n))
                    assert all(sh[p] == sh[p - 1] + bt[p - 1] for p in range(1, n))
                    assert sum(bt[p] for p in range(n)) == kb
                    assert all(bt[p] == 0 for p in range(n, 4))
                    assert max(bt[:n]) - min(bt[:n]) <= max_bits   # even split: no degenerate 1-bit tail pass
    # keys wider than the ordering allows are clamped (device ids are 16-bit by construction)
    np_, sh, bt = C.c_uint32(), (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    assert lib.kvg_debug_radix_plan(0xffffffff, 16, 11, C.byref(np_), sh, bt) == 0
    assert np_.value == 2 and list(bt) == [8, 8, 0, 0]
    assert lib.kvg_debug_radix_plan(0, 32, 11, C.byref(np_), sh, bt) == 0 and np_.value == 1


def test_header_is_plain_c99_and_links(tmp_path):
    """include/kvgpu.h is the contract a cgo / C caller compiles against: it must be valid C99 (no C++
    types, no torch types) and a C program that takes the address of EVERY declared entry point must
    link against libkvgpu.so.  Without a GPU the program's kvg_ctx_create fails loudly (exit code 3)."""
    import subprocess
    import kvgpu
    syms = sorted(kvgpu.declared_symbols())
    src = tmp_path / "all_symbols.c"
    src.write_text('#include "kvgpu.h"\n#include <stdio.h>\n'
                   "typedef void (*any_fn)(void);\n"
                   "static const any_fn table[] = {%s};\n" % ", ".join("(any_fn)%s" % s for s in syms) +
                   "int main(void) {\n  kvg_ctx *c = 0;\n  if (sizeof table / sizeof *table != %d) return 2;\n" % len(syms) +
                   "  if (sizeof(kvg_pci_rec) != 16 || sizeof(kvg_mdev_rec) != 32 || sizeof(kvg_pci_surv) != 16) return 4;\n"
                   "  if (kvg_ctx_create(0, &c) != KVG_OK) { puts(kvg_last_error(0)); return 3; }\n"
                   "  kvg_ctx_destroy(c);\n  return 0;\n}\n")
    exe = tmp_path / "all_symbols"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic",
                        "-I", os.path.join(conftest.ROOT, "include"), str(src), "-L", conftest.PKG, "-lkvgpu",
                        "-Wl,-rpath," + conftest.PKG, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == (0 if conftest.HAS_GPU else 3), (r.returncode, r.stdout)
    if not conftest.HAS_GPU:
        assert "CUDA" in r.stdout or "device" in r.stdout


def test_module_entry_point_fails_loudly_without_gpu():
    """python -m kvgpu is InitiateDevicePlugin(): with no CUDA device it must refuse (exit 2, message),
    never fall back to a CPU scan."""
    import subprocess
    import sys
    env = dict(os.environ, PYTHONPATH=conftest.PKG)
    r = subprocess.run([sys.executable, "-m", "kvgpu", "--help"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "--once" in r.stdout and "InitiateDevicePlugin" in r.stdout
    if conftest.HAS_GPU:
        return
    r = subprocess.run([sys.executable, "-m", "kvgpu", "--once", "--dump"], capture_output=True, text=True, env=env)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr and r.stdout == ""

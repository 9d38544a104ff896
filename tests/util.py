"""Shared helpers for the test-suite: fixtures, synthetic sysfs trees, pci.ids text."""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def ginkgo():
    with open(os.path.join(GOLDEN, "ginkgo_vectors.json")) as f:
        return json.load(f)


def pciids_text() -> bytes:
    with gzip.open(os.path.join(GOLDEN, "pci.ids.gz"), "rb") as f:
        return f.read()


def pciids_names():
    with open(os.path.join(GOLDEN, "pciids_names.json")) as f:
        return json.load(f)


def make_pci_tree(root, entries, via_symlink=True):
    """Build <root>/devices/<addr> like the Ginkgo test does (device_plugin_test.go:281-301):
    every device entry is a SYMLINK to a real directory elsewhere, so filepath.Walk (Lstat) sees a
    non-directory and visits it.  entries: addr -> dict(vendor, device, driver, iommu_group,
    numa_node) where a missing/None value means "that sysfs read fails"."""
    base = os.path.join(root, "devices")
    real = os.path.join(root, "real")
    links = os.path.join(root, "targets")
    os.makedirs(base, exist_ok=True)
    os.makedirs(real, exist_ok=True)
    for addr, e in entries.items():
        d = os.path.join(real, addr) if via_symlink else os.path.join(base, addr)
        os.makedirs(d, exist_ok=True)
        for prop in ("vendor", "device"):
            if e.get(prop) is not None:
                v = e[prop]
                with open(os.path.join(d, prop), "w") as f:
                    f.write(v if v.startswith("0x") else "0x" + v + e.get("nl", "\n"))
        if e.get("numa_node") is not None:
            with open(os.path.join(d, "numa_node"), "w") as f:
                f.write(e["numa_node"])
        for link, sub in (("driver", "drivers"), ("iommu_group", "iommu_groups")):
            if e.get(link) is not None:
                tgt = os.path.join(links, sub, e[link])
                os.makedirs(tgt, exist_ok=True)
                os.symlink(tgt, os.path.join(d, link))
        if via_symlink:
            os.symlink(d, os.path.join(base, addr))
    return base


def make_mdev_tree(root, parents, mdevs):
    """<root>/mdev/<uuid> -> symlink to <root>/pci/<parent>/<uuid> (real dir holding
    mdev_type/name); <root>/pci/<parent>/numa_node.  mdevs: uuid -> dict(type, parent) with None
    meaning the read fails.  parents: bdf -> numa_node content or None."""
    pci = os.path.join(root, "pci")
    mdev = os.path.join(root, "mdev")
    os.makedirs(pci, exist_ok=True)
    os.makedirs(mdev, exist_ok=True)
    for p, numa in parents.items():
        os.makedirs(os.path.join(pci, p), exist_ok=True)
        if numa is not None:
            with open(os.path.join(pci, p, "numa_node"), "w") as f:
                f.write(numa)
    for uuid, e in mdevs.items():
        if e.get("parent") is not None:
            d = os.path.join(pci, e["parent"], uuid)
            os.makedirs(d, exist_ok=True)
            if e.get("type") is not None:
                os.makedirs(os.path.join(d, "mdev_type"), exist_ok=True)
                with open(os.path.join(d, "mdev_type", "name"), "w") as f:
                    f.write(e["type"])
            os.symlink(d, os.path.join(mdev, uuid))
        else:
            # a plain file: Readlink fails -> readGpuIDForVgpu error
            d = os.path.join(mdev, uuid)
            if e.get("type") is not None:
                os.makedirs(os.path.join(root, "orphans", uuid, "mdev_type"), exist_ok=True)
                with open(os.path.join(root, "orphans", uuid, "mdev_type", "name"), "w") as f:
                    f.write(e["type"])
            with open(d, "w") as f:
                f.write("")
    return mdev, pci


def c1_tree_entries():
    """BASELINE.json config 1: 8 vfio-pci Tesla P40 entries plus decoys."""
    ent = {}
    for k, bus in enumerate(["04", "05", "06", "07", "84", "85", "86", "87"]):
        ent["0000:%s:00.0" % bus] = dict(vendor="10de", device="1b38", driver="vfio-pci",
                                         iommu_group=str(40 + k),
                                         numa_node="0\n" if k < 4 else "1\n")
    ent["0000:01:00.0"] = dict(vendor="8086", device="1572", driver="i40e", iommu_group="3",
                               numa_node="0\n")
    ent["0000:08:00.0"] = dict(vendor="10de", device="1b38", driver="nvidia", iommu_group="48",
                               numa_node="0\n")
    ent["0000:04:00.1"] = dict(vendor="10de", device="10f0", driver="vfio-pci", iommu_group="40",
                               numa_node="-1\n")
    ent["0000:09:00.0"] = dict(vendor="10de", device="1b38", iommu_group="49", numa_node="0\n")
    return ent

#!/usr/bin/env python3
"""Regenerate tests/golden/* from the mounted reference (run in the build container only).

  python tests/golden/make_golden.py [/root/reference]

Outputs (all committed, none read from /root/reference at test time):
  ginkgo_vectors.json  known-answer vectors LIFTED from the reference's own Ginkgo suite
                       (pkg/device_plugin/device_plugin_test.go).  The pci.ids fixture text and the
                       getDeviceName expectations are extracted from the Go source by regex, so a
                       transcription error cannot creep in; the sysfs-fixture matrices (which the Go
                       tests express as fake reader functions) are transcribed with file:line.
  pci.ids.gz           the reference's input data file utils/pci.ids (public PCI ID database,
                       v2025.07.11), gzip'd; BASELINE.json config 2 parses exactly this file.
  pciids_names.json    restatement-derived (NOT Go-derived) names of all NVIDIA ids + digests, for
                       regression; cross-checked against SURVEY.md 8c's independent digest.
"""
import gzip
import hashlib
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    test_go = open(os.path.join(REF, "pkg/device_plugin/device_plugin_test.go")).read()

    # ---- getDeviceName fixture + expectations (device_plugin_test.go:373-426)
    m = re.search(r"message := \[\]byte\(`(.*?)`\)", test_go, re.S)
    fixture = m.group(1)
    ctx = test_go[test_go.index('Context("getDeviceName() Tests"'):]
    ctx = ctx[:ctx.index('Context("GetDevicePluginOptions() Tests"')]
    kats = []
    for blk in re.findall(r"It\((.*?)\n\t\t\}\)", ctx, re.S):
        title = re.match(r'"(.*?)"', blk).group(1)
        key = re.search(r'getDeviceName\("(.*?)"\)', blk).group(1)
        want = re.search(r'Expect\(deviceName\)\.To\(Equal\("(.*?)"\)\)', blk).group(1)
        missing_file = 'filepath.Join(workDir, "fake")' in blk
        kats.append({"title": title, "key": key, "want": want, "missing_file": missing_file})
    assert len(kats) == 6, kats

    def const(name):
        return re.search(r'var %s = "(.*?)"' % name, test_go).group(1)

    vectors = {
        "source": "NVIDIA/kubevirt-gpu-device-plugin pkg/device_plugin/device_plugin_test.go",
        "get_device_name": {"cite": "device_plugin_test.go:373-426", "fixture": fixture,
                            "kats": kats},
        "read_link": {"cite": "device_plugin_test.go:136-165",
                      "link_target_basename": "vfio-pci", "missing_is_error": True},
        "read_id_from_file": {"cite": "device_plugin_test.go:168-189",
                              "file_content": "0x10de", "want": "10de"},
        "read_numa_node": {"cite": "device_plugin_test.go:191-219",
                           "cases": [{"content": "3\n", "want": 3, "err": False},
                                     {"content": "-1\n", "want": 0, "err": False},
                                     {"content": None, "want": 0, "err": True}]},
        "read_vgpu_id_from_file": {"cite": "device_plugin_test.go:221-244",
                                   "content": "GRID P100X-1B", "want": "GRID_P100X-1B"},
        "read_gpu_id_for_vgpu": {"cite": "device_plugin_test.go:246-277",
                                 "rule": "second-to-last component of the link target"},
        # fakes :54-100 turned into a real tree by tests/test_oracle_golden.py
        "create_iommu_device_map": {
            "cite": "device_plugin_test.go:279-323 with fakes :54-100",
            "entries": {
                const("deviceAddress1"): {"vendor": "10de", "device": const("deviceName"),
                                          "driver": "vfio-pci", "iommu_group": "io_1"},
                const("deviceAddress2"): {"vendor": "10de", "device": const("deviceName1"),
                                          "driver": "nvgrace_gpu_vfio_pci", "iommu_group": "io_2"},
                const("deviceAddress3"): {"vendor": "10de", "device": None,
                                          "driver": "vfio-pci", "iommu_group": "io_3"},
                const("deviceAddress4"): {"vendor": "10de", "device": const("deviceName"),
                                          "driver": None, "iommu_group": None},
                const("deviceAddress5"): {"vendor": "10de", "device": const("deviceName"),
                                          "driver": "vfio-pci", "iommu_group": None},
                const("deviceAddress6"): {"vendor": None, "device": None, "driver": None,
                                          "iommu_group": None},
            },
            "expect": {"iommuMap": {"io_1": [["1", 0]], "io_2": [["2", 0]]},
                       "deviceMap": {"1b80": [["1", 0]], "1b81": [["2", 0]]},
                       "bdfToIommuMap": {"1": "io_1", "2": "io_2"}},
        },
        "create_vgpu_id_map": {
            "cite": "device_plugin_test.go:325-371 with fakes :109-123",
            "parent_dir": "GpuId", "parent_numa_content": "2\n",
            "entries": {"1": {"type": const("vgpuDeviceName"), "parent": "GpuId"},
                        "2": {"type": const("vgpuDeviceName"), "parent": "GpuId"},
                        "3": {"type": const("vgpuDeviceName1"), "parent": "GpuId"},
                        "4": {"type": const("vgpuDeviceName1"), "parent": None},
                        "5": {"type": None, "parent": None}},
            "expect": {"gpuVgpuMap": {"GpuId": ["1", "2", "3"]},
                       "vGpuMap": {"vGPUId": [["1", 2], ["2", 2]], "vGPUId1": [["3", 2]]}},
        },
        "is_supported_vfio_driver": {"cite": "device_plugin_test.go:535-541",
                                     "yes": ["vfio-pci", "nvgrace_gpu_vfio_pci"],
                                     "no": ["nvidia", "vfio_pci", ""]},
    }
    with open(os.path.join(HERE, "ginkgo_vectors.json"), "w") as f:
        json.dump(vectors, f, indent=1, sort_keys=True)

    # ---- the reference's input data file
    raw = open(os.path.join(REF, "utils/pci.ids"), "rb").read()
    with open(os.path.join(HERE, "pci.ids.gz"), "wb") as f:
        f.write(gzip.compress(raw, 9, mtime=0))

    # ---- restatement-derived names of every NVIDIA id (regression vectors)
    from oracle import oracle as O
    ids = O.nv_ids(raw)
    names = {"%04x" % i: O.get_device_name(raw, "%04x" % i) for i in ids}
    table = "".join("%s %s\n" % (k, names[k]) for k in sorted(names)).encode()
    derived = {
        "note": "restatement-derived (oracle/kvg_oracle.c), NOT produced by the Go reference",
        "pciids_sha256": hashlib.sha256(raw).hexdigest(),
        "pciids_len": len(raw),
        "n_ids": len(ids),
        "ids_in_file_order": ["%04x" % i for i in ids],
        "table_sha256": hashlib.sha256(table).hexdigest(),
        "survey_table_sha256": "a1c4f74e04b980735d2949ba776579c67eeaf872ac418988a1e3eccbc4aa5aca",
        "names": names,
        "extra": {k: O.get_device_name(raw, k) for k in ["ffff", "abcd", "1b3", "", "10de"]},
    }
    assert derived["table_sha256"] == derived["survey_table_sha256"]
    with open(os.path.join(HERE, "pciids_names.json"), "w") as f:
        json.dump(derived, f, indent=0, sort_keys=True)
    print("wrote ginkgo_vectors.json, pci.ids.gz (%d B), pciids_names.json" % len(raw))


if __name__ == "__main__":
    main()

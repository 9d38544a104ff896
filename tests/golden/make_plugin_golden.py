#!/usr/bin/env python3
"""Regenerate tests/golden/plugin_vectors.json from the mounted reference (build container only).

  python tests/golden/make_plugin_golden.py [/root/reference]

Known-answer vectors for the CONSUMER side of the scan: GetPreferredAllocation,
GetDevicePluginOptions and Allocate of the passthrough plugin, lifted from the reference's own
Ginkgo suites.  Request / expectation literals are extracted from the Go source by regex; the fake
reader matrices (Go closures) are transcribed with file:line.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def strs(go_list):
    return re.findall(r'"(.*?)"', go_list)


def main():
    t1 = open(os.path.join(REF, "pkg/device_plugin/device_plugin_test.go")).read()
    t2 = open(os.path.join(REF, "pkg/device_plugin/generic_device_plugin_test.go")).read()

    # ---- GetPreferredAllocation (device_plugin_test.go:438-533)
    ctx = t1[t1.index('Context("GetPreferredAllocation() Tests"'):]
    ctx = ctx[:ctx.index('Context("isSupportedVfioDriver() Tests"')]
    devs = [{"id": i, "numa": int(n)} for i, n in re.findall(r'buildDevice\("(\w+)", (\d+)\)', ctx)]
    cases = []
    for blk in re.findall(r"\n\t\tIt\((.*?)\n\t\t\}\)", ctx, re.S):
        title = re.match(r'"(.*?)"', blk).group(1)
        case = {"title": title,
                "available": strs(re.search(r"AvailableDeviceIDs:\s*\[\]string\{(.*?)\}", blk).group(1)),
                "must_include": strs(re.search(r"MustIncludeDeviceIDs:\s*\[\]string\{(.*?)\}", blk).group(1)),
                "size": int(re.search(r"AllocationSize:\s*(\d+)", blk).group(1))}
        m = re.search(r"DeviceIDs\)\.To\(Equal\(\[\]string\{(.*?)\}\)\)", blk)
        if m:
            case["want"] = strs(m.group(1))
        m = re.search(r"ConsistOf\((.*?)\)\)", blk)
        if m:
            case["want_set"] = strs(m.group(1))
        m = re.search(r'MatchError\("(.*?)"\)', blk)
        if m:
            case["want_error"] = m.group(1)
        cases.append(case)
    assert len(devs) == 4 and len(cases) == 4, (devs, cases)

    def const(name):
        return re.search(r'var %s = "(.*?)"' % name, t2).group(1)

    a = {k: const(k) for k in ("iommuGroup1", "iommuGroup2", "iommuGroup3",
                               "pciAddress1", "pciAddress2", "pciAddress3", "pciAddress4", "nvVendorID")}
    A1, A2, A3, A4 = a["pciAddress1"], a["pciAddress2"], a["pciAddress3"], a["pciAddress4"]
    G1, G2, G3 = a["iommuGroup1"], a["iommuGroup2"], a["iommuGroup3"]
    vectors = {
        "source": "NVIDIA/kubevirt-gpu-device-plugin pkg/device_plugin/{device_plugin,generic_device_plugin}_test.go",
        "options": {"cite": "device_plugin_test.go:428-436", "pre_start_required": False,
                    "get_preferred_allocation_available": True},
        "preferred_allocation": {"cite": "device_plugin_test.go:438-533", "devs": devs, "cases": cases},
        "allocate": {
            "cite": "generic_device_plugin_test.go:64-123 (fakes), :125-172 (fixture), :180-331 (cases)",
            "device_name": "foo",
            "plugin_devs": [A1, A2],                                   # :153-160
            "iommu_map": {G1: [{"addr": A1, "numa": 0}], G2: [{"addr": A2, "numa": 1}],
                          G3: [{"addr": A3, "numa": 2}]},              # :64-71
            "bdf_to_iommu": {A1: G1, A2: G2, A3: G3},                  # :73-79
            "read_link": {A1: G1, A2: G2},                             # :100-108, others error
            "read_vendor_default": {A1: a["nvVendorID"]},              # :110-116, others error
            "read_vendor_shared_egm": {A1: a["nvVendorID"], A2: a["nvVendorID"]},   # :118-123
            "egm_shared": [{"dev_path": "/dev/egm4", "gpus": [A1, A2]}],          # :85-90
            "egm_multi_socket": [{"dev_path": "/dev/egm4", "gpus": [A1, A2]},
                                 {"dev_path": "/dev/egm5", "gpus": [A3, A4]}],  # :92-97
            "env_key": "PCI_RESOURCE_NVIDIA_COM_FOO",                  # gpuPrefix + "_FOO", :182
            "cases": [
                {"cite": ":180-196", "request": [A1], "vendor": "default", "egm": None,
                 "want_env": A1, "want_devices": ["/dev/vfio/vfio", "/dev/vfio/" + G1]},
                {"cite": ":198-217", "request": [A1, A2], "vendor": "shared_egm", "egm": "egm_shared",
                 "want_env": A1 + "," + A2, "want_host_path_count": {"/dev/egm4": 1}},
                {"cite": ":219-233", "request": [A1], "vendor": "default", "egm": "egm_shared",
                 "want_env": A1, "want_absent": ["/dev/egm4"]},
                {"cite": ":235-253", "request": [A1, A2], "vendor": "shared_egm", "egm": "egm_multi_socket",
                 "want_host_path_count": {"/dev/egm4": 1}, "want_absent": ["/dev/egm5"]},
                {"cite": ":255-271", "request": [A1], "vendor": "default", "egm": "error",
                 "want_env": A1, "want_absent": ["/dev/egm4"]},
                {"cite": ":273-300", "request": [A1], "vendor": "default", "egm": None, "iommufd": "vfio3",
                 "want_env": A1,
                 "want_devices": ["/dev/vfio/devices/vfio3", "/dev/vfio/vfio", "/dev/vfio/" + G1, "/dev/iommu"]},
                {"cite": ":302-310", "request": [A2], "vendor": "default", "egm": None, "want_error": True},
                {"cite": ":312-320", "request": [A3], "vendor": "default", "egm": None, "want_error": True},
                {"cite": ":322-331", "request": [A4], "vendor": "default", "egm": None, "want_error": True,
                 "want_error_text": "invalid allocation request: unknown device: " + A4},
            ],
            "permissions": "mrw",
        },
    }
    with open(os.path.join(HERE, "plugin_vectors.json"), "w") as f:
        json.dump(vectors, f, indent=1, sort_keys=True)
    print("wrote plugin_vectors.json:", len(cases), "preferred-allocation cases,",
          len(vectors["allocate"]["cases"]), "allocate cases")


if __name__ == "__main__":
    main()

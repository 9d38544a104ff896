"""The consumer side of the scan (kvgpu.serve) against the reference's own vectors:
GetDevicePluginOptions / GetPreferredAllocation (device_plugin_test.go:428-533), Allocate with the
fake maps and readers of generic_device_plugin_test.go:64-331, ListAndWatch health flips (:347-375)
and a full Register -> ListAndWatch -> Allocate replay against a mock kubelet over unix sockets.

CPU-only: the scan results are hand-built Maps and the classification inside the re-validation is
a numpy test double (tests are the only place that is allowed); tests/test_serve_gpu.py runs the
same flow on the real libkvgpu.so."""
import json
import os
import shutil
import tempfile
import threading

import numpy as np
import pytest

import conftest  # noqa: F401  (sys.path)
import kvgpu
from kvgpu import dpapi, serve

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def V():
    return json.load(open(os.path.join(HERE, "golden", "plugin_vectors.json")))


class _Res:
    def __init__(self, survivors):
        self.survivors = survivors


def fake_scan_pci(recs):
    """Test double for Context.scan_pci: K3's acceptance rule (device_plugin.go:203-238) in numpy,
    survivors in input order."""
    drop = kvgpu._lib.PF_VENDOR_ERR | kvgpu._lib.PF_DRIVER_ERR | kvgpu._lib.PF_IOMMU_ERR | kvgpu._lib.PF_DEVICE_ERR
    keep = (recs["vendor"] == 0x10de) & ((recs["flags"] & drop) == 0) & ((recs["driver"] == 1) | (recs["driver"] == 2))
    surv = np.zeros(int(keep.sum()), dtype=kvgpu.PCI_SURV)
    surv["addr"], surv["iommu_group"], surv["device"] = recs["addr"][keep], recs["iommu_group"][keep], recs["device"][keep]
    return _Res(surv)


def dict_readers(links, vendors):
    """The reference's fake readers (generic_device_plugin_test.go:100-123) as dict lookups."""
    def read_link(base, addr, link):
        return (links[addr], False) if addr in links else ("", True)

    def read_id(base, addr, prop):
        return (vendors[addr], False) if addr in vendors else ("", True)
    return read_link, read_id


def maps_from_vectors(a):
    m = kvgpu.Maps()
    for group, devs in a["iommu_map"].items():
        m.iommuMap[group] = [kvgpu.NvidiaGpuDevice(d["addr"], d["numa"]) for d in devs]
    m.bdfToIommuMap = dict(a["bdf_to_iommu"])
    return m


def plugin_for_case(a, case, root):
    rl, ri = dict_readers(a["read_link"], a["read_vendor_" + case["vendor"]])
    egm = case["egm"]
    if egm == "error":
        def discover():
            raise RuntimeError("egm discovery failed")
    elif egm:
        def discover():
            return [serve.EGMDeviceInfo(e["dev_path"], e["gpus"]) for e in a[egm]]
    else:
        def discover():
            return []
    devs = [dpapi.Device(ID=i, health=dpapi.HEALTHY) for i in a["plugin_devs"]]
    base = os.path.join(root, "bus")
    os.makedirs(base, exist_ok=True)
    if case.get("iommufd"):       # generic_device_plugin_test.go:274-279
        os.makedirs(os.path.join(root, "dev"), exist_ok=True)
        open(os.path.join(root, "dev", "iommu"), "w").close()
        os.makedirs(os.path.join(base, case["request"][0], "vfio-dev", case["iommufd"]))
    reval = serve.BatchRevalidator(fake_scan_pci, base, read_link=rl, read_id=ri)
    return serve.GenericDevicePlugin(a["device_name"], root + "/", devs, maps_from_vectors(a), revalidate=reval,
                                     base_path=base, root_path=root, discover_egm=discover, socket_dir=root)


def allocate(plugin, *requests):
    req = dpapi.AllocateRequest(container_requests=[dpapi.ContainerAllocateRequest(devices_ids=r) for r in requests])
    return plugin.Allocate(req, None)


# ------------------------------------------------------------------------------------------------
def test_device_plugin_options(V):
    o = serve.GenericDevicePlugin("x", "/", [], kvgpu.Maps()).GetDevicePluginOptions(dpapi.Empty(), None)
    assert o.pre_start_required is V["options"]["pre_start_required"]
    assert o.get_preferred_allocation_available is V["options"]["get_preferred_allocation_available"]
    o = serve.GenericVGpuDevicePlugin("x", "/", []).GetDevicePluginOptions(dpapi.Empty(), None)
    assert not o.pre_start_required and not o.get_preferred_allocation_available   # vgpu plugin :252-257


def test_preferred_allocation_vectors(V):
    p = V["preferred_allocation"]
    devs = [(d["id"], d["numa"]) for d in p["devs"]]
    plugin = serve.GenericDevicePlugin("test", "/", [
        dpapi.Device(ID=i, health=dpapi.HEALTHY, topology=dpapi.TopologyInfo(nodes=[dpapi.NUMANode(ID=n)]))
        for i, n in devs], kvgpu.Maps())
    for c in p["cases"]:
        req = dpapi.PreferredAllocationRequest(container_requests=[dpapi.ContainerPreferredAllocationRequest(
            available_deviceIDs=c["available"], must_include_deviceIDs=c["must_include"], allocation_size=c["size"])])
        if "want_error" in c:
            with pytest.raises(serve.AllocateError, match=c["want_error"].replace("(", r"\(").replace(")", r"\)")):
                plugin.GetPreferredAllocation(req, None)
            with pytest.raises(serve.AllocateError):
                serve.preferred_allocation(devs, c["available"], c["must_include"], c["size"])
            continue
        got = list(plugin.GetPreferredAllocation(req, None).container_responses[0].deviceIDs)
        assert got == serve.preferred_allocation(devs, c["available"], c["must_include"], c["size"])
        if "want" in c:
            assert got == c["want"], c["title"]
        else:
            assert sorted(got) == sorted(c["want_set"]) and len(got) == len(c["want_set"]), c["title"]


def test_preferred_allocation_edges():
    devs = [("a", 0), ("b", 0), ("c", 1), ("d", None), ("e", None)]
    pa = serve.preferred_allocation
    assert pa(devs, ["c", "a", "b"], [], 2) == ["a", "b"]              # first node that can hold 2, not first seen
    assert pa(devs, ["c", "a", "b"], [], 1) == ["c"]                   # node order = first appearance
    assert pa(devs, ["a", "c"], [], 2) == ["a", "c"]                   # no single node: kubelet order
    assert pa(devs, ["d", "e", "a", "b"], [], 2) == ["d", "e"]         # the -1 pseudo node stops the search
    assert pa(devs, ["a", "b", "c"], ["c", "c"], 1) == ["c"]           # duplicate must-include counted once
    assert pa(devs, ["a", "b"], [], 0) == []
    assert pa(devs, ["a"], [], 3) == ["a"]                             # fewer devices than asked: what there is
    assert pa(devs, ["a", "b", "c"], ["x"], 2) == ["x", "a"]           # unknown must-include is kept verbatim


def test_allocate_vectors(V, tmp_path):
    a = V["allocate"]
    for k, case in enumerate(a["cases"]):
        root = str(tmp_path / ("case%d" % k))
        os.makedirs(root)
        plugin = plugin_for_case(a, case, root)
        if case.get("want_error"):
            with pytest.raises(serve.AllocateError) as e:
                allocate(plugin, case["request"])
            if "want_error_text" in case:
                assert str(e.value) == case["want_error_text"]
            assert str(e.value).startswith("invalid allocation request: unknown device: ")
            continue
        resp = allocate(plugin, case["request"]).container_responses[0]
        hosts = [d.host_path for d in resp.devices]
        assert all(d.container_path == d.host_path and d.permissions == a["permissions"] for d in resp.devices)
        if "want_env" in case:
            assert dict(resp.envs) == {a["env_key"]: case["want_env"]}, case["cite"]
        if "want_devices" in case:
            assert hosts == case["want_devices"], case["cite"]
        for path, n in case.get("want_host_path_count", {}).items():
            assert hosts.count(path) == n, case["cite"]
        for path in case.get("want_absent", []):
            assert path not in hosts, case["cite"]


def test_allocate_env_accumulates_across_container_requests(V, tmp_path):
    """envList is declared outside the per-container loop (generic_device_plugin.go:361): the second
    container's env value also carries the first container's devices."""
    a = V["allocate"]
    case = dict(a["cases"][1], egm=None)     # both devices validate
    plugin = plugin_for_case(a, case, str(tmp_path))
    r = allocate(plugin, [a["plugin_devs"][0]], [a["plugin_devs"][1]]).container_responses
    assert r[0].envs[a["env_key"]] == a["plugin_devs"][0]
    assert r[1].envs[a["env_key"]] == ",".join(a["plugin_devs"])
    assert [d.host_path for d in r[1].devices] == ["/dev/vfio/vfio", "/dev/vfio/2"]


def test_allocate_first_error_in_reference_order_wins(tmp_path):
    m = kvgpu.Maps()
    m.iommuMap = {"7": [kvgpu.NvidiaGpuDevice("a", 0)], "8": [kvgpu.NvidiaGpuDevice("b", 0)]}
    m.bdfToIommuMap = {"a": "7", "b": "8", "ghost": "7"}      # "ghost" maps to a group that does not list it
    rl, ri = dict_readers({"a": "7", "b": "9"}, {"a": "10de", "b": "10de"})   # b's group moved 8 -> 9
    base = str(tmp_path)
    p = serve.GenericDevicePlugin("n", "/", [], m, revalidate=serve.BatchRevalidator(fake_scan_pci, base, rl, ri),
                                  base_path=base, root_path=base, discover_egm=lambda: [])
    with pytest.raises(serve.AllocateError, match="unknown device: ghost$"):
        allocate(p, ["ghost", "b"])        # bdf 0: members validate, requested id not among them (:413-415)
    with pytest.raises(serve.AllocateError, match="unknown device: b$"):
        allocate(p, ["b", "ghost"])        # bdf 0 already fails its re-validation (:389-392)
    with pytest.raises(serve.AllocateError, match="unknown device: nope$"):
        allocate(p, ["a", "nope"])         # unknown BDF (:378-380)
    with pytest.raises(serve.AllocateError, match="unknown device: b$"):
        allocate(p, ["b", "nope"])         # validation failure of bdf 0 comes before the lookup of bdf 1
    assert dict(allocate(p, ["a"]).container_responses[0].envs) == {"PCI_RESOURCE_NVIDIA_COM_N": "a"}


def test_batch_revalidator_checks_only_group_and_vendor(tmp_path):
    links = {"a": "1", "b": "1", "c": "2", "d": "2"}
    vendors = {"a": "10de", "b": "10de", "c": "8086"}       # d: vendor read fails
    rv = serve.BatchRevalidator(fake_scan_pci, str(tmp_path), *dict_readers(links, vendors))
    assert rv([]) is None
    assert rv([("a", "1"), ("b", "1")]) is None
    assert rv([("a", "1"), ("b", "2")]) == 1                # link points elsewhere
    assert rv([("c", "2")]) == 0                            # vendor changed
    assert rv([("a", "1"), ("d", "2"), ("c", "2")]) == 1    # vendor unreadable; first failure reported
    assert rv([("zz", "1")]) == 0                           # link unreadable
    # a short vendor file makes the reference's reader panic — but only if it gets that far
    def read_id(base, addr, prop):
        if addr == "p":
            raise kvgpu.ReferencePanic("slice bounds out of range")
        return vendors[addr], False
    rv = serve.BatchRevalidator(fake_scan_pci, str(tmp_path), dict_readers(dict(links, p="1"), {})[0], read_id)
    with pytest.raises(kvgpu.ReferencePanic):
        rv([("a", "1"), ("p", "1")])
    assert rv([("a", "2"), ("p", "1")]) == 0                # an earlier rejection returns before the panic
    assert rv([("p", "3")]) == 0                            # link mismatch: the vendor file is never read


def test_vgpu_allocate_skips_foreign_types(tmp_path):
    mdev, pci = __import__("util").make_mdev_tree(str(tmp_path), {"0000:06:00.0": "0\n"}, {
        "u1": dict(type="GRID P100X-1B\n", parent="0000:06:00.0"),
        "u2": dict(type="GRID  P100X-1B", parent="0000:06:00.0"),      # same label after \s+ -> _
        "u3": dict(type="GRID P100X-2B\n", parent="0000:06:00.0"),     # another type: skipped (:221-224)
        "u4": dict(type=None, parent="0000:06:00.0")})                 # unreadable: skipped
    p = serve.GenericVGpuDevicePlugin("GRID_P100X-1B", "vgpu", [], vgpu_base_path=mdev)
    r = allocate(p, ["u1", "u2", "u3", "u4", "missing"]).container_responses[0]
    assert dict(r.envs) == {"MDEV_PCI_RESOURCE_NVIDIA_COM_GRID_P100X-1B": "u1,u2"}
    assert [(d.host_path, d.container_path, d.permissions) for d in r.devices] == [("/dev/vfio", "/dev/vfio", "mrw")]
    r = allocate(p, ["u3"]).container_responses[0]
    assert dict(r.envs) == {} and len(r.devices) == 1
    assert len(p.GetPreferredAllocation(dpapi.PreferredAllocationRequest(), None).container_responses) == 0


def test_egm_discovery_and_selection(tmp_path):
    root = str(tmp_path)
    for name, gpus, node in (("egm5", "0000:0b:00.0\n0000:0c:00.0\n", True), ("egm4", "0000:09:00.0 0000:0a:00.0", True),
                             ("egm7", "", True), ("egm8", "0000:0d:00.0", False), ("other", "x", True)):
        os.makedirs(os.path.join(root, "sys/class/egm", name))
        open(os.path.join(root, "sys/class/egm", name, "gpu_devices"), "w").write(gpus)
        if node:
            os.makedirs(os.path.join(root, "dev"), exist_ok=True)
            open(os.path.join(root, "dev", name), "w").close()
    os.makedirs(os.path.join(root, "sys/class/egm", "egm9"))            # no gpu_devices file: skipped
    egm = serve.discover_egm_devices(root)
    assert [(e.dev_path, e.gpu_bdfs) for e in egm] == [("/dev/egm4", ["0000:09:00.0", "0000:0a:00.0"]),
                                                       ("/dev/egm5", ["0000:0b:00.0", "0000:0c:00.0"])]
    assert serve.discover_egm_devices(os.path.join(root, "nowhere")) == []
    sel = serve.egm_paths_for_allocated_gpus
    assert sel(["0000:09:00.0"], egm) == []
    assert sel([" 0000:0A:00.0", "0000:09:00.0 "], egm) == ["/dev/egm4"]          # trimmed, case-folded
    assert sel(["0000:0c:00.0", "0000:0b:00.0", "0000:09:00.0", "0000:0a:00.0"], egm) == ["/dev/egm4", "/dev/egm5"]
    assert sel(["x"], None) == []
    assert serve.supports_iommufd(root) is False
    open(os.path.join(root, "dev", "iommu"), "w").close()
    assert serve.supports_iommufd(root) is True


def test_list_and_watch_reacts_to_health_changes(V):
    """generic_device_plugin_test.go:347-375 with a fake stream."""
    a = V["allocate"]
    p = serve.GenericDevicePlugin("foo", "/", [dpapi.Device(ID=i, health=dpapi.HEALTHY) for i in a["plugin_devs"]],
                                  kvgpu.Maps())
    stream = p.ListAndWatch(dpapi.Empty(), None)
    first = next(stream)
    assert [(d.ID, d.health) for d in first.devices] == [(a["plugin_devs"][0], "Healthy"), (a["plugin_devs"][1], "Healthy")]
    p.unhealthy(a["plugin_devs"][1])
    assert [(d.ID, d.health) for d in next(stream).devices] == [(a["plugin_devs"][0], "Healthy"),
                                                                (a["plugin_devs"][1], "Unhealthy")]
    p.healthy(a["plugin_devs"][1])
    assert [d.health for d in next(stream).devices] == ["Healthy", "Healthy"]
    p._stop.set()
    assert list(stream) == []


def test_health_rescan_feed_routes_transitions():
    class Delta:
        def __init__(self, changed):
            self.changed = np.array(changed, dtype=np.uint32)
    p1 = serve.GenericDevicePlugin("one", "/", [dpapi.Device(ID="a", health="Healthy"), dpapi.Device(ID="b", health="Healthy")],
                                   kvgpu.Maps())
    p2 = serve.GenericDevicePlugin("two", "/", [dpapi.Device(ID="c", health="Healthy")], kvgpu.Maps())
    script = iter([[0 << 1 | 1, 1 << 1 | 1, 2 << 1 | 1], [], [1 << 1 | 0, 2 << 1 | 0, 3 << 1 | 1], [1 << 1 | 1]])
    feed = serve.HealthRescanFeed(lambda recs: Delta(next(script)), lambda: (None, ["a", "b", "c", "stranger"]), [p1, p2])
    assert feed.tick() == 0                      # the first tick only primes the alive set
    assert feed.tick() == 0
    assert feed.tick() == 2                      # b and c went away; "stranger" belongs to nobody
    assert [p1._events.get_nowait(), p2._events.get_nowait()] == [("unhealthy", "b"), ("unhealthy", "c")]
    assert feed.tick() == 1 and p1._events.get_nowait() == ("healthy", "b")


# ------------------------------------------------------------------------------------------------
# Register -> ListAndWatch -> Allocate over real gRPC / unix sockets against the mock kubelet
# ------------------------------------------------------------------------------------------------
def c1_maps():
    """What the scan returns for BASELINE.json config 1 (8 Tesla P40 on vfio-pci + the audio function)."""
    m = kvgpu.Maps()
    buses = ["04", "05", "06", "07", "84", "85", "86", "87"]
    m.deviceMap["1b38"] = [kvgpu.NvidiaGpuDevice("0000:%s:00.0" % b, 0 if k < 4 else 1) for k, b in enumerate(buses)]
    m.deviceMap["10f0"] = [kvgpu.NvidiaGpuDevice("0000:04:00.1", 0)]
    for k, b in enumerate(buses):
        m.iommuMap[str(40 + k)] = [kvgpu.NvidiaGpuDevice("0000:%s:00.0" % b, 0 if k < 4 else 1)]
        m.bdfToIommuMap["0000:%s:00.0" % b] = str(40 + k)
    m.iommuMap["40"].append(kvgpu.NvidiaGpuDevice("0000:04:00.1", 0))
    m.bdfToIommuMap["0000:04:00.1"] = "40"
    m.deviceNames = {"1b38": "GP102GL_TESLA_P40", "10f0": "GP102_HDMI_AUDIO_CONTROLLER"}
    m.vGpuMap["GRID_P100X-1B"] = [kvgpu.NvidiaGpuDevice("3f4c2b1a-0000-4000-8000-000000000001", 1)]
    m.deviceNames["GRID_P100X-1B"] = ""
    return m


def test_register_list_allocate_against_mock_kubelet():
    import grpc
    sockdir = tempfile.mkdtemp(prefix="kvg", dir="/tmp")     # unix socket paths are limited to 107 bytes
    kubelet = serve.MockKubelet(sockdir).start()
    maps = c1_maps()
    links = {a: g for a, g in maps.bdfToIommuMap.items()}
    vendors = {a: "10de" for a in links}
    reval = serve.BatchRevalidator(fake_scan_pci, sockdir, *dict_readers(links, vendors))
    specs = kvgpu.plugin_specs_from_maps(maps)
    plugins = serve.plugins_from_specs(specs, maps, reval, socket_dir=sockdir, base_path=sockdir, root_path=sockdir,
                                       discover_egm=lambda: [], vgpu_base_path=sockdir)
    try:
        for p in plugins:
            p.start()
        regs = kubelet.wait_for(3)
        assert [(r.version, r.endpoint, r.resource_name) for r in regs] == [
            ("v1beta1", "kubevirt-GP102GL_TESLA_P40.sock", "nvidia.com/GP102GL_TESLA_P40"),
            ("v1beta1", "kubevirt-GP102_HDMI_AUDIO_CONTROLLER.sock", "nvidia.com/GP102_HDMI_AUDIO_CONTROLLER"),
            ("v1beta1", "kubevirt-GRID_P100X-1B.sock", "nvidia.com/GRID_P100X-1B")]
        c = kubelet.connect(regs[0])
        assert c.options().get_preferred_allocation_available is True
        stream = c.list_and_watch()
        first = next(stream)
        assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in first.devices] == [
            (d.addr, "Healthy", d.numaNode) for d in maps.deviceMap["1b38"]]
        plugins[0].unhealthy("0000:85:00.0")
        second = next(stream)
        assert [d.health for d in second.devices] == ["Healthy"] * 5 + ["Unhealthy"] + ["Healthy"] * 2
        stream.cancel()

        # preferred allocation: must-include on node 1, fill from node 1
        r = c.preferred_allocation([d.addr for d in maps.deviceMap["1b38"]], ["0000:86:00.0"], 3)
        assert list(r.container_responses[0].deviceIDs) == ["0000:86:00.0", "0000:84:00.0", "0000:85:00.0"]

        # Allocate: group 40 holds the GPU and its audio function -> both addresses in the env value
        r = c.allocate(["0000:04:00.0"], ["0000:87:00.0"]).container_responses
        assert dict(r[0].envs) == {"PCI_RESOURCE_NVIDIA_COM_GP102GL_TESLA_P40": "0000:04:00.0,0000:04:00.1"}
        assert [d.host_path for d in r[0].devices] == ["/dev/vfio/vfio", "/dev/vfio/40"]
        assert r[1].envs["PCI_RESOURCE_NVIDIA_COM_GP102GL_TESLA_P40"] == "0000:04:00.0,0000:04:00.1,0000:87:00.0"
        assert [d.host_path for d in r[1].devices] == ["/dev/vfio/vfio", "/dev/vfio/47"]

        # the group link of 0000:06:00.0 changes under us -> Allocate is refused with the reference's text
        links["0000:06:00.0"] = "99"
        with pytest.raises(grpc.RpcError) as e:
            c.allocate(["0000:06:00.0"])
        assert e.value.code() == grpc.StatusCode.UNKNOWN
        assert e.value.details() == "invalid allocation request: unknown device: 0000:06:00.0"
        with pytest.raises(grpc.RpcError) as e:
            c.allocate(["0000:ff:00.0"])
        assert e.value.details() == "invalid allocation request: unknown device: 0000:ff:00.0"
        c.close()

        # the vGPU plugin: no preferred allocation, /dev/vfio as the only device spec
        v = kubelet.connect(regs[2])
        assert v.options().get_preferred_allocation_available is False
        assert [d.ID for d in next(v.list_and_watch()).devices] == ["3f4c2b1a-0000-4000-8000-000000000001"]
        v.close()

        # kubelet restart: the plugin re-registers on restart() (:276-287, :669-676)
        plugins[1].restart()
        assert kubelet.wait_for(4)[3].endpoint == "kubevirt-GP102_HDMI_AUDIO_CONTROLLER.sock"
    finally:
        for p in plugins:
            p.stop()
        kubelet.stop()
        leftovers = os.listdir(sockdir)
        shutil.rmtree(sockdir, ignore_errors=True)
    assert leftovers == []        # Stop removes the plugin sockets (:271), the mock kubelet its own


def test_start_twice_and_register_without_kubelet():
    sockdir = tempfile.mkdtemp(prefix="kvg", dir="/tmp")
    p = serve.GenericDevicePlugin("solo", "/", [], kvgpu.Maps(), socket_dir=sockdir)
    try:
        with pytest.raises(Exception):
            p.start()                     # no kubelet socket: Register fails like :289-293
    finally:
        p.stop()
        shutil.rmtree(sockdir, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
# native twins in libkvghost.so (kvg_host.cpp): same vectors, plus a differential fuzz
# ------------------------------------------------------------------------------------------------
def _host_lib():
    import ctypes as C
    L = C.CDLL(os.path.join(conftest.PKG, "libkvghost.so"))
    L.kvgh_preferred_allocation.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_size_t)]
    L.kvgh_egm_paths_for_allocated.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.kvgh_free.argtypes = [C.c_void_p]
    return L


def _native_text(L, fn, *args):
    import ctypes as C
    out, n = C.c_void_p(), C.c_size_t()
    rc = fn(*args, C.byref(out), C.byref(n))
    text = C.string_at(out, n.value).decode() if out.value else ""
    L.kvgh_free(out)
    return rc, text


def native_preferred(L, devs, available, must, size):
    blob = "".join("%s\t%d\n" % (i, n) if n is not None else "%s\n" % i for i, n in devs).encode()
    rc, text = _native_text(L, L.kvgh_preferred_allocation, blob, "\n".join(available).encode(),
                            "\n".join(must).encode(), size)
    if rc == -101:
        raise serve.AllocateError(text)
    assert rc == 0
    return text.split("\n")[:-1] if text else []


def test_native_preferred_allocation_matches_vectors_and_python(V):
    L = _host_lib()
    p = V["preferred_allocation"]
    devs = [(d["id"], d["numa"]) for d in p["devs"]]
    for c in p["cases"]:
        if "want_error" in c:
            with pytest.raises(serve.AllocateError) as e:
                native_preferred(L, devs, c["available"], c["must_include"], c["size"])
            assert str(e.value) == c["want_error"]
        else:
            got = native_preferred(L, devs, c["available"], c["must_include"], c["size"])
            assert got == serve.preferred_allocation(devs, c["available"], c["must_include"], c["size"])
            assert got == c["want"] if "want" in c else sorted(got) == sorted(c["want_set"])
    rng = np.random.default_rng(11)
    pool = ["g%d" % i for i in range(12)]
    for trial in range(400):
        devs = [(i, (int(rng.integers(0, 3)) if rng.random() < 0.8 else None)) for i in pool[:int(rng.integers(1, 12))]]
        ids = [d[0] for d in devs] + ["ghost"]
        available = [str(x) for x in rng.permutation(ids)[:int(rng.integers(0, len(ids) + 1))]]
        must = [str(x) for x in rng.choice(ids, size=int(rng.integers(0, 4)))]
        size = int(rng.integers(0, 6))
        try:
            want = serve.preferred_allocation(devs, available, must, size)
        except serve.AllocateError as e:
            with pytest.raises(serve.AllocateError) as e2:
                native_preferred(L, devs, available, must, size)
            assert str(e2.value) == str(e)
            continue
        assert native_preferred(L, devs, available, must, size) == want, (devs, available, must, size)


def test_native_egm_selection_matches_python():
    L = _host_lib()
    egm = [serve.EGMDeviceInfo("/dev/egm5", ["0000:0b:00.0", "0000:0c:00.0"]),
           serve.EGMDeviceInfo("/dev/egm4", ["0000:09:00.0", "0000:0A:00.0"])]
    blob = "".join("%s\t%s\n" % (e.dev_path, " ".join(e.gpu_bdfs)) for e in egm).encode()
    for alloc in ([], ["0000:09:00.0"], ["0000:0a:00.0 ", " 0000:09:00.0"],
                  ["0000:0c:00.0", "0000:0b:00.0", "0000:09:00.0", "0000:0A:00.0"]):
        rc, text = _native_text(L, L.kvgh_egm_paths_for_allocated, "\n".join(alloc).encode(), blob)
        assert rc == 0
        assert (text.split("\n")[:-1] if text else []) == serve.egm_paths_for_allocated_gpus(alloc, egm)


def test_device_node_watcher_like_the_reference_health_check(V):
    """generic_device_plugin.go:611-690 / generic_device_plugin_test.go:333-345: removing a device node
    marks its devices unhealthy, re-creating it marks them healthy; removing the plugin socket (kubelet
    restart) makes the plugin re-register."""
    a = V["allocate"]
    sockdir = tempfile.mkdtemp(prefix="kvg", dir="/tmp")
    devdir = os.path.join(sockdir, "vfio")
    os.makedirs(devdir)
    for g in ("1", "2"):
        open(os.path.join(devdir, g), "w").close()
    kubelet = serve.MockKubelet(sockdir).start()
    maps = maps_from_vectors(a)
    p = serve.GenericDevicePlugin("foo", devdir, [dpapi.Device(ID=i, health=dpapi.HEALTHY) for i in a["plugin_devs"]],
                                  maps, socket_dir=sockdir)
    w = None
    try:
        p.start()
        assert len(kubelet.wait_for(1)) == 1
        w = serve.DeviceNodeWatcher(p)
        assert w.path_devices == {os.path.join(devdir, "1"): ["11"], os.path.join(devdir, "2"): ["22"]}
        stream = p.ListAndWatch(dpapi.Empty(), None)
        next(stream)
        os.remove(os.path.join(devdir, "2"))
        assert w.poll_once() == 1
        assert [(d.ID, d.health) for d in next(stream).devices] == [("11", "Healthy"), ("22", "Unhealthy")]
        open(os.path.join(devdir, "2"), "w").close()
        assert w.poll_once() == 1
        assert [d.health for d in next(stream).devices] == ["Healthy", "Healthy"]
        os.rename(os.path.join(devdir, "1"), os.path.join(devdir, "1.gone"))     # fsnotify.Rename
        assert w.poll_once() == 1
        assert [d.health for d in next(stream).devices] == ["Unhealthy", "Healthy"]
        open(os.path.join(devdir, "unrelated"), "w").close()
        assert w.poll_once() == 0
        # kubelet restart: the device-plugin directory is wiped, our socket disappears
        os.remove(p.socket_path)
        assert w.poll_once() == 1 and w.restarted.is_set()
        regs = kubelet.wait_for(2)
        assert [r.endpoint for r in regs] == ["kubevirt-foo.sock"] * 2 and os.path.exists(p.socket_path)
        # a SECOND kubelet restart: the watcher is still there and re-registers again
        os.remove(p.socket_path)
        assert w.poll_once() == 1 and w.restarts == 2
        regs = kubelet.wait_for(3)
        assert len(regs) == 3 and os.path.exists(p.socket_path)
        # ... and it still reports device nodes
        os.remove(os.path.join(devdir, "2"))
        assert w.poll_once() == 1
    finally:
        if w:
            w.stop()
        p.stop()
        kubelet.stop()
        shutil.rmtree(sockdir, ignore_errors=True)


def test_xid_events_mark_every_vgpu_of_the_gpu():
    """generic_vgpu_device_plugin.go:330-339 + watchXIDsFunc :387-433 with the reference's own fixture
    (generic_vgpu_device_plugin_test.go:45-76: one GPU "busID", gpuVgpuMap["busID"] = ["1"], devs "1", "2")."""
    devs = [dpapi.Device(ID="1", health=dpapi.HEALTHY), dpapi.Device(ID="2", health=dpapi.HEALTHY)]
    p = serve.GenericVGpuDevicePlugin("vGPUId", "/nonexistent/", devs, socket_dir=tempfile.mkdtemp())
    r = serve.XidEventRouter([("GPU-a", "busID"), ("GPU-b", "other")], {"busID": ["1"], "other": ["2", "zz"]}, [p])
    stream = p.ListAndWatch(dpapi.Empty(), None)
    assert [d.health for d in next(stream).devices] == ["Healthy", "Healthy"]
    for xid in serve.XID_APPLICATION_ERRORS:                      # application errors: ignored
        assert r.on_event(xid, "GPU-a") == 0
    assert r.on_event(79, "GPU-a") == 1                           # fallen off the bus: its vGPU goes unhealthy
    assert [(d.ID, d.health) for d in next(stream).devices] == [("1", "Unhealthy"), ("2", "Healthy")]
    p.healthy("1")
    next(stream)
    assert r.on_event(48, None) == 2                              # no UUID: every GPU ("zz" is nobody's device)
    assert [d.health for d in next(stream).devices] == ["Unhealthy", "Healthy"]
    assert [d.health for d in next(stream).devices] == ["Unhealthy", "Unhealthy"]
    assert r.on_event(48, "GPU-unknown") == 0
    p.healthy("2")
    next(stream)
    assert r.on_unsupported("GPU-b") == 1                         # too old to register: marked at once
    assert [d.health for d in next(stream).devices] == ["Unhealthy", "Unhealthy"]

"""k8s device-plugin API v1beta1 — messages and service tables built programmatically.

The wire contract is the public kubelet API the reference vendors
(vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto:24-139 and constants.go:19-32).
There is no `protoc` in this image, so the FileDescriptorProto is assembled by hand here; field
names and numbers are the proto's (the gogoproto options only affect generated Go identifiers).

    from kvgpu import dpapi
    dev = dpapi.Device(ID="0000:04:00.0", health=dpapi.HEALTHY,
                       topology=dpapi.TopologyInfo(nodes=[dpapi.NUMANode(ID=0)]))
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

# constants.go:19-32
HEALTHY = "Healthy"
UNHEALTHY = "Unhealthy"
VERSION = "v1beta1"
DEVICE_PLUGIN_PATH = "/var/lib/kubelet/device-plugins/"
KUBELET_SOCKET = DEVICE_PLUGIN_PATH + "kubelet.sock"

_T = descriptor_pb2.FieldDescriptorProto
_SCALAR = {"string": _T.TYPE_STRING, "bool": _T.TYPE_BOOL, "int32": _T.TYPE_INT32, "int64": _T.TYPE_INT64}

# message -> [(field name, number, type, repeated)]; "map" = map<string,string>
_MESSAGES = {
    "DevicePluginOptions": [("pre_start_required", 1, "bool", False),
                            ("get_preferred_allocation_available", 2, "bool", False)],
    "RegisterRequest": [("version", 1, "string", False), ("endpoint", 2, "string", False),
                        ("resource_name", 3, "string", False), ("options", 4, "DevicePluginOptions", False)],
    "Empty": [],
    "ListAndWatchResponse": [("devices", 1, "Device", True)],
    "TopologyInfo": [("nodes", 1, "NUMANode", True)],
    "NUMANode": [("ID", 1, "int64", False)],
    "Device": [("ID", 1, "string", False), ("health", 2, "string", False), ("topology", 3, "TopologyInfo", False)],
    "PreStartContainerRequest": [("devices_ids", 1, "string", True)],
    "PreStartContainerResponse": [],
    "PreferredAllocationRequest": [("container_requests", 1, "ContainerPreferredAllocationRequest", True)],
    "ContainerPreferredAllocationRequest": [("available_deviceIDs", 1, "string", True),
                                            ("must_include_deviceIDs", 2, "string", True),
                                            ("allocation_size", 3, "int32", False)],
    "PreferredAllocationResponse": [("container_responses", 1, "ContainerPreferredAllocationResponse", True)],
    "ContainerPreferredAllocationResponse": [("deviceIDs", 1, "string", True)],
    "AllocateRequest": [("container_requests", 1, "ContainerAllocateRequest", True)],
    "ContainerAllocateRequest": [("devices_ids", 1, "string", True)],
    "CDIDevice": [("name", 1, "string", False)],
    "AllocateResponse": [("container_responses", 1, "ContainerAllocateResponse", True)],
    "ContainerAllocateResponse": [("envs", 1, "map", True), ("mounts", 2, "Mount", True),
                                  ("devices", 3, "DeviceSpec", True), ("annotations", 4, "map", True),
                                  ("cdi_devices", 5, "CDIDevice", True)],
    "Mount": [("container_path", 1, "string", False), ("host_path", 2, "string", False),
              ("read_only", 3, "bool", False)],
    "DeviceSpec": [("container_path", 1, "string", False), ("host_path", 2, "string", False),
                   ("permissions", 3, "string", False)],
}

# service -> {method: (request, response, server_streaming)}      api.proto:24-26, 51-77
SERVICES = {
    "Registration": {"Register": ("RegisterRequest", "Empty", False)},
    "DevicePlugin": {
        "GetDevicePluginOptions": ("Empty", "DevicePluginOptions", False),
        "ListAndWatch": ("Empty", "ListAndWatchResponse", True),
        "GetPreferredAllocation": ("PreferredAllocationRequest", "PreferredAllocationResponse", False),
        "Allocate": ("AllocateRequest", "AllocateResponse", False),
        "PreStartContainer": ("PreStartContainerRequest", "PreStartContainerResponse", False),
    },
}
PACKAGE = "v1beta1"


def _camel(name):
    return "".join(p[:1].upper() + p[1:] for p in name.split("_"))


def _build():
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name = "kvgpu/deviceplugin_v1beta1.proto"
    fdp.package = PACKAGE
    fdp.syntax = "proto3"
    for mname, fields in _MESSAGES.items():
        msg = fdp.message_type.add()
        msg.name = mname
        for fname, num, ftype, rep in fields:
            f = msg.field.add()
            f.name, f.number = fname, num
            f.label = _T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL
            if ftype == "map":
                entry = msg.nested_type.add()
                entry.name = _camel(fname) + "Entry"
                entry.options.map_entry = True
                for en, enum_ in (("key", 1), ("value", 2)):
                    ef = entry.field.add()
                    ef.name, ef.number, ef.label, ef.type = en, enum_, _T.LABEL_OPTIONAL, _T.TYPE_STRING
                f.type = _T.TYPE_MESSAGE
                f.type_name = ".%s.%s.%s" % (PACKAGE, mname, entry.name)
            elif ftype in _SCALAR:
                f.type = _SCALAR[ftype]
            else:
                f.type = _T.TYPE_MESSAGE
                f.type_name = ".%s.%s" % (PACKAGE, ftype)
    for sname, methods in SERVICES.items():
        svc = fdp.service.add()
        svc.name = sname
        for mname, (req, resp, stream) in methods.items():
            m = svc.method.add()
            m.name = mname
            m.input_type = ".%s.%s" % (PACKAGE, req)
            m.output_type = ".%s.%s" % (PACKAGE, resp)
            m.server_streaming = stream
    pool = descriptor_pool.DescriptorPool()   # private pool: never collides with a generated module
    pool.Add(fdp)
    return {name: message_factory.GetMessageClass(pool.FindMessageTypeByName("%s.%s" % (PACKAGE, name)))
            for name in _MESSAGES}


_CLASSES = _build()
globals().update(_CLASSES)
MESSAGES = dict(_CLASSES)


def method_path(service: str, method: str) -> str:
    return "/%s.%s/%s" % (PACKAGE, service, method)


def service_name(service: str) -> str:
    return "%s.%s" % (PACKAGE, service)

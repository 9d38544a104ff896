"""CPU: pin the oracle (oracle/kvg_oracle.c) to the reference's own Ginkgo vectors
(pkg/device_plugin/device_plugin_test.go) and to the derived pci.ids vectors."""
import hashlib
import os

import pytest

import util
from oracle import oracle as O

G = util.ginkgo()


# ---- getDeviceName() Tests (device_plugin_test.go:373-426) -----------------------------------
@pytest.mark.parametrize("kat", G["get_device_name"]["kats"], ids=lambda k: k["key"])
def test_get_device_name_kats(kat, tmp_path):
    fixture = G["get_device_name"]["fixture"].encode()
    if kat["missing_file"]:
        assert O.get_device_name_file(str(tmp_path / "fake"), kat["key"]) == kat["want"]
        return
    assert O.get_device_name(fixture, kat["key"]) == kat["want"]
    p = tmp_path / "pci.ids"
    p.write_bytes(fixture)
    assert O.get_device_name_file(str(p), kat["key"]) == kat["want"]


def test_shipped_pciids_digest():
    text = util.pciids_text()
    d = util.pciids_names()
    assert hashlib.sha256(text).hexdigest() == d["pciids_sha256"]
    ids = O.nv_ids(text)
    assert len(ids) == d["n_ids"] == 1931
    assert ["%04x" % i for i in ids] == d["ids_in_file_order"]
    table = "".join("%04x %s\n" % (i, O.get_device_name(text, "%04x" % i))
                    for i in sorted(ids)).encode()
    # SURVEY.md 8c digest, computed there by an independent throw-away Python restatement
    assert hashlib.sha256(table).hexdigest() == d["survey_table_sha256"]
    for k, v in d["extra"].items():
        assert O.get_device_name(text, k) == v
    assert O.get_device_name(text, "1b38") == "GP102GL_TESLA_P40"
    assert O.get_device_name(text, "1b3") == "0_GP102GL_QUADRO_P6000"  # prefix, not equality
    assert O.get_device_name(text, "") == "0008_NV1_STG2000XB_SERIES"


# ---- reader tests (device_plugin_test.go:136-277) ---------------------------------------------
def test_read_link(tmp_path):
    link_dir = tmp_path / "dp-test"
    (link_dir / "vfio-pci").mkdir(parents=True)
    work = tmp_path / "kubevirt-test"
    (work / "1").mkdir(parents=True)
    os.symlink(link_dir / "vfio-pci", work / "1" / "driver")
    assert O.read_link(str(work), "1", "driver") == ("vfio-pci", False)
    assert O.read_link(str(work), "1", "iommu_group") == ("", True)


def test_read_id_from_file(tmp_path):
    (tmp_path / "1").mkdir()
    (tmp_path / "1" / "vendor").write_text(G["read_id_from_file"]["file_content"])
    assert O.read_id_from_file(str(tmp_path), "1", "vendor") == (G["read_id_from_file"]["want"], False)
    assert O.read_id_from_file(str(tmp_path), "1", "iommu_group") == ("", True)
    (tmp_path / "1" / "short").write_text("0")
    with pytest.raises(IndexError):  # data[2:] on a 1-byte file panics in Go
        O.read_id_from_file(str(tmp_path), "1", "short")


@pytest.mark.parametrize("case", G["read_numa_node"]["cases"])
def test_read_numa_node(case, tmp_path):
    (tmp_path / "1").mkdir()
    if case["content"] is not None:
        (tmp_path / "1" / "numa_node").write_text(case["content"])
    assert O.read_numa_node(str(tmp_path), "1") == (case["want"], case["err"])


def test_read_numa_node_parse_errors(tmp_path):
    (tmp_path / "1").mkdir()
    for content, want in [("abc\n", (0, True)), ("", (0, True)), (" 7 \n", (7, False)),
                          ("+5", (5, False)), ("9223372036854775808", (0, True)),
                          ("-9223372036854775808", (0, False)), ("1_0", (0, True))]:
        (tmp_path / "1" / "numa_node").write_text(content)
        assert O.read_numa_node(str(tmp_path), "1") == want, content


def test_read_vgpu_id_from_file(tmp_path):
    (tmp_path / "1").mkdir()
    (tmp_path / "1" / "name").write_text(G["read_vgpu_id_from_file"]["content"])
    assert O.read_vgpu_id_from_file(str(tmp_path), "1", "name") == (
        G["read_vgpu_id_from_file"]["want"], False)
    assert O.read_vgpu_id_from_file(str(tmp_path), "1", "error") == ("", True)
    (tmp_path / "1" / "name").write_text("\n\nGRID  \t A100\r\n-4C\n\n")
    assert O.read_vgpu_id_from_file(str(tmp_path), "1", "name") == ("GRID_A100_-4C", False)


def test_read_gpu_id_for_vgpu(tmp_path):
    link_dir = tmp_path / "dp-test"
    (link_dir / "vfio-pci").mkdir(parents=True)
    work = tmp_path / "kubevirt-test"
    (work / "1").mkdir(parents=True)
    os.symlink(link_dir / "vfio-pci", work / "1" / "driver")
    assert O.read_gpu_id_for_vgpu(str(work), "1/driver") == ("dp-test", False)
    assert O.read_gpu_id_for_vgpu(str(work), "1/error") == ("", True)


def test_is_supported_vfio_driver():
    for d in G["is_supported_vfio_driver"]["yes"]:
        assert O.is_supported_vfio_driver(d)
    for d in G["is_supported_vfio_driver"]["no"]:
        assert not O.is_supported_vfio_driver(d)


# ---- createIommuDeviceMap() Tests (device_plugin_test.go:279-323) -----------------------------
def test_create_iommu_device_map_ginkgo(tmp_path):
    spec = G["create_iommu_device_map"]
    base = util.make_pci_tree(str(tmp_path), spec["entries"])
    m = O.Maps()
    assert m.create_iommu_device_map_tree(base) == 0
    dump = m.dump(None).decode()
    want = ("D 1b80 - nvidia.com/1b80 1\n  1 0\n"
            "D 1b81 - nvidia.com/1b81 1\n  2 0\n"
            "I io_1 1\n  1 0\n"
            "I io_2 1\n  2 0\n"
            "B 1 io_1\n"
            "B 2 io_2\n")
    assert dump == want
    exp = spec["expect"]
    assert exp["iommuMap"]["io_1"][0][0] == "1" and exp["deviceMap"]["1b81"][0][0] == "2"


# ---- createVgpuIDMap() Tests (device_plugin_test.go:325-371) ---------------------------------
def test_create_vgpu_id_map_ginkgo(tmp_path):
    spec = G["create_vgpu_id_map"]
    mdev, pci = util.make_mdev_tree(str(tmp_path), {spec["parent_dir"]: spec["parent_numa_content"]},
                                    spec["entries"])
    m = O.Maps()
    assert m.create_vgpu_id_map_tree(mdev, pci) == 0
    dump = m.dump(None).decode()
    want = ("V vGPUId - nvidia.com/vGPUId 2\n  1 2\n  2 2\n"
            "V vGPUId1 - nvidia.com/vGPUId1 1\n  3 2\n"
            "G GpuId 3\n  1\n  2\n  3\n")
    assert dump == want


def test_walk_descends_real_dirs_and_sorts(tmp_path):
    """filepath.Walk: lexical order, real sub-directories are descended and their files become
    candidate 'devices' (which then fail the vendor read) — SURVEY.md 4, createVgpuIDMap row."""
    ent = {"0000:0b:00.0": dict(vendor="10de", device="1b38", driver="vfio-pci",
                                iommu_group="7", numa_node="1\n"),
           "0000:0a:00.0": dict(vendor="10de", device="1b38", driver="vfio-pci",
                                iommu_group="7", numa_node="0\n")}
    base = util.make_pci_tree(str(tmp_path), ent)
    os.makedirs(os.path.join(base, "realdir"))
    with open(os.path.join(base, "realdir", "vendor"), "w") as f:
        f.write("0x10de\n")
    m = O.Maps()
    m.create_iommu_device_map_tree(base)
    text = util.pciids_text()
    assert m.dump(text).decode() == (
        "D 1b38 GP102GL_TESLA_P40 nvidia.com/GP102GL_TESLA_P40 2\n"
        "  0000:0a:00.0 0\n  0000:0b:00.0 1\n"
        "I 7 2\n  0000:0a:00.0 0\n  0000:0b:00.0 1\n"
        "B 0000:0a:00.0 7\nB 0000:0b:00.0 7\n")


def test_config1_tree(tmp_path):
    """BASELINE.json config 1: 8 Tesla P40 on vfio-pci + decoys."""
    base = util.make_pci_tree(str(tmp_path), util.c1_tree_entries())
    m = O.Maps()
    m.create_iommu_device_map_tree(base)
    dump = m.dump(util.pciids_text()).decode()
    assert dump.startswith("D 10f0 GP104_HIGH_DEFINITION_AUDIO_CONTROLLER nvidia.com/GP104_HIGH_DEFINITION_AUDIO_CONTROLLER 1\n"
                           "  0000:04:00.1 0\n"
                           "D 1b38 GP102GL_TESLA_P40 nvidia.com/GP102GL_TESLA_P40 8\n"
                           "  0000:04:00.0 0\n  0000:05:00.0 0\n")
    assert "0000:08:00.0" not in dump and "0000:09:00.0" not in dump and "0000:01:00.0" not in dump
    assert "I 40 2\n  0000:04:00.0 0\n  0000:04:00.1 0\n" in dump
    assert m.counts()["bdfs"] == 9


def test_missing_base_path_gives_empty_maps(tmp_path):
    m = O.Maps()
    assert m.create_iommu_device_map_tree(str(tmp_path / "nope")) == 0
    assert m.dump(None) == b""


# ---- Go stdlib restatements -------------------------------------------------------------------
def test_trim_space_unicode():
    assert O.trim_space(b" \t\v name \r\n") == b"name"
    assert O.trim_space(b" \xc2\xa0 name\xe2\x80\x80 ") == b"name"
    assert O.trim_space(b"\xe3\x80\x80x\xc2\x85") == b"x"
    assert O.trim_space(b"\xffx\xff") == b"\xffx\xff"          # invalid bytes are not spaces
    assert O.trim_space(b"x\xe2\x80\x80\x80") == b"x\xe2\x80\x80\x80"  # bad trailing sequence
    assert O.trim_space(b"   ") == b""


def test_to_upper_simple_mapping():
    assert O.to_upper(b"gk104.gl") == b"GK104.GL"
    assert O.to_upper(b"a\xc4\xb1b\xc5\xbf") == b"AIBS"
    assert O.to_upper(b"\xff") == b"\xef\xbf\xbd"


def test_sanitiser_pass_order():
    mk = lambda name: b"10de  NVIDIA\n\t1234" + name + b"\n"
    assert O.get_device_name(mk(b"  a / b"), "1234") == "A___B"
    assert O.get_device_name(mk(b"  a \xc3\xa9 b"), "1234") == "A__B"
    assert O.get_device_name(mk(b" \xc2\xa0 a.b/c\vd"), "1234") == "A_B_CD"
    assert O.get_device_name(mk(b"  x\r"), "1234") == "X"
    assert O.get_device_name(b"10de\n#c\n\t1234  n1\n\t1234  n2\n", "1234") == "N1"
    assert O.get_device_name(b"10de\n\n\t1234  n1\n", "1234") == ""       # blank line ends block
    assert O.get_device_name(b"10de\n8086 x\n\t1234  n1\n", "1234") == ""  # next vendor ends block
    assert O.get_device_name(b"10de\n\t1234  last-no-newline", "1234") == "LASTNONEWLINE"


def test_scanner_token_limit():
    long_ok = b"x" * 65535
    long_bad = b"x" * 65536
    tail = b"10de  NVIDIA\n\t1234  name\n"
    assert O.get_device_name(long_ok + b"\n" + tail, "1234") == "NAME"
    assert O.get_device_name(long_bad + b"\n" + tail, "1234") == ""   # bufio.ErrTooLong
    assert O.get_device_name(b"10de\n\t" + long_bad + b"\n\t1234  name\n", "1234") == ""
    assert O.get_device_name(b"10de\n\t1234  name\n\t" + long_bad, "1234") == "NAME"


def test_sha256_matches_hashlib():
    for n in (0, 1, 55, 56, 63, 64, 65, 1000, 100000):
        data = bytes((i * 131 + 7) & 0xFF for i in range(n))
        assert O.sha256(data) == hashlib.sha256(data).hexdigest()

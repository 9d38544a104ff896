// kvg_host.cpp — native host layer above the C-ABI (libkvghost.so + the kvg-discover CLI).
//
// The reference's host is compiled Go and no Go toolchain exists in this image, so this is the C++
// statement of what go/kvgpu_cgo.go does: mirror the reference's scan interface — same function
// names, argument meaning and error behaviour — with every filter / join / bucketing decision made by
// libkvgpu.so on the GPU.  Reference: pkg/device_plugin/device_plugin.go
//
//   kvgh_create_iommu_device_map   createIommuDeviceMap   :187-247   (walk + readers -> records -> kvg_scan_pci)
//   kvgh_create_vgpu_id_map        createVgpuIDMap        :255-291
//   kvgh_get_device_name           getDeviceName          :371-422
//   kvgh_device_plugins            payload half of createDevicePlugins :99-157 (resource name,
//                                  socket path, env key, device list per key)
//   kvgh_dump                      canonical parity dump — byte-identical to the oracle's
//
// Nothing here decides which device survives: read failures travel to the GPU as flag bits.
// This file never includes, links or calls anything under oracle/.
#include <dirent.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <ctype.h>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kvgpu.h"

namespace {

struct Dev {  // NvidiaGpuDevice :50-53
  std::string addr;
  long long numa;
};

struct Scan {
  std::string pci_ids_path, base_path, vgpu_base_path;  // pciIdsFilePath :79, basePath :70, vGpuBasePath :74
  kvg_ctx* ctx = nullptr;
  std::string loaded_path;
  bool loaded = false;
  std::string err;
  // the five maps :55-68 (+ the resolved names)
  std::map<std::string, std::vector<Dev>> iommuMap, deviceMap, vGpuMap;
  std::map<std::string, std::string> bdfToIommuMap, deviceNames;
  std::map<std::string, std::vector<std::string>> gpuVgpuMap;
};

bool read_file(const std::string& path, std::string* out) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  out->clear();
  char buf[65536];
  for (;;) {
    ssize_t r = read(fd, buf, sizeof buf);
    if (r < 0) {
      close(fd);
      return false;
    }
    if (r == 0) break;
    out->append(buf, (size_t)r);
  }
  close(fd);
  return true;
}

// filepath.Walk: lexical order, Lstat, real directories descended, a stat/readdir error aborts.
// fn(name, is_dir) returns false to abort.
template <class F>
bool walk_rec(const std::string& path, const std::string& name, const struct stat& st, F& fn) {
  if (!S_ISDIR(st.st_mode)) return fn(name, false);
  DIR* d = opendir(path.c_str());
  if (!d) return false;  // :193-196
  std::vector<std::string> names;
  while (dirent* de = readdir(d)) {
    if (!strcmp(de->d_name, ".") || !strcmp(de->d_name, "..")) continue;
    names.push_back(de->d_name);
  }
  closedir(d);
  std::sort(names.begin(), names.end());
  if (!fn(name, true)) return false;
  for (auto& n : names) {
    std::string child = path + "/" + n;
    struct stat cst;
    if (lstat(child.c_str(), &cst) != 0) return false;
    if (!walk_rec(child, n, cst, fn)) return false;
  }
  return true;
}
template <class F>
void walk(const std::string& root, F fn) {
  struct stat st;
  if (lstat(root.c_str(), &st) != 0) return;
  size_t s = root.rfind('/');
  walk_rec(root, s == std::string::npos ? root : root.substr(s + 1), st, fn);
}

// readIDFromFileFunc :294-302.  rc: 0 ok, 1 error, 2 the Go code would panic (len < 2)
int read_id(const std::string& base, const std::string& addr, const char* prop, std::string* out) {
  std::string data;
  if (!read_file(base + "/" + addr + "/" + prop, &data)) return 1;
  if (data.size() < 2) return 2;
  size_t a = 2, b = data.size();
  while (a < b && data[a] == '\n') a++;
  while (b > a && data[b - 1] == '\n') b--;
  *out = data.substr(a, b - a);
  return 0;
}
// readLinkFunc :323-331
bool read_link(const std::string& base, const std::string& addr, const char* link, std::string* out) {
  char target[4096];
  ssize_t n = readlink((base + "/" + addr + "/" + link).c_str(), target, sizeof target - 1);
  if (n < 0) return false;
  target[n] = 0;
  const char* slash = strrchr(target, '/');
  *out = slash ? slash + 1 : target;
  return true;
}
bool go_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
// readNUMANodeFunc :304-320 (raw value; the clamp <0 -> 0 happens on the GPU).  ASCII TrimSpace is
// enough here: a numa_node file holding Unicode spaces fails ParseInt on both sides only if the
// reference's Unicode trim would not have removed them; sysfs prints "%d\n".
bool read_numa(const std::string& base, const std::string& addr, long long* out) {
  std::string data;
  *out = 0;
  if (!read_file(base + "/" + addr + "/numa_node", &data)) return false;
  size_t a = 0, b = data.size();
  while (a < b && go_space((unsigned char)data[a])) a++;
  while (b > a && go_space((unsigned char)data[b - 1])) b--;
  if (a == b) return false;
  size_t i = a;
  bool neg = false;
  if (data[i] == '+' || data[i] == '-') {
    neg = data[i] == '-';
    if (++i == b) return false;
  }
  unsigned long long v = 0;
  const unsigned long long cutoff = neg ? (1ull << 63) : (1ull << 63) - 1;
  for (; i < b; i++) {
    if (data[i] < '0' || data[i] > '9') return false;
    unsigned long long dgt = (unsigned long long)(data[i] - '0');
    if (v > (cutoff - dgt) / 10) return false;
    v = v * 10 + dgt;
  }
  *out = neg ? -(long long)v : (long long)v;
  return true;
}
// readGpuIDForVgpuFunc :347-357. rc 0 ok, 1 error, 2 panic (no '/')
int read_gpu_id_for_vgpu(const std::string& base, const std::string& addr, std::string* out) {
  char target[4096];
  ssize_t n = readlink((base + "/" + addr).c_str(), target, sizeof target - 1);
  if (n < 0) return 1;
  target[n] = 0;
  char* last = strrchr(target, '/');
  if (!last) return 2;
  *last = 0;
  char* prev = strrchr(target, '/');
  std::string comp = prev ? prev + 1 : target;
  size_t a = 0, b = comp.size();
  while (a < b && comp[a] == '\n') a++;
  while (b > a && comp[b - 1] == '\n') b--;
  *out = comp.substr(a, b - a);
  return 0;
}
bool hex4(const std::string& s, uint16_t* v) {
  if (s.size() != 4) return false;
  unsigned x = 0;
  for (char c : s) {
    unsigned d;
    if (c >= '0' && c <= '9') d = (unsigned)(c - '0');
    else if (c >= 'a' && c <= 'f') d = (unsigned)(c - 'a' + 10);
    else return false;
    x = x * 16 + d;
  }
  *v = (uint16_t)x;
  return true;
}

int ensure_table(Scan* s) {
  if (s->loaded && s->loaded_path == s->pci_ids_path) return KVG_OK;
  std::string data;
  bool ok = read_file(s->pci_ids_path, &data);
  if (!ok) {
    fprintf(stderr, "Error opening pci ids file %s\n", s->pci_ids_path.c_str());  // :375
    data.clear();  // unreadable file: every name is "" (:374-377)
  }
  int rc = kvg_pciids_load(s->ctx, (const uint8_t*)data.data(), data.size());
  if (rc != KVG_OK) {
    s->err = std::string("kvg_pciids_load: ") + kvg_last_error(s->ctx);
    return rc;
  }
  s->loaded = true;
  s->loaded_path = s->pci_ids_path;
  return KVG_OK;
}

std::string pool_name(const kvg_pci_result* r, uint32_t slot) {
  if (slot == KVG_NO_NAME) return "";
  size_t n = r->name_pool[slot] | ((size_t)r->name_pool[slot + 1] << 8);
  return std::string((const char*)r->name_pool + slot + 2, n);
}

}  // namespace

extern "C" {

typedef struct kvgh_scan kvgh_scan;
#define KVGH_EPANIC (-100) /* the Go reference would panic on this sysfs content */

int kvgh_create(const char* pci_ids_path, const char* base_path, const char* vgpu_base_path, int device,
                kvgh_scan** out) {
  if (!out || !pci_ids_path || !base_path || !vgpu_base_path) return KVG_EINVAL;
  Scan* s = new Scan();
  s->pci_ids_path = pci_ids_path;
  s->base_path = base_path;
  s->vgpu_base_path = vgpu_base_path;
  int rc = kvg_ctx_create(device, &s->ctx);
  if (rc != KVG_OK) {  // no CPU fallback: the scan cannot exist without the GPU library
    delete s;
    return rc;
  }
  *out = (kvgh_scan*)s;
  return KVG_OK;
}
void kvgh_destroy(kvgh_scan* h) {
  Scan* s = (Scan*)h;
  if (!s) return;
  kvg_ctx_destroy(s->ctx);
  delete s;
}
const char* kvgh_last_error(kvgh_scan* h) { return h ? ((Scan*)h)->err.c_str() : kvg_last_error(nullptr); }
void kvgh_set_paths(kvgh_scan* h, const char* pci_ids, const char* base, const char* vgpu_base) {
  Scan* s = (Scan*)h;
  if (pci_ids) s->pci_ids_path = pci_ids;
  if (base) s->base_path = base;
  if (vgpu_base) s->vgpu_base_path = vgpu_base;
}

int kvgh_get_device_name(kvgh_scan* h, const char* id, size_t idlen, char* out, size_t cap, size_t* outlen) {
  Scan* s = (Scan*)h;
  int rc = ensure_table(s);
  if (rc) return rc;
  return kvg_name_lookup(s->ctx, id, idlen, out, cap, outlen);
}

// Snapshot only (exposed so the CPU tests can check it without a GPU): records in Walk order, names
// as a '\0'-separated blob, interned iommu-group strings likewise.  `device` contents: when every one that was
// read is "%04x" the records carry the number and the devices blob is empty; otherwise the whole column is in
// index mode (readIDFromFile returns string(data[2:]) whatever it is, :294-302): the records carry interned ids
// in first-appearance order and the blob holds the strings.  Free the blobs with free().
int kvgh_snapshot_pci(const char* base_path, kvg_pci_rec** recs_out, size_t* n_out, char** names_out,
                      size_t* names_len, char** groups_out, size_t* groups_len, char** devices_out,
                      size_t* devices_len) {
  std::string base = base_path;
  std::vector<kvg_pci_rec> recs;
  std::string names, groups;
  std::vector<std::string> dev_str;  // per record: what the `device` file held ("" = not read / unreadable)
  std::vector<bool> dev_read;
  std::unordered_map<std::string, uint32_t> gid;
  bool panic = false;
  walk(base, [&](const std::string& name, bool is_dir) {
    if (is_dir) return true;  // :197-200
    kvg_pci_rec r;
    memset(&r, 0, sizeof r);
    r.addr = (uint32_t)recs.size();  // index mode
    r.vendor = 0xffff;
    std::string vendor, driver, group, device;
    bool have_dev = false;
    int vrc = read_id(base, name, "vendor", &vendor);  // :202
    if (vrc == 2) {
      panic = true;
      return false;
    }
    uint16_t hv;
    if (vrc)
      r.flags |= KVG_PF_VENDOR_ERR;
    else if (hex4(vendor, &hv))
      r.vendor = hv;
    if (!vrc && vendor == "10de") {  // :209 — later files are only touched when the reference would
      if (!read_link(base, name, "driver", &driver)) {
        r.flags |= KVG_PF_DRIVER_ERR;  // :213-216
      } else {
        r.driver = driver == "vfio-pci" ? KVG_DRV_VFIO_PCI
                   : driver == "nvgrace_gpu_vfio_pci" ? KVG_DRV_NVGRACE : KVG_DRV_OTHER;
        if (r.driver != KVG_DRV_OTHER) {  // isSupportedVfioDriver :217-220
          if (!read_link(base, name, "iommu_group", &group)) {
            r.flags |= KVG_PF_IOMMU_ERR;  // :222-225
          } else {
            auto it = gid.find(group);
            if (it == gid.end()) {
              it = gid.emplace(group, (uint32_t)gid.size()).first;
              groups.append(group);
              groups.push_back('\0');
            }
            r.iommu_group = it->second;
            long long numa;
            if (!read_numa(base, name, &numa)) r.flags |= KVG_PF_NUMA_ERR;  // :226-230
            r.numa = (int16_t)std::max(-32768ll, std::min(32767ll, numa));
            int drc = read_id(base, name, "device", &device);  // :234
            if (drc == 2) {
              panic = true;
              return false;
            }
            if (drc) r.flags |= KVG_PF_DEVICE_ERR;
            have_dev = !drc;
          }
        }
      }
    }
    dev_read.push_back(have_dev);
    dev_str.push_back(have_dev ? device : std::string());
    recs.push_back(r);
    names.append(name);
    names.push_back('\0');
    return true;
  });
  if (panic) return KVGH_EPANIC;
  std::string devices;
  {
    bool numeric = true;
    uint16_t hd;
    for (size_t i = 0; i < recs.size(); i++)
      if (dev_read[i] && !hex4(dev_str[i], &hd)) numeric = false;
    std::unordered_map<std::string, uint32_t> did;
    for (size_t i = 0; i < recs.size(); i++) {
      if (!dev_read[i]) continue;
      if (numeric) {
        hex4(dev_str[i], &hd);
        recs[i].device = hd;
      } else {
        auto it = did.find(dev_str[i]);
        if (it == did.end()) {
          if (did.size() > 0xffff) return KVG_ERANGE;  // more than 65,536 distinct non-canonical strings
          it = did.emplace(dev_str[i], (uint32_t)did.size()).first;
          devices.append(dev_str[i]);
          devices.push_back('\0');
        }
        recs[i].device = (uint16_t)it->second;
      }
    }
  }
  if (devices_out) {
    *devices_out = (char*)malloc(devices.size() + 1);
    memcpy(*devices_out, devices.data(), devices.size());
    *devices_len = devices.size();
  }
  *n_out = recs.size();
  *recs_out = (kvg_pci_rec*)malloc(std::max<size_t>(1, recs.size()) * sizeof(kvg_pci_rec));
  memcpy(*recs_out, recs.data(), recs.size() * sizeof(kvg_pci_rec));
  *names_out = (char*)malloc(names.size() + 1);
  memcpy(*names_out, names.data(), names.size());
  *names_len = names.size();
  *groups_out = (char*)malloc(groups.size() + 1);
  memcpy(*groups_out, groups.data(), groups.size());
  *groups_len = groups.size();
  return KVG_OK;
}

static std::vector<std::string> split0(const char* blob, size_t len) {
  std::vector<std::string> v;
  size_t a = 0;
  for (size_t i = 0; i < len; i++)
    if (blob[i] == 0) {
      v.emplace_back(blob + a, i - a);
      a = i + 1;
    }
  return v;
}

int kvgh_create_iommu_device_map(kvgh_scan* h) {
  Scan* s = (Scan*)h;
  s->iommuMap.clear();  // :188-190
  s->deviceMap.clear();
  s->bdfToIommuMap.clear();
  int rc = ensure_table(s);
  if (rc) return rc;
  kvg_pci_rec* recs = nullptr;
  size_t n = 0, nl = 0, gl = 0, dl = 0;
  char *nb = nullptr, *gb = nullptr, *db = nullptr;
  rc = kvgh_snapshot_pci(s->base_path.c_str(), &recs, &n, &nb, &nl, &gb, &gl, &db, &dl);
  if (rc) return rc;
  std::vector<std::string> names = split0(nb, nl), groups = split0(gb, gl), devices = split0(db, dl);
  free(db);
  kvg_pci_result* res = nullptr;
  rc = kvg_scan_pci(s->ctx, recs, n, &res);
  free(recs);
  free(nb);
  free(gb);
  if (rc) {  // maps stay empty, like the reference after a failed walk (:193-196)
    s->err = std::string("kvg_scan_pci: ") + kvg_last_error(s->ctx);
    return rc;
  }
  auto dev = [&](uint32_t i) { return Dev{names[res->survivors[i].addr], (long long)res->survivors[i].numa}; };
  for (uint32_t k = 0; k < res->n_dev_keys; k++) {
    std::string key;
    if (devices.empty()) {
      char hex[8];
      snprintf(hex, sizeof hex, "%04x", res->dev_keys[k]);
      key = hex;
    } else {
      key = devices[res->dev_keys[k]];  // index mode: the id is a handle, the key is the string itself
    }
    auto& v = s->deviceMap[key];
    for (uint32_t j = res->dev_off[k]; j < res->dev_off[k + 1]; j++) v.push_back(dev(res->dev_perm[j]));  // :240
    if (devices.empty()) {
      s->deviceNames[key] = pool_name(res, res->dev_name_slot[k]);                                   // :124
    } else {  // getDeviceName with the exact bytes (prefix semantics and all)
      char name[512];
      size_t nlen = 0;
      int lrc = kvg_name_lookup(s->ctx, key.data(), key.size(), name, sizeof name, &nlen);
      s->deviceNames[key] = lrc == KVG_OK ? std::string(name, nlen) : std::string();
    }
  }
  for (uint32_t k = 0; k < res->n_groups; k++) {
    auto& v = s->iommuMap[groups[res->grp_keys[k]]];
    for (uint32_t j = res->grp_off[k]; j < res->grp_off[k + 1]; j++) v.push_back(dev(res->grp_perm[j]));  // :241-242
  }
  for (uint64_t i = 0; i < res->n_survivors; i++)
    s->bdfToIommuMap[names[res->survivors[i].addr]] = groups[res->survivors[i].iommu_group];  // :243
  kvg_result_free(res);
  return KVG_OK;
}

int kvgh_create_vgpu_id_map(kvgh_scan* h) {
  Scan* s = (Scan*)h;
  s->vGpuMap.clear();  // :256-257
  s->gpuVgpuMap.clear();
  int rc = ensure_table(s);
  if (rc) return rc;
  std::vector<kvg_mdev_rec> recs;
  std::vector<std::string> names, parents, raw_types;
  std::unordered_map<std::string, uint32_t> pid, tid;
  bool panic = false;
  walk(s->vgpu_base_path, [&](const std::string& name, bool is_dir) {
    if (is_dir) return true;
    kvg_mdev_rec r;
    memset(&r, 0, sizeof r);
    uint32_t idx = (uint32_t)recs.size();
    r.uuid[0] = (uint8_t)(idx >> 24);
    r.uuid[1] = (uint8_t)(idx >> 16);
    r.uuid[2] = (uint8_t)(idx >> 8);
    r.uuid[3] = (uint8_t)idx;
    std::string raw, parent;
    if (!read_file(s->vgpu_base_path + "/" + name + "/mdev_type/name", &raw)) {  // :269 (raw: GPU sanitises)
      r.flags |= KVG_MF_TYPE_ERR;
    } else {
      auto it = tid.find(raw);
      if (it == tid.end()) {
        it = tid.emplace(raw, (uint32_t)tid.size()).first;
        raw_types.push_back(raw);
      }
      r.type_idx = (uint16_t)it->second;
      int prc = read_gpu_id_for_vgpu(s->vgpu_base_path, name, &parent);  // :275
      if (prc == 2) {
        panic = true;
        return false;
      }
      if (prc) {
        r.flags |= KVG_MF_PARENT_ERR;
      } else {
        auto pt = pid.find(parent);
        if (pt == pid.end()) {
          pt = pid.emplace(parent, (uint32_t)pid.size()).first;
          parents.push_back(parent);
        }
        r.parent = pt->second;
        long long numa;
        if (!read_numa(s->base_path, parent, &numa)) r.flags |= KVG_MF_NUMA_ERR;  // :280-284
        r.parent_numa = (int16_t)std::max(-32768ll, std::min(32767ll, numa));
      }
    }
    recs.push_back(r);
    names.push_back(name);
    return true;
  });
  if (panic) return KVGH_EPANIC;
  std::vector<uint32_t> off(raw_types.size() + 1, 0);
  std::string blob;
  for (size_t i = 0; i < raw_types.size(); i++) {
    blob += raw_types[i];
    off[i + 1] = (uint32_t)blob.size();
  }
  blob.push_back('\0');
  kvg_type_dict dict{(uint32_t)raw_types.size(), off.data(), (const uint8_t*)blob.data()};
  kvg_mdev_result* res = nullptr;
  rc = kvg_scan_mdev(s->ctx, recs.data(), recs.size(), &dict, &res);
  if (rc) {
    s->err = std::string("kvg_scan_mdev: ") + kvg_last_error(s->ctx);
    return rc;
  }
  for (uint32_t k = 0; k < res->n_type_keys; k++) {
    uint32_t t = res->type_keys[k];
    std::string label((const char*)res->label_bytes + res->label_off[t], res->label_off[t + 1] - res->label_off[t]);
    auto& v = s->vGpuMap[label];
    for (uint32_t j = res->type_off[k]; j < res->type_off[k + 1]; j++) {
      const kvg_mdev_surv& sv = res->survivors[res->type_perm[j]];
      v.push_back(Dev{names[sv.src], (long long)sv.numa});  // :288
    }
    s->deviceNames[label] = std::string((const char*)res->type_name_bytes + res->type_name_off[t],
                                        res->type_name_off[t + 1] - res->type_name_off[t]);  // :152
  }
  for (uint32_t k = 0; k < res->n_parents; k++) {
    auto& v = s->gpuVgpuMap[parents[res->par_keys[k]]];
    for (uint32_t j = res->par_off[k]; j < res->par_off[k + 1]; j++)
      v.push_back(names[res->survivors[res->par_perm[j]].src]);  // :287
  }
  kvg_result_free(res);
  return KVG_OK;
}

// canonical dump, byte-identical to oracle kvo_dump (std::map iterates keys byte-wise ascending)
int kvgh_dump(kvgh_scan* h, char** out, size_t* outlen) {
  Scan* s = (Scan*)h;
  std::string b;
  auto dev_section = [&](char tag, const std::map<std::string, std::vector<Dev>>& m) {
    for (auto& kv : m) {
      auto it = s->deviceNames.find(kv.first);
      std::string name = it == s->deviceNames.end() ? "" : it->second;
      b += tag;
      b += ' ';
      b += kv.first + " " + (name.empty() ? "-" : name) + " nvidia.com/" + (name.empty() ? kv.first : name) + " " +
           std::to_string(kv.second.size()) + "\n";
      for (auto& d : kv.second) b += "  " + d.addr + " " + std::to_string(d.numa) + "\n";
    }
  };
  dev_section('D', s->deviceMap);
  for (auto& kv : s->iommuMap) {
    b += "I " + kv.first + " " + std::to_string(kv.second.size()) + "\n";
    for (auto& d : kv.second) b += "  " + d.addr + " " + std::to_string(d.numa) + "\n";
  }
  for (auto& kv : s->bdfToIommuMap) b += "B " + kv.first + " " + kv.second + "\n";
  dev_section('V', s->vGpuMap);
  for (auto& kv : s->gpuVgpuMap) {
    b += "G " + kv.first + " " + std::to_string(kv.second.size()) + "\n";
    for (auto& u : kv.second) b += "  " + u + "\n";
  }
  *out = (char*)malloc(b.size() + 1);
  memcpy(*out, b.data(), b.size());
  (*out)[b.size()] = 0;
  *outlen = b.size();
  return KVG_OK;
}

// payload half of createDevicePlugins (:99-157): one line per plugin the Go host would start
//   P <key> <deviceName> <resource> <socket> <envkey> <n>   then   "  <ID> <Health> <NUMA>"
int kvgh_device_plugins(kvgh_scan* h, char** out, size_t* outlen) {
  Scan* s = (Scan*)h;
  std::string b;
  auto upper = [](std::string x) {
    for (auto& c : x)
      if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return x;
  };
  auto emit = [&](const std::map<std::string, std::vector<Dev>>& m, const char* prefix) {
    for (auto& kv : m) {
      auto it = s->deviceNames.find(kv.first);
      std::string name = (it == s->deviceNames.end() || it->second.empty()) ? kv.first : it->second;  // :125-128
      b += "P " + kv.first + " " + name + " nvidia.com/" + name +                    // generic_device_plugin.go:299
           " /var/lib/kubelet/device-plugins/kubevirt-" + name + ".sock " +          // :87
           prefix + "_" + upper(name) + " " + std::to_string(kv.second.size()) + "\n";  // :420 / vgpu :223
      for (auto& d : kv.second) b += "  " + d.addr + " Healthy " + std::to_string(d.numa) + "\n";  // :111-123
    }
  };
  emit(s->deviceMap, "PCI_RESOURCE_NVIDIA_COM");
  emit(s->vGpuMap, "MDEV_PCI_RESOURCE_NVIDIA_COM");
  *out = (char*)malloc(b.size() + 1);
  memcpy(*out, b.data(), b.size());
  (*out)[b.size()] = 0;
  *outlen = b.size();
  return KVG_OK;
}

void kvgh_free(void* p) { free(p); }

// ---- consumer-side host logic, native twins of kvgpu/serve.py ------------
// Lists cross the boundary as '\n'-separated text; results are malloc'd (kvgh_free).
#define KVGH_EALLOC (-101) /* the reference returns an error for this request; *out holds its text */

static std::vector<std::string> split_lines(const char* text) {
  std::vector<std::string> v;
  if (!text) return v;
  std::string cur;
  for (const char* p = text; *p; p++) {
    if (*p == '\n') {
      v.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(*p);
    }
  }
  if (!cur.empty()) v.push_back(cur);
  return v;
}
static int give(const std::string& s, char** out, size_t* outlen) {
  char* b = (char*)malloc(s.size() + 1);
  if (!b) return KVG_ENOMEM;
  memcpy(b, s.c_str(), s.size() + 1);
  *out = b;
  if (outlen) *outlen = s.size();
  return KVG_OK;
}

// GetPreferredAllocation for one container request (generic_device_plugin.go:470-608).
//   devs          "id\tnuma" per line ("id" alone: a device without topology)
//   available / must_include   one id per line
//   out           the preferred ids, one per line — or the error text with KVGH_EALLOC
int kvgh_preferred_allocation(const char* devs, const char* available, const char* must_include, int allocation_size,
                              char** out, size_t* outlen) {
  if (!out) return KVG_EINVAL;
  std::map<std::string, long long> numa_of;
  for (const std::string& line : split_lines(devs)) {
    size_t t = line.find('\t');
    if (t != std::string::npos) numa_of[line.substr(0, t)] = atoll(line.c_str() + t + 1);
  }
  auto node = [&](const std::string& id) -> long long {
    auto it = numa_of.find(id);
    return it == numa_of.end() ? -1 : it->second;
  };
  const std::vector<std::string> avail = split_lines(available), must = split_lines(must_include);
  std::map<long long, std::vector<std::string>> by_node;
  std::vector<long long> node_order;
  for (const std::string& id : avail) {
    long long n = node(id);
    if (!by_node.count(n)) node_order.push_back(n);
    by_node[n].push_back(id);
  }
  std::vector<std::string> preferred;
  std::set<std::string> chosen;
  std::map<long long, int> selected;
  auto add = [&](const std::string& id) {
    if (!chosen.insert(id).second) return;
    selected[node(id)]++;
    preferred.push_back(id);
  };
  std::vector<long long> selected_order;
  for (const std::string& id : must) {
    if (chosen.count(id)) continue;
    add(id);
    long long n = node(id);
    if (std::find(selected_order.begin(), selected_order.end(), n) == selected_order.end()) selected_order.push_back(n);
  }
  if ((long long)preferred.size() > allocation_size) {
    char msg[128];
    snprintf(msg, sizeof msg, "number of MustIncludeDeviceIDs (%zu) exceeds allocation size (%d)", preferred.size(),
             allocation_size);
    int rc = give(msg, out, outlen);
    return rc ? rc : KVGH_EALLOC;
  }
  if ((long long)preferred.size() < allocation_size) {
    std::vector<long long> cand = selected_order;
    for (long long n : node_order)
      if (std::find(selected_order.begin(), selected_order.end(), n) == selected_order.end()) cand.push_back(n);
    long long target = -1;  // -1 doubles as "devices without topology": never a target (:552-575)
    for (long long n : cand) {
      int freecnt = 0;
      for (const std::string& id : by_node[n]) freecnt += !chosen.count(id);
      if (selected[n] + freecnt >= allocation_size) {
        target = n;
        break;
      }
    }
    if (target != -1)
      for (const std::string& id : by_node[target]) {
        if ((long long)preferred.size() >= allocation_size) break;
        add(id);
      }
  }
  for (const std::string& id : avail) {
    if ((long long)preferred.size() >= allocation_size) break;
    add(id);
  }
  std::string joined;
  for (const std::string& id : preferred) joined += id + "\n";
  return give(joined, out, outlen);
}

// egmPathsForAllocatedGPUs (:159-184).  egm: "devpath\tbdf bdf ..." per line; allocated: one BDF per line.
int kvgh_egm_paths_for_allocated(const char* allocated, const char* egm, char** out, size_t* outlen) {
  if (!out) return KVG_EINVAL;
  auto norm = [](std::string s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    s = s.substr(a, b - a);
    for (char& c : s) c = (char)tolower((unsigned char)c);
    return s;
  };
  std::set<std::string> have;
  for (const std::string& b : split_lines(allocated)) have.insert(norm(b));
  std::vector<std::string> paths;
  for (const std::string& line : split_lines(egm)) {
    size_t t = line.find('\t');
    if (t == std::string::npos) continue;
    bool all = true;
    std::string cur;
    const std::string gpus = line.substr(t + 1) + " ";
    for (char c : gpus) {
      if (c == ' ') {
        if (!cur.empty() && !have.count(norm(cur))) all = false;
        cur.clear();
      } else {
        cur.push_back(c);
      }
    }
    if (all) paths.push_back(line.substr(0, t));
  }
  std::sort(paths.begin(), paths.end());
  std::string joined;
  for (const std::string& p : paths) joined += p + "\n";
  return give(joined, out, outlen);
}

}  // extern "C"

#ifdef KVG_DISCOVER_MAIN
// kvg-discover: the scan half of InitiateDevicePlugin (:89-96) as a command — what the plugin would
// register with the kubelet on this host.
int main(int argc, char** argv) {
  const char* ids = "/usr/pci.ids";
  const char* base = "/sys/bus/pci/devices";
  const char* vbase = "/sys/bus/mdev/devices";
  bool dump = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--pci-ids") && i + 1 < argc) ids = argv[++i];
    else if (!strcmp(argv[i], "--sysfs-pci") && i + 1 < argc) base = argv[++i];
    else if (!strcmp(argv[i], "--sysfs-mdev") && i + 1 < argc) vbase = argv[++i];
    else if (!strcmp(argv[i], "--dump")) dump = true;
    else {
      fprintf(stderr, "usage: kvg-discover [--pci-ids F] [--sysfs-pci D] [--sysfs-mdev D] [--dump]\n");
      return 2;
    }
  }
  kvgh_scan* s = nullptr;
  int rc = kvgh_create(ids, base, vbase, 0, &s);
  if (rc) {
    fprintf(stderr, "kvg-discover: %s (no CPU fallback)\n", kvg_last_error(nullptr));
    return 1;
  }
  if ((rc = kvgh_create_iommu_device_map(s)) || (rc = kvgh_create_vgpu_id_map(s))) {
    fprintf(stderr, "kvg-discover: scan failed (%d): %s\n", rc, kvgh_last_error(s));
    return 1;
  }
  char* out = nullptr;
  size_t n = 0;
  if (dump) kvgh_dump(s, &out, &n);
  else kvgh_device_plugins(s, &out, &n);
  fwrite(out, 1, n, stdout);
  kvgh_free(out);
  kvgh_destroy(s);
  return 0;
}
#endif

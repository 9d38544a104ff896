/*
 * kvg_oracle.h — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the discovery-and-classification scan of
 * NVIDIA/kubevirt-gpu-device-plugin (pkg/device_plugin/device_plugin.go:187-438).  It exists so the
 * CUDA path in libkvgpu.so can be checked bit-for-bit.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product never does.
 *
 * Parity status: PINNED against the reference's own Ginkgo vectors
 * (pkg/device_plugin/device_plugin_test.go:153-426, lifted into tests/golden/ginkgo_vectors.json by
 * tests/golden/make_golden.py).  The Go reference itself cannot be built here (no Go toolchain,
 * no network), so there is no oracle/_ref binary.
 */
#ifndef KVG_ORACLE_H
#define KVG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/kvgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Go stdlib restatements used by the path (exposed for unit tests) ---------------------- */
/* strings.TrimSpace: returns the [start,end) range of s */
void kvo_trim_space(const uint8_t *s, size_t n, size_t *start, size_t *end);
/* strings.ToUpper (simple case mapping); out must hold 3*n bytes; returns the new length */
size_t kvo_to_upper(const uint8_t *s, size_t n, uint8_t *out);

/* ---- getDeviceName (device_plugin.go:371-438) ---------------------------------------------- */
/* returns the name length (0 == ""), or -1 if cap is too small */
long kvo_get_device_name(const uint8_t *text, size_t len, const uint8_t *key, size_t keylen,
                         uint8_t *out, size_t cap);
/* same, opening `path` first like the reference (:373-377): unreadable file -> "" */
long kvo_get_device_name_file(const char *path, const uint8_t *key, size_t keylen, uint8_t *out,
                              size_t cap);
/* device ids of the first "10de" section in file order (generator input); returns the count */
size_t kvo_nv_ids(const uint8_t *text, size_t len, uint16_t *out, size_t cap);

/* ---- sysfs readers (device_plugin.go:294-357). rc: >=0 length / 0 ok, -1 error, -2 = the Go
 *      code would panic (slice out of range) ------------------------------------------------- */
long kvo_read_id_from_file(const char *base, const char *addr, const char *prop, char *out,
                           size_t cap);
int kvo_read_numa_node(const char *base, const char *addr, int64_t *out);
long kvo_read_link(const char *base, const char *addr, const char *link, char *out, size_t cap);
long kvo_read_vgpu_id_from_file(const char *base, const char *addr, const char *prop, char *out,
                                size_t cap);
long kvo_read_gpu_id_for_vgpu(const char *base, const char *addr, char *out, size_t cap);
int kvo_is_supported_vfio_driver(const char *driver); /* :249-252 */

/* ---- the five maps (device_plugin.go:55-68) ------------------------------------------------- */
typedef struct kvo_maps kvo_maps;
kvo_maps *kvo_maps_new(void);
void kvo_maps_free(kvo_maps *m);

/* createIommuDeviceMap over a real directory tree (:187-247); rc 0, or -2 on a Go panic */
int kvo_create_iommu_device_map_tree(kvo_maps *m, const char *base_path);
/* createVgpuIDMap over a real tree (:255-291); pci_base is `basePath` used for numa (:280) */
int kvo_create_vgpu_id_map_tree(kvo_maps *m, const char *vgpu_base, const char *pci_base);
/* the same two functions fed from the flat snapshot: every record is turned back into the strings
 * the readers would have returned, then run through the identical per-entry logic */
int kvo_create_iommu_device_map_flat(kvo_maps *m, const kvg_pci_rec *recs, size_t n);
int kvo_create_vgpu_id_map_flat(kvo_maps *m, const kvg_mdev_rec *recs, size_t n,
                                const kvg_type_dict *types);

/* Canonical dump: keys sorted byte-wise, members in reference order.  Names are
 * resolved with kvo_get_device_name(pciids); the fallback name=="" -> key is applied
 * (device_plugin.go:125-128, :153-155).  *out is malloc'd; free with kvo_free. */
int kvo_dump(const kvo_maps *m, const uint8_t *pciids, size_t pciids_len, char **out,
             size_t *outlen);
void kvo_free(void *p);
void kvo_sha256(const void *data, size_t len, uint8_t digest[32]);

/* counts, for quick checks */
void kvo_maps_counts(const kvo_maps *m, uint64_t *n_dev_keys, uint64_t *n_groups, uint64_t *n_bdf,
                     uint64_t *n_types, uint64_t *n_parents);

/* ---- synthetic snapshots (CUDA twins live in csrc/kvg_kernels.cu) ------------ */
uint64_t kvo_mix(uint64_t x);
void kvo_gen_pci(kvg_pci_rec *out, uint64_t first, size_t n, const uint16_t *nv_ids,
                 uint32_t n_nv_ids, uint32_t group_bits);
void kvo_gen_mdev(kvg_mdev_rec *out, uint64_t first, size_t n);
/* raw type dictionary entry k: "NVIDIA SYN%02X-%d%c\n"; returns length */
size_t kvo_gen_type_name(uint32_t k, char *out, size_t cap);
void kvo_format_bdf(uint32_t packed, char out[16]);
void kvo_format_uuid(const uint8_t uuid[16], char out[40]);

/* ---- CPU baselines timed by bench.py ------------------------------------------------------- */
/* A: faithful cost, 1 thread: flat createIommuDeviceMap + one getDeviceName per distinct key.
 * B: best effort: parse pci.ids once, `threads` workers over record ranges, merged counts.
 * Both return seconds of wall time and write the survivor count. */
double kvo_bench_faithful(const kvg_pci_rec *recs, size_t n, const uint8_t *pciids, size_t len,
                          uint64_t *survivors, uint64_t *name_hits);
double kvo_bench_threads(const kvg_pci_rec *recs, size_t n, const uint8_t *pciids, size_t len,
                         int threads, uint64_t *survivors, uint64_t *name_hits);

#ifdef __cplusplus
}
#endif
#endif

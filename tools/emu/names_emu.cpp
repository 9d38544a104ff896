// names_emu.cpp — getDeviceName end to end from kernel source on the CPU: the DEFAULT parse
// (k_pciids_parse of csrc/kvg_parse.cuh, parser = 1) or the barrier-free one (csrc/kvg_parse_v2.cuh,
// parser = 2), followed by the name path of csrc/kvg_parse.cuh exactly as the library sequences it
// (kvg_api.cu: parse_enqueue / parse_enqueue_v2, table_publish, kvg_name_lookup, lookup_general):
//   hash path      k_nv_index -> k_pciids_sanitise_lines -> pool[slot] = u16 length + bytes
//   general path   k_section_lines -> k_lookup_general -> k_sanitise_matches   (non-canonical keys)
#define KVG_HOST_EMU 1
#include "warp_emu.h"
namespace kvg {
#include "emu_parse_all.inc"   // all of csrc/kvg_parse.cuh: the default parse (TMA ring) and the name path
}
#include "../../kubevirt-gpu-device-plugin_b200/csrc/kvg_parse_v2.cuh"
using namespace kvg;

static bool canonical_key(const uint8_t* k, uint32_t n, uint32_t* v) {  // kvg_api.cu: 4 lower-case hex digits
  if (n != 4) return false;
  uint32_t x = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t h = hexval(k[i]);
    if (h > 15) return false;
    x = (x << 4) | h;
  }
  *v = x;
  return true;
}

extern "C" {

// text: one image padded like kvg_text_pad.  keys: blob + n_keys+1 offsets.  names_out: n_keys x name_cap
// bytes, names_len: n_keys.  Returns 0.
int emu_get_device_names(int parser, const uint8_t* text, uint32_t len, uint32_t cap_log2, const uint8_t* keys,
                         const uint32_t* key_off, uint32_t n_keys, uint8_t* names_out, uint32_t name_cap, uint32_t* names_len,
                         uint32_t* info_out, uint32_t* nv_index_out, uint8_t* pool_out, uint32_t pool_cap, uint32_t* pool_len) {
  const uint32_t spf = (len + V2_SPAN - 1) / V2_SPAN, n_spans = spf;
  const size_t cap = (size_t)1 << cap_log2;
  std::vector<uint64_t> table(cap, P_EMPTY);
  PciIdsInfo info;
  memset(&info, 0, sizeof info);
  info.v_off = P_NONE;
  std::vector<uint32_t> arrays(3 * (size_t)n_spans + 1), state(2 * (size_t)n_spans + 1), pending((size_t)n_spans * V2_PEND_CAP + 1);
  ParseV2Args A;
  A.text = text;
  A.stride = 0;
  A.len = len;
  A.n_files = 1;
  A.spans_per_file = spf;
  A.n_spans = n_spans;
  A.tables = table.data();
  A.cap_mask = (uint32_t)cap - 1;
  A.cap_shift = 32 - cap_log2;
  A.info = &info;
  A.span_first_hdr = arrays.data();
  A.span_first_nl = arrays.data() + n_spans;
  A.span_last_nl = arrays.data() + 2 * (size_t)n_spans;
  A.span_state = state.data();
  A.pend_cnt = state.data() + n_spans;
  A.pending = pending.data();
  if (parser == 2) {
    const unsigned grid = (n_spans + V2_WARPS - 1) / V2_WARPS;
    if (n_spans) {
      emu_launch(k_pciids_scan_v2, dim3(grid), V2_WARPS * 32, A);
      emu_launch(k_pciids_resolve_v2, dim3(grid), V2_WARPS * 32, A);
    }
    ParseArgs F;
    memset(&F, 0, sizeof F);
    F.text = text;
    F.len = len;
    F.n_files = 1;
    F.tiles_per_file = spf;
    F.n_tiles = n_spans;
    F.info = &info;
    F.tile_first_hdr = A.span_first_hdr;
    F.tile_first_nl = A.span_first_nl;
    F.tile_last_nl = A.span_last_nl;
    emu_launch(k_pciids_finalize_v2, dim3(1), KVG_BLOCK, F);
  } else {
    // parse_enqueue: ONE persistent CTA walks every tile in order (any grid <= n_tiles is legal on the GPU;
    // a sequential emulation can only honour the look-back of a single CTA)
    const uint32_t tpf = (len + P_TILE - 1) / P_TILE;
    std::vector<uint32_t> tile_arrays(3 * (size_t)tpf + 1, 0xdeadbeefu);
    std::vector<uint64_t> tile_state(tpf + 1, 0);
    ParseArgs P;
    memset(&P, 0, sizeof P);
    P.text = text;
    P.stride = 0;
    P.len = len;
    P.n_files = 1;
    P.tiles_per_file = tpf;
    P.n_tiles = tpf;
    P.tables = table.data();
    P.cap_mask = (uint32_t)cap - 1;
    P.cap_shift = 32 - cap_log2;
    P.info = &info;
    P.tile_first_hdr = tile_arrays.data();
    P.tile_first_nl = tile_arrays.data() + tpf;
    P.tile_last_nl = tile_arrays.data() + 2 * (size_t)tpf;
    P.tile_state = tile_state.data();
    P.epoch = 5;
    if (tpf) emu_launch(k_pciids_parse, dim3(1), KVG_BLOCK, P);
    emu_launch(k_pciids_finalize, dim3(1), KVG_BLOCK, P);
  }
  memcpy(info_out, &info, sizeof info);
  if (info.overflow) return 1;
  // ---- table_publish: nv_index + named lines, sanitised pool, candidate lines of the general lookup
  std::vector<uint32_t> nv_index(65536), nv_lines(65536 + 8, 0);
  emu_launch(k_nv_index, dim3(256), 256, (const uint64_t*)table.data(), A.cap_mask, A.cap_shift, (const PciIdsInfo*)&info,
             nv_index.data(), nv_lines.data(), nv_lines.data() + 65536);
  const size_t sec = info.v_off == P_NONE ? 0 : (size_t)info.sec_end - info.v_off;
  std::vector<uint8_t> pool(sec + 16, 0);
  if (info.v_off != P_NONE)
    emu_launch(k_pciids_sanitise_lines, dim3(64), KVG_BLOCK, text, len, (const PciIdsInfo*)&info, (const uint32_t*)nv_lines.data(),
               (const uint32_t*)(nv_lines.data() + 65536), pool.data());
  const uint32_t sec_cap = (uint32_t)(sec / 2 + 8);
  std::vector<uint32_t> sec_lines(sec_cap + 1, 0);
  if (sec)
    emu_launch(k_section_lines, dim3(8), KVG_BLOCK, text, (const PciIdsInfo*)&info, sec_lines.data(), sec_lines.data() + sec_cap,
               sec_cap);
  if (nv_index_out) {   // table export for the end-to-end scan test: what the scans join against
    memcpy(nv_index_out, nv_index.data(), sizeof(uint32_t) * 65536);
    if (pool.size() > pool_cap) return 3;
    memcpy(pool_out, pool.data(), pool.size());
    *pool_len = (uint32_t)pool.size();
  }
  // ---- kvg_name_lookup per key
  for (uint32_t k = 0; k < n_keys; k++) {
    const uint8_t* key = keys + key_off[k];
    const uint32_t klen = key_off[k + 1] - key_off[k];
    uint8_t* out = names_out + (size_t)k * name_cap;
    names_len[k] = 0;
    uint32_t v;
    if (canonical_key(key, klen, &v)) {
      uint32_t slot = P_NONE;
      emu_launch(k_probe_keys, dim3(1), 32, (const uint64_t*)table.data(), A.cap_mask, A.cap_shift, (const PciIdsInfo*)&info, v, 1u,
                 &slot);
      if (slot == P_NONE) continue;
      const uint32_t n = pool[slot] | ((uint32_t)pool[slot + 1] << 8);
      if (n > name_cap) return 2;
      memcpy(out, &pool[slot + 2], n);
      names_len[k] = n;
      continue;
    }
    uint32_t match = P_NONE, off2[2] = {0, klen}, n = 0;
    if (sec)
      emu_launch(k_lookup_general, dim3(4, 1), KVG_BLOCK, text, len, (const uint32_t*)sec_lines.data(),
                 (const uint32_t*)(sec_lines.data() + sec_cap), key, (const uint32_t*)off2, &match);
    emu_launch(k_sanitise_matches, dim3(1), 64, text, len, (const uint32_t*)off2, (const uint32_t*)&match, 1u, out, name_cap, &n);
    if (n > name_cap) return 2;
    names_len[k] = n;
  }
  return 0;
}

}  // extern "C"

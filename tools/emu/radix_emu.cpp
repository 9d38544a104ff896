// radix_emu.cpp — the radix family of csrc/kvg_scan.cuh (k_radix_hist, k_radix_tilescan,
// k_radix_scatter<8|11>) and the experimental k_radix_tilescan_warp of csrc/kvg_radix_exp.cuh, compiled
// for the CPU from their real source on top of warp_emu.h.  The pass loop below is the one of
// enqueue_orderings (kvg_api.cu): passes 0..nsets-1, ping-pong between two pair buffers.
#define KVG_HOST_EMU 1
#include "warp_emu.h"
#include "kvgpu.h"
namespace kvg {
#include "emu_radix.inc"
}
#include "../../kubevirt-gpu-device-plugin_b200/csrc/kvg_radix_exp.cuh"

using namespace kvg;

extern "C" {

// Stable sort of {key, index} pairs the way the device does it.  pairs_io: n x {key, index}; on return
// the sorted pairs.  variant: 0 = default kernels, 1 = k_radix_tilescan_warp, 2 = k_radix_scatter_c (11-bit only),
// 3 = both experimental kernels.  Returns the pass count
// the device-side plan chose, or a negative number.
int emu_radix_sort(uint2* pairs_io, uint32_t n, uint32_t key_bits_max, uint32_t max_bits, int variant) {
  if (max_bits != 8 && max_bits != RADIX_MAX_BITS) return -1;
  uint32_t max_key = 0;
  for (uint32_t i = 0; i < n; i++) max_key = pairs_io[i].x > max_key ? pairs_io[i].x : max_key;
  const size_t T = n ? (n + C_TILE - 1) / C_TILE : 1;
  std::vector<uint2> p0(n + 1), p1(n + 1);
  memcpy(p1.data(), pairs_io, sizeof(uint2) * n);   // pass 0 reads p1 (SRC_PAIRS), like the owned-pairs path
  std::vector<uint32_t> tile_hist(RADIX_MAX_DIGITS * T, 0xdeadbeefu), bin_total(RADIX_MAX_DIGITS, 0xdeadbeefu);
  const int nsets = (int)((key_bits_max + max_bits - 1) / max_bits);
  for (int p = 0; p < nsets; p++) {
    RadixArgs a;
    a.n_ptr = &n;
    a.max_key = &max_key;
    a.src_records = nullptr;
    a.pairs_in = (p & 1) ? p0.data() : p1.data();
    a.pairs_out = (p & 1) ? p1.data() : p0.data();
    a.tile_hist = tile_hist.data();
    a.bin_total = bin_total.data();
    a.pass = (uint32_t)p;
    a.key_bits_max = key_bits_max;
    a.max_bits = max_bits;
    a.src = SRC_PAIRS;
    RadixArgs2 aa;
    aa.o[0] = a;
    aa.o[1] = a;
    emu_launch(k_radix_hist, dim3((unsigned)T, 1), KVG_BLOCK, aa);
    if (variant & 1)
      emu_launch(k_radix_tilescan_warp, dim3(RADIX_MAX_DIGITS / TS_WARPS, 1), TS_WARPS * 32, aa);
    else
      emu_launch(k_radix_tilescan, dim3(KVG_BLOCK, 1), KVG_BLOCK, aa);
    if (max_bits == 8)
      emu_launch(k_radix_scatter<8>, dim3((unsigned)T, 1), KVG_BLOCK, aa);
    else if (variant & 2)
      emu_launch(k_radix_scatter_c, dim3((unsigned)T, 1), KVG_BLOCK, aa);
    else
      emu_launch(k_radix_scatter<RADIX_MAX_BITS>, dim3((unsigned)T, 1), KVG_BLOCK, aa);
  }
  const uint32_t np = radix_plan(max_key, key_bits_max, 0, max_bits).npass;
  const std::vector<uint2>& fin = ((np - 1) & 1) ? p1 : p0;   // order_final_buf
  memcpy(pairs_io, fin.data(), sizeof(uint2) * n);
  return (int)np;
}

// One whole ordering the way enqueue_orderings runs it: radix passes, then k_order_final<false> ->
// k_tile_offsets -> k_order_final<true>.  surv: the survivor records the pairs index (head_name gathers
// surv[idx].w).  Outputs: perm[n], seg_key / seg_off / seg_name [n_seg (+1 for seg_off)]; returns n_seg.
int emu_ordering(uint2* pairs_io, uint32_t n, const uint4* surv, uint32_t key_bits_max, uint32_t max_bits, uint32_t* perm,
                 uint32_t* seg_key, uint32_t* seg_off, uint32_t* seg_name) {
  if (emu_radix_sort(pairs_io, n, key_bits_max, max_bits, 0) < 0) return -1;
  uint32_t max_key = 0;
  for (uint32_t i = 0; i < n; i++) max_key = pairs_io[i].x > max_key ? pairs_io[i].x : max_key;
  const size_t T = n ? (n + C_TILE - 1) / C_TILE : 1;
  // the device picks the final ping-pong buffer from the plan: put the sorted pairs where it will look
  std::vector<uint2> p0(n + 1), p1(n + 1);
  const uint32_t np = radix_plan(max_key, key_bits_max, 0, max_bits).npass;
  memcpy((((np - 1) & 1) ? p1 : p0).data(), pairs_io, sizeof(uint2) * n);
  std::vector<uint32_t> tile_heads(T + 1, 0xdeadbeefu), tile_off(T + 2, 0xdeadbeefu);
  std::vector<uint64_t> state(T + 2, 0);
  ScanCtrl ctrl;
  memset(&ctrl, 0, sizeof ctrl);
  ctrl.n_surv = n;
  ctrl.max_group = max_key;
  OrderFinalArgs a;
  a.p0 = p0.data();
  a.p1 = p1.data();
  a.max_key = &ctrl.max_group;
  a.key_bits_max = key_bits_max;
  a.max_bits = max_bits;
  a.n_ptr = &ctrl.n_surv;
  a.perm = perm;
  a.tile_heads = tile_heads.data();
  a.tile_off = tile_off.data();
  a.seg_key = seg_key;
  a.seg_off = seg_off;
  a.n_seg = &ctrl.n_groups;
  a.head_surv = surv;
  a.head_name = seg_name;
  OrderFinalArgs2 ff;
  ff.o[0] = a;
  ff.o[1] = a;
  TileOffsetsArgs2 tt;
  tt.o[0] = {tile_heads.data(), nullptr, &ctrl.n_surv, 0, tile_off.data(), &ctrl.n_groups, state.data()};
  tt.o[1] = tt.o[0];
  emu_launch(k_order_final<false>, dim3((unsigned)T, 1), KVG_BLOCK, ff);
  emu_launch(k_tile_offsets, dim3((unsigned)((T + C_TILE - 1) / C_TILE), 1), KVG_BLOCK, tt, &ctrl, 9u);
  emu_launch(k_order_final<true>, dim3((unsigned)T, 1), KVG_BLOCK, ff);
  return (int)ctrl.n_groups;
}

}  // extern "C"

// classify_emu.cpp — the record-classification pipeline of csrc/kvg_scan.cuh compiled for the CPU from
// its real source on top of warp_emu.h: the split form the large inputs and the pipelined host entry
// point use (k_classify_ragged -> k_tile_offsets -> k_pack_survivors) and the one-launch look-back form
// used below 2 M records (k_classify_oneshot).  Launch shapes are those of enqueue_classify (kvg_api.cu).
#define KVG_HOST_EMU 1
#include "warp_emu.h"
#include "kvgpu.h"
namespace kvg {
#include "emu_order.inc"
}
#include "../../kubevirt-gpu-device-plugin_b200/csrc/kvg_order.cuh"   // tile constants
namespace kvg {
#include "emu_classify.inc"
}
using namespace kvg;

extern "C" {

// recs: n x 16 B records; nv_index: 65536 name slots; surv_out: room for n survivors.
// variant 0 = ragged / offsets / pack, 1 = oneshot.  ctrl_out: {n_surv, max_group, max_devkey}.
int emu_classify_pci(const uint4* recs, uint32_t n, const uint32_t* nv_index, int variant, uint4* surv_out,
                     uint32_t* ctrl_out) {
  constexpr int T = 128, R = 8;
  const size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
  ScanCtrl ctrl;
  memset(&ctrl, 0, sizeof ctrl);
  std::vector<uint4> ragged((tiles + 1) * T * R, uint4{0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu});
  std::vector<uint32_t> tile_count(tiles + 2, 0xdeadbeefu), tile_off(tiles + 3, 0xdeadbeefu);
  std::vector<uint2> tile_max(tiles + 2);
  std::vector<uint64_t> state(tiles + 4, 0);
  PciClassifyOp op;
  op.recs = recs;
  op.n = n;
  op.ctrl = &ctrl;
  op.nv_index = nv_index;
  op.local_max_group = 0;
  op.local_max_dev = 0;
  const uint32_t epoch = 7;
  if (variant == 1) {
    op.out = (kvg_pci_surv*)surv_out;
    emu_launch(k_classify_oneshot<PciClassifyOp, T, R>, dim3((unsigned)(tiles ? tiles : 1)), T, op, state.data(), epoch);
  } else if (tiles) {
    op.out = (kvg_pci_surv*)ragged.data();
    emu_launch(k_classify_ragged<PciClassifyOp, T, R>, dim3((unsigned)tiles), T, op, tile_count.data(), tile_max.data());
    TileOffsetsArgs2 tt;
    tt.o[0] = {tile_count.data(), tile_max.data(), nullptr, (uint32_t)tiles, tile_off.data(), &ctrl.n_surv, state.data()};
    tt.o[1] = tt.o[0];
    emu_launch(k_tile_offsets, dim3((unsigned)((tiles + C_TILE - 1) / C_TILE)), KVG_BLOCK, tt, &ctrl, epoch);
    emu_launch(k_pack_survivors<1>, dim3((unsigned)tiles), 128, (const uint4*)ragged.data(),
               (const uint32_t*)tile_off.data(), (uint32_t)(T * R), surv_out);
  }
  ctrl_out[0] = ctrl.n_surv;
  ctrl_out[1] = ctrl.max_group;
  ctrl_out[2] = ctrl.max_devkey;
  return 0;
}

// K6 (kvg_health_rescan): alive-set diff against the previous scan, transitions in record order.
// alive_prev: n bytes, updated in place.  changed_out: room for n words.  ctrl_out: {n_changed, n_alive}.
// The grid is one CTA per tile — what compact_grid() picks whenever the tiles fit the GPU, and the only
// shape a sequential emulation of a look-back kernel can run.
int emu_health_rescan(const uint4* recs, uint32_t n, uint8_t* alive_prev, uint32_t* changed_out, uint32_t* ctrl_out) {
  const size_t tiles = (n + C_TILE - 1) / C_TILE;
  ScanCtrl ctrl;
  memset(&ctrl, 0, sizeof ctrl);
  std::vector<uint64_t> state(tiles + 4, 0);
  HealthOp op;
  op.recs = recs;
  op.n = n;
  op.alive_prev = alive_prev;
  op.changed = changed_out;
  op.ctrl = &ctrl;
  op.local_alive = 0;
  emu_launch(k_compact<HealthOp>, dim3((unsigned)(tiles ? tiles : 1)), KVG_BLOCK, op, state.data(), 11u);
  ctrl_out[0] = ctrl.n_changed;
  ctrl_out[1] = ctrl.n_alive;
  return 0;
}

// K6 at poll-loop sizes (kvg_health_rescan, n <= 32,768): one CTA, transitions + counters written where the
// host reads them.  hdr_out: {n_alive, n_changed}.
int emu_health_small(const uint4* recs, uint32_t n, uint8_t* alive_prev, uint32_t* changed_out, uint32_t* hdr_out) {
  if (n == 0 || n > HEALTH_SMALL_MAX) return -1;
  emu_launch(k_health_small, dim3(1), HEALTH_SMALL_THREADS, recs, n, alive_prev, changed_out, hdr_out, 7u);
  return 0;
}

// K5 (kvg_dev_scan_mdev up to the survivor list): type dictionary -> labels -> canonical ids, then the
// 32-byte mdev records through k_classify_ragged<MdevClassifyOp,128,4> -> k_tile_offsets -> k_pack_survivors<2>.
// raw / raw_off: the dictionary blob with n_types+1 offsets.  Outputs: label bytes per entry (at raw_off),
// label_len, canon; surv_out: 2 x uint4 per survivor; ctrl_out: {n_surv, max_parent, max_type}.
int emu_scan_mdev(const uint4* recs, uint32_t n, const uint8_t* raw, const uint32_t* raw_off, uint32_t n_types, uint8_t* label,
                  uint32_t* label_len, uint16_t* canon, uint4* surv_out, uint32_t* ctrl_out) {
  std::vector<uint64_t> label_hash(n_types + 1);
  if (n_types) {
    emu_launch(k_mdev_labels, dim3((n_types + 127) / 128), 128, raw, raw_off, n_types, label, label_len, label_hash.data());
    emu_launch(k_mdev_canon, dim3((n_types + 127) / 128), 128, (const uint8_t*)label, raw_off, (const uint32_t*)label_len,
               (const uint64_t*)label_hash.data(), n_types, canon);
  }
  constexpr int T = 128, R = 4;
  const size_t tiles = (n + (size_t)T * R - 1) / ((size_t)T * R);
  ScanCtrl ctrl;
  memset(&ctrl, 0, sizeof ctrl);
  std::vector<uint4> ragged(2 * (tiles + 1) * T * R);
  std::vector<uint32_t> tile_count(tiles + 2), tile_off(tiles + 3);
  std::vector<uint2> tile_max(tiles + 2);
  std::vector<uint64_t> state(tiles + 4, 0);
  MdevClassifyOp op;
  op.recs = recs;
  op.n = n;
  op.out = ragged.data();
  op.ctrl = &ctrl;
  op.type_canon = canon;
  op.n_types = n_types;
  op.local_max_parent = 0;
  op.local_max_type = 0;
  if (tiles) {
    emu_launch(k_classify_ragged<MdevClassifyOp, T, R>, dim3((unsigned)tiles), T, op, tile_count.data(), tile_max.data());
    TileOffsetsArgs2 tt;
    tt.o[0] = {tile_count.data(), tile_max.data(), nullptr, (uint32_t)tiles, tile_off.data(), &ctrl.n_surv, state.data()};
    tt.o[1] = tt.o[0];
    emu_launch(k_tile_offsets, dim3((unsigned)((tiles + C_TILE - 1) / C_TILE)), KVG_BLOCK, tt, &ctrl, 13u);
    emu_launch(k_pack_survivors<2>, dim3((unsigned)tiles), 128, (const uint4*)ragged.data(),
               (const uint32_t*)tile_off.data(), (uint32_t)(T * R), surv_out);
  }
  ctrl_out[0] = ctrl.n_surv;
  ctrl_out[1] = ctrl.max_group;
  ctrl_out[2] = ctrl.max_devkey;
  return 0;
}

}  // extern "C"

// kvg_scan.cuh — record classification and stable compaction.
//
//   k_classify_ragged<Op>  K3/K5 at bandwidth-bound sizes: every CTA classifies one tile and writes its
//                          survivors at a TILE-LOCAL base (no cross-tile dependency); k_tile_offsets scans
//                          the tile counts, k_pack_survivors makes the list dense
//   k_classify_oneshot<Op> K3 at latency-bound sizes: one launch, decoupled look-back on the tile counts
//        PciClassifyOp       createIommuDeviceMap's filter (device_plugin.go:201-244) + name join
//        MdevClassifyOp      createVgpuIDMap's filter (:268-289)
//   k_compact<HealthOp>    K6: alive-set diff against the previous scan
//   k_mdev_labels / _canon K5: label rule (:341-342) + merge of equal labels
//   k_gen_*                counter-based synthetic snapshots (twins of oracle/kvg_oracle.c kvo_gen_*)
#pragma once
#include "../../include/kvgpu.h"
#include "kvg_common.cuh"
#include "kvg_parse.cuh"
#include "kvg_order.cuh"

namespace kvg {

// device-resident control block of one scan (zeroed by one memset per step)
struct ScanCtrl {
  uint32_t n_own[2];      // sharded scans: records in the owned list of ordering 0 / 1
  uint32_t own_max[2];    // sharded scans: their largest keys (radix plan)
  uint32_t n_gathered;    // sharded scans, NCCL mode: length of the all-gathered survivor list
  uint32_t reserved[3];
  uint32_t n_surv;        // survivors of the classify kernel
  uint32_t max_group;     // max iommu group / parent among survivors (radix pass count)
  uint32_t max_devkey;    // max device / type key among survivors
  uint32_t n_dev_keys;    // distinct keys found by the heads kernels
  uint32_t n_groups;
  uint32_t n_alive;       // health
  uint32_t n_changed;
  uint32_t pad0;
  uint32_t reserved2[16];
};

// ------------------------------------------------------------------------------------------------
// generic stable compaction
//   Op::Item                        what a thread holds per input element
//   uint32_t count()                number of input items (may read device memory)
//   Item load(uint32_t i, bool ok)  ok == false -> any value that fails pred
//   bool pred(const Item&, i)
//   void emit(uint32_t pos, const Item&, i)
//   void warp_epilogue(...)         optional per-warp reduction hook (called once per tile)
//   void finish(uint32_t total)     called by one thread of the last tile
// ------------------------------------------------------------------------------------------------
template <class Op>
__global__ void __launch_bounds__(KVG_BLOCK) k_compact(Op op, uint64_t* tile_state, uint32_t epoch) {
  pdl_enter();
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_wtot[KVG_WARPS], s_woff[KVG_WARPS];
  op.begin();
  const uint32_t n = op.count();
  const uint32_t n_tiles = (n + C_TILE - 1) / C_TILE;
  const uint32_t lane = lane_id(), warp = warp_id();
  if (n_tiles == 0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) op.finish(0);
    return;
  }
  // Persistent, co-resident grid; tile = blockIdx + k*gridDim.  A look-back predecessor is owned by
  // a resident CTA that reaches it no later than this CTA reaches its own tile (no ticket needed).
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t base = tile * C_TILE + warp * C_WARP_ITEMS;
    typename Op::Item item[C_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
      uint32_t i = base + k * 32 + lane;
      item[k] = op.load(i, i < n);
    }
    uint32_t bal[C_ROWS], aux[C_ROWS];
    uint32_t wtot = 0;
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
      uint32_t i = base + k * 32 + lane;
      bool p = i < n && op.pred(item[k], i);
      bal[k] = __ballot_sync(KVG_FULL, p);
      wtot += __popc(bal[k]);
      aux[k] = p ? op.prepare(item[k]) : 0u;  // dependent loads overlap the look-back below
    }
    if (lane == 0) s_wtot[warp] = wtot;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < KVG_WARPS ? s_wtot[lane] : 0;
      uint32_t wi = warp_incl_sum(w);
      if (lane < KVG_WARPS) s_woff[lane] = wi - w;
      uint32_t tile_total = __shfl_sync(KVG_FULL, wi, KVG_WARPS - 1);
      uint32_t excl = lookback_sum(tile_state, tile, tile_total, epoch);
      if (lane == 0) {
        s_base = excl;
        if (tile == n_tiles - 1) op.finish(excl + tile_total);
      }
    }
    __syncthreads();
    uint32_t off = s_base + s_woff[warp];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
      uint32_t i = base + k * 32 + lane;
      if ((bal[k] >> lane) & 1u) op.emit(off + __popc(bal[k] & lanemask_lt()), item[k], i, aux[k]);
      off += __popc(bal[k]);
    }
    op.tile_epilogue();
  }
}

// ------------------------------------------------------------------------------------------------
// K3/K5, one-tile-per-CTA form: small CTAs (THREADS x ROWS records kept in registers), thousands
// of them, dispatched in blockIdx order by the hardware.  Many resident CTAs per SM hide the
// count -> look-back -> write-out latency chain of each other; predecessors were dispatched
// earlier, so the classic decoupled look-back usually finds an inclusive prefix close by.
// (Measured alternative, round 2: every tile publishes its count and sums ALL earlier counts itself, a thread
// per earlier tile — no chain, but 8 dependent L2 round trips per thread for the last tiles: 20.1 vs 14.9 us,
// config-2 step 0.108 vs 0.097 ms.  The chained scan stays.)
// ------------------------------------------------------------------------------------------------
template <class Op, int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) k_classify_oneshot(Op op, uint64_t* tile_state, uint32_t epoch) {
  pdl_enter();
  constexpr uint32_t TILE = THREADS * ROWS;
  constexpr uint32_t NW = THREADS / 32;
  constexpr uint32_t WARP_ITEMS = 32 * ROWS;
  __shared__ uint32_t s_wtot[NW], s_woff[NW];
  __shared__ uint32_t s_base;
  op.begin();
  const uint32_t n = op.count();
  const uint32_t n_tiles = (n + TILE - 1) / TILE;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x;
  if (n_tiles == 0) {
    if (tile == 0 && threadIdx.x == 0) op.finish(0);
    return;
  }
  if (tile >= n_tiles) return;
  const uint32_t base = tile * TILE + warp * WARP_ITEMS;
  typename Op::Item item[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    item[k] = op.load(i, i < n);
  }
  uint32_t bal[ROWS], aux[ROWS];
  uint32_t wtot = 0;
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    bool p = i < n && op.pred(item[k], i);
    bal[k] = __ballot_sync(KVG_FULL, p);
    wtot += __popc(bal[k]);
    aux[k] = p ? op.prepare(item[k]) : 0u;
  }
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NW ? s_wtot[lane] : 0;
    uint32_t wi = warp_incl_sum(w);
    if (lane < NW) s_woff[lane] = wi - w;
    uint32_t tile_total = __shfl_sync(KVG_FULL, wi, NW - 1);
    uint32_t excl = lookback_sum(tile_state, tile, tile_total, epoch);
    if (lane == 0) {
      s_base = excl;
      if (tile == n_tiles - 1) op.finish(excl + tile_total);
    }
  }
  __syncthreads();
  uint32_t off = s_base + s_woff[warp];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    if ((bal[k] >> lane) & 1u) op.emit(off + __popc(bal[k] & lanemask_lt()), item[k], i, aux[k]);
    off += __popc(bal[k]);
  }
  op.tile_epilogue();
}

// ---- K3: PCI classify ---------------------------------------------------------------------------
// record = {addr, vendor | device<<16, iommu_group, driver | flags<<8 | numa<<16}
// device_plugin.go:203-238: any of the vendor/driver/iommu/device read errors drops the entry,
// vendor must be "10de" (:209), driver must be in supportedVfioDrivers (:217, :75-78)
__device__ __forceinline__ bool pci_record_alive(const uint4& r) {
  // low 12 bits of r.w = driver | (drop flags << 8): alive iff they equal 1 or 2 exactly
  static_assert(KVG_DRV_VFIO_PCI == 1 && KVG_DRV_NVGRACE == 2, "driver codes");
  static_assert((KVG_PF_VENDOR_ERR | KVG_PF_DRIVER_ERR | KVG_PF_IOMMU_ERR | KVG_PF_DEVICE_ERR) == 0xf, "flags");
  return (r.y & 0xffffu) == 0x10deu && ((r.w & 0x0fffu) - 1u) < 2u;
}
struct PciClassifyOp {
  using Item = uint4;
  static constexpr uint32_t REC_BYTES = 16;
  const uint4* recs;
  uint32_t n;
  kvg_pci_surv* out;
  ScanCtrl* ctrl;
  const uint32_t* nv_index;                 // device id -> name pool slot (K1's k_pciids_names)
  uint32_t local_max_group, local_max_dev;  // per-thread running maxima (registers)

  __device__ __forceinline__ void begin() {}
  // the name join: one load per survivor from the flattened table of vendor 10de
  // (nv_index == NULL: the parse is still running on its own stream; the final ordering kernel joins the names)
  __device__ __forceinline__ uint32_t prepare(const Item& r) const { return nv_index ? __ldg(&nv_index[r.y >> 16]) : P_NONE; }
  __device__ __forceinline__ uint32_t count() const { return n; }
  __device__ __forceinline__ Item load(uint32_t i, bool ok) const {
    return ok ? ld_stream(recs + i) : make_uint4(0, 0, 0, 0xff00u);
  }
  // record = {addr, vendor | device<<16, iommu_group, driver | flags<<8 | numa<<16}
  // device_plugin.go:203-238: any of vendor/driver/iommu/device read errors drops the entry,
  // vendor must be "10de" (:209), driver in supportedVfioDrivers (:217, :75-78)
  __device__ __forceinline__ bool pred(const Item& r, uint32_t) const { return pci_record_alive(r); }
  static constexpr int UNITS = 1;  // 16-byte units per survivor
  // the survivor record (kvgpu.h kvg_pci_surv)
  __device__ __forceinline__ void make(const Item& r, uint32_t, uint32_t name_slot, uint4* s) const {
    uint32_t device = r.y >> 16;
    uint32_t flags = (r.w >> 8) & 0xffu;
    int32_t numa = (int32_t)r.w >> 16;  // sign-extended int16
    if ((flags & KVG_PF_NUMA_ERR) || numa < 0) numa = 0;  // :227-230, :316-318
    s[0].x = r.x;
    s[0].y = r.z;
    s[0].z = device | ((uint32_t)numa << 16);
    s[0].w = name_slot;
  }
  // keys of the two group-by maps: {device id (deviceMap), iommu group (iommuMap)}
  __device__ __forceinline__ uint2 keys(const Item& r, uint32_t) const { return make_uint2(r.y >> 16, r.z); }
  __device__ __forceinline__ void emit(uint32_t pos, const Item& r, uint32_t i, uint32_t name_slot) {
    uint4 s[1];
    make(r, i, name_slot, s);
    st_stream(reinterpret_cast<uint4*>(out) + pos, s[0]);
    local_max_group = max(local_max_group, r.z);
    local_max_dev = max(local_max_dev, r.y >> 16);
  }
  __device__ __forceinline__ void tile_epilogue() {
    uint32_t g = warp_max(local_max_group), d = warp_max(local_max_dev);
    if (lane_id() == 0) {
      if (g) atomicMax(&ctrl->max_group, g);
      if (d) atomicMax(&ctrl->max_devkey, d);
    }
    local_max_group = 0;
    local_max_dev = 0;
  }
  __device__ __forceinline__ void finish(uint32_t total) { ctrl->n_surv = total; }
  __device__ __forceinline__ uint2 take_maxima() {
    uint2 m = make_uint2(local_max_group, local_max_dev);
    local_max_group = local_max_dev = 0;
    return m;
  }
};
// the deferred name join of a dense PCI survivor list no ordering walks (the rank's own shard of a sharded scan):
// runs on the side stream behind the parse, beside the exchange and the orderings
__global__ void __launch_bounds__(KVG_BLOCK) k_join_names(uint4* __restrict__ recs, const uint32_t* __restrict__ n_ptr,
                                                          const uint32_t* __restrict__ nv_index) {
  pdl_enter();
  const uint32_t n = *n_ptr;
  for (uint32_t i = blockIdx.x * KVG_BLOCK + threadIdx.x; i < n; i += gridDim.x * KVG_BLOCK) {
    uint32_t* w = reinterpret_cast<uint32_t*>(recs + i);
    w[3] = __ldg(&nv_index[w[2] & 0xffffu]);
  }
}

// ---- K5: mdev classify --------------------------------------------------------------------------
struct MdevItem {
  uint4 lo, hi;
};
struct MdevClassifyOp {
  using Item = MdevItem;
  static constexpr uint32_t REC_BYTES = 32;
  const uint4* recs;  // 2 x uint4 per record
  uint32_t n;
  uint4* out;
  ScanCtrl* ctrl;
  const uint16_t* type_canon;  // [n_types] canonical id per raw dictionary entry
  uint32_t n_types;
  uint32_t local_max_parent, local_max_type;

  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ uint32_t prepare(const Item& r) const { return type_canon[r.hi.y & 0xffffu]; }
  __device__ __forceinline__ uint32_t count() const { return n; }
  __device__ __forceinline__ Item load(uint32_t i, bool ok) const {
    Item it;
    if (ok) {
      it.lo = ld_stream(recs + 2 * (size_t)i);
      it.hi = ld_stream(recs + 2 * (size_t)i + 1);
    } else {
      it.lo = make_uint4(0, 0, 0, 0);
      it.hi = make_uint4(0, 0xffu << 16, 0, 0);
    }
    return it;
  }
  // hi = {parent, type_idx | flags<<16 | pad<<24, parent_numa | pad.., pad}
  __device__ __forceinline__ bool pred(const Item& r, uint32_t) const {
    uint32_t flags = (r.hi.y >> 16) & 0xffu;
    uint32_t type_idx = r.hi.y & 0xffffu;
    return (flags & (KVG_MF_TYPE_ERR | KVG_MF_PARENT_ERR)) == 0 && type_idx < n_types;
  }
  static constexpr int UNITS = 2;  // 16-byte units per survivor
  // the survivor record (kvgpu.h kvg_mdev_surv)
  __device__ __forceinline__ void make(const Item& r, uint32_t i, uint32_t canon, uint4* s) const {
    uint32_t flags = (r.hi.y >> 16) & 0xffu;
    int32_t numa = (int32_t)(int16_t)(r.hi.z & 0xffffu);
    if ((flags & KVG_MF_NUMA_ERR) || numa < 0) numa = 0;  // :281-284, :316-318
    s[0] = r.lo;
    s[1].x = r.hi.x;
    s[1].y = canon | ((uint32_t)numa << 16);
    s[1].z = i;
    s[1].w = 0;
  }
  // keys of the two group-by maps: {canonical type (vGpuMap), parent GPU (gpuVgpuMap)}
  __device__ __forceinline__ uint2 keys(const Item& r, uint32_t canon) const { return make_uint2(canon, r.hi.x); }
  __device__ __forceinline__ void emit(uint32_t pos, const Item& r, uint32_t i, uint32_t canon) {
    uint4 s[2];
    make(r, i, canon, s);
    st_stream(out + 2 * (size_t)pos, s[0]);
    st_stream(out + 2 * (size_t)pos + 1, s[1]);
    local_max_parent = max(local_max_parent, r.hi.x);
    local_max_type = max(local_max_type, canon);
  }
  __device__ __forceinline__ void tile_epilogue() {
    uint32_t g = warp_max(local_max_parent), d = warp_max(local_max_type);
    if (lane_id() == 0) {
      if (g) atomicMax(&ctrl->max_group, g);
      if (d) atomicMax(&ctrl->max_devkey, d);
    }
    local_max_parent = 0;
    local_max_type = 0;
  }
  __device__ __forceinline__ void finish(uint32_t total) { ctrl->n_surv = total; }
  __device__ __forceinline__ uint2 take_maxima() {
    uint2 m = make_uint2(local_max_parent, local_max_type);
    local_max_parent = local_max_type = 0;
    return m;
  }
};

// ---- K6: health diff ----------------------------------------------------------------------------
struct HealthOp {
  using Item = uint4;
  const uint4* recs;
  uint32_t n;
  uint8_t* alive_prev;  // one byte per record, updated in place
  uint32_t* changed;
  ScanCtrl* ctrl;
  uint32_t local_alive;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ uint32_t prepare(const Item&) const { return 0; }
  __device__ __forceinline__ uint32_t count() const { return n; }
  __device__ __forceinline__ Item load(uint32_t i, bool ok) const {
    if (!ok) return make_uint4(0, 0, 0, 0);
    uint4 r = ld_stream(recs + i);
    uint32_t alive = pci_record_alive(r) ? 1u : 0u;
    r.x = alive | ((uint32_t)alive_prev[i] << 1);
    return r;
  }
  __device__ __forceinline__ bool pred(const Item& r, uint32_t) {
    local_alive += r.x & 1u;
    return (r.x & 1u) != (r.x >> 1);
  }
  __device__ __forceinline__ void emit(uint32_t pos, const Item& r, uint32_t i, uint32_t) {
    changed[pos] = (i << 1) | (r.x & 1u);
    alive_prev[i] = (uint8_t)(r.x & 1u);
  }
  __device__ __forceinline__ void tile_epilogue() {
    uint32_t a = warp_sum(local_alive);
    if (lane_id() == 0 && a) atomicAdd(&ctrl->n_alive, a);
    local_alive = 0;
  }
  __device__ __forceinline__ void finish(uint32_t total) { ctrl->n_changed = total; }
};

// K6 at poll-loop sizes (BASELINE.json config 5: 10,000 devices at 1 kHz): ONE CTA, one launch, one host
// synchronisation.  The records are read where the host left them (mapped pinned memory: zero-copy over PCIe,
// every load of a thread in flight at once), the transitions are written — in record order — straight into
// the host-visible result block, and the two counters follow.  No staging copy, no look-back, no second
// device-to-host copy.
constexpr uint32_t HEALTH_SMALL_THREADS = 1024;
constexpr uint32_t HEALTH_SMALL_ROWS = 32;                                        // rows of 1024 records
constexpr uint32_t HEALTH_SMALL_MAX = HEALTH_SMALL_THREADS * HEALTH_SMALL_ROWS;  // 32,768 records
constexpr uint32_t HEALTH_STAGE_ROWS = 12;                                        // rows staged per round
constexpr uint32_t HEALTH_SMALL_SMEM = HEALTH_STAGE_ROWS * HEALTH_SMALL_THREADS * 16;  // 192 KiB
__global__ void __launch_bounds__(HEALTH_SMALL_THREADS) k_health_small(const uint4* __restrict__ recs, uint32_t n,
                                                                       uint8_t* __restrict__ alive_prev,
                                                                       uint32_t* __restrict__ changed_host,
                                                                       uint32_t* __restrict__ hdr_host, uint32_t seq) {
  pdl_enter();
  constexpr uint32_t NW = HEALTH_SMALL_THREADS / 32;
#ifndef KVG_HOST_EMU
  extern __shared__ __align__(128) uint8_t hs_smem[];
#else
  static __attribute__((aligned(128))) uint8_t hs_smem[HEALTH_SMALL_SMEM];
#endif
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_bal[HEALTH_SMALL_ROWS][NW];  // "changed" ballot of (row, warp) -> its position in the list
  __shared__ uint32_t s_now[HEALTH_SMALL_ROWS][NW];  // "alive now" ballot of (row, warp)
  __shared__ uint32_t s_scan[NW], s_alv[NW];
  const uint4* stage = reinterpret_cast<const uint4*>(hs_smem);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rows = (n + HEALTH_SMALL_THREADS - 1) / HEALTH_SMALL_THREADS;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  uint32_t n_alive = 0, phase = 0;
  for (uint32_t r0 = 0; r0 < rows; r0 += HEALTH_STAGE_ROWS) {
    // the whole round (<= 192 KiB of the snapshot) is requested at once by the TMA unit: over PCIe what counts
    // is bytes in flight, and a bulk copy keeps them in flight without a register per load
    const uint32_t first = r0 * HEALTH_SMALL_THREADS;
    const uint32_t cnt = min(n - first, HEALTH_STAGE_ROWS * HEALTH_SMALL_THREADS);
    if (tid == 0) {
      mbar_arrive_expect_tx(&s_bar, cnt * 16);
      for (uint32_t off = 0; off < cnt * 16; off += 16384)
        tma_load_1d(hs_smem + off, reinterpret_cast<const uint8_t*>(recs + first) + off, min(16384u, cnt * 16 - off), &s_bar);
    }
    uint32_t was[HEALTH_STAGE_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < HEALTH_STAGE_ROWS; k++) {  // the previous state (device memory) meanwhile
      const uint32_t i = first + k * HEALTH_SMALL_THREADS + tid;
      was[k] = i < n ? alive_prev[i] : 0;
    }
    mbar_wait(&s_bar, phase);
    phase ^= 1;
#pragma unroll
    for (uint32_t k = 0; k < HEALTH_STAGE_ROWS; k++) {
      if (r0 + k < rows) {  // uniform
        const uint32_t i = first + k * HEALTH_SMALL_THREADS + tid;
        const bool in = i < n;
        const bool now = in && pci_record_alive(stage[k * HEALTH_SMALL_THREADS + tid]);
        const bool chg = in && (now ? 1u : 0u) != was[k];
        if (chg) alive_prev[i] = now ? 1 : 0;
        const uint32_t bc = __ballot_sync(KVG_FULL, chg), bn = __ballot_sync(KVG_FULL, now);
        if (lane == 0) {
          s_bal[r0 + k][warp] = bc;
          s_now[r0 + k][warp] = bn;
        }
        n_alive += now ? 1u : 0u;
      }
    }
    __syncthreads();  // every read of the stage is done before the next round's copy lands in it
  }
  // 32 rows x 32 warps = one counter per thread, in record order: ONE block scan places every transition
  const uint32_t crow = tid >> 5, cw = tid & 31;
  const uint32_t mybal = crow < rows ? s_bal[crow][cw] : 0;
  const uint32_t mycnt = __popc(mybal);
  const uint32_t incl = warp_incl_sum(mycnt);
  const uint32_t wal = warp_sum(n_alive);
  if (lane == 31) s_scan[warp] = incl;
  if (lane == 0) s_alv[warp] = wal;
  __syncthreads();
  uint32_t base = 0, total = 0, alive = 0;
#pragma unroll
  for (uint32_t w = 0; w < NW; w++) {
    const uint32_t c = s_scan[w];
    if (w < warp) base += c;
    total += c;
    alive += s_alv[w];
  }
  // thread (crow, cw) writes the transitions of warp cw in row crow: lane order == record order
  uint32_t pos = base + incl - mycnt;
  const uint32_t nowb = crow < rows ? s_now[crow][cw] : 0;
  for (uint32_t m = mybal; m; m &= m - 1) {
    const uint32_t l = (uint32_t)__ffs((int)m) - 1;
    const uint32_t i = crow * HEALTH_SMALL_THREADS + cw * 32 + l;
    changed_host[pos++] = (i << 1) | ((nowb >> l) & 1u);
  }
  __threadfence_system();  // the list entries of every thread are on their way before the flag
  __syncthreads();
  if (tid == 0) {
    hdr_host[0] = alive;
    hdr_host[1] = total;
    __threadfence_system();
    *((volatile uint32_t*)&hdr_host[2]) = seq;  // the host polls this word: no driver call on the way back
  }
}

// ------------------------------------------------------------------------------------------------
// mdev type dictionary: label = Trim(raw, "\n") then \s+ -> "_"  (device_plugin.go:341-342);
// canonical id = smallest raw index with an identical label (they are ONE vGpuMap key).
// ------------------------------------------------------------------------------------------------
__global__ void k_mdev_labels(const uint8_t* __restrict__ raw, const uint32_t* __restrict__ raw_off,
                              uint32_t n_types, uint8_t* __restrict__ label,
                              uint32_t* __restrict__ label_len, uint64_t* __restrict__ label_hash) {
  pdl_enter();
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_types) return;
  uint32_t a = raw_off[k], b = raw_off[k + 1];
  while (a < b && raw[a] == '\n') a++;
  while (b > a && raw[b - 1] == '\n') b--;
  uint8_t* out = label + raw_off[k];  // sanitised text is never longer than the raw text
  uint32_t o = 0;
  uint64_t h = 1469598103934665603ull;  // FNV-1a of the label: cheap first-level equality test
  for (uint32_t i = a; i < b;) {
    uint8_t c;
    if (d_re2_space(raw[i])) {
      c = '_';
      while (i < b && d_re2_space(raw[i])) i++;
    } else {
      c = raw[i++];
    }
    out[o++] = c;
    h = (h ^ c) * 1099511628211ull;
  }
  label_len[k] = o;
  label_hash[k] = h;
}
// canonical id = smallest raw index with an identical label: hash + length first (independent,
// pipelined loads), bytes only on a hash match
__global__ void k_mdev_canon(const uint8_t* __restrict__ label, const uint32_t* __restrict__ raw_off,
                             const uint32_t* __restrict__ label_len,
                             const uint64_t* __restrict__ label_hash, uint32_t n_types,
                             uint16_t* __restrict__ canon) {
  pdl_enter();
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_types) return;
  const uint32_t len = label_len[k];
  const uint64_t h = label_hash[k];
  const uint8_t* mine = label + raw_off[k];
  uint32_t c = k;
  for (uint32_t j = 0; j < k; j++) {
    if (label_hash[j] != h || label_len[j] != len) continue;
    const uint8_t* other = label + raw_off[j];
    bool eq = true;
    for (uint32_t t = 0; t < len && eq; t++) eq = other[t] == mine[t];
    if (eq) {
      c = j;
      break;
    }
  }
  canon[k] = (uint16_t)c;
}

// ------------------------------------------------------------------------------------------------
// synthetic snapshots (splitmix64, counter based) — identical to kvo_gen_pci / kvo_gen_mdev
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_gen_pci(uint4* __restrict__ out, uint64_t first, uint32_t n,
                          const uint16_t* __restrict__ nv_ids, uint32_t n_nv_ids,
                          uint32_t group_bits) {
  pdl_enter();
  const uint64_t SEED = 0x10DE000020250711ull;
  const uint16_t other[8] = {0x8086, 0x1002, 0x15b3, 0x1022, 0x144d, 0x14e4, 0x1af4, 0x10df};
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    uint64_t i = first + k;
    uint64_t r0 = mix64(SEED + 2 * i), r1 = mix64(SEED + 2 * i + 1);
    bool nvidia = (r0 & 0xFF) < 128;
    uint32_t vendor = nvidia ? 0x10deu : other[(r0 >> 8) & 7];
    uint32_t device;
    if (nvidia && ((r0 >> 16) & 0xFF) < 230 && n_nv_ids)
      device = nv_ids[(uint32_t)((r0 >> 24) & 0xFFFFFF) % n_nv_ids];
    else
      device = (uint32_t)((r0 >> 24) & 0xFFFF);
    uint32_t d = (uint32_t)((r0 >> 48) & 0xFF);
    uint32_t flags = 0, driver;
    if (d < 154)
      driver = KVG_DRV_VFIO_PCI;
    else if (d < 179)
      driver = KVG_DRV_NVGRACE;
    else if (d < 218)
      driver = 3;
    else if (d < 231)
      driver = 4;
    else {
      driver = KVG_DRV_NONE;
      flags |= KVG_PF_DRIVER_ERR;
    }
    uint32_t g = (uint32_t)(i >> 1);
    if (group_bits) {
      uint32_t mask = group_bits >= 32 ? 0xFFFFFFFFu : ((1u << group_bits) - 1);
      uint32_t hi = g & ~mask, lo = g & mask;
      lo = (lo * 0x9E3779B1u) & mask;
      lo ^= lo >> (group_bits / 2 + 1);
      g = hi | (lo & mask);
    }
    int32_t numa = (int32_t)(r1 & 7) - 1;
    if (((r1 >> 8) & 0xFF) == 0) flags |= KVG_PF_VENDOR_ERR;
    if (((r1 >> 16) & 0xFF) == 0) flags |= KVG_PF_IOMMU_ERR;
    if (((r1 >> 24) & 0xFF) == 0) flags |= KVG_PF_DEVICE_ERR;
    if (((r1 >> 32) & 0xFF) == 0) flags |= KVG_PF_NUMA_ERR;
    uint4 r;
    r.x = (uint32_t)i;
    r.y = vendor | (device << 16);
    r.z = g;
    r.w = driver | (flags << 8) | (((uint32_t)numa & 0xffffu) << 16);
    out[k] = r;
  }
}
__global__ void k_gen_mdev(uint4* __restrict__ out, uint64_t first, uint32_t n) {
  pdl_enter();
  const uint64_t SEED = 0x4D44455600010000ull;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    uint64_t j = first + k;
    uint64_t r0 = mix64(SEED + 2 * j), r1 = mix64(SEED + 2 * j + 1);
    // uuid = BE32(j) | BE64(r0) | BE32(r1 low 32)   (little-endian words hold big-endian bytes)
    uint4 lo;
    lo.x = __byte_perm((uint32_t)j, 0, 0x0123);
    lo.y = __byte_perm((uint32_t)(r0 >> 32), 0, 0x0123);
    lo.z = __byte_perm((uint32_t)r0, 0, 0x0123);
    lo.w = __byte_perm((uint32_t)r1, 0, 0x0123);
    uint32_t flags = 0;
    if (((r1 >> 40) & 0xFF) == 0) flags |= KVG_MF_TYPE_ERR;
    if (((r1 >> 48) & 0xFF) == 0) flags |= KVG_MF_PARENT_ERR;
    if (((r1 >> 56) & 0xFF) == 0) flags |= KVG_MF_NUMA_ERR;
    int32_t numa = (int32_t)((j >> 5) & 3) - 1;
    uint4 hi;
    hi.x = (uint32_t)(j >> 5);
    hi.y = (uint32_t)(r0 >> 56) | (flags << 16);
    hi.z = (uint32_t)numa & 0xffffu;
    hi.w = 0;
    out[2 * (size_t)k] = lo;
    out[2 * (size_t)k + 1] = hi;
  }
}

// L2 flush helper: stream zeros through a buffer larger than L2
__global__ void k_fill(uint4* __restrict__ p, size_t n16, uint32_t v) {
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(v, v, v, v);
}
__global__ void k_fill32(uint32_t* __restrict__ p, size_t n, uint32_t v) {
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// K3/K5 default form: classification without any cross-tile dependency.
//   k_classify_ragged  every CTA owns one tile (THREADS x ROWS records, 128-bit streaming loads, all
//                      in flight at once), evaluates the predicate, joins the name and writes its
//                      survivors — already in output format — at the TILE-LOCAL base tile*TILE of a
//                      scratch array, plus one count per tile.  Reads each record once, writes each
//                      survivor once, no waiting on other CTAs: this is the HBM-roofline kernel.
//   k_tile_offsets     exclusive scan of the tile counts (one CTA; n_tiles is ~N/1024)
//   k_pack_survivors   dense, order-preserving copy scratch -> survivors using the known offsets
// A single-pass look-back compaction (k_classify_oneshot / _tma / _ws, kept for A/B measurements,
// KVG_CLASSIFY=oneshot|tma|ws) moves fewer bytes but spends ~2/3 of each CTA's lifetime waiting
// for its base offset; measured on B200 it is ~1.7x slower end to end than this split form.
// ------------------------------------------------------------------------------------------------
template <class Op, int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) k_classify_ragged(Op op, uint32_t* __restrict__ tile_count,
                                                             uint2* __restrict__ tile_max) {
  pdl_enter();
  constexpr uint32_t TILE = THREADS * ROWS;
  constexpr uint32_t NW = THREADS / 32;
  constexpr uint32_t WARP_ITEMS = 32 * ROWS;
  __shared__ uint32_t s_wtot[NW];
  __shared__ uint2 s_wmax[NW];
  op.begin();
  const uint32_t n = op.count();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t base = tile * TILE + warp * WARP_ITEMS;
  typename Op::Item item[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    item[k] = op.load(i, i < n);
  }
  uint32_t bal[ROWS], aux[ROWS];
  uint32_t wtot = 0;
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    bool p = i < n && op.pred(item[k], i);
    bal[k] = __ballot_sync(KVG_FULL, p);
    wtot += __popc(bal[k]);
    aux[k] = p ? op.prepare(item[k]) : 0u;
  }
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  uint32_t off = tile * TILE;  // tile-local base: survivors of a tile stay contiguous and ordered
#pragma unroll
  for (uint32_t w = 0; w < NW; w++) {
    uint32_t c = s_wtot[w];
    if (w < warp) off += c;
    if (w == NW - 1 && threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (uint32_t v = 0; v < NW; v++) tot += s_wtot[v];
      tile_count[tile] = tot;
    }
  }
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    if ((bal[k] >> lane) & 1u) op.emit(off + __popc(bal[k] & lanemask_lt()), item[k], i, aux[k]);
    off += __popc(bal[k]);
  }
  // largest keys of the tile (they bound the radix pass counts): per-tile slot, no global atomics
  uint2 m = op.take_maxima();
  m.x = warp_max(m.x);
  m.y = warp_max(m.y);
  if (lane == 0) s_wmax[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint2 t = s_wmax[0];
#pragma unroll
    for (uint32_t w = 1; w < NW; w++) {
      t.x = max(t.x, s_wmax[w].x);
      t.y = max(t.y, s_wmax[w].y);
    }
    tile_max[tile] = t;
  }
}

// exclusive scan of tile_count[0..n_tiles) -> tile_off[0..n_tiles], total -> ctrl->n_surv, and the
// reduction of the per-tile maxima -> ctrl->max_group / max_devkey.  Chained scan: 2048 counts per
// CTA, coalesced, base by decoupled look-back (a few dozen CTAs at most, launched in order).
struct TileOffsetsArgs {
  const uint32_t* tile_count;
  const uint2* tile_max;        // optional per-tile maxima to reduce into ctrl
  const uint32_t* n_items_ptr;  // device-side item count (C_TILE items per tile) or NULL
  uint32_t n_tiles_host;
  uint32_t* tile_off;
  uint32_t* total_out;
  uint64_t* state;
};
struct TileOffsetsArgs2 {
  TileOffsetsArgs o[2];
};
__global__ void __launch_bounds__(KVG_BLOCK) k_tile_offsets(TileOffsetsArgs2 aa, ScanCtrl* ctrl,
                                                            uint32_t epoch) {
  pdl_enter();
  const TileOffsetsArgs A = blockIdx.y ? aa.o[1] : aa.o[0];
  const uint32_t* __restrict__ tile_count = A.tile_count;
  const uint2* __restrict__ tile_max = A.tile_max;
  const uint32_t* n_items_ptr = A.n_items_ptr;
  const uint32_t n_tiles_host = A.n_tiles_host;
  uint32_t* __restrict__ tile_off = A.tile_off;
  uint32_t* total_out = A.total_out;
  uint64_t* state = A.state;
  // n_tiles is either known on the host or derived from a device-side item count (C_TILE items/tile)
  const uint32_t n_tiles = n_items_ptr ? (*n_items_ptr + C_TILE - 1) / C_TILE : n_tiles_host;
  if (blockIdx.x * C_TILE >= n_tiles) {
    if (n_tiles == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
      *total_out = 0;
      tile_off[0] = 0;
    }
    return;
  }
  const uint32_t last_chunk = (n_tiles - 1) / C_TILE;
  __shared__ uint32_t scratch[KVG_WARPS + 1];
  __shared__ uint32_t s_base;
  const uint32_t chunk = blockIdx.x;
  const uint32_t i0 = chunk * C_TILE + threadIdx.x * C_ROWS;
  uint32_t v[C_ROWS], sum = 0, mg = 0, md = 0;
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = i0 + k;
    v[k] = i < n_tiles ? tile_count[i] : 0;
    sum += v[k];
    if (tile_max && i < n_tiles) {
      uint2 m = tile_max[i];
      mg = max(mg, m.x);
      md = max(md, m.y);
    }
  }
  uint32_t total;
  uint32_t excl = block_excl_sum(sum, scratch, &total);
  if (warp_id() == 0) {
    uint32_t base = lookback_sum(state, chunk, total, epoch);
    if (lane_id() == 0) {
      s_base = base;
      if (chunk == last_chunk) {
        *total_out = base + total;
        tile_off[n_tiles] = base + total;
      }
    }
  }
  mg = warp_max(mg);
  md = warp_max(md);
  if (lane_id() == 0) {
    if (mg) atomicMax(&ctrl->max_group, mg);
    if (md) atomicMax(&ctrl->max_devkey, md);
  }
  __syncthreads();
  uint32_t run = s_base + excl;
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = i0 + k;
    if (i < n_tiles) tile_off[i] = run;
    run += v[k];
  }
}

// dense, order-preserving pack: CTA = one tile, 16-byte units, fully coalesced on both sides
template <int UNITS_PER_ITEM>
__global__ void __launch_bounds__(128) k_pack_survivors(const uint4* __restrict__ ragged,
                                                        const uint32_t* __restrict__ tile_off,
                                                        uint32_t tile_items, uint4* __restrict__ dense) {
  pdl_enter();
  const uint32_t tile = blockIdx.x;
  const uint32_t o0 = tile_off[tile], o1 = tile_off[tile + 1];
  const uint32_t units = (o1 - o0) * UNITS_PER_ITEM;
  const uint4* src = ragged + (size_t)tile * tile_items * UNITS_PER_ITEM;
  uint4* dst = dense + (size_t)o0 * UNITS_PER_ITEM;
  for (uint32_t u = threadIdx.x; u < units; u += blockDim.x) st_stream(dst + u, ld_stream(src + u));
}

}  // namespace kvg

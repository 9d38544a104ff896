// kvg_scan.cuh — record classification, stable compaction and stable bucketing.
//
//   k_compact<Op>      single-pass stable compaction: 2048-item tiles, 8 x 128-bit loads in flight
//                      per thread, warp ballots, block scan, decoupled look-back on tile counts.
//                      Instantiated for
//        PciClassifyOp   K3: createIommuDeviceMap's filter (device_plugin.go:201-244) + name join
//        MdevClassifyOp  K5: createVgpuIDMap's filter (:268-289)
//        HeadsOp         segment heads of a sorted key array (distinct map keys + offsets)
//        HealthOp        K6: alive-set diff against the previous scan
//   k_radix_*          K4: stable LSD radix sort of (key, index) pairs, 8-bit digits:
//                      per-tile histogram -> per-digit tile scan -> ranked scatter
//   k_gen_*            counter-based synthetic snapshots (twins of oracle/kvg_oracle.c kvo_gen_*)
#pragma once
#include "../../include/kvgpu.h"
#include "kvg_common.cuh"
#include "kvg_parse.cuh"

namespace kvg {

constexpr uint32_t C_ROWS = 8;                      // items per thread
constexpr uint32_t C_TILE = KVG_BLOCK * C_ROWS;     // 2048 items per tile
constexpr uint32_t C_WARP_ITEMS = 32 * C_ROWS;      // 256 contiguous items per warp

// device-resident control block of one scan (zeroed by one memset per step)
struct ScanCtrl {
  uint32_t n_own[2];      // sharded scans: survivors whose key this rank owns, per ordering
  uint32_t reserved[6];
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
// K3/K5 hot form: the same stable compaction for fixed-size RECORDS, software-pipelined so HBM
// loads never stop:
//   * records arrive through a STAGES-deep ring of TMA bulk copies (cp.async.bulk -> UBLKCP), so
//     bytes stay in flight while the CTA is in its barrier / look-back phases;
//   * the look-back + write-out of tile i-1 run one iteration LATE, behind the ballots of tile i:
//     by then every predecessor published its count, so the look-back resolves without spinning
//     (its state words are even prefetched before the ballots).
// Iteration i:  [prefetch look-back words of tile i-1] -> wait TMA(i) -> LDS, predicate, ballots,
//   probes -> sync -> warp 0: publish count(i), resolve base(i-1) -> sync -> emit tile i-1.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t LB_KMAX = 16;             // look-back loads per lane: covers grids up to 512 CTAs
constexpr uint32_t CLASSIFY_MAX_GRID = 32 * LB_KMAX;

template <class Op, int ROWS, int STAGES>
__global__ void __launch_bounds__(KVG_BLOCK, 3) k_classify_tma(Op op, uint64_t* tile_agg,
                                                            uint64_t* round_incl, uint32_t epoch) {
  pdl_enter();
  // Tile t = b + r*G (CTA b, round r).  Its base offset is
  //     round_incl[r-1]  +  sum of tile_agg[r*G + k] for k < b
  // i.e. ONE batch of independent loads (prefetched a phase early) instead of a serial walk:
  // with co-resident CTAs running in lockstep every tile of a round resolves at the same time,
  // so a classic look-back would crawl through ~G/2 not-yet-inclusive predecessors.
  constexpr uint32_t TILE = KVG_BLOCK * ROWS;
  constexpr uint32_t RB = Op::REC_BYTES;
  constexpr uint32_t STAGE_BYTES = TILE * RB;
  constexpr uint32_t WARP_ITEMS = 32 * ROWS;
  extern __shared__ __align__(128) uint8_t c_smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ uint32_t s_wtot[KVG_WARPS];
  __shared__ uint32_t s_woff[2][KVG_WARPS];
  __shared__ uint32_t s_base;

  op.begin();
  const uint32_t n = op.count();
  const uint32_t n_tiles = (n + TILE - 1) / TILE;
  const uint32_t lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const uint32_t G = gridDim.x, b = blockIdx.x;
  if (n_tiles == 0) {
    if (b == 0 && tid == 0) op.finish(0);
    return;
  }
  if (b >= n_tiles) return;
  const uint32_t my_count = (n_tiles - b + G - 1) / G;  // rounds in which this CTA has a tile
  const uint8_t* src = reinterpret_cast<const uint8_t*>(op.src());
  const uint32_t tag = epoch & 0x3fffffffu;

  auto issue = [&](uint32_t i) {  // thread 0: TMA for my round-i tile into stage i % STAGES
    if (i >= my_count) return;
    uint32_t tile = b + i * G;
    uint32_t items = min(TILE, n - tile * TILE);
    uint32_t st = i % STAGES;
    mbar_arrive_expect_tx(&full_bar[st], items * RB);
    tma_load_1d(c_smem + st * STAGE_BYTES, src + (size_t)tile * STAGE_BYTES, items * RB, &full_bar[st]);
  };
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < STAGES; s++) issue((uint32_t)s);

  uint32_t prev_bal[ROWS], prev_aux[ROWS];
  uint32_t prev_total = 0;  // warp 0 only
#pragma unroll
  for (int k = 0; k < ROWS; k++) prev_bal[k] = prev_aux[k] = 0;

  for (uint32_t i = 0; i <= my_count; ++i) {
    // -- warp 0: prefetch everything the base of my round-(i-1) tile needs
    const uint32_t r = i - 1;                // round being resolved (valid when i > 0)
    const uint32_t ptile = b + r * G;
    uint64_t w[LB_KMAX], wr = 0;
    if (i > 0 && warp == 0) {
#pragma unroll
      for (uint32_t k = 0; k < LB_KMAX; k++) {
        uint32_t j = lane + 32 * k;
        w[k] = j < b ? ld_relaxed_u64(&tile_agg[r * G + j]) : 0;
      }
      if (r > 0) wr = ld_relaxed_u64(&round_incl[r - 1]);
    }
    uint32_t bal[ROWS], aux[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; k++) bal[k] = aux[k] = 0;
    if (i < my_count) {
      const uint32_t tile = b + i * G;
      const uint32_t st = i % STAGES;
      mbar_wait(&full_bar[st], (i / STAGES) & 1);
      const uint8_t* stage = c_smem + st * STAGE_BYTES;
      uint32_t wtot = 0;
#pragma unroll
      for (int k = 0; k < ROWS; k++) {
        uint32_t j = warp * WARP_ITEMS + k * 32 + lane;  // item index inside the tile
        bool ok = tile * TILE + j < n;
        typename Op::Item it = op.from_smem(stage + (size_t)j * RB);
        bool p = ok && op.pred(it, tile * TILE + j);
        bal[k] = __ballot_sync(KVG_FULL, p);
        wtot += __popc(bal[k]);
        aux[k] = p ? op.prepare(it) : 0u;
      }
      if (lane == 0) s_wtot[warp] = wtot;
    }
    __syncthreads();  // A: counts of tile i visible; everyone finished emitting tile i-2
    if (tid == 0 && i >= 2) issue(i - 2 + STAGES);  // stage (i-2) % STAGES is free again
    if (warp == 0) {
      uint32_t my_total = 0;
      if (i < my_count) {
        const uint32_t tile = b + i * G;
        uint32_t wv = lane < KVG_WARPS ? s_wtot[lane] : 0;
        uint32_t wi = warp_incl_sum(wv);
        if (lane < KVG_WARPS) s_woff[i & 1][lane] = wi - wv;
        my_total = __shfl_sync(KVG_FULL, wi, KVG_WARPS - 1);
        if (lane == 0) st_relaxed_u64(&tile_agg[tile], ((uint64_t)tag << 34) | my_total);
      }
      if (i > 0) {
        uint32_t part = 0;
#pragma unroll
        for (uint32_t k = 0; k < LB_KMAX; k++) {
          uint32_t j = lane + 32 * k;
          if (j < b) {
            uint64_t v = w[k];
            while ((uint32_t)(v >> 34) != tag) v = ld_relaxed_u64(&tile_agg[r * G + j]);
            part += (uint32_t)v;
          }
        }
        uint32_t excl = warp_sum(part);
        if (r > 0) {
          while ((uint32_t)(wr >> 34) != tag) wr = ld_relaxed_u64(&round_incl[r - 1]);
          excl += (uint32_t)wr;
        }
        if (lane == 0) {
          s_base = excl;
          if (b == G - 1) st_relaxed_u64(&round_incl[r], ((uint64_t)tag << 34) | (excl + prev_total));
          if (ptile == n_tiles - 1) op.finish(excl + prev_total);
        }
      }
      prev_total = my_total;
    }
    __syncthreads();  // B: base of tile i-1 visible
    if (i > 0) {
      const uint8_t* stage = c_smem + ((i - 1) % STAGES) * STAGE_BYTES;
      uint32_t off = s_base + s_woff[(i - 1) & 1][warp];
#pragma unroll
      for (int k = 0; k < ROWS; k++) {
        if ((prev_bal[k] >> lane) & 1u) {
          uint32_t j = warp * WARP_ITEMS + k * 32 + lane;
          typename Op::Item it = op.from_smem(stage + (size_t)j * RB);
          op.emit(off + __popc(prev_bal[k] & lanemask_lt()), it, ptile * TILE + j, prev_aux[k]);
        }
        off += __popc(prev_bal[k]);
      }
      op.tile_epilogue();
    }
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
      prev_bal[k] = bal[k];
      prev_aux[k] = aux[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3/K5, warp-specialised form (the one the scans launch).  9 warps per CTA:
//   warp 8  "scan" warp: issues the TMA ring, publishes the tile count, resolves the tile's base
//           offset (round prefix + same-round aggregates, independent loads) and hands it over;
//   warps 0..7 compute: wait TMA -> LDS -> predicate -> ballots (phase 1 of tile i), then write
//           out tile i-1 whose base the scan warp produced meanwhile.  They never spin on global
//           memory and never wait for the look-back.
// Hand-off: named barriers X[p] (counts ready) and Y[p] (base ready), p = tile parity; stage
// recycling through an mbarrier per stage that the compute warps arrive on after their write-out.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t WS_THREADS = KVG_BLOCK + 32;

template <class Op, int ROWS, int STAGES>
__global__ void __launch_bounds__(WS_THREADS) k_classify_ws(Op op, uint64_t* tile_agg,
                                                            uint64_t* round_incl, uint32_t epoch) {
  pdl_enter();
  constexpr uint32_t TILE = KVG_BLOCK * ROWS;
  constexpr uint32_t RB = Op::REC_BYTES;
  constexpr uint32_t STAGE_BYTES = TILE * RB;
  constexpr uint32_t WARP_ITEMS = 32 * ROWS;
  extern __shared__ __align__(128) uint8_t c_smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
  __shared__ uint32_t s_wtot[2][KVG_WARPS];
  __shared__ uint32_t s_woff[2][KVG_WARPS];
  __shared__ uint32_t s_base[2];

  op.begin();
  const uint32_t n = op.count();
  const uint32_t n_tiles = (n + TILE - 1) / TILE;
  const uint32_t lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const uint32_t G = gridDim.x, b = blockIdx.x;
  if (n_tiles == 0) {
    if (b == 0 && tid == 0) op.finish(0);
    return;
  }
  if (b >= n_tiles) return;
  const uint32_t my_count = (n_tiles - b + G - 1) / G;
  const uint32_t tag = epoch & 0x3fffffffu;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], KVG_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == KVG_WARPS) {
    // ============================== scan / producer warp =====================================
    const uint8_t* src = reinterpret_cast<const uint8_t*>(op.src());
    auto issue = [&](uint32_t i) {  // lane 0
      if (i >= my_count) return;
      uint32_t tile = b + i * G;
      uint32_t items = min(TILE, n - tile * TILE);
      uint32_t st = i % STAGES;
      mbar_arrive_expect_tx(&full_bar[st], items * RB);
      tma_load_1d(c_smem + st * STAGE_BYTES, src + (size_t)tile * STAGE_BYTES, items * RB, &full_bar[st]);
    };
    if (lane == 0)
      for (int s = 0; s < STAGES; s++) issue((uint32_t)s);
    for (uint32_t i = 0; i < my_count; ++i) {
      const uint32_t tile = b + i * G;
      named_bar_sync(1 + (i & 1), WS_THREADS);  // X: the 8 warp counts of tile i are in s_wtot
      uint32_t wv = lane < KVG_WARPS ? s_wtot[i & 1][lane] : 0;
      uint32_t wi = warp_incl_sum(wv);
      if (lane < KVG_WARPS) s_woff[i & 1][lane] = wi - wv;
      const uint32_t total = __shfl_sync(KVG_FULL, wi, KVG_WARPS - 1);
      if (lane == 0) st_relaxed_u64(&tile_agg[tile], ((uint64_t)tag << 34) | total);
      // base = round_incl[i-1] + sum of this round's aggregates of CTAs 0..b-1
      uint64_t w[LB_KMAX], wr = 0;
#pragma unroll
      for (uint32_t k = 0; k < LB_KMAX; k++) {
        uint32_t j = lane + 32 * k;
        w[k] = j < b ? ld_relaxed_u64(&tile_agg[i * G + j]) : 0;
      }
      if (i > 0) wr = ld_relaxed_u64(&round_incl[i - 1]);
      uint32_t part = 0;
#pragma unroll
      for (uint32_t k = 0; k < LB_KMAX; k++) {
        uint32_t j = lane + 32 * k;
        if (j < b) {
          uint64_t v = w[k];
          while ((uint32_t)(v >> 34) != tag) v = ld_relaxed_u64(&tile_agg[i * G + j]);
          part += (uint32_t)v;
        }
      }
      uint32_t excl = warp_sum(part);
      if (i > 0) {
        while ((uint32_t)(wr >> 34) != tag) wr = ld_relaxed_u64(&round_incl[i - 1]);
        excl += (uint32_t)wr;
      }
      if (lane == 0) {
        s_base[i & 1] = excl;
        if (b == G - 1) st_relaxed_u64(&round_incl[i], ((uint64_t)tag << 34) | (excl + total));
        if (tile == n_tiles - 1) op.finish(excl + total);
      }
      named_bar_arrive(3 + (i & 1), WS_THREADS);  // Y: base + warp offsets of tile i are ready
      // recycle the stage of tile i-1 once every compute warp has written that tile out
      if (i >= 1 && i - 1 + STAGES < my_count) {
        mbar_wait(&empty_bar[(i - 1) % STAGES], ((i - 1) / STAGES) & 1);
        if (lane == 0) issue(i - 1 + STAGES);
      }
    }
    return;
  }

  // ================================== compute warps ==========================================
  uint32_t prev_bal[ROWS], prev_aux[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) prev_bal[k] = prev_aux[k] = 0;
  for (uint32_t i = 0; i <= my_count; ++i) {
    uint32_t bal[ROWS], aux[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; k++) bal[k] = aux[k] = 0;
    if (i < my_count) {
      const uint32_t tile = b + i * G;
      const uint32_t st = i % STAGES;
      mbar_wait(&full_bar[st], (i / STAGES) & 1);
      const uint8_t* stage = c_smem + st * STAGE_BYTES;
      uint32_t wtot = 0;
#pragma unroll
      for (int k = 0; k < ROWS; k++) {
        uint32_t j = warp * WARP_ITEMS + k * 32 + lane;
        bool ok = tile * TILE + j < n;
        typename Op::Item it = op.from_smem(stage + (size_t)j * RB);
        bool p = ok && op.pred(it, tile * TILE + j);
        bal[k] = __ballot_sync(KVG_FULL, p);
        wtot += __popc(bal[k]);
        aux[k] = p ? op.prepare(it) : 0u;
      }
      if (lane == 0) s_wtot[i & 1][warp] = wtot;
      named_bar_arrive(1 + (i & 1), WS_THREADS);  // X
    }
    if (i > 0) {
      const uint32_t pi = i - 1, ptile = b + pi * G;
      named_bar_sync(3 + (pi & 1), WS_THREADS);  // Y
      const uint8_t* stage = c_smem + (pi % STAGES) * STAGE_BYTES;
      uint32_t off = s_base[pi & 1] + s_woff[pi & 1][warp];
#pragma unroll
      for (int k = 0; k < ROWS; k++) {
        if ((prev_bal[k] >> lane) & 1u) {
          uint32_t j = warp * WARP_ITEMS + k * 32 + lane;
          typename Op::Item it = op.from_smem(stage + (size_t)j * RB);
          op.emit(off + __popc(prev_bal[k] & lanemask_lt()), it, ptile * TILE + j, prev_aux[k]);
        }
        off += __popc(prev_bal[k]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[pi % STAGES]);
    }
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
      prev_bal[k] = bal[k];
      prev_aux[k] = aux[k];
    }
  }
  op.tile_epilogue();  // running maxima -> one atomic per warp
}

// ------------------------------------------------------------------------------------------------
// K3/K5, one-tile-per-CTA form: small CTAs (THREADS x ROWS records kept in registers), thousands
// of them, dispatched in blockIdx order by the hardware.  Many resident CTAs per SM hide the
// count -> look-back -> write-out latency chain of each other; predecessors were dispatched
// earlier, so the classic decoupled look-back usually finds an inclusive prefix close by.
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
  __device__ __forceinline__ const void* src() const { return recs; }
  __device__ __forceinline__ Item from_smem(const uint8_t* p) const {
    return *reinterpret_cast<const uint4*>(p);
  }
  const uint4* recs;
  uint32_t n;
  kvg_pci_surv* out;
  ScanCtrl* ctrl;
  const uint64_t* table;
  uint32_t cap_mask, cap_shift;
  const PciIdsInfo* info;
  const uint32_t* nv_index;                 // device id -> name pool slot (NULL: probe the hash)
  uint32_t local_max_group, local_max_dev;  // per-thread running maxima (registers)
  uint32_t v_off, sec_end;                  // section bounds, read once per thread

  __device__ __forceinline__ void begin() {
    v_off = info ? info->v_off : P_NONE;
    sec_end = info ? info->sec_end : 0;
  }
  // the name join.  nv_index is the pci.ids hash table flattened for vendor 10de by
  // k_probe_keys right after every parse (65,536 x u32 pool slots): one load per survivor.
  __device__ __forceinline__ uint32_t prepare(const Item& r) const {
    if (nv_index) return __ldg(&nv_index[r.y >> 16]);
    uint32_t off = table_probe(table, cap_mask, cap_shift, (0x10deu << 16) | (r.y >> 16));
    return (off != P_NONE && v_off != P_NONE && off > v_off && off < sec_end) ? off - v_off : P_NONE;
  }
  __device__ __forceinline__ uint32_t count() const { return n; }
  __device__ __forceinline__ Item load(uint32_t i, bool ok) const {
    return ok ? ld_stream(recs + i) : make_uint4(0, 0, 0, 0xff00u);
  }
  // record = {addr, vendor | device<<16, iommu_group, driver | flags<<8 | numa<<16}
  // device_plugin.go:203-238: any of vendor/driver/iommu/device read errors drops the entry,
  // vendor must be "10de" (:209), driver in supportedVfioDrivers (:217, :75-78)
  __device__ __forceinline__ bool pred(const Item& r, uint32_t) const { return pci_record_alive(r); }
  __device__ __forceinline__ void emit(uint32_t pos, const Item& r, uint32_t, uint32_t name_slot) {
    uint32_t device = r.y >> 16;
    uint32_t flags = (r.w >> 8) & 0xffu;
    int32_t numa = (int32_t)r.w >> 16;  // sign-extended int16
    if ((flags & KVG_PF_NUMA_ERR) || numa < 0) numa = 0;  // :227-230, :316-318
    uint4 s;
    s.x = r.x;
    s.y = r.z;
    s.z = device | ((uint32_t)numa << 16);
    s.w = name_slot;
    st_stream(reinterpret_cast<uint4*>(out) + pos, s);
    local_max_group = max(local_max_group, r.z);
    local_max_dev = max(local_max_dev, device);
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

// ---- K5: mdev classify --------------------------------------------------------------------------
struct MdevItem {
  uint4 lo, hi;
};
struct MdevClassifyOp {
  using Item = MdevItem;
  static constexpr uint32_t REC_BYTES = 32;
  __device__ __forceinline__ const void* src() const { return recs; }
  __device__ __forceinline__ Item from_smem(const uint8_t* p) const {
    Item it;
    it.lo = *reinterpret_cast<const uint4*>(p);
    it.hi = *reinterpret_cast<const uint4*>(p + 16);
    return it;
  }
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
  __device__ __forceinline__ void emit(uint32_t pos, const Item& r, uint32_t i, uint32_t canon) {
    uint32_t flags = (r.hi.y >> 16) & 0xffu;
    int32_t numa = (int32_t)(int16_t)(r.hi.z & 0xffffu);
    if ((flags & KVG_MF_NUMA_ERR) || numa < 0) numa = 0;  // :281-284, :316-318
    uint4 hi;
    hi.x = r.hi.x;
    hi.y = canon | ((uint32_t)numa << 16);
    hi.z = i;
    hi.w = 0;
    st_stream(out + 2 * (size_t)pos, r.lo);
    st_stream(out + 2 * (size_t)pos + 1, hi);
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

// ------------------------------------------------------------------------------------------------
// K4: stable LSD radix sort of (key, survivor index) pairs, 8-bit digits, tiles of 2048 pairs.
//   k_radix_hist      per-tile digit histogram: one MATCH-aggregated shared-memory add per distinct
//                     digit per warp row (device ids are heavily skewed: naive atomics serialise)
//   k_radix_tilescan  per digit: exclusive scan of the per-tile counts (one CTA per digit)
//   k_radix_scatter   stable ranks (match + per-warp counters), pairs staged in shared memory in
//                     tile-sorted order, then written out so that consecutive threads write
//                     consecutive addresses of a bucket (coalesced runs instead of 4-byte scatters)
//   k_order_count / k_tile_offsets / k_order_emit
//                     final permutation + distinct keys (segment heads) without any look-back:
//                     count heads per tile -> scan -> emit at known offsets
// Passes whose digit is above the largest key (device-side knowledge) return immediately; the
// ping-pong parity then tells consumers which buffer is final.
// ------------------------------------------------------------------------------------------------
enum : int { SRC_PAIRS = 0, SRC_PCI_GROUP = 1, SRC_PCI_DEVICE = 2, SRC_MDEV_PARENT = 3, SRC_MDEV_TYPE = 4 };

// Digit width is decided ON THE DEVICE from the largest key of the ordering: the fewest passes of at
// most RADIX_MAX_BITS bits, the key bits split evenly between them (19-bit keys: 2 passes of 10 bits
// instead of 3 of 8; 16-bit keys: 2 x 8; 23-bit: 3 x 8).  Every CTA of every kernel of a pass derives
// the same plan from the same word, so nothing about it crosses the host.
constexpr uint32_t RADIX_MAX_BITS = 11;
constexpr uint32_t RADIX_MAX_DIGITS = 1u << RADIX_MAX_BITS;  // 2048
constexpr uint32_t RADIX_CHUNKS = RADIX_MAX_DIGITS / KVG_BLOCK;  // digit chunks of one thread each
struct RadixPlan {
  uint32_t npass, shift, bits;  // of the queried pass; bits == 0: the pass does not exist
};
__host__ __device__ __forceinline__ RadixPlan radix_plan(uint32_t max_key, uint32_t key_bits_max, uint32_t pass,
                                                         uint32_t max_bits) {
#ifdef __CUDA_ARCH__
  uint32_t kb = max_key ? 32u - (uint32_t)__clz((int)max_key) : 1u;
#else
  uint32_t kb = 1;
  while (kb < 32 && (max_key >> kb) != 0) kb++;
#endif
  if (kb > key_bits_max) kb = key_bits_max;
  RadixPlan r;
  r.npass = (kb + max_bits - 1) / max_bits;
  const uint32_t w = (kb + r.npass - 1) / r.npass;
  r.shift = pass * w;
  r.bits = pass < r.npass ? (kb - r.shift < w ? kb - r.shift : w) : 0;
  return r;
}

struct RadixArgs {
  const uint32_t* n_ptr;      // element count (device)
  const uint32_t* max_key;    // largest key (device): decides the plan
  const void* src_records;    // survivors (SRC_* != PAIRS)
  const uint2* pairs_in;      // {key, val}
  uint2* pairs_out;
  uint32_t* tile_hist;        // [digits][T]  (T = ceil(n / C_TILE)), digit-major
  uint32_t* bin_total;        // [RADIX_MAX_DIGITS] for this ordering (rewritten by every pass)
  uint32_t pass;              // 0xff: this ordering has no such pass
  uint32_t key_bits_max;      // 16 (device id / type) or 32 (iommu group / parent)
  uint32_t max_bits;          // widest digit: 11 (latency-bound sizes) or 8 (large inputs: 6 CTAs/SM)
  int src;                    // where pass-0 keys come from
};

// both orderings (device id, iommu group) run their passes in the SAME launches: blockIdx.y
// selects the ordering
struct RadixArgs2 {
  RadixArgs o[2];
};
__device__ __forceinline__ RadixPlan radix_pass(const RadixArgs& a) {
  if (a.pass == 0xffu) {
    RadixPlan r = {0, 0, 0};
    return r;
  }
  return radix_plan(*a.max_key, a.key_bits_max, a.pass, a.max_bits);
}
__device__ __forceinline__ uint2 radix_load(const RadixArgs& a, uint32_t i) {
  switch (a.src) {
    case SRC_PCI_GROUP:
      return make_uint2(reinterpret_cast<const kvg_pci_surv*>(a.src_records)[i].iommu_group, i);
    case SRC_PCI_DEVICE:
      return make_uint2(reinterpret_cast<const kvg_pci_surv*>(a.src_records)[i].device, i);
    case SRC_MDEV_PARENT:
      return make_uint2(reinterpret_cast<const kvg_mdev_surv*>(a.src_records)[i].parent, i);
    case SRC_MDEV_TYPE:
      return make_uint2(reinterpret_cast<const kvg_mdev_surv*>(a.src_records)[i].type_key, i);
    default: return a.pairs_in[i];
  }
}

__global__ void __launch_bounds__(KVG_BLOCK) k_radix_hist(RadixArgs2 aa) {
  pdl_enter();
  const RadixArgs a = blockIdx.y ? aa.o[1] : aa.o[0];  // static indices: parameters stay in the constant bank
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const RadixPlan pl = radix_pass(a);
  if (!pl.bits) return;
  const uint32_t dmask = (1u << pl.bits) - 1;
  const uint32_t nj = ((1u << pl.bits) + KVG_BLOCK - 1) / KVG_BLOCK;  // digit chunks in use
  __shared__ uint32_t h[RADIX_MAX_DIGITS];
  const uint32_t lane = lane_id();
  // tile loop: launched with one CTA per tile for the always-active passes, with a small grid for the
  // high passes that are usually ruled out by the device-side max key (they then cost ~nothing)
  for (uint32_t tile = blockIdx.x; tile < T; tile += gridDim.x) {
    for (uint32_t j = 0; j < nj; j++) h[j * KVG_BLOCK + threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * C_TILE + warp_id() * C_WARP_ITEMS;
    uint32_t d[C_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {  // all loads in flight before the first shared atomic
      uint32_t i = base + k * 32 + lane;
      d[k] = i < n ? ((radix_load(a, i).x >> pl.shift) & dmask) : 0xffffffffu;
    }
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++)
      if (d[k] != 0xffffffffu) atomicAdd(&h[d[k]], 1u);
    __syncthreads();
    for (uint32_t j = 0; j < nj; j++) {
      const uint32_t dg = j * KVG_BLOCK + threadIdx.x;
      a.tile_hist[(size_t)dg * T + tile] = h[dg];
    }
    __syncthreads();
  }
}

// one CTA per digit (and per 256-digit chunk): exclusive scan of that digit's per-tile counts, in place
__global__ void __launch_bounds__(KVG_BLOCK) k_radix_tilescan(RadixArgs2 aa) {
  pdl_enter();
  const RadixArgs a = blockIdx.y ? aa.o[1] : aa.o[0];  // static indices: parameters stay in the constant bank
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const RadixPlan pl = radix_pass(a);
  if (T == 0 || !pl.bits) return;
  const uint32_t nj = ((1u << pl.bits) + KVG_BLOCK - 1) / KVG_BLOCK;
  __shared__ uint32_t scratch[KVG_WARPS + 1];
  for (uint32_t j = 0; j < nj; j++) {  // grid.x == KVG_BLOCK: CTA b owns digits b, b + 256, ...
    const uint32_t dg = j * KVG_BLOCK + blockIdx.x;
    uint32_t* row = a.tile_hist + (size_t)dg * T;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < T; b += KVG_BLOCK) {
      uint32_t i = b + threadIdx.x;
      uint32_t v = i < T ? row[i] : 0;
      uint32_t total;
      uint32_t e = block_excl_sum(v, scratch, &total);
      if (i < T) row[i] = carry + e;
      carry += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) a.bin_total[dg] = carry;  // total of this digit (was: atomics in the histogram)
  }
}

// dynamic shared memory of k_radix_scatter: per-warp digit counts (u16: a warp owns 256 items),
// tile-local digit starts, global run offsets, the staged tile
// (11-bit digits: 32 + 4 + 8 + 16 = 60 KiB, 3 CTAs/SM; 8-bit digits: 4 + 0.5 + 1 + 16 KiB, 5 CTAs/SM)
template <uint32_t MAXB>
struct RadixScatterCfg {
  static constexpr uint32_t DIGITS = 1u << MAXB;
  static constexpr uint32_t CHUNKS = DIGITS / KVG_BLOCK;
  static constexpr uint32_t CNT_BYTES = KVG_WARPS * DIGITS * 2;
  static constexpr uint32_t START_BYTES = DIGITS * 2;
  static constexpr uint32_t GOFF_BYTES = DIGITS * 4;
  static constexpr uint32_t STAGE_BYTES = C_TILE * 8;
  static constexpr uint32_t SMEM = CNT_BYTES + START_BYTES + GOFF_BYTES + STAGE_BYTES;
  static constexpr int MIN_CTAS = MAXB <= 8 ? 5 : 3;  // 6 would cap registers at 40 and spill
};

template <uint32_t MAXB>
__global__ void __launch_bounds__(KVG_BLOCK, RadixScatterCfg<MAXB>::MIN_CTAS) k_radix_scatter(RadixArgs2 aa) {
  using Cfg = RadixScatterCfg<MAXB>;
  constexpr uint32_t RADIX_CHUNKS = Cfg::CHUNKS, RADIX_MAX_DIGITS = Cfg::DIGITS;
  constexpr uint32_t RS_CNT_BYTES = Cfg::CNT_BYTES, RS_START_BYTES = Cfg::START_BYTES, RS_GOFF_BYTES = Cfg::GOFF_BYTES;
  pdl_enter();
  const RadixArgs a = blockIdx.y ? aa.o[1] : aa.o[0];  // static indices: parameters stay in the constant bank
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const RadixPlan pl = radix_pass(a);
  if (!pl.bits) return;
  const uint32_t dmask = (1u << pl.bits) - 1;
  const uint32_t nj = RADIX_CHUNKS == 1 ? 1u : ((1u << pl.bits) + KVG_BLOCK - 1) / KVG_BLOCK;
  const uint32_t lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  extern __shared__ __align__(16) uint8_t rs_smem[];
  uint16_t (*s_cnt)[RADIX_MAX_DIGITS] = reinterpret_cast<uint16_t (*)[RADIX_MAX_DIGITS]>(rs_smem);
  uint16_t* s_start = reinterpret_cast<uint16_t*>(rs_smem + RS_CNT_BYTES);  // tile-local exclusive start of each digit
  int32_t* s_goff = reinterpret_cast<int32_t*>(rs_smem + RS_CNT_BYTES + RS_START_BYTES);  // global run position - local start
  uint2* s_stage = reinterpret_cast<uint2*>(rs_smem + RS_CNT_BYTES + RS_START_BYTES + RS_GOFF_BYTES);
  __shared__ uint32_t scratch[KVG_WARPS + 1];
  uint32_t total;
  uint32_t bin_base[RADIX_CHUNKS];
  bool have_base = false;
  for (uint32_t tile = blockIdx.x; tile < T; tile += gridDim.x) {
  const uint32_t base = tile * C_TILE + warp * C_WARP_ITEMS;
  // every global load of the tile is issued before anything waits: the pairs, the scanned tile
  // counts of this thread's digits and (first tile only) the digit totals -> ONE memory latency
  uint32_t tile_prefix[RADIX_CHUNKS];
#pragma unroll
  for (uint32_t j = 0; j < RADIX_CHUNKS; j++)
    tile_prefix[j] = j < nj ? a.tile_hist[(size_t)(j * KVG_BLOCK + tid) * T + tile] : 0;
  uint2 kv[C_ROWS];
  uint32_t rank[C_ROWS];
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    kv[k] = i < n ? radix_load(a, i) : make_uint2(0, 0);
  }
  __syncthreads();  // previous tile's stage fully written out
  for (uint32_t j = 0; j < nj; j++) {
#pragma unroll
    for (uint32_t w = 0; w < KVG_WARPS; w++) s_cnt[w][j * KVG_BLOCK + tid] = 0;
  }
  if (!have_base) {  // exclusive scan of the digit totals, digit chunks in order
    uint32_t carry = 0;
#pragma unroll
    for (uint32_t j = 0; j < RADIX_CHUNKS; j++) {
      bin_base[j] = 0;
      if (j < nj) {
        const uint32_t mine = a.bin_total[j * KVG_BLOCK + tid];
        bin_base[j] = carry + block_excl_sum(mine, scratch, &total);  // syncs inside
        carry += total;
        __syncthreads();
      }
    }
    have_base = true;
  }
  __syncthreads();
  // stable rank inside the warp: rows in order, lanes in order within a row
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    bool ok = i < n;
    uint32_t d = ok ? ((kv[k].x >> pl.shift) & dmask) : (0x10000u + lane);  // inactive lanes: unique
    uint32_t peers = __match_any_sync(KVG_FULL, d);
    uint32_t leader = (uint32_t)__ffs(peers) - 1;
    uint32_t before = 0;
    if (ok && lane == leader) {
      before = s_cnt[warp][d];
      s_cnt[warp][d] = (uint16_t)(before + __popc(peers));
    }
    before = __shfl_sync(KVG_FULL, before, leader);
    rank[k] = before + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: exclusive prefix over warps, tile-local start, global run offset
    uint32_t carry = 0;
    for (uint32_t j = 0; j < nj; j++) {
      const uint32_t dg = j * KVG_BLOCK + tid;
      uint32_t dtot = 0;
#pragma unroll
      for (uint32_t w = 0; w < KVG_WARPS; w++) {
        uint32_t c = s_cnt[w][dg];
        s_cnt[w][dg] = (uint16_t)dtot;
        dtot += c;
      }
      const uint32_t lstart = carry + block_excl_sum(dtot, scratch, &total);  // syncs inside
      carry += total;
      s_start[dg] = (uint16_t)lstart;
      uint32_t bb = 0, tp = 0;
#pragma unroll
      for (uint32_t q = 0; q < RADIX_CHUNKS; q++)  // static register indices
        if (q == j) {
          bb = bin_base[q];
          tp = tile_prefix[q];
        }
      s_goff[dg] = (int32_t)(bb + tp) - (int32_t)lstart;
      __syncthreads();
    }
  }
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    if (i < n) {
      uint32_t d = (kv[k].x >> pl.shift) & dmask;
      s_stage[(uint32_t)s_start[d] + s_cnt[warp][d] + rank[k]] = kv[k];
    }
  }
  __syncthreads();
  const uint32_t cnt = min(C_TILE, n - tile * C_TILE);
  for (uint32_t j = tid; j < cnt; j += KVG_BLOCK) {
    uint2 e = s_stage[j];
    uint32_t d = (e.x >> pl.shift) & dmask;
    a.pairs_out[(uint32_t)(s_goff[d] + (int32_t)j)] = e;
  }
  }  // tile loop
}

// ---- final permutation + distinct keys of one ordering, look-back free --------------------------
struct OrderFinalArgs {
  const uint2* p0;            // ping-pong buffers of the radix passes
  const uint2* p1;
  const uint32_t* max_key;
  uint32_t key_bits_max, max_bits;
  const uint32_t* n_ptr;
  uint32_t* perm;             // [n] survivor indices in key order (stable)
  uint32_t* tile_heads;       // [T] number of segment heads in each tile
  const uint32_t* tile_off;   // [T+1] exclusive scan of tile_heads (emit only)
  uint32_t* seg_key;
  uint32_t* seg_off;          // [n_seg + 1]
  const uint32_t* n_seg;      // total heads (emit only)
  const uint4* head_surv;     // optional: survivors, to publish the joined name slot of each segment
  uint32_t* head_name;        // [n_seg] name slot of the segment's first member (NULL: skip)
};
__device__ __forceinline__ const uint2* order_final_buf(const OrderFinalArgs& a) {
  const uint32_t np = radix_plan(*a.max_key, a.key_bits_max, 0, a.max_bits).npass;
  return ((np - 1) & 1) ? a.p1 : a.p0;
}
struct OrderFinalArgs2 {
  OrderFinalArgs o[2];
};
template <bool EMIT>
__global__ void __launch_bounds__(KVG_BLOCK) k_order_final(OrderFinalArgs2 aa) {
  pdl_enter();
  const OrderFinalArgs a = blockIdx.y ? aa.o[1] : aa.o[0];
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const uint32_t tile = blockIdx.x;
  if (tile >= T) {
    if (EMIT && n == 0 && tile == 0 && threadIdx.x == 0) a.seg_off[0] = 0;
    return;
  }
  const uint2* pairs = order_final_buf(a);
  const uint32_t lane = lane_id(), warp = warp_id();
  const uint32_t base = tile * C_TILE + warp * C_WARP_ITEMS;
  __shared__ uint32_t s_w[KVG_WARPS];
  uint32_t bal[C_ROWS], key[C_ROWS], wtot = 0;
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    bool head = false;
    key[k] = 0;
    if (i < n) {
      uint2 e = pairs[i];
      key[k] = e.x;
      head = i == 0 || pairs[i - 1].x != e.x;
      if (!EMIT) a.perm[i] = e.y;
    }
    bal[k] = __ballot_sync(KVG_FULL, head);
    wtot += __popc(bal[k]);
  }
  if (lane == 0) s_w[warp] = wtot;
  __syncthreads();
  if (!EMIT) {
    if (threadIdx.x == 0) {
      uint32_t t = 0;
#pragma unroll
      for (uint32_t w = 0; w < KVG_WARPS; w++) t += s_w[w];
      a.tile_heads[tile] = t;
    }
    return;
  }
  uint32_t off = a.tile_off[tile];
#pragma unroll
  for (uint32_t w = 0; w < KVG_WARPS; w++)
    if (w < warp) off += s_w[w];
#pragma unroll
  for (uint32_t k = 0; k < C_ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    if ((bal[k] >> lane) & 1u) {
      uint32_t pos = off + __popc(bal[k] & lanemask_lt());
      a.seg_key[pos] = key[k];
      a.seg_off[pos] = i;
      // all members of a device-id bucket share the name (same id): take the first member's slot
      if (a.head_name) a.head_name[pos] = __ldg(&a.head_surv[pairs[i].y].w);
    }
    off += __popc(bal[k]);
  }
  if (tile == T - 1 && threadIdx.x == 0) a.seg_off[*a.n_seg] = n;
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
__global__ void k_fill64(uint64_t* __restrict__ p, size_t n, uint64_t v) {
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
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

// ---- multi-GPU: key-partitioned bucketing -------------------------------------------------------
// After the allgatherv every rank holds the full survivor list; rank r orders only the survivors
// whose key (device id for ordering 0, iommu group for ordering 1) satisfies key % nranks == r.
// Key sets are disjoint, so the union of the per-rank buckets is the global map and the per-rank
// ordering work stays constant as GPUs are added.  This op selects the owned {key, index} pairs
// (same ragged -> offsets -> pack machinery as the classification).
struct OwnedPairOp {
  using Item = uint4;  // one kvg_pci_surv
  const uint4* surv;
  const uint32_t* n_ptr;  // gathered survivor count lives on the device (peer-memory path: never on the host)
  uint2* out;
  uint32_t field;  // 0: device id (ordering 0), 1: iommu group (ordering 1)
  uint32_t nranks, rank;
  uint32_t local_max;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ uint32_t count() const { return *n_ptr; }
  __device__ __forceinline__ Item load(uint32_t i, bool ok) const {
    return ok ? ld_stream(surv + i) : make_uint4(0, 0, 0, 0);
  }
  __device__ __forceinline__ uint32_t key_of(const Item& r) const { return field ? r.y : (r.z & 0xffffu); }
  __device__ __forceinline__ bool pred(const Item& r, uint32_t) const { return key_of(r) % nranks == rank; }
  __device__ __forceinline__ uint32_t prepare(const Item& r) const { return key_of(r); }
  __device__ __forceinline__ void emit(uint32_t pos, const Item&, uint32_t i, uint32_t key) {
    out[pos] = make_uint2(key, i);
    local_max = max(local_max, key);
  }
  __device__ __forceinline__ void tile_epilogue() {}
  __device__ __forceinline__ void finish(uint32_t) {}
  __device__ __forceinline__ uint2 take_maxima() {
    uint2 m = field ? make_uint2(local_max, 0) : make_uint2(0, local_max);
    local_max = 0;
    return m;
  }
};
// dense, order-preserving pack of 8-byte pairs (tile_items pairs of scratch per tile)
__global__ void __launch_bounds__(128) k_pack_pairs(const uint2* __restrict__ ragged,
                                                    const uint32_t* __restrict__ tile_off,
                                                    uint32_t tile_items, uint2* __restrict__ dense) {
  pdl_enter();
  const uint32_t tile = blockIdx.x;
  const uint32_t o0 = tile_off[tile], o1 = tile_off[tile + 1];
  const uint2* src = ragged + (size_t)tile * tile_items;
  for (uint32_t u = threadIdx.x; u < o1 - o0; u += blockDim.x) dense[o0 + u] = src[u];
}

// ---- multi-GPU over peer memory (NVLink, CUDA IPC): compaction fused with the all-gather -----------
// Every rank owns a "gather window" with one region per rank.  k_pack_to_peers is the dense pack of
// the classification AND the collective: each 16-byte survivor is stored into region `rank` of every
// peer's window (P-1 NVLink stores + 1 local) at its final rank-local position.  Counts and completion
// travel as release/acquire flags in the peers' control blocks; nothing returns to the host.
// Two windows alternate by step parity and a consumed-ack per rank protects their reuse.
constexpr int P2P_MAX_RANKS = 16;
struct P2PCtrl {  // lives at the head of every rank's window allocation; written by the peers
  unsigned long long flag[2][P2P_MAX_RANKS];  // [window][src rank] = step whose region is complete
  unsigned long long ack[P2P_MAX_RANKS];      // [rank] = last step that rank has consumed
  uint32_t count[2][P2P_MAX_RANKS];           // [window][src rank] survivors in the region
};
struct P2PPeers {
  uint4* win[P2P_MAX_RANKS];     // window base (of the step's parity) on every rank, peer-mapped
  P2PCtrl* ctrl[P2P_MAX_RANKS];  // control block of every rank, peer-mapped
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr long long P2P_SPIN_LIMIT = 20000000000ll;  // ~10 s of SM clocks, then give up loudly

// before reusing a window: every peer must have consumed the step that used it two steps ago
__global__ void k_p2p_wait_acks(const P2PCtrl* mine, uint32_t P, unsigned long long need, uint32_t* err) {
  pdl_enter();
  const uint32_t q = threadIdx.x;
  if (q >= P) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(&mine->ack[q]) < need) {
    if (clock64() - t0 > P2P_SPIN_LIMIT) {
      atomicExch(err, 1u);
      return;
    }
  }
}
__global__ void __launch_bounds__(128) k_pack_to_peers(const uint4* __restrict__ ragged,
                                                       const uint32_t* __restrict__ tile_off,
                                                       uint32_t tile_items, P2PPeers peers, uint32_t P,
                                                       size_t region_off) {
  pdl_enter();
  const uint32_t tile = blockIdx.x;
  const uint32_t o0 = tile_off[tile], o1 = tile_off[tile + 1];
  const uint4* src = ragged + (size_t)tile * tile_items;
  for (uint32_t u = threadIdx.x; u < o1 - o0; u += blockDim.x) {
    const uint4 v = ld_stream(src + u);
    for (uint32_t q = 0; q < P; q++) peers.win[q][region_off + o0 + u] = v;  // NVLink stores
  }
}
__global__ void k_p2p_signal(P2PPeers peers, uint32_t P, uint32_t rank, uint32_t w, unsigned long long step,
                             const uint32_t* n_local) {
  pdl_enter();
  const uint32_t q = threadIdx.x;
  if (q >= P) return;
  peers.ctrl[q]->count[w][rank] = *n_local;
  __threadfence_system();
  st_release_sys(&peers.ctrl[q]->flag[w][rank], step);
}
// wait for every region of this step, then publish the region bases and the total
__global__ void k_p2p_wait_gather(const P2PCtrl* mine, uint32_t P, uint32_t w, unsigned long long step,
                                  uint32_t* gather_base, ScanCtrl* ctrl, uint32_t* err) {
  pdl_enter();
  __shared__ uint32_t cnt[P2P_MAX_RANKS];
  const uint32_t q = threadIdx.x;
  if (q < P) {
    const long long t0 = clock64();
    bool ok = true;
    while (ld_acquire_sys(&mine->flag[w][q]) != step) {
      if (clock64() - t0 > P2P_SPIN_LIMIT) {
        atomicExch(err, 1u);
        ok = false;
        break;
      }
    }
    cnt[q] = ok ? *((volatile const uint32_t*)&mine->count[w][q]) : 0;
  }
  __syncthreads();
  if (q == 0) {
    uint32_t run = 0;
    for (uint32_t r = 0; r < P; r++) {
      gather_base[r] = run;
      run += cnt[r];
    }
    gather_base[P] = run;
    ctrl->n_surv = run;
  }
}
// window regions (rank order == Walk order) -> the dense survivor array every later kernel uses
__global__ void __launch_bounds__(KVG_BLOCK) k_p2p_copy_regions(const uint4* __restrict__ window, size_t cap,
                                                                const uint32_t* __restrict__ gather_base,
                                                                uint4* __restrict__ dense) {
  pdl_enter();
  const uint32_t q = blockIdx.y;
  const uint32_t b0 = gather_base[q], n = gather_base[q + 1] - b0;
  const uint4* src = window + (size_t)q * cap;
  for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x)
    st_stream(dense + b0 + u, ld_stream(src + u));
}
__global__ void k_p2p_ack(P2PPeers peers, uint32_t P, uint32_t rank, unsigned long long step) {
  pdl_enter();
  const uint32_t q = threadIdx.x;
  if (q < P) st_release_sys(&peers.ctrl[q]->ack[rank], step);
}

// Diagnostic decomposition of the classify kernel (kvg_dev_debug_classify):
//   mode 0  read + predicate + count only (one atomicAdd per tile)
//   mode 1  + write survivors at a TILE-LOCAL base (no cross-tile dependency, output not compact)
//   mode 2  + direct-index name join
template <int THREADS, int ROWS>
__global__ void __launch_bounds__(THREADS) k_debug_classify(const uint4* __restrict__ recs, uint32_t n,
                                                             uint4* __restrict__ out,
                                                             const uint32_t* __restrict__ nv_index,
                                                             uint32_t* counter, int mode) {
  pdl_enter();
  constexpr uint32_t TILE = THREADS * ROWS;
  constexpr uint32_t NW = THREADS / 32;
  __shared__ uint32_t s_wtot[NW], s_woff[NW];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t base = tile * TILE + warp * 32 * ROWS;
  uint4 item[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    uint32_t i = base + k * 32 + lane;
    item[k] = i < n ? ld_stream(recs + i) : make_uint4(0, 0, 0, 0xff00u);
  }
  uint32_t bal[ROWS], aux[ROWS], wtot = 0;
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    bool p = pci_record_alive(item[k]);
    bal[k] = __ballot_sync(KVG_FULL, p);
    wtot += __popc(bal[k]);
    aux[k] = (mode >= 2 && p) ? __ldg(&nv_index[item[k].y >> 16]) : 0u;
  }
  if (lane == 0) s_wtot[warp] = wtot;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NW ? s_wtot[lane] : 0;
    uint32_t wi = warp_incl_sum(w);
    if (lane < NW) s_woff[lane] = wi - w;
    if (lane == NW - 1) atomicAdd(counter, wi);
  }
  if (mode == 0) return;
  __syncthreads();
  uint32_t off = tile * TILE + s_woff[warp];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    if ((bal[k] >> lane) & 1u) {
      uint4 sv = item[k];
      sv.w = aux[k];
      st_stream(out + off + __popc(bal[k] & lanemask_lt()), sv);
    }
    off += __popc(bal[k]);
  }
}

}  // namespace kvg

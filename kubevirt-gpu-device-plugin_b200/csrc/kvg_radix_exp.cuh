// kvg_radix_exp.cuh — experimental replacements for kernels of the radix family (kvg_scan.cuh), each
// selected by an environment variable and OFF by default until it has been measured on a GPU.
// Functional checks without a GPU: tools/emu/ runs this source on the CPU (tests/test_radix_emu.py).
//
//   KVG_TILESCAN=warp   k_radix_tilescan_warp: one WARP per digit row instead of one CTA looping over
//                       digit chunks with a block-wide scan (two __syncthreads per 256 tiles).  At
//                       config 2 a digit row is 169 tile counts: six 32-wide warp scans, no barrier.
//                       The default kernel costs ~12 us per launch there (24 us of the 127 us step).
//   KVG_SCATTER=c       k_radix_scatter_c: every thread OWNS 8 CONSECUTIVE digits instead of one digit per
//                       256-digit chunk, so the two digit scans of a tile (digit totals -> global bases,
//                       per-tile digit counts -> tile-local starts) are ONE block scan each instead of one
//                       per chunk: at 10-bit digits 4 + 2 barriers per tile instead of 17.  Same shared-memory
//                       layout, same ranking, same staged write-out as k_radix_scatter<11>.
#pragma once
#ifndef KVG_HOST_EMU
#include "kvg_scan.cuh"
#define KVG_DYN_SMEM(name, bytes) extern __shared__ __align__(16) uint8_t name[]
#else  // tools/emu/: blocks run one at a time, a static buffer of the launch size stands in
#define KVG_DYN_SMEM(name, bytes) static __attribute__((aligned(16))) uint8_t name[bytes]
#endif

namespace kvg {

constexpr uint32_t TS_WARPS = 8;  // digit rows per CTA

// per digit (and per ordering, blockIdx.y): exclusive scan of that digit's per-tile counts, in place;
// bin_total[digit] <- the digit's total.  Same contract as k_radix_tilescan.
__global__ void __launch_bounds__(TS_WARPS * 32) k_radix_tilescan_warp(RadixArgs2 aa) {
  pdl_enter();
  const RadixArgs a = blockIdx.y ? aa.o[1] : aa.o[0];  // static indices: parameters stay in the constant bank
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const RadixPlan pl = radix_pass(a);
  if (T == 0 || !pl.bits) return;
  const uint32_t digits = (((1u << pl.bits) + KVG_BLOCK - 1) / KVG_BLOCK) * KVG_BLOCK;  // rows the histogram wrote
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t dg = blockIdx.x * TS_WARPS + warp; dg < digits; dg += gridDim.x * TS_WARPS) {
    uint32_t* row = a.tile_hist + (size_t)dg * T;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < T; b += 32) {
      const uint32_t i = b + lane;
      const uint32_t v = i < T ? row[i] : 0;
      const uint32_t incl = warp_incl_sum(v);
      if (i < T) row[i] = carry + incl - v;
      carry += __shfl_sync(KVG_FULL, incl, 31);
    }
    if (lane == 0) a.bin_total[dg] = carry;
  }
}

// ---- KVG_SCATTER=c ---------------------------------------------------------------------------------
constexpr uint32_t SC_DPT = RADIX_MAX_DIGITS / KVG_BLOCK;  // 8 consecutive digits per thread
static_assert(SC_DPT == 8, "one uint4 of u16 counters per thread and warp row");

__global__ void __launch_bounds__(KVG_BLOCK, RadixScatterCfg<RADIX_MAX_BITS>::MIN_CTAS) k_radix_scatter_c(RadixArgs2 aa) {
  using Cfg = RadixScatterCfg<RADIX_MAX_BITS>;
  pdl_enter();
  const RadixArgs a = blockIdx.y ? aa.o[1] : aa.o[0];  // static indices: parameters stay in the constant bank
  const uint32_t n = *a.n_ptr;
  const uint32_t T = (n + C_TILE - 1) / C_TILE;
  const RadixPlan pl = radix_pass(a);
  if (!pl.bits) return;
  const uint32_t dmask = (1u << pl.bits) - 1;
  const uint32_t digits = (((1u << pl.bits) + KVG_BLOCK - 1) / KVG_BLOCK) * KVG_BLOCK;  // rows hist / tilescan wrote
  const uint32_t lane = lane_id(), warp = warp_id(), tid = threadIdx.x;
  const uint32_t d0 = tid * SC_DPT;          // my digits: d0 .. d0 + 7
  const bool mine = d0 < digits;             // (digits is a multiple of 256 >= 8: all or none of my eight)
  KVG_DYN_SMEM(rs_smem, Cfg::SMEM);
  uint16_t (*s_cnt)[Cfg::DIGITS] = reinterpret_cast<uint16_t (*)[Cfg::DIGITS]>(rs_smem);
  uint16_t* s_start = reinterpret_cast<uint16_t*>(rs_smem + Cfg::CNT_BYTES);
  int32_t* s_goff = reinterpret_cast<int32_t*>(rs_smem + Cfg::CNT_BYTES + Cfg::START_BYTES);
  uint2* s_stage = reinterpret_cast<uint2*>(rs_smem + Cfg::CNT_BYTES + Cfg::START_BYTES + Cfg::GOFF_BYTES);
  __shared__ uint32_t scratch[KVG_WARPS + 1];
  uint32_t total;
  uint32_t bin_base[SC_DPT];
  bool have_base = false;
  for (uint32_t tile = blockIdx.x; tile < T; tile += gridDim.x) {
    const uint32_t base = tile * C_TILE + warp * C_WARP_ITEMS;
    // all global loads of the tile up front: the scanned tile counts of my digits, the pairs
    uint32_t tile_prefix[SC_DPT];
#pragma unroll
    for (uint32_t q = 0; q < SC_DPT; q++) tile_prefix[q] = mine ? a.tile_hist[(size_t)(d0 + q) * T + tile] : 0;
    uint2 kv[C_ROWS];
    uint32_t rank[C_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
      uint32_t i = base + k * 32 + lane;
      kv[k] = i < n ? radix_load(a, i) : make_uint2(0, 0);
    }
    __syncthreads();  // previous tile's stage fully written out
    if (mine) {
#pragma unroll
      for (uint32_t w = 0; w < KVG_WARPS; w++) *reinterpret_cast<uint4*>(&s_cnt[w][d0]) = make_uint4(0, 0, 0, 0);
    }
    if (!have_base) {  // exclusive scan of the digit totals: local prefix over my eight + ONE block scan
      uint32_t loc[SC_DPT], sum = 0;
#pragma unroll
      for (uint32_t q = 0; q < SC_DPT; q++) {
        loc[q] = sum;
        sum += mine ? a.bin_total[d0 + q] : 0;
      }
      const uint32_t excl = block_excl_sum(sum, scratch, &total);  // syncs inside
#pragma unroll
      for (uint32_t q = 0; q < SC_DPT; q++) bin_base[q] = excl + loc[q];
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
    {  // my eight digits: exclusive prefix over the warps, totals, then ONE block scan for the tile-local starts
      uint32_t dtot[SC_DPT];
#pragma unroll
      for (uint32_t q = 0; q < SC_DPT; q++) dtot[q] = 0;
      if (mine) {
#pragma unroll
        for (uint32_t w = 0; w < KVG_WARPS; w++) {
          // my eight u16 counters of this warp row are ONE 128-bit word: conflict-free LDS.128 / STS.128
          uint4* row = reinterpret_cast<uint4*>(&s_cnt[w][d0]);
          const uint4 c = *row;
          const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
          uint32_t pw[4];
#pragma unroll
          for (uint32_t h = 0; h < 4; h++) {  // counters 2h (low half) and 2h + 1 (high half)
            pw[h] = (dtot[2 * h] & 0xffffu) | (dtot[2 * h + 1] << 16);
            dtot[2 * h] += cw[h] & 0xffffu;
            dtot[2 * h + 1] += cw[h] >> 16;
          }
          *row = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        }
      }
      uint32_t loc[SC_DPT], sum = 0;
#pragma unroll
      for (uint32_t q = 0; q < SC_DPT; q++) {
        loc[q] = sum;
        sum += dtot[q];
      }
      const uint32_t excl = block_excl_sum(sum, scratch, &total);  // syncs inside
      if (mine) {
#pragma unroll
        for (uint32_t q = 0; q < SC_DPT; q++) {
          const uint32_t lstart = excl + loc[q];
          s_start[d0 + q] = (uint16_t)lstart;
          s_goff[d0 + q] = (int32_t)(bin_base[q] + tile_prefix[q]) - (int32_t)lstart;
        }
      }
      __syncthreads();
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

}  // namespace kvg

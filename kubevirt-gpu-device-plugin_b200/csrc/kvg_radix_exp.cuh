// kvg_radix_exp.cuh — experimental replacements for kernels of the radix family (kvg_scan.cuh), each
// selected by an environment variable and OFF by default until it has been measured on a GPU.
// Functional checks without a GPU: tools/emu/ runs this source on the CPU (tests/test_radix_emu.py).
//
//   KVG_TILESCAN=warp   k_radix_tilescan_warp: one WARP per digit row instead of one CTA looping over
//                       digit chunks with a block-wide scan (two __syncthreads per 256 tiles).  At
//                       config 2 a digit row is 169 tile counts: six 32-wide warp scans, no barrier.
//                       The default kernel costs ~12 us per launch there (24 us of the 127 us step).
#pragma once
#ifndef KVG_HOST_EMU
#include "kvg_scan.cuh"
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

}  // namespace kvg

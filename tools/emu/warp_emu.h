// warp_emu.h — run CUDA kernel SOURCE on the CPU, one OS thread per CUDA thread, blocks one after
// another.  Just enough of the CUDA surface for kernels that are written in the warp-synchronous
// style of csrc/kvg_parse_v2.cuh: full-mask warp collectives, __syncwarp, __syncthreads, global and
// shared atomics, static __shared__ arrays.  Collectives are real rendezvous (std::barrier), so a lane
// that skips one, or lanes that disagree about how many they execute, deadlock here exactly like a
// mis-synchronised kernel misbehaves on the GPU (the harness aborts after a timeout).
//
// Test infrastructure only (tests/test_parse_v2_emu.py); never part of the product.
#pragma once
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct EmuDim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct EmuBlock {
  unsigned n_threads;
  std::barrier<> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> xchg;  // [warp][lane]
  explicit EmuBlock(unsigned n) : n_threads(n), block_bar(n), xchg(n) {
    for (unsigned w = 0; w < (n + 31) / 32; w++) {
      unsigned lanes = (w + 1) * 32 <= n ? 32 : n - w * 32;
      warp_bar.emplace_back(new std::barrier<>(lanes));
    }
  }
};
static thread_local EmuBlock* emu_block = nullptr;

#define __global__
#define __device__ static inline
#define __host__
#define __forceinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__ __restrict
#define KVG_FULL 0xffffffffu
constexpr uint32_t KVG_BLOCK = 256;

static inline uint32_t lane_id() { return threadIdx.x & 31u; }
static inline uint32_t warp_id() { return threadIdx.x >> 5; }
static inline void pdl_enter() {}
static inline uint4 ld_stream(const uint4* p) { return *p; }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }

static inline void __syncthreads() { emu_block->block_bar.arrive_and_wait(); }
static inline void __syncwarp() { emu_block->warp_bar[warp_id()]->arrive_and_wait(); }
// every lane publishes, everybody reads, everybody leaves: two rendezvous per collective
static inline uint64_t emu_exchange(uint64_t mine, uint32_t src_lane) {
  uint64_t* slot = &emu_block->xchg[warp_id() * 32];
  slot[lane_id()] = mine;
  __syncwarp();
  uint64_t got = slot[src_lane & 31u];
  __syncwarp();
  return got;
}
static inline uint32_t __shfl_sync(uint32_t, uint32_t v, uint32_t src) { return (uint32_t)emu_exchange(v, src); }
static inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, uint32_t d) {
  uint32_t l = lane_id();
  return (uint32_t)emu_exchange(v, l >= d ? l - d : l);
}
static inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, uint32_t m) { return (uint32_t)emu_exchange(v, lane_id() ^ m); }
static inline uint32_t __ballot_sync(uint32_t, bool p) {
  uint64_t* slot = &emu_block->xchg[warp_id() * 32];
  slot[lane_id()] = p ? 1 : 0;
  __syncwarp();
  uint32_t m = 0;
  for (uint32_t l = 0; l < 32; l++) m |= (uint32_t)(slot[l] & 1) << l;
  __syncwarp();
  return m;
}
static inline bool __any_sync(uint32_t mask, bool p) { return __ballot_sync(mask, p) != 0; }

// the reductions of kvg_common.cuh, on top of the emulated shuffles
static inline uint32_t warp_sum(uint32_t v) {
  for (uint32_t o = 16; o; o >>= 1) v += __shfl_xor_sync(KVG_FULL, v, o);
  return v;
}
static inline uint32_t warp_min(uint32_t v) {
  for (uint32_t o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(KVG_FULL, v, o));
  return v;
}
static inline uint32_t warp_max(uint32_t v) {
  for (uint32_t o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(KVG_FULL, v, o));
  return v;
}
static inline uint32_t warp_incl_max(uint32_t v) {
  for (uint32_t o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(KVG_FULL, v, o);
    if (lane_id() >= o) v = max(v, t);
  }
  return v;
}

static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicExch(uint32_t* p, uint32_t v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
  uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;  // the old value either way
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}

// kernel<<<grid, block>>>(args): blocks run one after another, threads of a block concurrently
template <class Args>
static void emu_launch(void (*kernel)(Args), unsigned grid, unsigned block, Args args) {
  for (unsigned b = 0; b < grid; b++) {
    EmuBlock blk(block);
    std::vector<std::thread> ts;
    ts.reserve(block);
    for (unsigned t = 0; t < block; t++)
      ts.emplace_back([&, t] {
        emu_block = &blk;
        threadIdx.x = t;
        blockIdx.x = b;
        blockDim.x = block;
        gridDim.x = grid;
        kernel(args);
        // a thread that left early must not strand its warp / block at a later rendezvous
        blk.warp_bar[t >> 5]->arrive_and_drop();
        blk.block_bar.arrive_and_drop();
      });
    for (auto& th : ts) th.join();
  }
}

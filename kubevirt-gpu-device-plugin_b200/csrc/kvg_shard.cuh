// kvg_shard.cuh — the exchange step of the sharded scan (BASELINE.json config 4): one process per GPU,
// records range-sharded in Walk order, results partitioned BY KEY.
//
// Rank r classifies its shard (K3 / K5).  Its survivors stay on rank r (that is the rank's part of
// bdfToIommuMap: concatenating the shards in rank order is Walk order).  For the two group-by maps every
// survivor is sent to the OWNER of its key, owner = key % P, once per ordering:
//   ordering 0  device id (deviceMap)  /  mdev type   (vGpuMap)
//   ordering 1  iommu group (iommuMap) /  parent GPU  (gpuVgpuMap)
// so a rank sends 2 records per survivor whatever P is (constant volume per GPU: weak scaling can hold),
// and receives only the members of the keys it owns.  The first version replicated the whole survivor
// list on every rank: inbound bytes grew with P and efficiency fell to 0.385 at 8 GPUs.
//
// Transport: peer memory (CUDA IPC over NVLink / NVSwitch).  Every rank owns a receive WINDOW
//   [parity w][ordering o][source rank s][cap records]
// mapped by every peer.  The multisplit is stable and its stores ARE the collective:
//   k_shard_count   per 2048-survivor tile: how many survivors go to each (ordering, owner)
//   k_shard_scan    per (ordering, owner) row: exclusive scan over the tiles, totals
//   k_shard_send    every survivor is stored at its final position of region `me` in its owner's window
//                   (ballot ranks: stable), once per ordering; the last CTA to finish publishes the
//                   region counts and a release flag (st.release.sys) in every peer's control block
//   k_shard_gather  waits for the P flags of the step (ld.acquire.sys), concatenates the P regions of each
//                   ordering — source-rank order == Walk order — into the dense OWNED list the orderings
//                   read, reduces the largest owned keys (radix plan), and the last CTA acknowledges the
//                   window parity to every peer
// Nothing returns to the host between the classify launch and the last ordering kernel.  Spin loops carry
// a ~10 s clock bound and raise an error flag that the fetch turns into KVG_ENCCL.
//
// NCCL fallback (no peer access): the survivors are all-gathered with NCCL (grouped broadcasts — NCCL has no
// allgatherv) and the SAME kernels run in local mode on the gathered list: one source, only the records
// this rank owns are kept.
#pragma once
#ifndef KVG_HOST_EMU
#include "kvg_common.cuh"
#include "kvg_order.cuh"
#endif

namespace kvg {

constexpr int SH_MAX_RANKS = 16;
constexpr uint32_t SH_ALL = 0xffffffffu;
constexpr long long SH_SPIN_LIMIT = 20000000000ll;  // ~10 s of SM clocks, then give up loudly

struct ShardCtrl {  // lives at the head of every rank's window allocation; written by the peers
  unsigned long long flag[2][SH_MAX_RANKS];  // [parity][source rank] = step whose regions are complete
  unsigned long long ack[SH_MAX_RANKS];      // [rank] = last step that rank has consumed
  uint32_t count[2][2][SH_MAX_RANKS];        // [parity][ordering][source rank] records in the region
};
struct ShardPeers {
  uint4* win[SH_MAX_RANKS];        // window base (16-byte units) of every rank, peer-mapped
  ShardCtrl* ctrl[SH_MAX_RANKS];   // control block of every rank, peer-mapped
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#ifndef KVG_HOST_EMU
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#ifndef KVG_HOST_EMU
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#endif
}

// record kinds: U = 16-byte units per survivor; keys of ordering 0 / 1
//   PCI  (U = 1)  {addr, iommu_group, device | numa<<16, name_slot}
//   mdev (U = 2)  {uuid[16]}, {parent, type_key | numa<<16, src, pad}
template <int U>
__device__ __forceinline__ uint2 shard_keys(const uint4* rec) {
  if (U == 1) return make_uint2(rec[0].z & 0xffffu, rec[0].y);
  return make_uint2(rec[1].y & 0xffffu, rec[1].x);
}

struct ShardArgs {
  const uint4* list;       // dense survivor list of this rank (NCCL mode: the gathered list)
  const uint32_t* n_ptr;   // its length (device)
  uint32_t* tile_cnt;      // [2 orderings][P][T] per-tile counts -> exclusive offsets (T = tiles the launch covers)
  uint32_t* totals;        // [2][P]
  uint32_t* ticket;        // self-resetting "last CTA" counters: [0] send, [1] gather
  uint32_t T;              // row pitch of tile_cnt
  uint32_t P;              // owners (key % P)
  uint32_t me;             // this rank
  uint32_t only;           // SH_ALL: send to every owner; else keep only records owned by `only` (local mode)
  uint32_t n_src;          // regions per (parity, ordering): P, or 1 in local mode
  uint32_t src;            // my region index in the peers' windows: me, or 0 in local mode
  uint64_t region_cap;     // records per region
  uint32_t parity;
  unsigned long long step;
};

template <int U>
__global__ void __launch_bounds__(KVG_BLOCK) k_shard_count(ShardArgs A) {
  pdl_enter();
  const uint32_t n = *A.n_ptr;
  const uint32_t tile = blockIdx.x;
  __shared__ uint32_t s_cnt[2][SH_MAX_RANKS];
  if (threadIdx.x < 2 * SH_MAX_RANKS) (&s_cnt[0][0])[threadIdx.x] = 0;
  __syncthreads();
  if ((uint64_t)tile * C_TILE < n) {
    const uint32_t lane = lane_id();
    const uint32_t base = tile * C_TILE + warp_id() * C_WARP_ITEMS;
    uint32_t q0[C_ROWS], q1[C_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {  // all loads in flight first
      const uint32_t i = base + k * 32 + lane;
      q0[k] = q1[k] = SH_ALL;
      if (i < n) {
        uint4 rec[U];
#pragma unroll
        for (int u = 0; u < U; u++) rec[u] = A.list[(size_t)i * U + u];
        const uint2 key = shard_keys<U>(rec);
        q0[k] = key.x % A.P;
        q1[k] = key.y % A.P;
      }
    }
    for (uint32_t q = 0; q < A.P; q++) {  // P <= 16 ballots per row and ordering
      uint32_t c0 = 0, c1 = 0;
#pragma unroll
      for (uint32_t k = 0; k < C_ROWS; k++) {
        c0 += __popc(__ballot_sync(KVG_FULL, q0[k] == q));
        c1 += __popc(__ballot_sync(KVG_FULL, q1[k] == q));
      }
      if (lane == 0) {
        if (c0) atomicAdd(&s_cnt[0][q], c0);
        if (c1) atomicAdd(&s_cnt[1][q], c1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * A.P) {
    const uint32_t o = threadIdx.x / A.P, q = threadIdx.x - o * A.P;
    A.tile_cnt[((size_t)o * A.P + q) * A.T + tile] = s_cnt[o][q];
  }
}

// one warp per (ordering, owner) row: exclusive scan over the tiles in place, row total
__global__ void __launch_bounds__(KVG_BLOCK) k_shard_scan(ShardArgs A) {
  pdl_enter();
  const uint32_t n = *A.n_ptr;
  const uint32_t Tu = (n + C_TILE - 1) / C_TILE;
  const uint32_t lane = lane_id();
  for (uint32_t row = blockIdx.x * KVG_WARPS + warp_id(); row < 2 * A.P; row += gridDim.x * KVG_WARPS) {
    uint32_t* r = A.tile_cnt + (size_t)row * A.T;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < Tu; b += 32) {
      const uint32_t i = b + lane;
      const uint32_t v = i < Tu ? r[i] : 0;
      const uint32_t incl = warp_incl_sum(v);
      if (i < Tu) r[i] = carry + incl - v;
      carry += __shfl_sync(KVG_FULL, incl, 31);
    }
    if (lane == 0) A.totals[row] = carry;
  }
}

// window addressing (16-byte units): region (parity, ordering, source) of a window
__device__ __forceinline__ size_t shard_region(const ShardArgs& A, uint32_t o, uint32_t s, int U) {
  return (((size_t)A.parity * 2 + o) * A.n_src + s) * A.region_cap * (size_t)U;
}

template <int U>
__global__ void __launch_bounds__(KVG_BLOCK) k_shard_send(ShardArgs A, ShardPeers peers, const ShardCtrl* mine,
                                                          uint32_t* err) {
  pdl_enter();
  const uint32_t n = *A.n_ptr;
  const uint32_t tile = blockIdx.x;
  __shared__ uint32_t s_wcnt[KVG_WARPS][2][SH_MAX_RANKS];   // per warp: survivors of every (ordering, owner)
  __shared__ uint32_t s_wbase[KVG_WARPS][2][SH_MAX_RANKS];  // per warp: running position in the owner's region
  __shared__ uint32_t s_last;
  const uint32_t lane = lane_id(), warp = warp_id();
  // the window parity is rewritten: every owner must have consumed the step that used it two steps ago
  if (threadIdx.x < A.P && A.step > 2 && A.only == SH_ALL) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine->ack[threadIdx.x]) < A.step - 2) {
      if (clock64() - t0 > SH_SPIN_LIMIT) {
        atomicExch(err, 1u);
        break;
      }
    }
  }
  if ((uint64_t)tile * C_TILE < n) {
    const uint32_t base = tile * C_TILE + warp * C_WARP_ITEMS;
    uint4 rec[C_ROWS][U];
    uint32_t q0[C_ROWS], q1[C_ROWS];
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
      const uint32_t i = base + k * 32 + lane;
      q0[k] = q1[k] = SH_ALL;
      if (i < n) {
#pragma unroll
        for (int u = 0; u < U; u++) rec[k][u] = A.list[(size_t)i * U + u];
        const uint2 key = shard_keys<U>(rec[k]);
        q0[k] = key.x % A.P;
        q1[k] = key.y % A.P;
      }
    }
    // per warp: counts of every (ordering, owner) in my 256 survivors
    for (uint32_t q = 0; q < A.P; q++) {
      uint32_t c0 = 0, c1 = 0;
#pragma unroll
      for (uint32_t k = 0; k < C_ROWS; k++) {
        c0 += __popc(__ballot_sync(KVG_FULL, q0[k] == q));
        c1 += __popc(__ballot_sync(KVG_FULL, q1[k] == q));
      }
      if (lane == 0) {
        s_wcnt[warp][0][q] = c0;
        s_wcnt[warp][1][q] = c1;
      }
    }
    __syncthreads();
    // my warp's base = tile offset (scanned) + the warps in front of me
    if (lane < 2 * A.P) {
      const uint32_t o = lane / A.P, q = lane - o * A.P;
      uint32_t b = A.tile_cnt[((size_t)o * A.P + q) * A.T + tile];
      for (uint32_t w = 0; w < warp; w++) b += s_wcnt[w][o][q];
      s_wbase[warp][o][q] = b;
    }
    __syncwarp();
#pragma unroll
    for (uint32_t k = 0; k < C_ROWS; k++) {
#pragma unroll
      for (uint32_t o = 0; o < 2; o++) {
        const uint32_t q = o ? q1[k] : q0[k];
        // stable rank among the lanes of this row that go to the same owner
        uint32_t before = 0, total_q = 0;
        for (uint32_t t = 0; t < A.P; t++) {
          const uint32_t b = __ballot_sync(KVG_FULL, q == t);
          if (q == t) {
            before = __popc(b & lanemask_lt());
            total_q = __popc(b);
          }
        }
        if (q != SH_ALL) {
          const uint32_t pos = s_wbase[warp][o][q] + before;
          if (A.only == SH_ALL || q == A.only) {
            uint4* dst = peers.win[q] + shard_region(A, o, A.src, U) + (size_t)pos * U;
#pragma unroll
            for (int u = 0; u < U; u++) dst[u] = rec[k][u];  // NVLink store (or local)
          }
        }
        __syncwarp();
        if (q != SH_ALL && before == 0) s_wbase[warp][o][q] += total_q;  // the first lane of each owner advances the base
        __syncwarp();
      }
    }
  }
  // the last CTA to finish publishes the region counts and the release flag to every owner
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&A.ticket[0], 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < A.P) {
    const uint32_t q = threadIdx.x;
    if (A.only == SH_ALL || q == A.only) {
      ShardCtrl* c = peers.ctrl[q];
      c->count[A.parity][0][A.src] = *((volatile uint32_t*)&A.totals[q]);
      c->count[A.parity][1][A.src] = *((volatile uint32_t*)&A.totals[A.P + q]);
      __threadfence_system();
      st_release_sys(&c->flag[A.parity][A.src], A.step);
    }
  }
  if (threadIdx.x == 0) A.ticket[0] = 0;
}

struct GatherArgs {
  const uint4* window;     // my window (16-byte units)
  uint4* owned[2];         // dense owned list of ordering 0 / 1
  uint32_t* n_own;         // [2] -> ScanCtrl::n_own
  uint32_t* max_key;       // [2]: {max_devkey, max_group} of ScanCtrl (pre-zeroed)
};

template <int U>
__global__ void __launch_bounds__(KVG_BLOCK) k_shard_gather(ShardArgs A, GatherArgs G, ShardPeers peers,
                                                            const ShardCtrl* mine, uint32_t* err) {
  pdl_enter();
  __shared__ uint32_t s_cnt[2][SH_MAX_RANKS], s_off[2][SH_MAX_RANKS + 1];
  __shared__ uint32_t s_last;
  const uint32_t tid = threadIdx.x;
  if (tid < A.n_src) {  // wait for every source's regions of this step
    const long long t0 = clock64();
    bool ok = true;
    while (ld_acquire_sys(&mine->flag[A.parity][tid]) != A.step) {
      if (clock64() - t0 > SH_SPIN_LIMIT) {
        atomicExch(err, 1u);
        ok = false;
        break;
      }
    }
    s_cnt[0][tid] = ok ? *((volatile const uint32_t*)&mine->count[A.parity][0][tid]) : 0;
    s_cnt[1][tid] = ok ? *((volatile const uint32_t*)&mine->count[A.parity][1][tid]) : 0;
  }
  __syncthreads();
  if (tid < 2) {
    uint32_t run = 0;
    for (uint32_t s = 0; s < A.n_src; s++) {
      s_off[tid][s] = run;
      run += s_cnt[tid][s];
    }
    s_off[tid][A.n_src] = run;
    if (blockIdx.x == 0) G.n_own[tid] = run;
  }
  __syncthreads();
  // blockIdx.y = ordering; regions in source order == Walk order
  const uint32_t o = blockIdx.y;
  uint32_t mx = 0;
  for (uint32_t s = 0; s < A.n_src; s++) {
    const uint4* src = G.window + shard_region(A, o, s, U);
    uint4* dst = G.owned[o] + (size_t)s_off[o][s] * U;
    const uint32_t cnt = s_cnt[o][s];
    for (uint32_t i = blockIdx.x * KVG_BLOCK + tid; i < cnt; i += gridDim.x * KVG_BLOCK) {
      uint4 rec[U];
#pragma unroll
      for (int u = 0; u < U; u++) rec[u] = ld_stream(src + (size_t)i * U + u);
#pragma unroll
      for (int u = 0; u < U; u++) st_stream(dst + (size_t)i * U + u, rec[u]);
      const uint2 key = shard_keys<U>(rec);
      mx = max(mx, o ? key.y : key.x);
    }
  }
  mx = warp_max(mx);
  if (lane_id() == 0 && mx) atomicMax(&G.max_key[o], mx);
  // the last CTA acknowledges the window parity to every source
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&A.ticket[1], 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  if (tid < A.P && A.only == SH_ALL) st_release_sys(&peers.ctrl[tid]->ack[A.me], A.step);
  if (tid == 0) A.ticket[1] = 0;
}

}  // namespace kvg

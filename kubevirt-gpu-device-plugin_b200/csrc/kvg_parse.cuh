// kvg_parse.cuh — pci.ids on the GPU: getDeviceName / locateVendor of the reference
// (pkg/device_plugin/device_plugin.go:371-438) turned into a build-once table.
//
//   k_pciids_parse     K1  TMA-staged 16 KiB text tiles -> newline flags -> per-line classify ->
//                          vendor context (last-writer look-back across tiles) -> open-addressed
//                          (vendor<<16|device) -> line-offset hash of the lines under vendor 10de,
//                          first line wins (atomicMin)
//   k_pciids_finalize      section bounds of the FIRST "10de" line, bufio.Scanner 64 KiB limit
//   k_pciids_sanitise_lines  K2  name transform of :404-414, one warp per named line of the section
//   k_probe_keys           hash probe for 4-lower-hex keys (the join used by the scans)
//   k_lookup_general       exact prefix semantics of :388-402 for arbitrary key bytes
//   k_sanitise_matches     name transform for the lines found by k_lookup_general
#pragma once
#include "kvg_common.cuh"

namespace kvg {

constexpr uint32_t P_TILE = 8192;             // text bytes owned by one tile
constexpr uint32_t P_HALO = 16;               // bytes after the tile needed to classify its last line
constexpr uint32_t P_STAGE = P_TILE + P_HALO; // one TMA transaction
constexpr uint32_t P_STAGES = 4;
constexpr uint32_t P_SPAN = P_TILE / KVG_BLOCK;  // 32 text bytes per thread
constexpr uint32_t P_WSPAN = 32 * P_SPAN;        // 1 KiB of text per warp
constexpr uint32_t P_WCAP = P_WSPAN + 8;         // worst case: every byte of the warp span is '\n'
constexpr uint32_t P_SMEM = P_STAGES * P_STAGE + 2 * KVG_WARPS * P_WCAP * 2;  // ring + line lists
constexpr uint32_t P_NONE = 0xffffffffu;
constexpr uint64_t P_EMPTY = 0xffffffffffffffffull;
constexpr uint32_t SCAN_TOKEN_MAX = 65536;  // bufio.MaxScanTokenSize

static_assert(P_SPAN == 32, "one 32-bit newline mask per thread");
static_assert(P_TILE + 1 < (1u << 15), "tile-relative line start + 1 must fit 15 bits");

// per-image facts; v_off is filled by K1 (atomicMin), the rest by k_pciids_finalize
struct PciIdsInfo {
  uint32_t v_off;     // offset of the first line with prefix "10de" (locateVendor :424-431)
  uint32_t sec_end;   // first header-type line after it / EOF / scanner failure point
  uint32_t n_entries; // distinct (vendor,device) keys inserted
  uint32_t n_lines;
  uint32_t limit;     // start of the first line bufio.Scanner would reject (>= 64 KiB), or len
  uint32_t overflow;  // set when an insert found no free slot: the host re-parses with a larger table
  uint32_t pad[2];
};

struct ParseArgs {
  const uint8_t* text;  // image 0; image f at text + f*stride
  uint64_t stride;
  uint32_t len;
  uint32_t n_files;
  uint32_t tiles_per_file;
  uint32_t n_tiles;
  uint64_t* tables;  // n_files tables of (cap_mask+1) slots, pre-filled with P_EMPTY
  uint32_t cap_mask;
  uint32_t cap_shift;  // 32 - log2(cap)
  PciIdsInfo* info;    // [n_files]
  uint32_t* tile_first_hdr;  // [n_tiles] file offset of the first header-type line owned, or NONE
  uint32_t* tile_first_nl;   // [n_tiles] file offset of the first / last '\n' in the tile
  uint32_t* tile_last_nl;
  uint64_t* tile_state;  // [n_tiles] vendor-context look-back
  uint32_t epoch;
};

__device__ __forceinline__ uint32_t hash32(uint32_t key, uint32_t shift) {
  return (key * 0x9E3779B1u) >> shift;
}
// double hashing: an odd stride visits every slot of the power-of-two table and avoids the
// primary clustering of linear probing (each extra probe is a serialized L2 atomic round trip)
__device__ __forceinline__ uint32_t hash_step(uint32_t key, uint32_t shift) {
  return ((key * 0x85EBCA6Bu) >> shift) | 1u;
}

// 0x80 in every byte of w that equals '\n' (exact, no cross-byte carries)
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
  uint32_t x = w ^ 0x0A0A0A0Au;
  uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | x | 0x7F7F7F7Fu);
}
// gather the four 0x80 flags into a 4-bit mask (bit i = byte i)
__device__ __forceinline__ uint32_t nl_nibble(uint32_t w) { return (nl_flags(w) * 0x00204081u) >> 28; }
__device__ __forceinline__ uint32_t nl_mask16(const uint4& v) {
  return nl_nibble(v.x) | (nl_nibble(v.y) << 4) | (nl_nibble(v.z) << 8) | (nl_nibble(v.w) << 12);
}

// lower-case hex digit value, or 16 for anything else (upper-case is NOT hex here: the keys the
// reference builds come from sysfs "0x%04x" and the match is a byte compare, :388/:400)
__device__ __forceinline__ uint32_t hexval(uint32_t c) {
  uint32_t d = c - '0';
  uint32_t a = c - 'a';
  return d <= 9 ? d : (a <= 5 ? a + 10 : 16);
}
// parse 4 bytes at p -> (valid<<16)|value
__device__ __forceinline__ uint32_t parse_hex4(const uint8_t* p) {
  uint32_t h0 = hexval(p[0]), h1 = hexval(p[1]), h2 = hexval(p[2]), h3 = hexval(p[3]);
  uint32_t bad = (h0 | h1 | h2 | h3) & 16;
  return bad ? 0u : (0x10000u | (h0 << 12) | (h1 << 8) | (h2 << 4) | h3);
}

// first line wins: slot = key<<32 | line offset, atomicMin on a matching key.
// Probing is bounded by the table size; a full table raises *overflow instead of spinning.
__device__ __forceinline__ bool table_insert(uint64_t* table, uint32_t mask, uint32_t shift,
                                             uint32_t key, uint32_t off, uint32_t* overflow) {
  uint64_t item = ((uint64_t)key << 32) | off;
  uint32_t h = hash32(key, shift) & mask;
  const uint32_t step = hash_step(key, shift);
  for (uint32_t probes = 0;; probes++) {
    if (probes > mask) {
      atomicExch(overflow, 1u);
      return false;
    }
    uint64_t old = atomicCAS((unsigned long long*)&table[h], (unsigned long long)P_EMPTY,
                             (unsigned long long)item);
    if (old == P_EMPTY) return true;
    if ((uint32_t)(old >> 32) == key) {
      if ((uint32_t)old > off) atomicMin((unsigned long long*)&table[h], (unsigned long long)item);
      return false;
    }
    h = (h + step) & mask;
  }
}
__device__ __forceinline__ uint32_t table_probe(const uint64_t* __restrict__ table, uint32_t mask,
                                                uint32_t shift, uint32_t key) {
  uint32_t h = hash32(key, shift) & mask;
  const uint32_t step = hash_step(key, shift);
  for (uint32_t probes = 0; probes <= mask; probes++) {
    uint64_t s = __ldg((const unsigned long long*)&table[h]);
    if (s == P_EMPTY) return P_NONE;
    if ((uint32_t)(s >> 32) == key) return (uint32_t)s;
    h = (h + step) & mask;
  }
  return P_NONE;
}

// ------------------------------------------------------------------------------------------------
// K1.  Persistent co-resident CTAs (grid <= occupancy x SMs), tile = blockIdx + k*gridDim, 4-stage
// TMA ring of 8 KiB text tiles.  A tile owns the lines that START in (a, a+TILE] (plus offset 0 for
// the first tile of an image); 16 halo bytes let it classify a line starting on its last byte.
//
// Two phases per tile, phase 2 running ONE ITERATION LATE so the cross-tile vendor context
// (look-back) is already published when it is needed:
//   phase 1(i)  per thread: 32-byte span -> 32-bit newline mask -> line starts appended to the
//               WARP's list (shuffle prefix); warp-dense pass over the list: first byte tells
//               header-type lines (not '\t', not '#'), their 4-hex vendor is parsed, the warp's
//               last header is reduced.  Warp 0 then scans the 8 warp summaries, publishes the
//               tile's context state and the tile summaries used by k_pciids_finalize.
//   phase 2(i-1) warp-dense over the saved list: 32 lines per round with all lanes converged —
//               classify, intra-round header scan, and ONE batch of hash inserts per round.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_parse(ParseArgs A) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t p_smem[];
  uint8_t* stage_buf = p_smem;                                                      // ring
  uint16_t* lists = reinterpret_cast<uint16_t*>(p_smem + P_STAGES * P_STAGE);       // [2][8][P_WCAP]
  __shared__ __align__(8) uint64_t full_bar[P_STAGES];
  __shared__ uint32_t s_wcnt[2][KVG_WARPS], s_whdr[2][KVG_WARPS], s_wctx[2][KVG_WARPS];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_first_hdr, s_first_nl, s_last_nl;

  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const uint32_t G = gridDim.x, b = blockIdx.x;
  if (b >= A.n_tiles) return;
  const uint32_t my_count = (A.n_tiles - b + G - 1) / G;

  auto issue = [&](uint32_t i) {  // thread 0 only
    if (i >= my_count) return;
    uint32_t t = b + i * G;
    uint32_t f = t / A.tiles_per_file, j = t - f * A.tiles_per_file;
    const uint8_t* src = A.text + (uint64_t)f * A.stride + (uint64_t)j * P_TILE;
    uint32_t st = i % P_STAGES;
    mbar_arrive_expect_tx(&full_bar[st], P_STAGE);
    tma_load_1d(stage_buf + st * P_STAGE, src, P_STAGE, &full_bar[st]);
  };

  if (tid == 0) {
    for (uint32_t s = 0; s < P_STAGES; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
    s_first_hdr = P_NONE;
    s_first_nl = P_NONE;
    s_last_nl = 0;
  }
  __syncthreads();
  if (tid == 0)
    for (uint32_t s = 0; s < P_STAGES; s++) issue(s);

  uint32_t n_new = 0;                   // per-thread tally of first-time inserts
  bool prev_defines = false, prev_first = false;  // warp 0: facts about tile i-1

  for (uint32_t i = 0; i <= my_count; ++i) {
    // ---- warp 0: prefetch the look-back window of tile i-1
    const uint32_t ptile = b + (i - 1) * G;  // meaningful for i > 0
    uint64_t lbw = 0;
    if (i > 0 && warp == 0 && !prev_first) {
      int idx = (int)ptile - 1 - (int)lane;
      lbw = idx >= 0 ? ld_relaxed_u64(&A.tile_state[idx]) : lb_pack(A.epoch, LB_INCLUSIVE, 0);
    }

    // ---- phase 1 of tile i
    if (i < my_count) {
      const uint32_t tile = b + i * G;
      const uint32_t f = tile / A.tiles_per_file, j = tile - f * A.tiles_per_file;
      const uint32_t a = j * P_TILE;
      const uint32_t st = i % P_STAGES;
      mbar_wait(&full_bar[st], (i / P_STAGES) & 1);
      const uint8_t* sm = stage_buf + st * P_STAGE;
      uint16_t* L = lists + ((i & 1) * KVG_WARPS + warp) * P_WCAP;

      const uint32_t sp = tid * P_SPAN;
      uint32_t c0 = (tid >> 2) & 1;  // rotate the two 16-byte chunks: conflict-free LDS.128
      uint4 va = *reinterpret_cast<const uint4*>(sm + sp + c0 * 16);
      uint4 vb = *reinterpret_cast<const uint4*>(sm + sp + (c0 ^ 1) * 16);
      uint32_t ma = nl_mask16(va), mb = nl_mask16(vb);
      uint32_t mask = c0 ? (mb | (ma << 16)) : (ma | (mb << 16));  // bit q: byte sp+q is '\n'
      // bytes at or beyond EOF are padding
      uint32_t pos0 = a + sp;
      uint32_t keep = A.len > pos0 ? A.len - pos0 : 0;  // span bytes inside the file
      if (keep < 32) mask &= keep ? ((1u << keep) - 1) : 0u;
      // a newline that is the last byte of the file starts no line
      uint32_t ls_mask = mask;
      if (keep >= 1 && keep <= 32) ls_mask &= ~(1u << (keep - 1));
      {
        uint32_t l2 = warp_sum((uint32_t)__popc(mask));
        if (lane == 0 && l2) atomicAdd(&A.info[f].n_lines, l2);
      }
      uint32_t fn = mask ? pos0 + (uint32_t)__ffs(mask) - 1 : P_NONE;
      uint32_t lnl = mask ? pos0 + 31 - (uint32_t)__clz(mask) : 0;
      fn = warp_min(fn);
      lnl = warp_max(lnl);

      const uint32_t extra = (j == 0 && tid == 0 && A.len > 0) ? 1u : 0u;  // the line at offset 0
      uint32_t cnt = (uint32_t)__popc(ls_mask) + extra;
      uint32_t incl = warp_incl_sum(cnt);
      uint32_t o = incl - cnt;
      const uint32_t wcnt = __shfl_sync(KVG_FULL, incl, 31);
      if (extra) L[o++] = 0;
      for (uint32_t mm = ls_mask; mm; mm &= mm - 1) L[o++] = (uint16_t)(sp + (uint32_t)__ffs(mm));
      __syncwarp();

      uint32_t my_last = 0, my_first = P_NONE;
      for (uint32_t e = lane; e < wcnt; e += 32) {
        uint32_t p = L[e];
        uint32_t b0 = sm[p];
        if (b0 != '\t' && b0 != '#') {
          uint32_t hv = parse_hex4(sm + p);
          my_last = ((p + 1) << 17) | hv;  // e ascends per lane, so the last assignment wins
          my_first = min(my_first, a + p);
          if (hv == (0x10000u | 0x10deu)) atomicMin(&A.info[f].v_off, a + p);
        }
      }
      my_last = warp_max(my_last);
      my_first = warp_min(my_first);
      if (lane == 0) {
        s_wcnt[i & 1][warp] = wcnt;
        s_whdr[i & 1][warp] = my_last;
        if (my_first != P_NONE) atomicMin(&s_first_hdr, my_first);
        if (fn != P_NONE) {
          atomicMin(&s_first_nl, fn);
          atomicMax(&s_last_nl, lnl);
        }
      }
    }
    __syncthreads();  // A: warp summaries of tile i visible; phase 2 of tile i-2 finished
    if (tid == 0 && i >= 2) issue(i - 2 + P_STAGES);

    if (warp == 0) {
      bool defines = false, first = false;
      if (i < my_count) {
        const uint32_t tile = b + i * G;
        const uint32_t j = tile % A.tiles_per_file;
        uint32_t h = lane < KVG_WARPS ? s_whdr[i & 1][lane] : 0;
        uint32_t hi = warp_incl_max(h);
        uint32_t he = __shfl_up_sync(KVG_FULL, hi, 1);
        if (lane == 0) he = 0;
        if (lane < KVG_WARPS) s_wctx[i & 1][lane] = he;
        uint32_t tile_hdr = __shfl_sync(KVG_FULL, hi, KVG_WARPS - 1);
        defines = tile_hdr != 0;
        first = j == 0;
        if (lane == 0) {
          uint32_t status = (defines || first) ? LB_INCLUSIVE : LB_AGGREGATE;
          st_relaxed_u64(&A.tile_state[tile], lb_pack(A.epoch, status, tile_hdr & 0x1ffffu));
          A.tile_first_hdr[tile] = s_first_hdr;
          A.tile_first_nl[tile] = s_first_nl;
          A.tile_last_nl[tile] = s_last_nl;
          s_first_hdr = P_NONE;
          s_first_nl = P_NONE;
          s_last_nl = 0;
        }
      }
      if (i > 0) {  // vendor context carried into tile i-1
        uint32_t carry = 0;
        if (!prev_first) {
          int look = (int)ptile - 1;
          uint64_t w = lbw;
          for (;;) {
            uint32_t st = lb_status(w, A.epoch);
            uint32_t incl_mask = __ballot_sync(KVG_FULL, st == LB_INCLUSIVE);
            uint32_t inv_mask = __ballot_sync(KVG_FULL, st == LB_INVALID);
            uint32_t fi = incl_mask ? (uint32_t)__ffs(incl_mask) - 1 : 32;
            uint32_t need = fi >= 31 ? KVG_FULL : ((2u << fi) - 1);
            if (!(inv_mask & need)) {
              if (fi < 32) {
                carry = __shfl_sync(KVG_FULL, (uint32_t)w, fi);
                break;
              }
              look -= 32;
            }
            int idx = look - (int)lane;
            w = idx >= 0 ? ld_relaxed_u64(&A.tile_state[idx]) : lb_pack(A.epoch, LB_INCLUSIVE, 0);
          }
          if (!prev_defines && lane == 0)  // pass-through tile: shorten later look-backs
            st_relaxed_u64(&A.tile_state[ptile], lb_pack(A.epoch, LB_INCLUSIVE, carry));
        }
        if (lane == 0) s_carry = carry;
      }
      prev_defines = defines;
      prev_first = first;
    }
    __syncthreads();  // B: s_wctx of tile i and the carry of tile i-1 visible

    // ---- phase 2 of tile i-1: device lines "\t" + 4 lower-hex under a valid vendor -> hash
    if (i > 0) {
      const uint32_t f = ptile / A.tiles_per_file, j = ptile - f * A.tiles_per_file;
      const uint32_t a = j * P_TILE;
      const uint8_t* sm = stage_buf + ((i - 1) % P_STAGES) * P_STAGE;
      const uint16_t* L = lists + (((i - 1) & 1) * KVG_WARPS + warp) * P_WCAP;
      uint64_t* table = A.tables + (uint64_t)f * (A.cap_mask + 1);
      const uint32_t cnt = s_wcnt[(i - 1) & 1][warp];
      const uint32_t c0 = s_wctx[(i - 1) & 1][warp];
      uint32_t running = c0 ? (c0 & 0x1ffffu) : s_carry;  // valid<<16 | vendor
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t e = base + lane;
        uint32_t hval = 0, dv = 0, p = 0;
        if (e < cnt) {
          p = L[e];
          uint32_t b0 = sm[p];
          if (b0 == '\t')
            dv = parse_hex4(sm + p + 1);
          else if (b0 != '#')
            hval = ((e + 1) << 17) | parse_hex4(sm + p);
        }
        uint32_t sc = warp_incl_max(hval);  // last header at or before this line, this round
        uint32_t ctx = sc ? (sc & 0x1ffffu) : running;
        // only vendor 10de is ever looked up (getDeviceName opens the FIRST "10de" section, :424-431):
        // lines under any other vendor never reach the table, which keeps the serialised L2 atomics of
        // the inserts (the parse kernel's former bottleneck) to ~5 % of the device lines
        if (((dv & ctx) & 0x10000u) && (ctx & 0xffffu) == 0x10deu)
          n_new += table_insert(table, A.cap_mask, A.cap_shift,
                                ((ctx & 0xffffu) << 16) | (dv & 0xffffu), a + p, &A.info[f].overflow);
        uint32_t last = __shfl_sync(KVG_FULL, sc, 31);
        if (last) running = last & 0x1ffffu;
      }
      uint32_t e2 = warp_sum(n_new);
      if (lane == 0 && e2) atomicAdd(&A.info[f].n_entries, e2);
      n_new = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One CTA per image: section end and scanner limit.  (Cheap: a few hundred tiles at most.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_finalize(ParseArgs A) {
  pdl_enter();
  const uint32_t f = blockIdx.x;
  const uint8_t* text = A.text + (uint64_t)f * A.stride;
  const uint32_t t0 = f * A.tiles_per_file;
  PciIdsInfo* info = &A.info[f];
  __shared__ uint32_t s_end, s_limit, s_hdr_tile;
  if (threadIdx.x == 0) {
    s_end = A.len;
    s_limit = A.len;
    s_hdr_tile = P_NONE;
  }
  __syncthreads();
  // bufio.Scanner: a line with no '\n' in its first 64 KiB ends the scan with ErrTooLong.  Inside
  // one tile two newlines are < 16 KiB apart, so only a gap that spans tiles can be long: thread
  // per tile, gap = from the previous newline (scan back over newline-free tiles) to my first.
  for (uint32_t t = threadIdx.x; t <= A.tiles_per_file; t += blockDim.x) {
    uint32_t fn;  // first newline at or after tile t (the virtual last tile stands for EOF)
    if (t == A.tiles_per_file) fn = A.len;
    else {
      fn = A.tile_first_nl[t0 + t];
      if (fn == P_NONE) continue;
    }
    uint32_t line_start = 0;
    for (int u = (int)t - 1; u >= 0; u--) {
      if (A.tile_first_nl[t0 + u] != P_NONE) {
        line_start = A.tile_last_nl[t0 + u] + 1;
        break;
      }
    }
    if (line_start < A.len && fn - line_start >= SCAN_TOKEN_MAX) atomicMin(&s_limit, line_start);
  }
  __syncthreads();
  const uint32_t V = info->v_off;
  const uint32_t limit = s_limit;
  if (V == P_NONE || V >= limit) {  // vendor line never reached (:382-385)
    if (threadIdx.x == 0) {
      info->v_off = P_NONE;
      info->sec_end = P_NONE;
      info->limit = limit;
    }
    return;
  }
  // first header-type line after V: inside V's tile by a byte scan, else from the tile summaries
  const uint32_t tv = V == 0 ? 0 : (V - 1) / P_TILE;  // tile that owns the line starting at V
  const uint32_t tile_end = min(A.len, (tv + 1) * P_TILE + 1);  // line starts owned: <= (tv+1)*TILE
  for (uint32_t p = V + 1 + threadIdx.x; p < tile_end; p += blockDim.x) {
    if (text[p - 1] == '\n') {
      uint8_t b0 = text[p];
      if (b0 != '\t' && b0 != '#') atomicMin(&s_end, p);
    }
  }
  for (uint32_t t = tv + 1 + threadIdx.x; t < A.tiles_per_file; t += blockDim.x)
    if (A.tile_first_hdr[t0 + t] != P_NONE) atomicMin(&s_hdr_tile, t);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t e = s_end;
    if (e == A.len && s_hdr_tile != P_NONE) e = A.tile_first_hdr[t0 + s_hdr_tile];
    info->sec_end = min(e, limit);
    info->limit = limit;
  }
}

// ------------------------------------------------------------------------------------------------
// The name transform (:404-414) as one left-to-right pass.
//   TrimSpace (unicode.IsSpace, with Go's ASCII fast path) -> ToUpper (simple mapping: only U+0131
//   and U+017F land in ASCII) -> '/'->'_' -> '.'->'_' -> RE2 \s+ ([\t\n\f\r ]) -> '_' ->
//   delete [^a-zA-Z0-9_.]+.   Output alphabet is [A-Z0-9_].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool d_ascii_space(uint32_t c) {
  return c == ' ' || (c >= 9 && c <= 13);
}
__device__ __forceinline__ bool d_re2_space(uint32_t c) {
  return c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r';
}
__device__ uint32_t d_decode_rune(const uint8_t* p, uint32_t n, uint32_t* width) {
  *width = 1;
  if (n == 0) {
    *width = 0;
    return 0xFFFD;
  }
  uint32_t b0 = p[0];
  if (b0 < 0x80) return b0;
  if (b0 < 0xC2 || b0 > 0xF4) return 0xFFFD;
  if (b0 < 0xE0) {
    if (n < 2 || (p[1] & 0xC0) != 0x80) return 0xFFFD;
    *width = 2;
    return ((b0 & 0x1F) << 6) | (p[1] & 0x3F);
  }
  if (b0 < 0xF0) {
    if (n < 3) return 0xFFFD;
    uint32_t lo = b0 == 0xE0 ? 0xA0 : 0x80, hi = b0 == 0xED ? 0x9F : 0xBF;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80) return 0xFFFD;
    *width = 3;
    return ((b0 & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F);
  }
  if (n < 4) return 0xFFFD;
  uint32_t lo = b0 == 0xF0 ? 0x90 : 0x80, hi = b0 == 0xF4 ? 0x8F : 0xBF;
  if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) return 0xFFFD;
  *width = 4;
  return ((b0 & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}
__device__ uint32_t d_decode_last_rune(const uint8_t* p, uint32_t n, uint32_t* width) {
  *width = 1;
  if (n == 0) {
    *width = 0;
    return 0xFFFD;
  }
  int end = (int)n, start = end - 1;
  if (p[start] < 0x80) return p[start];
  int lim = end - 4 < 0 ? 0 : end - 4;
  for (start--; start >= lim; start--)
    if ((p[start] & 0xC0) != 0x80) break;
  if (start < 0) start = 0;
  uint32_t w;
  uint32_t r = d_decode_rune(p + start, (uint32_t)(end - start), &w);
  if (start + (int)w != end) return 0xFFFD;
  *width = w;
  return r;
}
__device__ __forceinline__ bool d_is_space_rune(uint32_t r) {
  if (r < 0x80) return d_ascii_space(r);
  return r == 0x85 || r == 0xA0 || r == 0x1680 || (r >= 0x2000 && r <= 0x200A) || r == 0x2028 ||
         r == 0x2029 || r == 0x202F || r == 0x205F || r == 0x3000;
}
__device__ void d_trim_space(const uint8_t* s, uint32_t n, uint32_t* pa, uint32_t* pb) {
  uint32_t a = 0, b = n;
  int uni = 0;
  for (; a < n; a++) {
    uint32_t c = s[a];
    if (c >= 0x80) {
      uni = 1;
      break;
    }
    if (!d_ascii_space(c)) break;
  }
  if (!uni) {
    for (; b > a; b--) {
      uint32_t c = s[b - 1];
      if (c >= 0x80) {
        uni = 2;
        break;
      }
      if (!d_ascii_space(c)) break;
    }
  }
  if (uni == 1) {
    while (a < n) {
      uint32_t w;
      uint32_t r = d_decode_rune(s + a, n - a, &w);
      if (!d_is_space_rune(r)) break;
      a += w;
    }
    b = n;
  }
  if (uni) {
    while (b > a) {
      uint32_t w;
      uint32_t r = d_decode_last_rune(s + a, b - a, &w);
      if (!d_is_space_rune(r)) break;
      b -= w;
    }
  }
  *pa = a;
  *pb = b;
}
// returns the output length; writes at most `cap` bytes (the caller sizes cap >= n)
__device__ uint32_t d_sanitise_name(const uint8_t* s, uint32_t n, uint8_t* out, uint32_t cap) {
  uint32_t a, b;
  d_trim_space(s, n, &a, &b);
  uint32_t o = 0;
  bool prev_ws = false;
  for (uint32_t i = a; i < b;) {
    uint32_t c = s[i];
    uint32_t e = 0;
    if (c < 0x80) {
      i++;
      if (d_re2_space(c)) {
        if (!prev_ws) e = '_';
        prev_ws = true;
      } else {
        prev_ws = false;
        if (c >= 'a' && c <= 'z')
          e = c - 32;
        else if ((c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_')
          e = c;
        else if (c == '/' || c == '.')
          e = '_';
      }
    } else {
      prev_ws = false;
      if (c == 0xC4 && i + 1 < b && s[i + 1] == 0xB1) {
        e = 'I';
        i += 2;
      } else if (c == 0xC5 && i + 1 < b && s[i + 1] == 0xBF) {
        e = 'S';
        i += 2;
      } else {
        i++;
      }
    }
    if (e) {
      if (o < cap) out[o] = (uint8_t)e;
      o++;
    }
  }
  return o;
}


// hash probe for canonical keys: keys[i] = (vendor<<16)|device -> pool slot (off - V) or P_NONE
__device__ __forceinline__ uint32_t probe_name_slot(const uint64_t* __restrict__ table,
                                                    uint32_t mask, uint32_t shift,
                                                    const PciIdsInfo* __restrict__ info,
                                                    uint32_t device) {
  uint32_t off = table_probe(table, mask, shift, (0x10deu << 16) | device);
  uint32_t V = info->v_off, E = info->sec_end;
  return (off != P_NONE && V != P_NONE && off > V && off < E) ? off - V : P_NONE;
}
// nv_index[d] for all 65,536 device ids of vendor 10de + the list of line offsets that have a name
// (order irrelevant: each entry is sanitised independently by k_pciids_sanitise_lines)
__global__ void k_nv_index(const uint64_t* __restrict__ table, uint32_t mask, uint32_t shift,
                           const PciIdsInfo* __restrict__ info, uint32_t* __restrict__ nv_index,
                           uint32_t* __restrict__ line_list, uint32_t* __restrict__ line_count) {
  pdl_enter();
  uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t slot = probe_name_slot(table, mask, shift, info, d & 0xffffu);
  nv_index[d] = slot;
  uint32_t hit = __ballot_sync(KVG_FULL, slot != P_NONE);
  if (hit) {
    uint32_t base = 0;
    if (lane_id() == 0) base = atomicAdd(line_count, (uint32_t)__popc(hit));
    base = __shfl_sync(KVG_FULL, base, 0);
    if (slot != P_NONE) line_list[base + __popc(hit & lanemask_lt())] = slot;  // offset - V
  }
}

// K2, one warp per NVIDIA device line: lanes classify 32 characters at a time.  ASCII-only lines
// (every line of the shipped file's NVIDIA block) take the lane-parallel path; a line with any byte
// >= 0x80 falls back to the exact serial routine (Unicode TrimSpace / ToUpper rules).
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_sanitise_lines(const uint8_t* __restrict__ text,
                                                                     uint32_t len,
                                                                     const PciIdsInfo* __restrict__ info,
                                                                     const uint32_t* __restrict__ line_list,
                                                                     const uint32_t* __restrict__ line_count,
                                                                     uint8_t* __restrict__ pool) {
  pdl_enter();
  const uint32_t V = info->v_off;
  if (V == P_NONE) return;
  const uint32_t n_lines = *line_count;
  const uint32_t lane = lane_id();
  for (uint32_t w = blockIdx.x * KVG_WARPS + warp_id(); w < n_lines; w += gridDim.x * KVG_WARPS) {
    const uint32_t slot = line_list[w];
    const uint32_t s0 = V + slot + 5;  // first byte after "\t" + 4 hex
    uint8_t* out = pool + slot;
    // line end and ASCII test
    uint32_t n = 0;
    bool ascii = true;
    for (uint32_t c0 = 0;; c0 += 32) {
      uint32_t pos = s0 + c0 + lane;
      uint32_t c = pos < len ? text[pos] : '\n';
      uint32_t nl = __ballot_sync(KVG_FULL, c == '\n');
      uint32_t hi = __ballot_sync(KVG_FULL, c >= 0x80);
      if (nl) {
        uint32_t k = (uint32_t)__ffs(nl) - 1;
        n = c0 + k;
        if (hi & ((1u << k) - 1)) ascii = false;
        break;
      }
      if (hi) ascii = false;
    }
    if (!ascii) {
      if (lane == 0) {
        uint32_t m = d_sanitise_name(text + s0, n, out + 2, n);
        out[0] = (uint8_t)(m & 0xff);
        out[1] = (uint8_t)(m >> 8);
      }
      continue;
    }
    // trim: first / last byte that is not ASCII white space (TrimSpace; includes \v)
    uint32_t first = n, last = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {
      uint32_t i = c0 + lane;
      bool ns = i < n && !d_ascii_space(text[s0 + i]);
      uint32_t b = __ballot_sync(KVG_FULL, ns);
      if (b) {
        if (first == n) first = c0 + (uint32_t)__ffs(b) - 1;
        last = c0 + 32 - (uint32_t)__clz(b);  // one past the last non-space
      }
    }
    uint32_t o = 0;
    if (first < last) {
      for (uint32_t c0 = first; c0 < last; c0 += 32) {
        uint32_t i = c0 + lane;
        uint32_t e = 0;
        if (i < last) {
          uint32_t c = text[s0 + i];
          if (d_re2_space(c)) {
            if (!(i > first && d_re2_space(text[s0 + i - 1]))) e = '_';  // one '_' per run
          } else if (c >= 'a' && c <= 'z') {
            e = c - 32;
          } else if ((c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_') {
            e = c;
          } else if (c == '/' || c == '.') {
            e = '_';
          }
        }
        uint32_t b = __ballot_sync(KVG_FULL, e != 0);
        if (e) out[2 + o + __popc(b & lanemask_lt())] = (uint8_t)e;
        o += __popc(b);
      }
    }
    if (lane == 0) {
      out[0] = (uint8_t)(o & 0xff);
      out[1] = (uint8_t)(o >> 8);
    }
  }
}

__global__ void k_probe_keys(const uint64_t* __restrict__ table, uint32_t mask, uint32_t shift,
                             const PciIdsInfo* __restrict__ info, uint32_t first, uint32_t count,
                             uint32_t* __restrict__ slots) {
  pdl_enter();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) slots[i] = probe_name_slot(table, mask, shift, info, (first + i) & 0xffffu);
}

// every line start of the NVIDIA section that begins with '\t' (candidates of the prefix match),
// collected once per table load; order is irrelevant (matches are reduced with atomicMin)
__global__ void __launch_bounds__(KVG_BLOCK) k_section_lines(const uint8_t* __restrict__ text,
                                                             const PciIdsInfo* __restrict__ info,
                                                             uint32_t* __restrict__ lines,
                                                             uint32_t* __restrict__ n_lines, uint32_t cap) {
  pdl_enter();
  const uint32_t V = info->v_off, E = info->sec_end;
  if (V == P_NONE) return;
  for (uint32_t b = V + 1 + blockIdx.x * blockDim.x + threadIdx.x; b < E; b += gridDim.x * blockDim.x) {
    if (text[b - 1] == '\n' && text[b] == '\t') {
      uint32_t k = atomicAdd(n_lines, 1u);
      if (k < cap) lines[k] = b;
    }
  }
}

// Exact :388-402 for arbitrary keys.  grid.y = key index; threads sweep the section's '\t' lines
// (E already excludes everything past the first non-'\t' non-'#' line; '#' lines are not listed).
//   match = HasPrefix(line, "\t"+key) on the token bufio.ScanLines returns (one trailing '\r'
//   dropped); the first matching line wins (atomicMin on its offset).
__global__ void __launch_bounds__(KVG_BLOCK) k_lookup_general(const uint8_t* __restrict__ text,
                                                              uint32_t len,
                                                              const uint32_t* __restrict__ lines,
                                                              const uint32_t* __restrict__ n_lines_ptr,
                                                              const uint8_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ key_off,
                                                              uint32_t* __restrict__ match_off) {
  pdl_enter();
  const uint32_t n_lines = *n_lines_ptr;
  const uint32_t kidx = blockIdx.y;
  const uint8_t* key = keys + key_off[kidx];
  const uint32_t klen = key_off[kidx + 1] - key_off[kidx];
  for (uint32_t li = blockIdx.x * blockDim.x + threadIdx.x; li < n_lines; li += gridDim.x * blockDim.x) {
    const uint32_t b = lines[li];
    bool ok = true;
    for (uint32_t k = 0; k < klen && ok; k++) {
      uint32_t pos = b + 1 + k;
      if (pos >= len) {
        ok = false;
        break;
      }
      uint8_t c = text[pos];
      if (c == '\n' || c != key[k]) ok = false;
      // a '\r' that is the last byte of the line is not part of the token
      else if (c == '\r' && (pos + 1 >= len || text[pos + 1] == '\n'))
        ok = false;
    }
    if (ok) atomicMin(&match_off[kidx], b);
  }
}
// one thread per key: sanitise the remainder of the matched line into out[k*cap ...]
__global__ void k_sanitise_matches(const uint8_t* __restrict__ text, uint32_t len,
                                   const uint32_t* __restrict__ key_off,
                                   const uint32_t* __restrict__ match_off, uint32_t n_keys,
                                   uint8_t* __restrict__ out, uint32_t cap,
                                   uint32_t* __restrict__ out_len) {
  pdl_enter();
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_keys) return;
  uint32_t b = match_off[k];
  if (b == P_NONE) {
    out_len[k] = 0;
    return;
  }
  uint32_t s = b + 1 + (key_off[k + 1] - key_off[k]);
  uint32_t e = s;
  while (e < len && text[e] != '\n') e++;
  if (e > s && text[e - 1] == '\r') e--;  // ScanLines dropCR (TrimSpace would drop it anyway)
  out_len[k] = d_sanitise_name(text + s, e - s, out + (size_t)k * cap, cap);
}

}  // namespace kvg

// kvg_parse_v2.cuh — barrier-free variant of the pci.ids parse (selected with KVG_PARSE=v2).
//
// STATUS: EXPERIMENTAL.  Written at the end of round 1 after the GPU budget was spent, so these kernels
// have NOT run on a GPU yet.  What has been done instead:
//   * the decomposition is pinned against the oracle by a Python model (tools/parse_v2_model.py,
//     tests/test_parse_v2_model.py);
//   * THIS SOURCE FILE is compiled for the CPU and executed under a warp emulator (tools/emu/: one OS
//     thread per CUDA thread, collectives as rendezvous, real atomics, poisoned scratch buffers) and
//     compared with the oracle and the model on the shipped pci.ids, the grammar fuzz, span-edge and
//     scanner-limit cases and a worst-case pending list (tests/test_parse_v2_emu.py);
//   * it compiles for sm_100a: 64 registers, 16.5 KB shared memory, no barrier.
// The default parse stays k_pciids_parse (kvg_parse.cuh).  First GPU call of the next round:
// tools/round2_first_run.sh (parity with KVG_PARSE=v2, then the v1 / v2 A/B of bench.py).
//
// Why: ncu on the 128-image launch of k_pciids_parse shows "barrier" as the top stall (5.8 warps per
// issue-active): every 8 KiB tile is a choreography of five __syncthreads between the line lists of
// phase 1, the look-back of warp 0 and the deferred phase 2.  Here a WARP owns a 4 KiB span end to end
// and never waits for another warp:
//
//   k_pciids_scan_v2     span -> newline masks (registers) -> header lines -> warp scan of the vendor
//                        context inside the span -> device lines under a 10de header of the span are
//                        inserted at once; device lines BEFORE the span's first header are written to
//                        the span's pending list (their context is decided by earlier spans); the
//                        span publishes its summaries (first/last newline, first header, last header)
//   k_pciids_resolve_v2  span with pending lines -> nearest earlier span that has a header (plain
//                        loads: the scan kernel has completed) -> insert if that header is 10de
//   k_pciids_finalize    unchanged (tile = span)
//
// No shared-memory barrier, no cross-CTA wait, no epoch-tagged state: everything a kernel reads was
// written by a COMPLETED kernel.  Semantics are those of kvg_parse.cuh (reference
// device_plugin.go:371-438): only device lines whose vendor context is a valid lower-hex "10de" enter
// the table, first line wins.
#pragma once
#ifndef KVG_HOST_EMU  // tools/emu/ compiles this file for the CPU on top of warp_emu.h instead
#include "kvg_parse.cuh"
#endif

namespace kvg {

constexpr uint32_t V2_SPAN = 4096;                // text bytes owned by one warp
constexpr uint32_t V2_ROWS = V2_SPAN / 512;       // 8 rows of 32 lanes x 16 bytes
constexpr uint32_t V2_WARPS = 4;                  // spans per CTA (independent of each other)
constexpr uint32_t V2_SMEM_SPAN = V2_SPAN + 16;   // span + halo: a line starting on the last byte is classified
constexpr uint32_t V2_PEND_CAP = 688;             // >= 4096 / 6 + 2: a device line is at least "\tabcd\n"
static_assert(V2_ROWS * 512 == V2_SPAN, "rows cover the span");
static_assert(P_TILE % V2_SPAN == 0, "the padded text (multiple of P_TILE + halo) covers whole spans");

struct ParseV2Args {
  const uint8_t* text;  // image 0; image f at text + f*stride
  uint64_t stride;
  uint32_t len;
  uint32_t n_files;
  uint32_t spans_per_file;
  uint32_t n_spans;
  uint64_t* tables;
  uint32_t cap_mask, cap_shift;
  PciIdsInfo* info;
  uint32_t* span_first_hdr;  // [n_spans] file offset of the first header-type line owned, or NONE
  uint32_t* span_first_nl;   // [n_spans] file offset of the first / last '\n' in the span
  uint32_t* span_last_nl;
  uint32_t* span_state;      // [n_spans] 0: no header in the span, else 0x80000000 | valid<<16 | vendor
  uint32_t* pend_cnt;        // [n_spans]
  uint32_t* pending;         // [n_spans][V2_PEND_CAP]  device<<16 | (line offset - span base)
};

// one line start at span-relative offset p (0 .. 4096): header key or device id
//   header-type line (first byte neither '\t' nor '#'): (p+1)<<17 | valid<<16 | vendor   (never 0)
__device__ __forceinline__ uint32_t v2_header_key(const uint8_t* sm, uint32_t p) {
  return ((p + 1) << 17) | parse_hex4(sm + p);
}

__global__ void __launch_bounds__(V2_WARPS * 32) k_pciids_scan_v2(ParseV2Args A) {
  pdl_enter();
  __shared__ __align__(16) uint8_t s_text[V2_WARPS][V2_SMEM_SPAN];
  __shared__ uint32_t s_pend[V2_WARPS];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t span = blockIdx.x * V2_WARPS + warp;
  if (span >= A.n_spans) return;  // warp-uniform; no block-wide barrier exists in this kernel
  const uint32_t f = span / A.spans_per_file, j = span - f * A.spans_per_file;
  const uint32_t a = j * V2_SPAN;
  const uint8_t* src = A.text + (uint64_t)f * A.stride + a;
  uint8_t* sm = s_text[warp];

  // ---- the span: 8 coalesced 512-byte rows, all in flight, kept in registers for the newline masks
  uint4 v[V2_ROWS];
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) v[r] = ld_stream(reinterpret_cast<const uint4*>(src + r * 512) + lane);
  uint4 halo = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
  if (lane == 0) halo = ld_stream(reinterpret_cast<const uint4*>(src + V2_SPAN));
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) *reinterpret_cast<uint4*>(sm + r * 512 + lane * 16) = v[r];
  if (lane == 0) {
    *reinterpret_cast<uint4*>(sm + V2_SPAN) = halo;
    s_pend[warp] = 0;
  }
  __syncwarp();

  // ---- newline masks; bytes at or beyond EOF are padding; a '\n' that is the last byte starts no line
  uint32_t ls[V2_ROWS];
  uint32_t nl_total = 0, fn = P_NONE, lnl = 0;
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) {
    const uint32_t pos0 = a + r * 512 + lane * 16;
    uint32_t mask = nl_mask16(v[r]);
    const uint32_t keep = A.len > pos0 ? A.len - pos0 : 0;
    if (keep < 16) mask &= keep ? ((1u << keep) - 1) : 0u;
    uint32_t lsm = mask;
    if (keep >= 1 && keep <= 16) lsm &= ~(1u << (keep - 1));
    nl_total += (uint32_t)__popc(mask);
    if (mask) {
      fn = min(fn, pos0 + (uint32_t)__ffs(mask) - 1);
      lnl = max(lnl, pos0 + 31 - (uint32_t)__clz(mask));
    }
    ls[r] = lsm;
  }
  nl_total = warp_sum(nl_total);
  fn = warp_min(fn);
  lnl = warp_max(lnl);
  const bool extra = j == 0 && lane == 0 && A.len > 0;  // the line at offset 0 of the image

  // ---- loop 1: header-type lines.  hk[r] = key of the LAST header among this lane's lines of row r
  uint32_t hk[V2_ROWS];
  uint32_t first_hdr = P_NONE;
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) {
    uint32_t k = 0;
    if (r == 0 && extra) {
      const uint32_t b0 = sm[0];
      if (b0 != '\t' && b0 != '#') {
        k = v2_header_key(sm, 0);
        first_hdr = min(first_hdr, 0u);
        if ((k & 0x1ffffu) == (0x10000u | 0x10deu)) atomicMin(&A.info[f].v_off, a);
      }
    }
    for (uint32_t mm = ls[r]; mm; mm &= mm - 1) {
      const uint32_t p = r * 512 + lane * 16 + (uint32_t)__ffs(mm);  // newline position + 1
      const uint32_t b0 = sm[p];
      if (b0 != '\t' && b0 != '#') {
        k = v2_header_key(sm, p);  // bits ascend: the last assignment is the last header
        first_hdr = min(first_hdr, p);
        if ((k & 0x1ffffu) == (0x10000u | 0x10deu)) atomicMin(&A.info[f].v_off, a + p);
      }
    }
    hk[r] = k;
  }
  first_hdr = warp_min(first_hdr);

  // ---- vendor context in front of every (row, lane): keys ascend with position, so max == latest
  uint32_t cb[V2_ROWS];
  uint32_t carry = 0;
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) {
    uint32_t before = carry;
    if (__any_sync(KVG_FULL, hk[r] != 0)) {
      const uint32_t incl = warp_incl_max(hk[r]);
      uint32_t excl = __shfl_up_sync(KVG_FULL, incl, 1);
      if (lane == 0) excl = 0;
      before = max(carry, excl);
      carry = max(carry, __shfl_sync(KVG_FULL, incl, 31));
    }
    cb[r] = before;
  }

  // ---- loop 2: device lines ("\t" + 4 lower hex).  Only lanes that can matter walk their lines again:
  // context unknown (pending), context 10de, or a header of their own in this row
  uint32_t n_new = 0;
  uint64_t* table = A.tables + (uint64_t)f * (A.cap_mask + 1);
#pragma unroll
  for (uint32_t r = 0; r < V2_ROWS; r++) {
    const bool has_lines = ls[r] != 0 || (r == 0 && extra);
    if (!has_lines) continue;
    uint32_t ctx = cb[r];
    if (ctx != 0 && (ctx & 0x1ffffu) != (0x10000u | 0x10deu) && hk[r] == 0) continue;
    bool first_iter = r == 0 && extra;
    uint32_t mm = ls[r];
    while (first_iter || mm) {
      uint32_t p;
      if (first_iter) {
        p = 0;
        first_iter = false;
      } else {
        p = r * 512 + lane * 16 + (uint32_t)__ffs(mm);
        mm &= mm - 1;
      }
      const uint32_t b0 = sm[p];
      if (b0 == '\t') {
        const uint32_t dv = parse_hex4(sm + p + 1);
        if (dv & 0x10000u) {
          if (ctx == 0) {
            const uint32_t at = atomicAdd(&s_pend[warp], 1u);
            if (at < V2_PEND_CAP) A.pending[(size_t)span * V2_PEND_CAP + at] = ((dv & 0xffffu) << 16) | p;
          } else if ((ctx & 0x1ffffu) == (0x10000u | 0x10deu)) {
            n_new += table_insert(table, A.cap_mask, A.cap_shift, (0x10deu << 16) | (dv & 0xffffu), a + p,
                                  &A.info[f].overflow);
          }
        }
      } else if (b0 != '#') {
        ctx = v2_header_key(sm, p);
      }
    }
  }
  __syncwarp();
  n_new = warp_sum(n_new);
  if (lane == 0) {
    if (nl_total) atomicAdd(&A.info[f].n_lines, nl_total);
    if (n_new) atomicAdd(&A.info[f].n_entries, n_new);
    A.span_first_nl[span] = fn;
    A.span_last_nl[span] = lnl;
    A.span_first_hdr[span] = first_hdr == P_NONE ? P_NONE : a + first_hdr;
    A.span_state[span] = carry ? (0x80000000u | (carry & 0x1ffffu)) : 0u;
    A.pend_cnt[span] = min(s_pend[warp], V2_PEND_CAP);
  }
}

// one warp per span: decide the pending device lines from the nearest earlier header of the image
__global__ void __launch_bounds__(V2_WARPS * 32) k_pciids_resolve_v2(ParseV2Args A) {
  pdl_enter();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t span = blockIdx.x * V2_WARPS + warp;
  if (span >= A.n_spans) return;
  const uint32_t cnt = A.pend_cnt[span];
  if (cnt == 0) return;
  const uint32_t f = span / A.spans_per_file, j = span - f * A.spans_per_file;
  uint32_t ctx = 0;  // no header before this span: no vendor context, nothing is inserted
  for (int base = (int)j - 1; base >= 0; base -= 32) {
    const int u = base - (int)lane;  // lane 0 looks at the nearest span
    const uint32_t st = u >= 0 ? A.span_state[f * A.spans_per_file + (uint32_t)u] : 0u;
    const uint32_t m = __ballot_sync(KVG_FULL, st != 0);
    if (m) {
      ctx = __shfl_sync(KVG_FULL, st, (uint32_t)__ffs(m) - 1);
      break;
    }
  }
  if ((ctx & 0x1ffffu) != (0x10000u | 0x10deu)) return;
  uint64_t* table = A.tables + (uint64_t)f * (A.cap_mask + 1);
  uint32_t n_new = 0;
  for (uint32_t e = lane; e < cnt; e += 32) {
    const uint32_t w = A.pending[(size_t)span * V2_PEND_CAP + e];
    n_new += table_insert(table, A.cap_mask, A.cap_shift, (0x10deu << 16) | (w >> 16), j * V2_SPAN + (w & 0xffffu),
                          &A.info[f].overflow);
  }
  n_new = warp_sum(n_new);
  if (lane == 0 && n_new) atomicAdd(&A.info[f].n_entries, n_new);
}

// k_pciids_finalize with tile = span (the default kernel's finalize is left untouched on purpose: it is
// the GPU-verified one).  ParseArgs carries the span summaries in its tile arrays.
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_finalize_v2(ParseArgs A) {
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
  const uint32_t tv = V == 0 ? 0 : (V - 1) / V2_SPAN;  // tile that owns the line starting at V
  const uint32_t tile_end = min(A.len, (tv + 1) * V2_SPAN + 1);  // line starts owned: <= (tv+1)*TILE
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

}  // namespace kvg

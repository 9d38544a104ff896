// kvg_parse.cuh — pci.ids on the GPU: getDeviceName / locateVendor of the reference
// (pkg/device_plugin/device_plugin.go:371-438) turned into a build-once table.
//
//   k_pciids_parse     K1  TMA-staged 16 KiB text tiles -> newline flags -> per-line classify ->
//                          vendor context (last-writer look-back across tiles) -> open-addressed
//                          (vendor<<16|device) -> line-offset hash, first line wins (atomicMin)
//   k_pciids_finalize      section bounds of the FIRST "10de" line, bufio.Scanner 64 KiB limit
//   k_pciids_sanitise  K2  name transform of :404-414 for every candidate line of the section
//   k_probe_keys           hash probe for 4-lower-hex keys (the join used by the scans)
//   k_lookup_general       exact prefix semantics of :388-402 for arbitrary key bytes
//   k_sanitise_matches     name transform for the lines found by k_lookup_general
#pragma once
#include "kvg_common.cuh"

namespace kvg {

constexpr uint32_t P_TILE = 16384;            // text bytes owned by one tile
constexpr uint32_t P_HALO = 16;               // bytes after the tile needed to classify its last line
constexpr uint32_t P_STAGE = P_TILE + P_HALO; // one TMA transaction
constexpr uint32_t P_STAGES = 3;
constexpr uint32_t P_SPAN = P_TILE / KVG_BLOCK;  // 64 text bytes per thread
constexpr uint32_t P_NONE = 0xffffffffu;
constexpr uint64_t P_EMPTY = 0xffffffffffffffffull;
constexpr uint32_t SCAN_TOKEN_MAX = 65536;  // bufio.MaxScanTokenSize

static_assert(P_SPAN == 64, "one 64-bit newline mask per thread");

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
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t table_probe(const uint64_t* __restrict__ table, uint32_t mask,
                                                uint32_t shift, uint32_t key) {
  uint32_t h = hash32(key, shift) & mask;
  for (uint32_t probes = 0; probes <= mask; probes++) {
    uint64_t s = __ldg((const unsigned long long*)&table[h]);
    if (s == P_EMPTY) return P_NONE;
    if ((uint32_t)(s >> 32) == key) return (uint32_t)s;
    h = (h + 1) & mask;
  }
  return P_NONE;
}

// ------------------------------------------------------------------------------------------------
// K1.  Persistent co-resident CTAs (grid <= occupancy x SMs), tile = blockIdx + k*gridDim: a
// look-back predecessor always belongs to a resident CTA that reaches it no later than we reach
// ours, and consecutive tiles sit in different CTAs so their chains overlap.  3-stage TMA ring.  A tile owns the lines that START in (a, a+TILE], plus the
// line at offset 0 for the first tile of an image; the 16 halo bytes let it classify a line that
// starts on its last byte.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_parse(ParseArgs A) {
  extern __shared__ __align__(128) uint8_t p_smem[];
  uint8_t* stage_buf = p_smem;  // P_STAGES * P_STAGE
  __shared__ __align__(8) uint64_t full_bar[P_STAGES];
  __shared__ uint32_t s_scratch[KVG_WARPS + 1];
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_first_hdr, s_first_nl, s_last_nl;

  const uint32_t tid = threadIdx.x;

  auto issue = [&](uint32_t stage, uint32_t it) {  // thread 0 only
    uint32_t t = blockIdx.x + it * gridDim.x;
    if (t < A.n_tiles) {
      uint32_t f = t / A.tiles_per_file, j = t - f * A.tiles_per_file;
      const uint8_t* src = A.text + (uint64_t)f * A.stride + (uint64_t)j * P_TILE;
      mbar_arrive_expect_tx(&full_bar[stage], P_STAGE);
      tma_load_1d(stage_buf + stage * P_STAGE, src, P_STAGE, &full_bar[stage]);
    }
  };

  if (tid == 0) {
    for (uint32_t s = 0; s < P_STAGES; s++) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0)
    for (uint32_t s = 0; s < P_STAGES; s++) issue(s, s);

  uint32_t n_new = 0, n_lines = 0;  // per-thread tallies, flushed per tile

  for (uint32_t it = 0;; ++it) {
    const uint32_t stage = it % P_STAGES;
    const uint32_t parity = (it / P_STAGES) & 1;
    const uint32_t tile = blockIdx.x + it * gridDim.x;
    if (tile >= A.n_tiles) break;
    const uint32_t f = tile / A.tiles_per_file, j = tile - f * A.tiles_per_file;
    const uint32_t a = j * P_TILE;  // file offset of the tile
    const uint8_t* sm = stage_buf + stage * P_STAGE;
    uint64_t* table = A.tables + (uint64_t)f * (A.cap_mask + 1);

    if (tid == 0) {
      s_first_hdr = P_NONE;
      s_first_nl = P_NONE;
      s_last_nl = 0;
    }
    mbar_wait(&full_bar[stage], parity);

    // ---- newline mask of this thread's 64-byte span (bank-conflict-free rotated chunk order)
    const uint32_t sp = tid * P_SPAN;
    uint64_t mask = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      uint32_t c = (k + (tid >> 1)) & 3;
      uint4 v = *reinterpret_cast<const uint4*>(sm + sp + c * 16);
      mask |= (uint64_t)nl_mask16(v) << (16 * c);
    }
    // newlines at or beyond EOF are padding, not line terminators of real lines
    if (a + sp + P_SPAN > A.len) {
      uint32_t keep = A.len > a + sp ? A.len - (a + sp) : 0;  // bytes of the span inside the file
      mask = keep >= 64 ? mask : (mask & ((1ull << keep) - 1));
    }
    __syncthreads();  // s_first_* initialised

    // ---- pass A: header-type lines (first byte neither '\t' nor '#') -> context for later lines
    uint32_t last_hdr = 0;  // ((p+1)<<17) | valid<<16 | vendor, 0 = none in this span
    {
      uint32_t my_first_hdr = P_NONE;
      auto visit = [&](uint32_t p) {
        if (a + p >= A.len) return;  // a line must start before EOF
        uint32_t b0 = sm[p];
        if (b0 != '\t' && b0 != '#') {
          uint32_t hv = parse_hex4(sm + p);
          last_hdr = ((p + 1) << 17) | hv;
          my_first_hdr = min(my_first_hdr, a + p);
          if (hv == (0x10000u | 0x10deu)) atomicMin(&A.info[f].v_off, a + p);
        }
      };
      if (j == 0 && tid == 0) visit(0);
      for (uint64_t mm = mask; mm; mm &= mm - 1) visit(sp + (uint32_t)__ffsll((long long)mm));
      if (my_first_hdr != P_NONE) atomicMin(&s_first_hdr, my_first_hdr);
      if (mask) {
        atomicMin(&s_first_nl, a + sp + (uint32_t)__ffsll((long long)mask) - 1);
        atomicMax(&s_last_nl, a + sp + 63 - (uint32_t)__clzll((long long)mask));
        n_lines += (uint32_t)__popcll(mask);
      }
    }
    uint32_t tile_hdr;
    uint32_t ctx = block_excl_max(last_hdr, s_scratch, &tile_hdr);  // two __syncthreads inside

    // ---- vendor context carried into the tile
    if (warp_id() == 0) {
      uint32_t carry = lookback_last(A.tile_state, tile, j == 0, tile_hdr != 0, tile_hdr & 0x1ffffu,
                                     A.epoch);
      if (lane_id() == 0) {
        s_carry = carry;
        A.tile_first_hdr[tile] = s_first_hdr;
        A.tile_first_nl[tile] = s_first_nl;
        A.tile_last_nl[tile] = s_last_nl;
      }
    }
    __syncthreads();
    ctx = ctx ? (ctx & 0x1ffffu) : s_carry;

    // ---- pass B: device lines "\t" + 4 lower-hex under a valid vendor -> hash insert
    {
      auto visit = [&](uint32_t p) {
        if (a + p >= A.len) return;
        uint32_t b0 = sm[p];
        if (b0 == '#') return;
        if (b0 != '\t') {
          ctx = parse_hex4(sm + p);
          return;
        }
        uint32_t dv = parse_hex4(sm + p + 1);
        if ((dv & ctx) & 0x10000u)
          n_new += table_insert(table, A.cap_mask, A.cap_shift, ((ctx & 0xffffu) << 16) | (dv & 0xffffu),
                                a + p, &A.info[f].overflow);
      };
      if (j == 0 && tid == 0) visit(0);
      for (uint64_t mm = mask; mm; mm &= mm - 1) visit(sp + (uint32_t)__ffsll((long long)mm));
    }

    // ---- flush tallies once per tile per warp
    {
      uint32_t e = warp_sum(n_new), l = warp_sum(n_lines);
      if (lane_id() == 0) {
        if (e) atomicAdd(&A.info[f].n_entries, e);
        if (l) atomicAdd(&A.info[f].n_lines, l);
      }
      n_new = 0;
      n_lines = 0;
    }
    __syncthreads();  // everyone is done with this stage
    if (tid == 0) issue(stage, it + P_STAGES);
  }
}

// ------------------------------------------------------------------------------------------------
// One CTA per image: section end and scanner limit.  (Cheap: a few hundred tiles at most.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_finalize(ParseArgs A) {
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

// K2: one thread per byte of the NVIDIA section; the thread sitting on the first byte of a
// candidate line ("\t" + 4 lower-hex) sanitises that line into pool[off - V] = u16 len + bytes.
__global__ void __launch_bounds__(KVG_BLOCK) k_pciids_sanitise(const uint8_t* __restrict__ text,
                                                               uint32_t len,
                                                               const PciIdsInfo* __restrict__ info,
                                                               uint8_t* __restrict__ pool) {
  const uint32_t V = info->v_off, E = info->sec_end;
  if (V == P_NONE) return;
  for (uint32_t b = V + 1 + blockIdx.x * blockDim.x + threadIdx.x; b < E;
       b += gridDim.x * blockDim.x) {
    if (text[b - 1] != '\n' || text[b] != '\t') continue;
    if (b + 5 > len) continue;
    if (!(parse_hex4(text + b + 1) & 0x10000u)) continue;
    uint32_t e = b + 5;
    while (e < len && text[e] != '\n') e++;
    uint8_t* slot = pool + (b - V);
    uint32_t n = d_sanitise_name(text + b + 5, e - (b + 5), slot + 2, e - (b + 5));
    slot[0] = (uint8_t)(n & 0xff);
    slot[1] = (uint8_t)(n >> 8);
  }
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
__global__ void k_probe_keys(const uint64_t* __restrict__ table, uint32_t mask, uint32_t shift,
                             const PciIdsInfo* __restrict__ info, uint32_t first, uint32_t count,
                             uint32_t* __restrict__ slots) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) slots[i] = probe_name_slot(table, mask, shift, info, (first + i) & 0xffffu);
}

// Exact :388-402 for arbitrary keys.  grid.y = key index; threads sweep the section bytes.
//   candidate = a line start b in (V, E) whose line is not a comment; E already excludes
//   everything past the first non-'\t' non-'#' line.  match = HasPrefix(line, "\t"+key) on the
//   token bufio.ScanLines returns (one trailing '\r' dropped).
__global__ void __launch_bounds__(KVG_BLOCK) k_lookup_general(const uint8_t* __restrict__ text,
                                                              uint32_t len,
                                                              const PciIdsInfo* __restrict__ info,
                                                              const uint8_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ key_off,
                                                              uint32_t* __restrict__ match_off) {
  const uint32_t V = info->v_off, E = info->sec_end;
  if (V == P_NONE) return;
  const uint32_t kidx = blockIdx.y;
  const uint8_t* key = keys + key_off[kidx];
  const uint32_t klen = key_off[kidx + 1] - key_off[kidx];
  for (uint32_t b = V + 1 + blockIdx.x * blockDim.x + threadIdx.x; b < E;
       b += gridDim.x * blockDim.x) {
    if (text[b - 1] != '\n' || text[b] != '\t') continue;  // '#' lines and non-starts skipped
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

/*
 * kvg_oracle.c — CPU ORACLE: test infrastructure, not the product (see kvg_oracle.h).
 *
 * Every function cites the reference lines it restates; paths are relative to the reference repo
 * (NVIDIA/kubevirt-gpu-device-plugin), file pkg/device_plugin/device_plugin.go unless noted.
 * Go standard-library behaviour (bufio.Scanner, strings.*, regexp, strconv, filepath.Walk) is not
 * vendored in the reference; it is restated here from its documented semantics and each
 * restatement says which stdlib function it stands for.
 */
#define _GNU_SOURCE
#include "kvg_oracle.h"

#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

/* ================================================================================================
 * small utilities
 * ============================================================================================== */

typedef struct {
  uint8_t *p;
  size_t n, cap;
} bytes_t;

static void bytes_reserve(bytes_t *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap * 2 : 256;
  while (nc < b->n + extra) nc *= 2;
  b->p = (uint8_t *)realloc(b->p, nc);
  b->cap = nc;
}
static void bytes_put(bytes_t *b, const void *s, size_t n) {
  bytes_reserve(b, n + 1);
  memcpy(b->p + b->n, s, n);
  b->n += n;
}
static void bytes_putc(bytes_t *b, uint8_t c) { bytes_put(b, &c, 1); }

/* arena for key / address strings so the maps never free piecemeal */
typedef struct arena_blk {
  struct arena_blk *next;
  size_t used, cap;
  char data[];
} arena_blk;
typedef struct {
  arena_blk *head;
} arena_t;
static char *arena_dup(arena_t *a, const char *s, size_t n) {
  if (!a->head || a->head->used + n + 1 > a->head->cap) {
    size_t cap = (n + 1 > (1u << 20)) ? n + 1 : (1u << 20);
    arena_blk *b = (arena_blk *)malloc(sizeof(arena_blk) + cap);
    b->next = a->head;
    b->used = 0;
    b->cap = cap;
    a->head = b;
  }
  char *p = a->head->data + a->head->used;
  memcpy(p, s, n);
  p[n] = 0;
  a->head->used += n + 1;
  return p;
}
static void arena_free(arena_t *a) {
  arena_blk *b = a->head;
  while (b) {
    arena_blk *n = b->next;
    free(b);
    b = n;
  }
  a->head = NULL;
}

void kvo_free(void *p) { free(p); }

/* ================================================================================================
 * Go stdlib restatements
 * ============================================================================================== */

#define RUNE_ERROR 0xFFFD

/* unicode/utf8.DecodeRune: returns the rune and its width; invalid -> (RuneError, 1) */
static uint32_t decode_rune(const uint8_t *p, size_t n, int *width) {
  if (n == 0) {
    *width = 0;
    return RUNE_ERROR;
  }
  uint8_t b0 = p[0];
  if (b0 < 0x80) {
    *width = 1;
    return b0;
  }
  *width = 1;
  if (b0 < 0xC2 || b0 > 0xF4) return RUNE_ERROR;
  if (b0 < 0xE0) { /* 2 bytes */
    if (n < 2 || (p[1] & 0xC0) != 0x80) return RUNE_ERROR;
    *width = 2;
    return ((uint32_t)(b0 & 0x1F) << 6) | (p[1] & 0x3F);
  }
  if (b0 < 0xF0) { /* 3 bytes */
    if (n < 3) return RUNE_ERROR;
    uint8_t lo = 0x80, hi = 0xBF;
    if (b0 == 0xE0) lo = 0xA0;
    if (b0 == 0xED) hi = 0x9F;
    if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80) return RUNE_ERROR;
    *width = 3;
    return ((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
  }
  if (n < 4) return RUNE_ERROR;
  uint8_t lo = 0x80, hi = 0xBF;
  if (b0 == 0xF0) lo = 0x90;
  if (b0 == 0xF4) hi = 0x8F;
  if (p[1] < lo || p[1] > hi || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) return RUNE_ERROR;
  *width = 4;
  return ((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) |
         ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}

/* unicode/utf8.DecodeLastRune */
static uint32_t decode_last_rune(const uint8_t *p, size_t n, int *width) {
  if (n == 0) {
    *width = 0;
    return RUNE_ERROR;
  }
  size_t end = n;
  long start = (long)end - 1;
  if (p[start] < 0x80) {
    *width = 1;
    return p[start];
  }
  long lim = (long)end - 4;
  if (lim < 0) lim = 0;
  for (start--; start >= lim; start--)
    if ((p[start] & 0xC0) != 0x80) break; /* utf8.RuneStart */
  if (start < 0) start = 0;
  int w;
  uint32_t r = decode_rune(p + start, end - (size_t)start, &w);
  if ((size_t)start + (size_t)w != end) {
    *width = 1;
    return RUNE_ERROR;
  }
  *width = w;
  return r;
}

/* unicode.IsSpace (White_Space property) */
static int is_space_rune(uint32_t r) {
  if (r == '\t' || r == '\n' || r == '\v' || r == '\f' || r == '\r' || r == ' ') return 1;
  if (r == 0x85 || r == 0xA0 || r == 0x1680) return 1;
  if (r >= 0x2000 && r <= 0x200A) return 1;
  if (r == 0x2028 || r == 0x2029 || r == 0x202F || r == 0x205F || r == 0x3000) return 1;
  return 0;
}
static int is_ascii_space(uint8_t c) {
  return c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r' || c == ' ';
}

/* strings.TrimSpace: ASCII fast path, falling back to TrimFunc(unicode.IsSpace) as soon as a
 * byte >= 0x80 is met on either side */
void kvo_trim_space(const uint8_t *s, size_t n, size_t *pstart, size_t *pend) {
  size_t start = 0, stop = n;
  int unicode_path = 0;
  for (; start < n; start++) {
    uint8_t c = s[start];
    if (c >= 0x80) {
      unicode_path = 1;
      break;
    }
    if (!is_ascii_space(c)) break;
  }
  if (!unicode_path) {
    for (; stop > start; stop--) {
      uint8_t c = s[stop - 1];
      if (c >= 0x80) {
        unicode_path = 2; /* TrimRightFunc only */
        break;
      }
      if (!is_ascii_space(c)) break;
    }
  }
  if (unicode_path == 1) { /* TrimLeftFunc on s[start:] */
    while (start < n) {
      int w;
      uint32_t r = decode_rune(s + start, n - start, &w);
      if (!is_space_rune(r)) break;
      start += (size_t)w;
    }
    stop = n;
  }
  if (unicode_path) { /* TrimRightFunc */
    while (stop > start) {
      int w;
      uint32_t r = decode_last_rune(s + start, stop - start, &w);
      if (!is_space_rune(r)) break;
      stop -= (size_t)w;
    }
  }
  *pstart = start;
  *pend = stop;
}

/* strings.ToUpper with Go's SIMPLE case mapping.  Only two non-ASCII runes map into ASCII:
 * U+0131 (dotless i) -> 'I' and U+017F (long s) -> 'S'.  Every other non-ASCII rune is copied
 * through unchanged instead of case-mapped: all of them are deleted two passes later by
 * `[^a-zA-Z0-9_.]+` (device_plugin.go:413-414), so the final name cannot differ.  Invalid bytes
 * become U+FFFD exactly like strings.Map does. */
size_t kvo_to_upper(const uint8_t *s, size_t n, uint8_t *out) {
  size_t o = 0, i = 0;
  while (i < n) {
    uint8_t c = s[i];
    if (c < 0x80) {
      out[o++] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
      i++;
      continue;
    }
    int w;
    uint32_t r = decode_rune(s + i, n - i, &w);
    if (r == RUNE_ERROR && w == 1) {
      out[o++] = 0xEF;
      out[o++] = 0xBF;
      out[o++] = 0xBD;
    } else if (r == 0x131) {
      out[o++] = 'I';
    } else if (r == 0x17F) {
      out[o++] = 'S';
    } else {
      memcpy(out + o, s + i, (size_t)w);
      o += (size_t)w;
    }
    i += (size_t)w;
  }
  return o;
}

/* RE2 \s == [\t\n\f\r ] (note: no \v) */
static int re2_space(uint8_t c) { return c == '\t' || c == '\n' || c == '\f' || c == '\r' || c == ' '; }

/* regexp.MustCompile("\\s+").ReplaceAllString(s, "_") */
static size_t replace_ws_runs(const uint8_t *s, size_t n, uint8_t *out) {
  size_t o = 0, i = 0;
  while (i < n) {
    if (re2_space(s[i])) {
      out[o++] = '_';
      while (i < n && re2_space(s[i])) i++;
    } else {
      out[o++] = s[i++];
    }
  }
  return o;
}

/* bufio.Scanner + bufio.ScanLines over an in-memory file: split on '\n', drop ONE trailing '\r',
 * return a final unterminated line, fail (ErrTooLong) on a line with no '\n' in its first
 * 64 KiB (bufio.MaxScanTokenSize) */
typedef struct {
  const uint8_t *buf;
  size_t len, pos;
  int err;
} scanner_t;
#define MAX_SCAN_TOKEN 65536u

static int scanner_scan(scanner_t *sc, const uint8_t **line, size_t *linelen) {
  if (sc->err || sc->pos >= sc->len) return 0;
  const uint8_t *p = sc->buf + sc->pos;
  size_t rem = sc->len - sc->pos;
  const uint8_t *nl = (const uint8_t *)memchr(p, '\n', rem);
  size_t content = nl ? (size_t)(nl - p) : rem;
  if (content >= MAX_SCAN_TOKEN) {
    sc->err = 1;
    return 0;
  }
  sc->pos += content + (nl ? 1 : 0);
  if (content > 0 && p[content - 1] == '\r') content--;
  *line = p;
  *linelen = content;
  return 1;
}

static int has_prefix(const uint8_t *s, size_t n, const uint8_t *pre, size_t pn) {
  return n >= pn && memcmp(s, pre, pn) == 0;
}

/* ================================================================================================
 * getDeviceName / locateVendor  (device_plugin.go:371-438)
 * ============================================================================================== */

static const uint8_t NVIDIA_VENDOR_ID[4] = {'1', '0', 'd', 'e'}; /* :45-47 */

/* the name transform of :404-414, pass by pass */
static size_t sanitise_name(const uint8_t *s, size_t n, uint8_t *out /* >= 3n+1 */) {
  size_t a, b;
  kvo_trim_space(s, n, &a, &b); /* :405 */
  uint8_t *t1 = (uint8_t *)malloc(3 * (b - a) + 1);
  size_t n1 = kvo_to_upper(s + a, b - a, t1); /* :406 */
  for (size_t i = 0; i < n1; i++)             /* :407 */
    if (t1[i] == '/') t1[i] = '_';
  for (size_t i = 0; i < n1; i++) /* :408 */
    if (t1[i] == '.') t1[i] = '_';
  uint8_t *t2 = (uint8_t *)malloc(n1 + 1);
  size_t n2 = replace_ws_runs(t1, n1, t2); /* :410-411 */
  size_t o = 0;                            /* :413-414  [^a-zA-Z0-9_.]+ -> "" */
  for (size_t i = 0; i < n2; i++) {
    uint8_t c = t2[i];
    if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' ||
        c == '.')
      out[o++] = c;
  }
  free(t1);
  free(t2);
  return o;
}

long kvo_get_device_name(const uint8_t *text, size_t len, const uint8_t *key, size_t keylen,
                         uint8_t *out, size_t cap) {
  scanner_t sc = {text, len, 0, 0};
  const uint8_t *line;
  size_t ll;
  /* locateVendor :424-438 */
  int found = 0;
  while (scanner_scan(&sc, &line, &ll)) {
    if (has_prefix(line, ll, NVIDIA_VENDOR_ID, 4)) {
      found = 1;
      break;
    }
  }
  if (!found) return 0; /* :382-385 */
  /* prefix := "\t" + deviceID  :388 */
  uint8_t *prefix = (uint8_t *)malloc(keylen + 1);
  prefix[0] = '\t';
  memcpy(prefix + 1, key, keylen);
  long result = 0;
  while (scanner_scan(&sc, &line, &ll)) {              /* :389 */
    if (ll >= 1 && line[0] == '#') continue;           /* :392-394 */
    if (!(ll >= 1 && line[0] == '\t')) {               /* :396-399 */
      result = 0;
      break;
    }
    if (!has_prefix(line, ll, prefix, keylen + 1)) continue; /* :400-402 */
    const uint8_t *rest = line + keylen + 1;                 /* :404 TrimPrefix */
    size_t rn = ll - (keylen + 1);
    uint8_t *tmp = (uint8_t *)malloc(3 * rn + 1);
    size_t on = sanitise_name(rest, rn, tmp);
    if (on > cap) {
      free(tmp);
      free(prefix);
      return -1;
    }
    memcpy(out, tmp, on);
    free(tmp);
    result = (long)on;
    break; /* :415 */
  }
  free(prefix);
  return result;
}

static uint8_t *read_whole_file(const char *path, size_t *len) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return NULL;
  bytes_t b = {0};
  for (;;) {
    bytes_reserve(&b, 65536);
    ssize_t r = read(fd, b.p + b.n, 65536);
    if (r < 0) {
      if (errno == EINTR) continue;
      close(fd);
      free(b.p);
      return NULL;
    }
    if (r == 0) break;
    b.n += (size_t)r;
  }
  close(fd);
  if (!b.p) b.p = (uint8_t *)malloc(1);
  *len = b.n;
  return b.p;
}

long kvo_get_device_name_file(const char *path, const uint8_t *key, size_t keylen, uint8_t *out,
                              size_t cap) {
  size_t len;
  uint8_t *text = read_whole_file(path, &len); /* os.Open :373 */
  if (!text) return 0;                         /* :374-377 */
  long r = kvo_get_device_name(text, len, key, keylen, out, cap);
  free(text);
  return r;
}

static int hexval_lower(uint8_t c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  return -1;
}

size_t kvo_nv_ids(const uint8_t *text, size_t len, uint16_t *out, size_t cap) {
  scanner_t sc = {text, len, 0, 0};
  const uint8_t *line;
  size_t ll, n = 0;
  int found = 0;
  while (scanner_scan(&sc, &line, &ll))
    if (has_prefix(line, ll, NVIDIA_VENDOR_ID, 4)) {
      found = 1;
      break;
    }
  if (!found) return 0;
  while (scanner_scan(&sc, &line, &ll)) {
    if (ll >= 1 && line[0] == '#') continue;
    if (!(ll >= 1 && line[0] == '\t')) break;
    if (ll < 5) continue;
    int v = 0, ok = 1;
    for (int i = 1; i <= 4; i++) {
      int h = hexval_lower(line[i]);
      if (h < 0) ok = 0;
      v = v * 16 + (h < 0 ? 0 : h);
    }
    if (!ok) continue;
    if (n < cap) out[n] = (uint16_t)v;
    n++;
  }
  return n;
}

/* ================================================================================================
 * sysfs readers (device_plugin.go:294-357)
 * ============================================================================================== */

static void join3(char *dst, size_t cap, const char *a, const char *b, const char *c) {
  if (c)
    snprintf(dst, cap, "%s/%s/%s", a, b, c);
  else
    snprintf(dst, cap, "%s/%s", a, b);
}

/* readIDFromFileFunc :294-302 */
long kvo_read_id_from_file(const char *base, const char *addr, const char *prop, char *out,
                           size_t cap) {
  char path[4096];
  join3(path, sizeof path, base, addr, prop);
  size_t len;
  uint8_t *data = read_whole_file(path, &len);
  if (!data) return -1; /* :296-299 */
  if (len < 2) {        /* data[2:] panics: slice bounds out of range */
    free(data);
    return -2;
  }
  /* strings.Trim(string(data[2:]), "\n")  :300 */
  size_t a = 2, b = len;
  while (a < b && data[a] == '\n') a++;
  while (b > a && data[b - 1] == '\n') b--;
  size_t n = b - a;
  if (n + 1 > cap) n = cap - 1;
  memcpy(out, data + a, n);
  out[n] = 0;
  free(data);
  return (long)n;
}

/* strconv.ParseInt(s, 10, 64) */
static int parse_int64(const uint8_t *s, size_t n, int64_t *out) {
  if (n == 0) return -1;
  size_t i = 0;
  int neg = 0;
  if (s[0] == '+' || s[0] == '-') {
    neg = s[0] == '-';
    i = 1;
    if (n == 1) return -1;
  }
  uint64_t v = 0;
  const uint64_t cutoff = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
  for (; i < n; i++) {
    if (s[i] < '0' || s[i] > '9') return -1;
    uint64_t d = (uint64_t)(s[i] - '0');
    if (v > (cutoff - d) / 10) return -1; /* range error */
    v = v * 10 + d;
  }
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  return 0;
}

/* readNUMANodeFunc :304-320 */
int kvo_read_numa_node(const char *base, const char *addr, int64_t *out) {
  char path[4096];
  join3(path, sizeof path, base, addr, "numa_node");
  size_t len;
  uint8_t *data = read_whole_file(path, &len);
  *out = 0;
  if (!data) return -1; /* :306-309 */
  size_t a, b;
  kvo_trim_space(data, len, &a, &b); /* :310 */
  int64_t v;
  int rc = parse_int64(data + a, b - a, &v); /* :311 */
  free(data);
  if (rc) return -1;  /* :312-315 */
  if (v < 0) v = 0;   /* :316-318 */
  *out = v;
  return 0;
}

/* readLinkFunc :323-331 — os.Readlink then filepath.Split: the part after the last '/' */
long kvo_read_link(const char *base, const char *addr, const char *link, char *out, size_t cap) {
  char path[4096], target[4096];
  join3(path, sizeof path, base, addr, link);
  ssize_t n = readlink(path, target, sizeof target - 1);
  if (n < 0) return -1;
  target[n] = 0;
  const char *slash = strrchr(target, '/');
  const char *file = slash ? slash + 1 : target;
  size_t fl = strlen(file);
  if (fl + 1 > cap) fl = cap - 1;
  memcpy(out, file, fl);
  out[fl] = 0;
  return (long)fl;
}

/* readVgpuIDFromFileFunc :334-344 */
static size_t vgpu_label(const uint8_t *data, size_t len, uint8_t *out /* >= len */) {
  size_t a = 0, b = len; /* strings.Trim(str, "\n") :341 */
  while (a < b && data[a] == '\n') a++;
  while (b > a && data[b - 1] == '\n') b--;
  return replace_ws_runs(data + a, b - a, out); /* :342 */
}
long kvo_read_vgpu_id_from_file(const char *base, const char *addr, const char *prop, char *out,
                                size_t cap) {
  char path[4096];
  join3(path, sizeof path, base, addr, prop);
  size_t len;
  uint8_t *data = read_whole_file(path, &len);
  if (!data) return -1;
  uint8_t *tmp = (uint8_t *)malloc(len + 1);
  size_t n = vgpu_label(data, len, tmp);
  if (n + 1 > cap) n = cap - 1;
  memcpy(out, tmp, n);
  out[n] = 0;
  free(tmp);
  free(data);
  return (long)n;
}

/* readGpuIDForVgpuFunc :347-357 — strings.Split(path, "/")[len-2], Trim "\n" */
long kvo_read_gpu_id_for_vgpu(const char *base, const char *addr, char *out, size_t cap) {
  char path[4096], target[4096];
  join3(path, sizeof path, base, addr, NULL);
  ssize_t n = readlink(path, target, sizeof target - 1);
  if (n < 0) return -1;
  target[n] = 0;
  char *last = strrchr(target, '/');
  if (!last) return -2; /* splitStr[-1]: index out of range panic */
  *last = 0;
  char *prev = strrchr(target, '/');
  const char *comp = prev ? prev + 1 : target;
  size_t a = 0, b = strlen(comp);
  while (a < b && comp[a] == '\n') a++;
  while (b > a && comp[b - 1] == '\n') b--;
  size_t l = b - a;
  if (l + 1 > cap) l = cap - 1;
  memcpy(out, comp + a, l);
  out[l] = 0;
  return (long)l;
}

/* supportedVfioDrivers :75-78, isSupportedVfioDriver :249-252 */
int kvo_is_supported_vfio_driver(const char *driver) {
  return strcmp(driver, "vfio-pci") == 0 || strcmp(driver, "nvgrace_gpu_vfio_pci") == 0;
}

/* ================================================================================================
 * the five maps (:55-68): string-keyed hash maps with insertion-ordered value slices
 * ============================================================================================== */

typedef struct {
  const char *p;
  uint32_t len;
} kstr;

typedef struct {
  kstr addr;
  int64_t numa;
} kvo_dev; /* NvidiaGpuDevice :50-53 */

typedef struct {
  kstr key;
  uint64_t hash;
  /* value: either a slice of devices, a slice of strings, or one string */
  kvo_dev *devs;
  kstr *strs;
  kstr one;
  uint32_t n, cap;
} map_entry;

typedef struct {
  map_entry *e;   /* dense entries in first-insertion order */
  uint32_t *slot; /* open-addressed index into e, 0 = empty, else idx+1 */
  size_t n, ecap, scap;
} strmap;

static uint64_t fnv1a(const char *p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) {
    h ^= (uint8_t)p[i];
    h *= 1099511628211ull;
  }
  return h;
}

static void strmap_rehash(strmap *m, size_t nscap) {
  free(m->slot);
  m->slot = (uint32_t *)calloc(nscap, sizeof(uint32_t));
  m->scap = nscap;
  for (size_t i = 0; i < m->n; i++) {
    size_t s = m->e[i].hash & (nscap - 1);
    while (m->slot[s]) s = (s + 1) & (nscap - 1);
    m->slot[s] = (uint32_t)i + 1;
  }
}

static map_entry *strmap_get(strmap *m, const char *k, size_t kl, arena_t *ar, int create) {
  uint64_t h = fnv1a(k, kl);
  if (m->scap) {
    size_t s = h & (m->scap - 1);
    while (m->slot[s]) {
      map_entry *e = &m->e[m->slot[s] - 1];
      if (e->hash == h && e->key.len == kl && memcmp(e->key.p, k, kl) == 0) return e;
      s = (s + 1) & (m->scap - 1);
    }
  }
  if (!create) return NULL;
  if (m->n == m->ecap) {
    m->ecap = m->ecap ? m->ecap * 2 : 64;
    m->e = (map_entry *)realloc(m->e, m->ecap * sizeof(map_entry));
  }
  if ((m->n + 1) * 2 > m->scap) strmap_rehash(m, m->scap ? m->scap * 2 : 128);
  map_entry *e = &m->e[m->n];
  memset(e, 0, sizeof *e);
  e->key.p = arena_dup(ar, k, kl);
  e->key.len = (uint32_t)kl;
  e->hash = h;
  size_t s = h & (m->scap - 1);
  while (m->slot[s]) s = (s + 1) & (m->scap - 1);
  m->slot[s] = (uint32_t)m->n + 1;
  m->n++;
  return e;
}

static void entry_append_dev(map_entry *e, kstr addr, int64_t numa) {
  if (e->n == e->cap) {
    e->cap = e->cap ? e->cap * 2 : 2;
    e->devs = (kvo_dev *)realloc(e->devs, e->cap * sizeof(kvo_dev));
  }
  e->devs[e->n].addr = addr;
  e->devs[e->n].numa = numa;
  e->n++;
}
static void entry_append_str(map_entry *e, kstr s) {
  if (e->n == e->cap) {
    e->cap = e->cap ? e->cap * 2 : 2;
    e->strs = (kstr *)realloc(e->strs, e->cap * sizeof(kstr));
  }
  e->strs[e->n++] = s;
}
static void strmap_free(strmap *m) {
  for (size_t i = 0; i < m->n; i++) {
    free(m->e[i].devs);
    free(m->e[i].strs);
  }
  free(m->e);
  free(m->slot);
  memset(m, 0, sizeof *m);
}

struct kvo_maps {
  strmap iommuMap;      /* :56 */
  strmap deviceMap;     /* :59 */
  strmap bdfToIommuMap; /* :62 */
  strmap vGpuMap;       /* :65 */
  strmap gpuVgpuMap;    /* :68 */
  arena_t arena;
};

kvo_maps *kvo_maps_new(void) { return (kvo_maps *)calloc(1, sizeof(kvo_maps)); }
void kvo_maps_free(kvo_maps *m) {
  if (!m) return;
  strmap_free(&m->iommuMap);
  strmap_free(&m->deviceMap);
  strmap_free(&m->bdfToIommuMap);
  strmap_free(&m->vGpuMap);
  strmap_free(&m->gpuVgpuMap);
  arena_free(&m->arena);
  free(m);
}
void kvo_maps_counts(const kvo_maps *m, uint64_t *a, uint64_t *b, uint64_t *c, uint64_t *d,
                     uint64_t *e) {
  if (a) *a = m->deviceMap.n;
  if (b) *b = m->iommuMap.n;
  if (c) *c = m->bdfToIommuMap.n;
  if (d) *d = m->vGpuMap.n;
  if (e) *e = m->gpuVgpuMap.n;
}

/* ---- the per-entry body of the createIommuDeviceMap walk callback, :201-244 ------------------
 * The five reader results arrive as (rc, value) pairs so the tree walk and the flat snapshot run
 * the very same code. */
typedef struct {
  int vendor_rc;
  const char *vendor;
  int driver_rc;
  const char *driver;
  int iommu_rc;
  const char *iommu;
  int numa_rc;
  int64_t numa;
  int device_rc;
  const char *device;
} pci_reads;

static void pci_entry(kvo_maps *m, const char *name, size_t name_len, const pci_reads *r) {
  if (r->vendor_rc) return;                              /* :203-206 */
  if (strcmp(r->vendor, "10de") != 0) return;            /* :209 */
  if (r->driver_rc) return;                              /* :213-216 */
  if (!kvo_is_supported_vfio_driver(r->driver)) return;  /* :217-220 */
  if (r->iommu_rc) return;                               /* :222-225 */
  int64_t numa = r->numa_rc ? 0 : r->numa;               /* :226-230 */
  if (r->device_rc) return;                              /* :234-238 */
  kstr addr = {arena_dup(&m->arena, name, name_len), (uint32_t)name_len};
  entry_append_dev(strmap_get(&m->deviceMap, r->device, strlen(r->device), &m->arena, 1), addr,
                   numa); /* :240 */
  map_entry *g = strmap_get(&m->iommuMap, r->iommu, strlen(r->iommu), &m->arena, 1);
  entry_append_dev(g, addr, numa);                                           /* :241-242 */
  map_entry *b = strmap_get(&m->bdfToIommuMap, name, name_len, &m->arena, 1); /* :243 */
  b->one = g->key;
}

/* ---- filepath.Walk: lexical order, Lstat (symlinks are not followed, so a symlinked device
 * directory is "not a dir" and gets visited), real sub-directories are descended ---------------- */
typedef int (*walk_fn)(void *ctx, const char *path, const char *name, int is_dir, int err);

static int cmp_names(const void *a, const void *b) {
  return strcmp(*(const char *const *)a, *(const char *const *)b);
}

static int walk_rec(const char *path, const char *name, const struct stat *st, walk_fn fn,
                    void *ctx) {
  if (!S_ISDIR(st->st_mode)) return fn(ctx, path, name, 0, 0);
  DIR *d = opendir(path);
  char **names = NULL;
  size_t nn = 0, cap = 0;
  int rderr = 0;
  if (!d) {
    rderr = 1;
  } else {
    struct dirent *de;
    while ((de = readdir(d))) {
      if (!strcmp(de->d_name, ".") || !strcmp(de->d_name, "..")) continue;
      if (nn == cap) {
        cap = cap ? cap * 2 : 64;
        names = (char **)realloc(names, cap * sizeof(char *));
      }
      names[nn++] = strdup(de->d_name);
    }
    closedir(d);
    qsort(names, nn, sizeof(char *), cmp_names);
  }
  int rc = fn(ctx, path, name, 1, rderr);
  if (rderr || rc) {
    for (size_t i = 0; i < nn; i++) free(names[i]);
    free(names);
    return rc ? rc : 0;
  }
  for (size_t i = 0; i < nn && !rc; i++) {
    char child[4096];
    snprintf(child, sizeof child, "%s/%s", path, names[i]);
    struct stat cst;
    if (lstat(child, &cst) != 0)
      rc = fn(ctx, child, names[i], 0, 1);
    else
      rc = walk_rec(child, names[i], &cst, fn, ctx);
  }
  for (size_t i = 0; i < nn; i++) free(names[i]);
  free(names);
  return rc;
}

static int walk(const char *root, walk_fn fn, void *ctx) {
  struct stat st;
  const char *slash = strrchr(root, '/');
  const char *name = slash ? slash + 1 : root;
  if (lstat(root, &st) != 0) return fn(ctx, root, name, 0, 1);
  return walk_rec(root, name, &st, fn, ctx);
}

typedef struct {
  kvo_maps *m;
  const char *base;     /* basePath / vGpuBasePath */
  const char *pci_base; /* basePath used for the parent's numa_node (:280) */
  int panicked;
} walk_ctx;

static int pci_walk_cb(void *vctx, const char *path, const char *name, int is_dir, int err) {
  (void)path;
  walk_ctx *w = (walk_ctx *)vctx;
  if (err) return 1;     /* :193-196 abort the walk */
  if (is_dir) return 0;  /* :197-200 */
  char vendor[256], driver[256], iommu[256], device[256];
  pci_reads r;
  memset(&r, 0, sizeof r);
  long rc = kvo_read_id_from_file(w->base, name, "vendor", vendor, sizeof vendor); /* :202 */
  if (rc == -2) {
    w->panicked = 1;
    return 1;
  }
  r.vendor_rc = rc < 0;
  r.vendor = vendor;
  if (r.vendor_rc || strcmp(vendor, "10de") != 0) { /* nothing else is read (:209) */
    pci_entry(w->m, name, strlen(name), &r);
    return 0;
  }
  rc = kvo_read_link(w->base, name, "driver", driver, sizeof driver); /* :212 */
  r.driver_rc = rc < 0;
  r.driver = driver;
  if (!r.driver_rc && kvo_is_supported_vfio_driver(driver)) {
    rc = kvo_read_link(w->base, name, "iommu_group", iommu, sizeof iommu); /* :221 */
    r.iommu_rc = rc < 0;
    r.iommu = iommu;
    if (!r.iommu_rc) {
      r.numa_rc = kvo_read_numa_node(w->base, name, &r.numa); /* :226 */
      rc = kvo_read_id_from_file(w->base, name, "device", device, sizeof device); /* :234 */
      if (rc == -2) {
        w->panicked = 1;
        return 1;
      }
      r.device_rc = rc < 0;
      r.device = device;
    }
  }
  pci_entry(w->m, name, strlen(name), &r);
  return 0;
}

static void maps_reset_pci(kvo_maps *m) { /* :188-190 */
  strmap_free(&m->iommuMap);
  strmap_free(&m->deviceMap);
  strmap_free(&m->bdfToIommuMap);
}
static void maps_reset_vgpu(kvo_maps *m) { /* :256-257 */
  strmap_free(&m->vGpuMap);
  strmap_free(&m->gpuVgpuMap);
}

int kvo_create_iommu_device_map_tree(kvo_maps *m, const char *base_path) {
  maps_reset_pci(m);
  walk_ctx w = {m, base_path, base_path, 0};
  walk(base_path, pci_walk_cb, &w); /* :192 */
  return w.panicked ? -2 : 0;
}

/* ---- createVgpuIDMap per-entry body, :268-289 ------------------------------------------------ */
static void mdev_entry(kvo_maps *m, const char *name, size_t name_len, int type_rc,
                       const char *type, size_t type_len, int parent_rc, const char *parent,
                       size_t parent_len, int numa_rc, int64_t numa) {
  if (type_rc) return;   /* :270-273 */
  if (parent_rc) return; /* :276-279 */
  if (numa_rc) numa = 0; /* :281-284 */
  kstr addr = {arena_dup(&m->arena, name, name_len), (uint32_t)name_len};
  entry_append_str(strmap_get(&m->gpuVgpuMap, parent, parent_len, &m->arena, 1), addr); /* :287 */
  entry_append_dev(strmap_get(&m->vGpuMap, type, type_len, &m->arena, 1), addr, numa);  /* :288 */
}

static int mdev_walk_cb(void *vctx, const char *path, const char *name, int is_dir, int err) {
  (void)path;
  walk_ctx *w = (walk_ctx *)vctx;
  if (err) return 1;    /* :260-263 */
  if (is_dir) return 0; /* :264-267 */
  char type[1024], parent[1024];
  long tl = kvo_read_vgpu_id_from_file(w->base, name, "mdev_type/name", type, sizeof type); /* :269 */
  if (tl < 0) return 0;
  long pl = kvo_read_gpu_id_for_vgpu(w->base, name, parent, sizeof parent); /* :275 */
  if (pl == -2) {
    w->panicked = 1;
    return 1;
  }
  if (pl < 0) return 0;
  int64_t numa = 0;
  int nrc = kvo_read_numa_node(w->pci_base, parent, &numa); /* :280 */
  mdev_entry(w->m, name, strlen(name), 0, type, (size_t)tl, 0, parent, (size_t)pl, nrc, numa);
  return 0;
}

int kvo_create_vgpu_id_map_tree(kvo_maps *m, const char *vgpu_base, const char *pci_base) {
  maps_reset_vgpu(m);
  walk_ctx w = {m, vgpu_base, pci_base, 0};
  walk(vgpu_base, mdev_walk_cb, &w); /* :259 */
  return w.panicked ? -2 : 0;
}

/* ================================================================================================
 * flat snapshots -> the same per-entry logic
 * ============================================================================================== */

static const char HEXD[] = "0123456789abcdef";
static void fmt_hex4(uint16_t v, char out[5]) {
  out[0] = HEXD[(v >> 12) & 15];
  out[1] = HEXD[(v >> 8) & 15];
  out[2] = HEXD[(v >> 4) & 15];
  out[3] = HEXD[v & 15];
  out[4] = 0;
}
static size_t fmt_u32(uint32_t v, char *out) {
  char tmp[12];
  size_t n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  for (size_t i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
  out[n] = 0;
  return n;
}
/* "dddd:bb:dd.f" from domain<<16 | bus<<8 | dev<<3 | fn */
void kvo_format_bdf(uint32_t p, char out[16]) {
  uint32_t dom = p >> 16, bus = (p >> 8) & 0xff, dev = (p >> 3) & 0x1f, fn = p & 7;
  out[0] = HEXD[(dom >> 12) & 15];
  out[1] = HEXD[(dom >> 8) & 15];
  out[2] = HEXD[(dom >> 4) & 15];
  out[3] = HEXD[dom & 15];
  out[4] = ':';
  out[5] = HEXD[bus >> 4];
  out[6] = HEXD[bus & 15];
  out[7] = ':';
  out[8] = HEXD[dev >> 4];
  out[9] = HEXD[dev & 15];
  out[10] = '.';
  out[11] = HEXD[fn];
  out[12] = 0;
}
void kvo_format_uuid(const uint8_t u[16], char out[40]) {
  size_t o = 0;
  for (int i = 0; i < 16; i++) {
    if (i == 4 || i == 6 || i == 8 || i == 10) out[o++] = '-';
    out[o++] = HEXD[u[i] >> 4];
    out[o++] = HEXD[u[i] & 15];
  }
  out[o] = 0;
}
static const char *driver_name(uint8_t code, char scratch[16]) {
  switch (code) {
    case KVG_DRV_NONE: return "";
    case KVG_DRV_VFIO_PCI: return "vfio-pci";
    case KVG_DRV_NVGRACE: return "nvgrace_gpu_vfio_pci";
    case 3: return "nvidia";
    case 4: return "nouveau";
    default: snprintf(scratch, 16, "other%u", (unsigned)code); return scratch;
  }
}

int kvo_create_iommu_device_map_flat(kvo_maps *m, const kvg_pci_rec *recs, size_t n) {
  maps_reset_pci(m);
  for (size_t i = 0; i < n; i++) {
    const kvg_pci_rec *rec = &recs[i];
    char name[16], vendor[5], device[5], iommu[12], scratch[16];
    kvo_format_bdf(rec->addr, name);
    fmt_hex4(rec->vendor, vendor);
    fmt_hex4(rec->device, device);
    fmt_u32(rec->iommu_group, iommu);
    pci_reads r;
    r.vendor_rc = (rec->flags & KVG_PF_VENDOR_ERR) != 0;
    r.vendor = vendor;
    r.driver_rc = (rec->flags & KVG_PF_DRIVER_ERR) != 0;
    r.driver = driver_name(rec->driver, scratch);
    r.iommu_rc = (rec->flags & KVG_PF_IOMMU_ERR) != 0;
    r.iommu = iommu;
    r.numa_rc = (rec->flags & KVG_PF_NUMA_ERR) != 0;
    r.numa = rec->numa < 0 ? 0 : rec->numa; /* readNUMANodeFunc :316-318 */
    r.device_rc = (rec->flags & KVG_PF_DEVICE_ERR) != 0;
    r.device = device;
    pci_entry(m, name, 12, &r);
  }
  return 0;
}

int kvo_create_vgpu_id_map_flat(kvo_maps *m, const kvg_mdev_rec *recs, size_t n,
                                const kvg_type_dict *types) {
  maps_reset_vgpu(m);
  uint8_t *label = (uint8_t *)malloc(65536);
  for (size_t i = 0; i < n; i++) {
    const kvg_mdev_rec *rec = &recs[i];
    char name[40], parent[16];
    kvo_format_uuid(rec->uuid, name);
    kvo_format_bdf(rec->parent, parent);
    int type_rc = (rec->flags & KVG_MF_TYPE_ERR) != 0 || rec->type_idx >= types->n_types;
    size_t ll = 0;
    if (!type_rc) {
      uint32_t a = types->off[rec->type_idx], b = types->off[rec->type_idx + 1];
      if (b - a > 65536) {
        free(label);
        return -1;
      }
      ll = vgpu_label(types->bytes + a, b - a, label); /* run per call, like :335-342 */
    }
    mdev_entry(m, name, 36, type_rc, (const char *)label, ll,
               (rec->flags & KVG_MF_PARENT_ERR) != 0, parent, 12,
               (rec->flags & KVG_MF_NUMA_ERR) != 0, rec->parent_numa < 0 ? 0 : rec->parent_numa);
  }
  free(label);
  return 0;
}

/* ================================================================================================
 * SHA-256 (FIPS 180-4)
 * ============================================================================================== */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(uint32_t h[8], const uint8_t *p) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) |
           ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
    uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g;
    g = f;
    f = e;
    e = d + t1;
    d = c;
    c = b;
    b = a;
    a = t1 + t2;
  }
  h[0] += a;
  h[1] += b;
  h[2] += c;
  h[3] += d;
  h[4] += e;
  h[5] += f;
  h[6] += g;
  h[7] += hh;
}
void kvo_sha256(const void *data, size_t len, uint8_t digest[32]) {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  const uint8_t *p = (const uint8_t *)data;
  size_t full = len / 64;
  for (size_t i = 0; i < full; i++) sha256_block(h, p + 64 * i);
  uint8_t tail[128];
  size_t rem = len - full * 64;
  memcpy(tail, p + full * 64, rem);
  tail[rem++] = 0x80;
  size_t padlen = rem <= 56 ? 64 : 128;
  memset(tail + rem, 0, padlen - rem);
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; i++) tail[padlen - 1 - i] = (uint8_t)(bits >> (8 * i));
  sha256_block(h, tail);
  if (padlen == 128) sha256_block(h, tail + 64);
  for (int i = 0; i < 8; i++) {
    digest[4 * i] = (uint8_t)(h[i] >> 24);
    digest[4 * i + 1] = (uint8_t)(h[i] >> 16);
    digest[4 * i + 2] = (uint8_t)(h[i] >> 8);
    digest[4 * i + 3] = (uint8_t)h[i];
  }
}

/* ================================================================================================
 * canonical dump
 * ============================================================================================== */

static int cmp_kstr(const kstr *a, const kstr *b) {
  size_t n = a->len < b->len ? a->len : b->len;
  int c = memcmp(a->p, b->p, n);
  if (c) return c;
  return (a->len > b->len) - (a->len < b->len);
}
static int cmp_entry_ptr(const void *a, const void *b) {
  const map_entry *ea = *(const map_entry *const *)a, *eb = *(const map_entry *const *)b;
  return cmp_kstr(&ea->key, &eb->key);
}
static const map_entry **sorted_entries(const strmap *m) {
  const map_entry **v = (const map_entry **)malloc((m->n + 1) * sizeof(*v));
  for (size_t i = 0; i < m->n; i++) v[i] = &m->e[i];
  qsort(v, m->n, sizeof(*v), cmp_entry_ptr);
  return v;
}
static void put_u64(bytes_t *b, uint64_t v) {
  char tmp[24];
  int n = snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)v);
  bytes_put(b, tmp, (size_t)n);
}
static void put_i64(bytes_t *b, int64_t v) {
  char tmp[24];
  int n = snprintf(tmp, sizeof tmp, "%lld", (long long)v);
  bytes_put(b, tmp, (size_t)n);
}
static void put_kstr(bytes_t *b, kstr s) { bytes_put(b, s.p, s.len); }

/* "X <key> <name|-> nvidia.com/<name or key> <n>" + members: the consumer-side join of
 * createDevicePlugins :124-130 / :152-157 and Register's ResourceName
 * (generic_device_plugin.go:299) */
static void dump_dev_section(bytes_t *b, char tag, const strmap *m, const uint8_t *pciids,
                             size_t plen) {
  const map_entry **v = sorted_entries(m);
  uint8_t *name = (uint8_t *)malloc(3 * 70000);
  for (size_t i = 0; i < m->n; i++) {
    const map_entry *e = v[i];
    long nl = pciids ? kvo_get_device_name(pciids, plen, (const uint8_t *)e->key.p, e->key.len,
                                           name, 3 * 70000)
                     : 0;
    if (nl < 0) nl = 0;
    bytes_putc(b, (uint8_t)tag);
    bytes_putc(b, ' ');
    put_kstr(b, e->key);
    bytes_putc(b, ' ');
    if (nl)
      bytes_put(b, name, (size_t)nl);
    else
      bytes_putc(b, '-');
    bytes_put(b, " nvidia.com/", 12);
    if (nl)
      bytes_put(b, name, (size_t)nl);
    else
      put_kstr(b, e->key); /* deviceName = k :125-128 */
    bytes_putc(b, ' ');
    put_u64(b, e->n);
    bytes_putc(b, '\n');
    for (uint32_t j = 0; j < e->n; j++) {
      bytes_put(b, "  ", 2);
      put_kstr(b, e->devs[j].addr);
      bytes_putc(b, ' ');
      put_i64(b, e->devs[j].numa);
      bytes_putc(b, '\n');
    }
  }
  free(name);
  free(v);
}

int kvo_dump(const kvo_maps *m, const uint8_t *pciids, size_t plen, char **out, size_t *outlen) {
  bytes_t b = {0};
  dump_dev_section(&b, 'D', &m->deviceMap, pciids, plen);
  {
    const map_entry **v = sorted_entries(&m->iommuMap);
    for (size_t i = 0; i < m->iommuMap.n; i++) {
      const map_entry *e = v[i];
      bytes_put(&b, "I ", 2);
      put_kstr(&b, e->key);
      bytes_putc(&b, ' ');
      put_u64(&b, e->n);
      bytes_putc(&b, '\n');
      for (uint32_t j = 0; j < e->n; j++) {
        bytes_put(&b, "  ", 2);
        put_kstr(&b, e->devs[j].addr);
        bytes_putc(&b, ' ');
        put_i64(&b, e->devs[j].numa);
        bytes_putc(&b, '\n');
      }
    }
    free(v);
  }
  {
    const map_entry **v = sorted_entries(&m->bdfToIommuMap);
    for (size_t i = 0; i < m->bdfToIommuMap.n; i++) {
      bytes_put(&b, "B ", 2);
      put_kstr(&b, v[i]->key);
      bytes_putc(&b, ' ');
      put_kstr(&b, v[i]->one);
      bytes_putc(&b, '\n');
    }
    free(v);
  }
  dump_dev_section(&b, 'V', &m->vGpuMap, pciids, plen);
  {
    const map_entry **v = sorted_entries(&m->gpuVgpuMap);
    for (size_t i = 0; i < m->gpuVgpuMap.n; i++) {
      const map_entry *e = v[i];
      bytes_put(&b, "G ", 2);
      put_kstr(&b, e->key);
      bytes_putc(&b, ' ');
      put_u64(&b, e->n);
      bytes_putc(&b, '\n');
      for (uint32_t j = 0; j < e->n; j++) {
        bytes_put(&b, "  ", 2);
        put_kstr(&b, e->strs[j]);
        bytes_putc(&b, '\n');
      }
    }
    free(v);
  }
  if (!b.p) b.p = (uint8_t *)malloc(1);
  *out = (char *)b.p;
  *outlen = b.n;
  return 0;
}

/* ================================================================================================
 * synthetic snapshots (counter-based; see DESIGN.md "Synthetic inputs")
 * ============================================================================================== */

uint64_t kvo_mix(uint64_t x) { /* splitmix64 finaliser */
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

#define PCI_SEED 0x10DE000020250711ull
#define MDEV_SEED 0x4D44455600010000ull
static const uint16_t OTHER_VENDORS[8] = {0x8086, 0x1002, 0x15b3, 0x1022,
                                          0x144d, 0x14e4, 0x1af4, 0x10df};

void kvo_gen_pci(kvg_pci_rec *out, uint64_t first, size_t n, const uint16_t *nv_ids,
                 uint32_t n_nv_ids, uint32_t group_bits) {
  for (size_t k = 0; k < n; k++) {
    uint64_t i = first + k;
    uint64_t r0 = kvo_mix(PCI_SEED + 2 * i), r1 = kvo_mix(PCI_SEED + 2 * i + 1);
    kvg_pci_rec r;
    r.addr = (uint32_t)i; /* ascending == Walk order of the formatted BDF strings */
    int nvidia = (r0 & 0xFF) < 128;
    r.vendor = nvidia ? 0x10de : OTHER_VENDORS[(r0 >> 8) & 7];
    if (nvidia && ((r0 >> 16) & 0xFF) < 230 && n_nv_ids)
      r.device = nv_ids[(uint32_t)((r0 >> 24) & 0xFFFFFF) % n_nv_ids];
    else
      r.device = (uint16_t)((r0 >> 24) & 0xFFFF);
    uint32_t d = (uint32_t)((r0 >> 48) & 0xFF);
    r.flags = 0;
    if (d < 154)
      r.driver = KVG_DRV_VFIO_PCI;
    else if (d < 179)
      r.driver = KVG_DRV_NVGRACE;
    else if (d < 218)
      r.driver = 3;
    else if (d < 231)
      r.driver = 4;
    else {
      r.driver = KVG_DRV_NONE;
      r.flags |= KVG_PF_DRIVER_ERR;
    }
    /* two functions per group; group numbers are a bijective scramble of i>>1 inside
       [0, 2^group_bits) so the group bucketing is a real permutation (group_bits==0: i>>1) */
    uint32_t g = (uint32_t)(i >> 1);
    if (group_bits) {
      uint32_t mask = group_bits >= 32 ? 0xFFFFFFFFu : ((1u << group_bits) - 1);
      uint32_t hi = g & ~mask, lo = g & mask;
      lo = (lo * 0x9E3779B1u) & mask; /* odd multiplier: a bijection mod 2^bits */
      lo ^= lo >> (group_bits / 2 + 1);
      g = hi | (lo & mask);
    }
    r.iommu_group = g;
    r.numa = (int16_t)((int)(r1 & 7) - 1);
    if (((r1 >> 8) & 0xFF) == 0) r.flags |= KVG_PF_VENDOR_ERR;
    if (((r1 >> 16) & 0xFF) == 0) r.flags |= KVG_PF_IOMMU_ERR;
    if (((r1 >> 24) & 0xFF) == 0) r.flags |= KVG_PF_DEVICE_ERR;
    if (((r1 >> 32) & 0xFF) == 0) r.flags |= KVG_PF_NUMA_ERR;
    out[k] = r;
  }
}

void kvo_gen_mdev(kvg_mdev_rec *out, uint64_t first, size_t n) {
  for (size_t k = 0; k < n; k++) {
    uint64_t j = first + k;
    uint64_t r0 = kvo_mix(MDEV_SEED + 2 * j), r1 = kvo_mix(MDEV_SEED + 2 * j + 1);
    kvg_mdev_rec r;
    memset(&r, 0, sizeof r);
    r.uuid[0] = (uint8_t)(j >> 24);
    r.uuid[1] = (uint8_t)(j >> 16);
    r.uuid[2] = (uint8_t)(j >> 8);
    r.uuid[3] = (uint8_t)j;
    for (int b = 0; b < 8; b++) r.uuid[4 + b] = (uint8_t)(r0 >> (56 - 8 * b));
    for (int b = 0; b < 4; b++) r.uuid[12 + b] = (uint8_t)(r1 >> (24 - 8 * b));
    r.type_idx = (uint16_t)(r0 >> 56);
    r.parent = (uint32_t)(j >> 5);
    r.parent_numa = (int16_t)((int)((j >> 5) & 3) - 1);
    if (((r1 >> 40) & 0xFF) == 0) r.flags |= KVG_MF_TYPE_ERR;
    if (((r1 >> 48) & 0xFF) == 0) r.flags |= KVG_MF_PARENT_ERR;
    if (((r1 >> 56) & 0xFF) == 0) r.flags |= KVG_MF_NUMA_ERR;
    out[k] = r;
  }
}

size_t kvo_gen_type_name(uint32_t k, char *out, size_t cap) {
  /* pairs (2m, 2m+1) differ only in the amount of whitespace, so they sanitise to the same label
     and must merge into one vGpuMap key */
  int n = snprintf(out, cap, "NVIDIA%sSYN%02X-%d%c\n", (k & 1) ? "  \t" : " ", (k >> 1) & 0x7F,
                   1 << ((k >> 1) & 3), "QCBA"[(k >> 3) & 3]);
  return (size_t)n;
}

/* ================================================================================================
 * CPU baselines
 * ============================================================================================== */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* A — faithful cost: the reference algorithm on the in-memory snapshot, logging off:
 * createIommuDeviceMap (string maps) then getDeviceName once per distinct key, each call
 * re-scanning the pci.ids buffer from byte 0 (:371-438), as createDevicePlugins does (:109-124). */
double kvo_bench_faithful(const kvg_pci_rec *recs, size_t n, const uint8_t *pciids, size_t len,
                          uint64_t *survivors, uint64_t *name_hits) {
  double t0 = now_s();
  kvo_maps *m = kvo_maps_new();
  kvo_create_iommu_device_map_flat(m, recs, n);
  uint64_t hits = 0;
  uint8_t name[4096];
  for (size_t i = 0; i < m->deviceMap.n; i++) {
    const map_entry *e = &m->deviceMap.e[i];
    long nl = kvo_get_device_name(pciids, len, (const uint8_t *)e->key.p, e->key.len, name,
                                  sizeof name);
    if (nl > 0) hits++;
  }
  double t1 = now_s();
  if (survivors) *survivors = m->bdfToIommuMap.n;
  if (name_hits) *name_hits = hits;
  kvo_maps_free(m);
  return t1 - t0;
}

/* B — best effort: pci.ids parsed once into a 65,536-entry presence table, `threads` workers
 * filter disjoint record ranges into private survivor lists, then a serial merge builds the
 * device-id and iommu-group orderings with counting / radix sorts. */
typedef struct {
  const kvg_pci_rec *recs;
  size_t lo, hi;
  kvg_pci_surv *out;
  size_t n_out;
} bw_t;

static void *best_worker(void *arg) {
  bw_t *w = (bw_t *)arg;
  size_t o = 0;
  for (size_t i = w->lo; i < w->hi; i++) {
    const kvg_pci_rec *r = &w->recs[i];
    if (r->flags & (KVG_PF_VENDOR_ERR | KVG_PF_DRIVER_ERR | KVG_PF_IOMMU_ERR | KVG_PF_DEVICE_ERR))
      continue;
    if (r->vendor != 0x10de) continue;
    if (r->driver != KVG_DRV_VFIO_PCI && r->driver != KVG_DRV_NVGRACE) continue;
    kvg_pci_surv s;
    s.addr = r->addr;
    s.iommu_group = r->iommu_group;
    s.device = r->device;
    s.numa = (r->flags & KVG_PF_NUMA_ERR) || r->numa < 0 ? 0 : (uint16_t)r->numa;
    s.name_slot = 0;
    w->out[o++] = s;
  }
  w->n_out = o;
  return NULL;
}

double kvo_bench_threads(const kvg_pci_rec *recs, size_t n, const uint8_t *pciids, size_t len,
                         int threads, uint64_t *survivors, uint64_t *name_hits) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  double t0 = now_s();
  /* parse once */
  uint8_t *present = (uint8_t *)calloc(65536, 1);
  {
    uint16_t *ids = (uint16_t *)malloc(65536 * sizeof(uint16_t));
    size_t k = kvo_nv_ids(pciids, len, ids, 65536);
    if (k > 65536) k = 65536;
    for (size_t i = 0; i < k; i++) present[ids[i]] = 1;
    free(ids);
  }
  pthread_t th[256];
  bw_t w[256];
  size_t per = (n + (size_t)threads - 1) / (size_t)threads;
  for (int t = 0; t < threads; t++) {
    w[t].recs = recs;
    w[t].lo = (size_t)t * per < n ? (size_t)t * per : n;
    w[t].hi = w[t].lo + per < n ? w[t].lo + per : n;
    w[t].out = (kvg_pci_surv *)malloc((w[t].hi - w[t].lo + 1) * sizeof(kvg_pci_surv));
    w[t].n_out = 0;
    pthread_create(&th[t], NULL, best_worker, &w[t]);
  }
  size_t S = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    S += w[t].n_out;
  }
  kvg_pci_surv *all = (kvg_pci_surv *)malloc((S + 1) * sizeof(kvg_pci_surv));
  size_t o = 0;
  for (int t = 0; t < threads; t++) {
    memcpy(all + o, w[t].out, w[t].n_out * sizeof(kvg_pci_surv));
    o += w[t].n_out;
    free(w[t].out);
  }
  /* deviceMap ordering: stable counting sort on the 16-bit id */
  uint32_t *cnt = (uint32_t *)calloc(65537, sizeof(uint32_t));
  for (size_t i = 0; i < S; i++) cnt[all[i].device + 1]++;
  uint64_t hits = 0;
  for (size_t k = 0; k < 65536; k++) {
    if (cnt[k + 1] && present[k]) hits++;
    cnt[k + 1] += cnt[k];
  }
  uint32_t *dev_perm = (uint32_t *)malloc((S + 1) * sizeof(uint32_t));
  for (size_t i = 0; i < S; i++) dev_perm[cnt[all[i].device]++] = (uint32_t)i;
  /* iommuMap ordering: stable LSD radix sort of (group, index), 4 x 8 bits */
  uint32_t *ka = (uint32_t *)malloc((S + 1) * 4), *kb = (uint32_t *)malloc((S + 1) * 4);
  uint32_t *va = (uint32_t *)malloc((S + 1) * 4), *vb = (uint32_t *)malloc((S + 1) * 4);
  for (size_t i = 0; i < S; i++) {
    ka[i] = all[i].iommu_group;
    va[i] = (uint32_t)i;
  }
  for (int pass = 0; pass < 4; pass++) {
    size_t c[257] = {0};
    int sh = 8 * pass;
    for (size_t i = 0; i < S; i++) c[((ka[i] >> sh) & 0xFF) + 1]++;
    for (int k = 0; k < 256; k++) c[k + 1] += c[k];
    for (size_t i = 0; i < S; i++) {
      size_t p = c[(ka[i] >> sh) & 0xFF]++;
      kb[p] = ka[i];
      vb[p] = va[i];
    }
    uint32_t *t = ka;
    ka = kb;
    kb = t;
    t = va;
    va = vb;
    vb = t;
  }
  double t1 = now_s();
  if (survivors) *survivors = S;
  if (name_hits) *name_hits = hits;
  free(ka);
  free(kb);
  free(va);
  free(vb);
  free(dev_perm);
  free(cnt);
  free(all);
  free(present);
  return t1 - t0;
}

/*
 * oracle/tables_oracle.c -- CPU restatement of the reference's edit-distance
 * table generator.  TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product path (iivision_b200/csrc) never calls into this file.
 *
 * PARITY STATUS: "parity unpinned" for absolute table values.  The reference
 * generator (transcoder/make_data_tables.py) cannot run offline: its arithmetic
 * lives in two third-party wheels absent from /root/reference:
 *   - colormath 3.0.0            (requirements.txt:6)  -> see oracle/cie2000.py
 *   - weighted-levenshtein 0.2.2 (requirements.txt:32) -> oracle_dam_lev() below
 * The reference's own tests hold no absolute table value
 * (make_data_tables_test.py:18-95 checks symmetry / sign / zero diagonal only).
 * What IS pinned against the imported reference code: to_dots and the nominal
 * colour pixel strings (tests/test_oracle_vs_reference.py, SHA-256 in
 * tests/golden/pixel_strings.json).
 *
 * Functions and the reference lines they follow:
 *   oracle_hgr_to_dots      screen.py:710-789  (HGRBitmap._double_pixels, to_dots)
 *   oracle_dhgr_to_dots     screen.py:982-990  (identity)
 *   oracle_pixel_string     colours.py:83-148  (rol + 4-bit sliding window)
 *   oracle_dam_lev          make_data_tables.py:92-108 -> weighted_levenshtein.dam_lev
 *                           (published algorithm of weighted-levenshtein 0.2.2:
 *                           true Damerau-Levenshtein with per-character last-row
 *                           table, (len+2)^2 float64 matrix, DBL_MAX border)
 *   oracle_chain_distance   the 1-D collapse of the above that holds because
 *                           insert/delete cost 1e5 (make_data_tables.py:35-36)
 *   oracle_build_table      make_data_tables.py:111-174 (compute_edit_distance)
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define ORACLE_MODE_HGR 0
#define ORACLE_MODE_DHGR 1

static const int kBits[2] = {14, 13};         /* MASKED_BITS  screen.py:617, 887 */
static const int kDots[2] = {18, 10};         /* MASKED_DOTS  screen.py:626, 891 */
static const int kOffsets[2] = {2, 4};        /* len(BYTE_MASKS) screen.py:632, 894 */
static const int kPhases[2][4] = {{1, 3, 0, 0}, {1, 0, 3, 2}}; /* screen.py:645, 919 */

/* No libgomp in this image: plain pthreads with a shared row counter. */
static int g_threads = 0; /* 0 = all online cores */
void oracle_set_threads(int n) { g_threads = n > 0 ? n : 0; }
int oracle_max_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef void (*row_fn)(int64_t row, void* ctx);
typedef struct {
  atomic_llong next;
  int64_t end;
  int64_t chunk;
  row_fn fn;
  void* ctx;
} row_pool;

static void* row_worker(void* p) {
  row_pool* pool = (row_pool*)p;
  for (;;) {
    int64_t b = atomic_fetch_add(&pool->next, pool->chunk);
    if (b >= pool->end) break;
    int64_t e = b + pool->chunk < pool->end ? b + pool->chunk : pool->end;
    for (int64_t r = b; r < e; ++r) pool->fn(r, pool->ctx);
  }
  return NULL;
}

static void parallel_rows(int64_t begin, int64_t end, int64_t chunk, row_fn fn,
                          void* ctx) {
  row_pool pool;
  atomic_init(&pool.next, begin);
  pool.end = end;
  pool.chunk = chunk;
  pool.fn = fn;
  pool.ctx = ctx;
  int nt = oracle_max_threads();
  if (nt > 256) nt = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nt; ++t)
    if (pthread_create(&th[started], NULL, row_worker, &pool) == 0) ++started;
  row_worker(&pool);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

int oracle_masked_bits(int mode) { return kBits[mode]; }
int oracle_masked_dots(int mode) { return kDots[mode]; }
int oracle_num_offsets(int mode) { return kOffsets[mode]; }
int oracle_phase(int mode, int o) { return kPhases[mode][o]; }

/* screen.py:710-739: bits 0..5 light two dots each, bit 6 lights three. */
static uint32_t double_pixels(uint32_t x7) {
  uint32_t out = 0;
  for (int k = 0; k < 6; ++k)
    if (x7 & (1u << k)) out |= 3u << (2 * k);
  if (x7 & 0x40u) out |= 7u << 12;
  return out;
}

/* screen.py:741-789 */
uint32_t oracle_hgr_to_dots(uint32_t v, int byte_offset) {
  uint32_t h = (v & 7u) << 5;
  uint32_t hp = (h & 0x80u) >> 7;
  uint32_t res = double_pixels(h & 0x7fu) >> (11 - hp);
  uint32_t b, bp;
  if (byte_offset == 0) {
    b = (v >> 3) & 0xffu;
    bp = (b & 0x80u) >> 7;
  } else {
    bp = (v >> 3) & 1u;
    b = ((v >> 4) & 0x7fu) ^ (bp << 7);
  }
  res &= ~(0x3fffu << (3 + bp));
  res ^= double_pixels(b & 0x7fu) << (3 + bp);
  /* Python precedence at :781-782: ((v>>12)&3) ^ (((v>>11)&1) << 7) */
  uint32_t f = ((v >> 12) & 3u) ^ (((v >> 11) & 1u) << 7);
  uint32_t fp = (f & 0x80u) >> 7;
  res &= ~(0xfu << (17 + fp));
  res ^= double_pixels(f & 0x7fu) << (17 + fp);
  return res & ((1u << 21) - 1);
}

uint32_t oracle_dhgr_to_dots(uint32_t v, int byte_offset) {
  (void)byte_offset;
  return v;
}

uint32_t oracle_to_dots(int mode, uint32_t v, int byte_offset) {
  return mode == ORACLE_MODE_HGR ? oracle_hgr_to_dots(v, byte_offset)
                                 : oracle_dhgr_to_dots(v, byte_offset);
}

/* colours.py:100-134: pixel_t = rol4((dots >> t) & 15, (phase + t) % 4) */
void oracle_pixel_string(uint32_t dots, int n, int init_phase, uint8_t* out) {
  int phase = init_phase;
  for (int t = 0; t < n; ++t) {
    uint32_t w = (dots >> t) & 0xfu;
    uint32_t r = w;
    for (int k = 0; k < phase; ++k) r = ((r & 7u) << 1) ^ ((r & 8u) >> 3);
    out[t] = (uint8_t)r;
    if (++phase == 4) phase = 0;
  }
}

/* All pixel strings of one mode: out[o][v][t], uint8, C order. */
void oracle_all_pixel_strings(int mode, uint8_t* out) {
  const int bits = kBits[mode], n = kDots[mode];
  for (int o = 0; o < kOffsets[mode]; ++o)
    for (uint32_t v = 0; v < (1u << bits); ++v)
      oracle_pixel_string(oracle_to_dots(mode, v, o), n, kPhases[mode][o],
                          out + ((size_t)o * (1u << bits) + v) * n);
}

void oracle_all_dots(int mode, uint32_t* out) {
  const int bits = kBits[mode];
  for (int o = 0; o < kOffsets[mode]; ++o)
    for (uint32_t v = 0; v < (1u << bits); ++v)
      out[(size_t)o * (1u << bits) + v] = oracle_to_dots(mode, v, o);
}

/*
 * weighted_levenshtein.dam_lev restated.  Strings are byte strings whose
 * characters index 128-entry insert/delete cost vectors and 128x128
 * substitute/transpose matrices (row-major).  d has a guard row/column at
 * index -1 holding DBL_MAX; row 0 / column 0 hold cumulative insert/delete
 * costs; da[c] is the last row of s1 that held character c, db the last column
 * of the current row where the characters matched.
 */
double oracle_dam_lev(const uint8_t* s1, int n1, const uint8_t* s2, int n2,
                      const double* ins, const double* del, const double* sub,
                      const double* tr) {
  const int W = n2 + 2;
  double* d = (double*)malloc(sizeof(double) * (size_t)(n1 + 2) * W);
  int da[128];
#define D(i, j) d[((i) + 1) * W + ((j) + 1)]
  memset(da, 0, sizeof(da));
  D(-1, -1) = DBL_MAX;
  for (int i = 0; i <= n1; ++i) D(i, -1) = DBL_MAX;
  for (int j = 0; j <= n2; ++j) D(-1, j) = DBL_MAX;
  D(0, 0) = 0.0;
  for (int i = 1; i <= n1; ++i) D(i, 0) = D(i - 1, 0) + del[s1[i - 1]];
  for (int j = 1; j <= n2; ++j) D(0, j) = D(0, j - 1) + ins[s2[j - 1]];
  for (int i = 1; i <= n1; ++i) {
    const uint8_t ci = s1[i - 1];
    int db = 0;
    for (int j = 1; j <= n2; ++j) {
      const uint8_t cj = s2[j - 1];
      const int k = da[cj];
      const int l = db;
      double cost;
      if (ci == cj) {
        cost = 0.0;
        db = j;
      } else {
        cost = sub[ci * 128 + cj];
      }
      double best = D(i - 1, j) + del[ci];
      double c2 = D(i, j - 1) + ins[cj];
      if (c2 < best) best = c2;
      c2 = D(i - 1, j - 1) + cost;
      if (c2 < best) best = c2;
      if (k > 0 && l > 0) {
        /* delete s1[k+1..i-1], transpose, insert s2[l+1..j-1] */
        const double del_range = D(i - 1, 0) - D(k, 0);
        const double ins_range = D(0, j - 1) - D(0, l);
        const double t = tr ? tr[s1[k - 1] * 128 + ci] : 1.0;
        c2 = D(k - 1, l - 1) + del_range + t + ins_range;
        if (c2 < best) best = c2;
      }
      D(i, j) = best;
    }
    da[ci] = i;
  }
  const double out = D(n1, n2);
#undef D
  free(d);
  return out;
}

/* a, b: n nibble-valued pixels; S: 16x16 integer substitution costs. */
int32_t oracle_chain_distance(const uint8_t* a, const uint8_t* b, int n,
                              const int32_t* S) {
  int32_t prev2 = 0, prev1 = 0;
  for (int t = 0; t < n; ++t) {
    int32_t cur = prev1 + (a[t] == b[t] ? 0 : S[a[t] * 16 + b[t]]);
    if (t >= 1 && a[t - 1] == b[t] && a[t] == b[t - 1] && prev2 + 1 < cur)
      cur = prev2 + 1;
    prev2 = prev1;
    prev1 = cur;
  }
  return prev1;
}

static const char kPixelChars[] = "0123456789ABCDEF"; /* make_data_tables.py:18 */

/* make_data_tables.py:30-52, 73-89: 128x128 float64 matrices at ASCII codes. */
static void fill_costs(const int32_t* lut, double* ins, double* del,
                       double* sub) {
  for (int i = 0; i < 128; ++i) ins[i] = del[i] = 100000.0;
  memset(sub, 0, sizeof(double) * 128 * 128);
  for (int i = 0; i < 16; ++i)
    for (int j = 0; j < 16; ++j) {
      const double c = (double)lut[i * 16 + j];
      sub[kPixelChars[i] * 128 + kPixelChars[j]] = c;
      sub[kPixelChars[j] * 128 + kPixelChars[i]] = c;
    }
}

typedef struct {
  int mode, algo, triangular;
  const int32_t* lut;
  const uint8_t* pix;
  const double *ins, *del, *sub;
  uint16_t* out;
  atomic_llong evaluated;
  int64_t row_begin, row_step;
} build_ctx;

static void build_row(int64_t k, void* p) {
  build_ctx* c = (build_ctx*)p;
  const int64_t i = c->row_begin + k * c->row_step;
  const int bits = kBits[c->mode], n = kDots[c->mode], noff = kOffsets[c->mode];
  const uint32_t N = 1u << bits;
  uint8_t sa[32], sb[32];
  int64_t evaluated = 0;
  for (int o = 0; o < noff; ++o) {
    const uint8_t* a = c->pix + ((size_t)o * N + i) * n;
    uint16_t* row = c->out + ((size_t)o << (2 * bits)) + ((size_t)i << bits);
    const uint32_t jend = c->triangular ? (uint32_t)i : N;
    if (c->algo == 0)
      for (int t = 0; t < n; ++t) sa[t] = (uint8_t)kPixelChars[a[t]];
    for (uint32_t j = 0; j < jend; ++j) {
      const uint8_t* b = c->pix + ((size_t)o * N + j) * n;
      double r;
      if (c->algo == 0) {
        for (int t = 0; t < n; ++t) sb[t] = (uint8_t)kPixelChars[b[t]];
        r = oracle_dam_lev(sa, n, sb, n, c->ins, c->del, c->sub, NULL);
      } else {
        r = (double)oracle_chain_distance(a, b, n, c->lut);
      }
      row[j] = (uint16_t)r; /* reference asserts 0 <= r < 2^16 (:107) */
      ++evaluated;
    }
  }
  atomic_fetch_add(&c->evaluated, evaluated);
}

/*
 * compute_edit_distance restated.  out: uint16[n_off][2^(2*bits)] (whole
 * table; only rows [row_begin,row_end) of the source index i are written).
 * algo 0 = faithful dam_lev on ASCII pixel strings, 1 = 1-D chain recurrence.
 * triangular != 0 stores only j < i (the reference's file layout, :156-172);
 * otherwise the full symmetric square.
 * Returns the number of entries evaluated.
 */
 /* rows row_begin, row_begin + row_step, ... < row_end */
int64_t oracle_build_table_strided(int mode, const int32_t* lut, uint16_t* out,
                                   uint32_t row_begin, uint32_t row_end,
                                   uint32_t row_step, int algo, int triangular) {
  const int bits = kBits[mode], n = kDots[mode], noff = kOffsets[mode];
  const uint32_t N = 1u << bits;
  uint8_t* pix = (uint8_t*)malloc((size_t)noff * N * n);
  oracle_all_pixel_strings(mode, pix);
  double ins[128], del[128];
  double* sub = (double*)malloc(sizeof(double) * 128 * 128);
  fill_costs(lut, ins, del, sub);
  build_ctx c;
  c.mode = mode;
  c.algo = algo;
  c.triangular = triangular;
  c.lut = lut;
  c.pix = pix;
  c.ins = ins;
  c.del = del;
  c.sub = sub;
  c.out = out;
  atomic_init(&c.evaluated, 0);
  if (row_step == 0) row_step = 1;
  c.row_begin = row_begin;
  c.row_step = row_step;
  const int64_t n_rows =
      row_end > row_begin ? ((int64_t)row_end - row_begin + row_step - 1) / row_step : 0;
  /* a faithful row is ~10^4 full DPs: hand rows out one at a time */
  parallel_rows(0, n_rows, algo == 0 ? 1 : 4, build_row, &c);
  free(sub);
  free(pix);
  return (int64_t)atomic_load(&c.evaluated);
}

int64_t oracle_build_table(int mode, const int32_t* lut, uint16_t* out,
                           uint32_t row_begin, uint32_t row_end, int algo,
                           int triangular) {
  return oracle_build_table_strided(mode, lut, out, row_begin, row_end, 1, algo,
                                    triangular);
}

typedef struct {
  int bits;
  uint16_t* t;
} sym_ctx;

static void sym_row(int64_t i, void* p) {
  sym_ctx* c = (sym_ctx*)p;
  const int bits = c->bits;
  uint16_t* t = c->t;
  uint16_t* dg = t + ((size_t)i << bits) + i;
  *dg = (uint16_t)(*dg + *dg);
  for (uint32_t j = 0; j < (uint32_t)i; ++j) {
    uint16_t* lo = t + ((size_t)i << bits) + j;
    uint16_t* up = t + ((size_t)j << bits) + i;
    const uint16_t s = (uint16_t)(*lo + *up);
    *lo = s;
    *up = s;
  }
}

/*
 * screen.py:343-367 loader: dist[o, transpose] += dist[o, identity], i.e.
 * new[y] = old[y] + old[T(y)] for every y (numpy evaluates the right-hand side
 * before scattering, and T is a permutation), so the diagonal doubles and a
 * lower-triangular input becomes symmetric.  uint16 wrap-around as in numpy.
 */
void oracle_symmetrise(int mode, uint16_t* table) {
  sym_ctx c;
  c.bits = kBits[mode];
  for (int o = 0; o < kOffsets[mode]; ++o) {
    c.t = table + ((size_t)o << (2 * c.bits));
    parallel_rows(0, (int64_t)1 << c.bits, 32, sym_row, &c);
  }
}

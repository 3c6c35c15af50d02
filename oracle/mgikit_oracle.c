/*
 * mgikit_oracle.c — TEST INFRASTRUCTURE ONLY (see mgikit_oracle.h).
 *
 * Plain-C restatement of the reference's demultiplex hot path.  It follows the
 * reference's own data structures (a dictionary of every <=m-substitution
 * variant of every index, looked up by the observed bytes) rather than the
 * XOR/popcount formulation of the CUDA path, so the two implementations share
 * no algorithmic shortcut.  Citations are relative to /root/reference/.
 *
 * Parity: pinned — see oracle/README.md (reference fixtures + reference binary).
 */
#define _POSIX_C_SOURCE 200809L
#include "mgikit_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* growable byte buffer = SampleReads::compression_buffer (src/sample_data.rs:331-337, :392-398) */
typedef struct {
  uint8_t* p;
  size_t n, cap;
} bytebuf;

static void bb_add(bytebuf* b, const void* data, size_t len) {
  if (b->n + len > b->cap) {
    size_t nc = b->cap ? b->cap * 2 : 4096;
    while (nc < b->n + len) nc *= 2;
    b->p = (uint8_t*)realloc(b->p, nc);
    b->cap = nc;
  }
  memcpy(b->p + b->n, data, len);
  b->n += len;
}

/* ------------------------------------------------------------------------- */
/* dictionary: observed index bytes -> (indexes at the minimum distance, that distance)
 * == the HashMap returned by get_all_mismatches(), src/sample_manager.rs:732-850 */
typedef struct {
  uint8_t key[MGK_MAX_INDEX_LEN];
  int klen;
  int level; /* -1 = empty slot */
  int n, cap;
  int* ids;
} dent;

typedef struct {
  dent* e;
  size_t cap, n;
} dict;

static uint64_t hash_bytes(const uint8_t* k, int len) {
  uint64_t h = 1469598103934665603ULL;
  for (int i = 0; i < len; ++i) {
    h ^= k[i];
    h *= 1099511628211ULL;
  }
  return h ^ (h >> 29);
}

static void dict_init(dict* d, size_t cap) {
  size_t c = 64;
  while (c < cap) c <<= 1;
  d->e = (dent*)malloc(c * sizeof(dent));
  for (size_t i = 0; i < c; ++i) {
    d->e[i].level = -1;
    d->e[i].ids = NULL;
    d->e[i].n = d->e[i].cap = 0;
  }
  d->cap = c;
  d->n = 0;
}

static void dict_free(dict* d) {
  if (!d->e) return;
  for (size_t i = 0; i < d->cap; ++i) free(d->e[i].ids);
  free(d->e);
  d->e = NULL;
  d->cap = d->n = 0;
}

static dent* dict_slot(dict* d, const uint8_t* k, int len) {
  size_t i = hash_bytes(k, len) & (d->cap - 1);
  for (;;) {
    dent* e = &d->e[i];
    if (e->level < 0) return e;
    if (e->klen == len && memcmp(e->key, k, (size_t)len) == 0) return e;
    i = (i + 1) & (d->cap - 1);
  }
}

static const dent* dict_get(const dict* d, const uint8_t* k, int len) {
  if (!d->e || d->n == 0 || len <= 0 || len > MGK_MAX_INDEX_LEN) return NULL;
  const dent* e = dict_slot((dict*)d, k, len);
  return e->level < 0 ? NULL : e;
}

static void dict_grow(dict* d) {
  dict nd;
  dict_init(&nd, d->cap * 2);
  for (size_t i = 0; i < d->cap; ++i) {
    dent* o = &d->e[i];
    if (o->level < 0) continue;
    dent* s = dict_slot(&nd, o->key, o->klen);
    *s = *o;
    nd.n++;
  }
  free(d->e);
  *d = nd;
}

/* Keep, per key, only the list at the minimum level — the reduction at
 * src/sample_manager.rs:837-847.  An index itself is level 0 and overrides
 * whatever was there (src/sample_manager.rs:750-761). */
static void dict_put(dict* d, const uint8_t* k, int len, int id, int level) {
  if ((d->n + 1) * 2 > d->cap) dict_grow(d);
  dent* e = dict_slot(d, k, len);
  if (e->level < 0) {
    memcpy(e->key, k, (size_t)len);
    e->klen = len;
    e->level = level;
    e->n = 0;
    d->n++;
  } else if (level < e->level) {
    e->level = level;
    e->n = 0;
  } else if (level > e->level) {
    return;
  }
  if (e->n == e->cap) {
    e->cap = e->cap ? e->cap * 2 : 2;
    e->ids = (int*)realloc(e->ids, (size_t)e->cap * sizeof(int));
  }
  e->ids[e->n++] = id;
}

/* every variant of `index` with exactly `k` substitutions at positions >= from,
 * letters from the dictionary alphabet (src/sample_manager.rs:740) */
static const char NEC_LS[5] = {'A', 'C', 'T', 'G', 'N'};

static void enum_variants(dict* d, uint8_t* work, const uint8_t* index, int len, int from,
                          int left, int level, int id) {
  if (left == 0) {
    dict_put(d, work, len, id, level);
    return;
  }
  for (int pos = from; pos <= len - left; ++pos) {
    for (int c = 0; c < 5; ++c) {
      if ((uint8_t)NEC_LS[c] == index[pos]) continue;
      work[pos] = (uint8_t)NEC_LS[c];
      enum_variants(d, work, index, len, pos + 1, left - 1, level, id);
    }
    work[pos] = index[pos];
  }
}

/* get_all_mismatches(), src/sample_manager.rs:732-850 */
static void get_all_mismatches(dict* d, const char* indexes, int n, int len, int m) {
  d->e = NULL;
  d->cap = d->n = 0;
  if (n == 0 || len <= 0) return; /* :736-738 */
  dict_init(d, 1024);
  uint8_t work[MGK_MAX_INDEX_LEN];
  for (int i = 0; i < n; ++i) {
    const uint8_t* idx = (const uint8_t*)indexes + (size_t)i * len;
    dict_put(d, idx, len, i, 0);
    for (int k = 1; k <= m && k <= len; ++k) {
      memcpy(work, idx, (size_t)len);
      enum_variants(d, work, idx, len, 0, k, k, i);
    }
  }
}

/* ------------------------------------------------------------------------- */
typedef struct {
  int32_t g[10];
  int has_i5;
  int n_i7, n_i5;
  char* i7;
  char* i5;
  int32_t* pair;
  dict d7, d5;
} otpl;

struct oslot {
  int state; /* 0 free, 1 submitted, 2 collected */
  uint64_t seq, order;
  bytebuf* bin_r2; /* [total] */
  bytebuf* bin_r1;
  bytebuf out_r2, out_r1;
  uint64_t *off2, *len2, *off1, *len1;
  uint32_t* unassigned;
  size_t n_unassigned, cap_unassigned;
  mgk_result res;
};

struct orc_ctx {
  mgk_config cfg;
  char* prefix;
  size_t prefix_len;
  int32_t* writing;
  int n_tpl;
  otpl* tpl;
  int S, total, BL, mism_w;
  int reformat;     /* process_buffer with demultiplex == false */
  char* rf_barcode; /* ReformatedSample::barcode, src/formater.rs:10-11 */
  size_t rf_barcode_len;
  uint64_t* stats; /* [total][10] */
  uint64_t* mism;  /* [total][2m+2] */
  uint64_t* red_stats; /* result of the last orc_allreduce_counters */
  uint64_t* red_mism;
  int reduced;
  /* result slots: cfg.n_slots chunks may be submitted before the first is released */
  struct oslot* slots;
  int n_slots;
  uint64_t submit_clock;
  char err[512];
};

static char g_create_err[512];

const char* orc_last_error(const orc_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int orc_create(const mgk_config* cfg, orc_ctx** out) {
  if (!cfg || !out || cfg->abi_version != MGK_ABI_VERSION || cfg->n_templates < (cfg->reformat ? 0 : 1) ||
      cfg->n_templates > MGK_MAX_TEMPLATES || cfg->n_samples < 1 ||
      (cfg->reformat && (cfg->n_samples != 1 || cfg->n_templates != 0 || cfg->reformat_umi_length < 0 ||
                         cfg->reformat_umi_length > MGK_MAX_BARCODE_LEN || !cfg->read2_has_sequence ||
                         cfg->out_mode == MGK_OUT_MGI_FULL_HEADER))) { /* src/lib.rs:2302-2322, run_manager.rs:137 */
    snprintf(g_create_err, sizeof g_create_err, "orc_create: bad config");
    return MGK_ERR_ARG;
  }
  orc_ctx* c = (orc_ctx*)calloc(1, sizeof *c);
  c->cfg = *cfg;
  c->S = cfg->n_samples;
  c->total = c->S + 2;
  c->mism_w = 2 * cfg->allowed_mismatches + 2; /* src/report_manager.rs:268 */
  c->prefix = strdup(cfg->illumina_prefix ? cfg->illumina_prefix : "");
  c->prefix_len = strlen(c->prefix);
  c->writing = (int32_t*)malloc(sizeof(int32_t) * (size_t)c->total);
  memcpy(c->writing, cfg->writing_sample, sizeof(int32_t) * (size_t)c->total);
  c->n_tpl = cfg->n_templates;
  c->tpl = (otpl*)calloc((size_t)(c->n_tpl ? c->n_tpl : 1), sizeof(otpl));
  for (int t = 0; t < c->n_tpl; ++t) {
    const mgk_template* s = &cfg->templates[t];
    otpl* o = &c->tpl[t];
    memcpy(o->g, s->geometry, sizeof o->g);
    o->has_i5 = s->has_i5;
    o->n_i7 = s->n_i7;
    o->n_i5 = s->n_i5;
    if (o->g[1] > MGK_MAX_INDEX_LEN || o->g[4] > MGK_MAX_INDEX_LEN || o->g[9] > MGK_MAX_BARCODE_LEN) {
      snprintf(g_create_err, sizeof g_create_err, "orc_create: index/barcode too long");
      return MGK_ERR_ARG;
    }
    size_t b7 = (size_t)o->n_i7 * (size_t)o->g[1], b5 = (size_t)o->n_i5 * (size_t)o->g[4];
    o->i7 = (char*)malloc(b7 + 1);
    memcpy(o->i7, s->i7, b7);
    o->i5 = (char*)malloc(b5 + 1);
    if (b5) memcpy(o->i5, s->i5, b5);
    size_t np = o->has_i5 ? (size_t)o->n_i7 * (size_t)o->n_i5 : (size_t)o->n_i7;
    o->pair = (int32_t*)malloc(sizeof(int32_t) * (np ? np : 1));
    memcpy(o->pair, s->pair_sample, sizeof(int32_t) * np);
    /* per worker, per template: src/lib.rs:1412-1415 */
    get_all_mismatches(&o->d7, o->i7, o->n_i7, o->g[1], cfg->allowed_mismatches);
    get_all_mismatches(&o->d5, o->i5, o->n_i5, o->g[4], cfg->allowed_mismatches);
  }
  c->reformat = cfg->reformat != 0;
  c->rf_barcode = strdup(cfg->reformat && cfg->reformat_barcode ? cfg->reformat_barcode : "");
  c->rf_barcode_len = strlen(c->rf_barcode);
  c->BL = c->reformat ? cfg->reformat_umi_length : c->tpl[0].g[9]; /* src/lib.rs:878-882 */
  c->stats = (uint64_t*)calloc((size_t)c->total * MGK_N_STATS, sizeof(uint64_t));
  c->mism = (uint64_t*)calloc((size_t)c->total * (size_t)c->mism_w, sizeof(uint64_t));
  c->red_stats = (uint64_t*)calloc((size_t)c->total * MGK_N_STATS, sizeof(uint64_t));
  c->red_mism = (uint64_t*)calloc((size_t)c->total * (size_t)c->mism_w, sizeof(uint64_t));
  c->n_slots = cfg->n_slots < 1 ? 1 : (cfg->n_slots > 8 ? 8 : cfg->n_slots);
  c->slots = (struct oslot*)calloc((size_t)c->n_slots, sizeof(struct oslot));
  for (int k = 0; k < c->n_slots; ++k) {
    struct oslot* s = &c->slots[k];
    s->bin_r2 = (bytebuf*)calloc((size_t)c->total, sizeof(bytebuf));
    s->bin_r1 = (bytebuf*)calloc((size_t)c->total, sizeof(bytebuf));
    s->off2 = (uint64_t*)calloc((size_t)c->total * 4, sizeof(uint64_t));
    s->len2 = s->off2 + c->total;
    s->off1 = s->len2 + c->total;
    s->len1 = s->off1 + c->total;
  }
  *out = c;
  return MGK_OK;
}

void orc_destroy(orc_ctx* c) {
  if (!c) return;
  for (int t = 0; t < c->n_tpl; ++t) {
    free(c->tpl[t].i7);
    free(c->tpl[t].i5);
    free(c->tpl[t].pair);
    dict_free(&c->tpl[t].d7);
    dict_free(&c->tpl[t].d5);
  }
  for (int k = 0; k < c->n_slots; ++k) {
    struct oslot* s = &c->slots[k];
    for (int b = 0; b < c->total; ++b) {
      free(s->bin_r2[b].p);
      free(s->bin_r1[b].p);
    }
    free(s->bin_r2);
    free(s->bin_r1);
    free(s->out_r2.p);
    free(s->out_r1.p);
    free(s->off2);
    free(s->unassigned);
  }
  free(c->slots);
  free(c->tpl);
  free(c->red_stats);
  free(c->red_mism);
  free(c->stats);
  free(c->mism);
  free(c->writing);
  free(c->prefix);
  free(c->rf_barcode);
  free(c);
}

/* ------------------------------------------------------------------------- */
/* translate an index id of template `from` into template `to` (only differs
 * when the dictionary iterator lags, see find_matching_sample below) */
static int xlat(const char* from_seqs, int from_len, int id, const char* to_seqs, int to_n, int to_len) {
  if (from_len != to_len) return -1;
  const char* s = from_seqs + (size_t)id * from_len;
  for (int j = 0; j < to_n; ++j)
    if (memcmp(to_seqs + (size_t)j * to_len, s, (size_t)to_len) == 0) return j;
  return -1;
}

typedef struct {
  int sample_id;
  int curr_mismatch;
  int bar_t; /* template whose geometry cut curr_barcode, -1 = empty string */
  int umi_t; /* template whose geometry cut curr_umi, -1 = empty string */
  int panicked;
} match_out;

/* find_matching_sample(), src/lib.rs:223-443.  `bar` = the last BL bytes of the
 * barcode read's sequence line (src/lib.rs:1012). */
static match_out find_matching_sample(const orc_ctx* c, const uint8_t* bar) {
  const int total = c->total;
  const int und = total - 2, amb = total - 1; /* :243-244 */
  const int m = c->cfg.allowed_mismatches, M = c->cfg.total_allowed_mismatches;
  const int BL = c->BL;
  match_out r;
  r.sample_id = und;
  r.bar_t = r.umi_t = -1;
  r.panicked = 0;
  long latest = -1; /* usize::MAX == "infinity" */
  long curr = -1;
#define INF_GT(a, b) (((a) < 0) ? ((b) >= 0) : ((b) >= 0 && (a) > (b))) /* a > b with -1 = inf */
#define INF_LT(a, b) (((a) < 0) ? 0 : ((b) < 0 ? 1 : (a) < (b)))          /* a < b */
  int di = 0; /* template_itr: the dictionary iterator (:242, :437) */
  for (int t = 0; t < c->n_tpl; ++t) {
    const otpl* T = &c->tpl[t];
    const otpl* D = &c->tpl[di];
    const int32_t* g = T->g;
    const int extract_umi = g[6] == 1; /* :260 */
    const dent* e7 = dict_get(&D->d7, bar + BL - g[2], g[1]); /* :262-265 */
    if (e7 && e7->level <= m) {                              /* :268 */
      if (T->has_i5) {                                       /* :269 */
        const dent* e5 = dict_get(&D->d5, bar + BL - g[5], g[4]); /* :270-273 */
        if (e5) {
          for (int a = 0; a < e7->n; ++a) {   /* :276 */
            for (int b = 0; b < e5->n; ++b) { /* :277 */
              long tot = e7->level + e5->level;
              if (e5->level > m || INF_LT(latest, tot)) continue; /* :278-282 */
              curr = tot;                                          /* :283 */
              if (curr <= M) {                                     /* :284 */
                int ia = e7->ids[a], ib = e5->ids[b];
                if (di != t) {
                  ia = xlat(D->i7, D->g[1], ia, T->i7, T->n_i7, T->g[1]);
                  if (ia < 0) { /* sample_info.get(..).unwrap() on None, :286-288 */
                    r.panicked = 1;
                    return r;
                  }
                  ib = xlat(D->i5, D->g[4], ib, T->i5, T->n_i5, T->g[4]);
                }
                int s = (ib < 0) ? -1 : T->pair[(size_t)ia * T->n_i5 + ib]; /* :286-291 */
                if (s >= 0) {
                  if (r.sample_id >= total || INF_GT(latest, curr)) { /* :293-295 */
                    r.sample_id = s;
                    latest = curr;
                    r.bar_t = t;                 /* :298-320 */
                    if (extract_umi) r.umi_t = t; /* :322-358 (i7_rc is always false, :1004) */
                  } else {
                    r.sample_id = amb; /* :360-361 */
                    break;
                  }
                }
              }
            }
          }
        }
      } else {
        if (INF_LT(latest, (long)e7->level)) { /* :375-376: `continue` skips template_itr += 1 */
          continue;
        } else if (INF_GT(latest, (long)e7->level) && e7->n < 2) { /* :377 */
          curr = e7->level;
          latest = curr;
          int ia = e7->ids[0];
          if (di != t) ia = xlat(D->i7, D->g[1], ia, T->i7, T->n_i7, T->g[1]);
          if (ia < 0) { /* "There should be a value here", :382-384 */
            r.panicked = 1;
            return r;
          }
          r.sample_id = T->pair[ia]; /* :380-381 */
          r.bar_t = t;               /* :387-393 */
          if (extract_umi) r.umi_t = t; /* :394-405 */
        } else {
          r.sample_id = amb; /* :425 */
        }
      }
    }
    if (r.sample_id != und && r.sample_id < total && !c->cfg.comprehensive_scan) break; /* :433-435 */
    di += 1; /* :437 */
  }
  if (r.sample_id >= und) curr = 0; /* :439-441 */
  r.curr_mismatch = (int)curr;
  return r;
}

/* ------------------------------------------------------------------------- */
/* write_illumina_header() free fn, src/sample_data.rs:569-654.
 * h = the header slice handed over: the MGI header INCLUDING its '\n' (reformat of raw MGI headers: without
 * its last two bytes, src/lib.rs:1246); returns bytes written to out.
 * sb_header (reformat): [":" umi] + what the input header holds from sep on (:627-637);
 * else (demultiplex): umi (its ':' included) + " " n ":N:0:" (:638-652). */
static size_t illumina_fields(uint8_t* out, const uint8_t* h, size_t hl, size_t L, size_t sep,
                              const uint8_t* umi, size_t umi_len, int sb_header) {
  size_t e = 0;
  out[e++] = h[L + 1];
  out[e++] = ':';
  for (size_t i = L + 10; i < sep; ++i) { /* :584-591 */
    if (h[i] != '0') {
      memcpy(out + e, h + i, sep - i);
      e += sep - i;
      break;
    }
  }
  out[e++] = ':';
  for (int f = 0; f < 2; ++f) { /* :596-626 */
    size_t o = L + 3 + 4 * (size_t)f;
    if (h[o] != '0') {
      out[e++] = h[o];
      out[e++] = h[o + 1];
      out[e++] = h[o + 2];
    } else if (h[o + 1] != '0') {
      out[e++] = h[o + 1];
      out[e++] = h[o + 2];
    } else {
      out[e++] = h[o + 2];
    }
    if (f == 0) out[e++] = ':';
  }
  if (sb_header) {
    if (umi_len) { /* :628-633 */
      out[e++] = ':';
      memcpy(out + e, umi, umi_len);
      e += umi_len;
    }
    memcpy(out + e, h + sep, hl - sep - 1); /* :634-637 */
    e += hl - sep - 1;
    return e;
  }
  if (umi_len) { /* :640-643 */
    memcpy(out + e, umi, umi_len);
    e += umi_len;
  }
  out[e++] = ' ';
  out[e++] = h[hl - 2]; /* :647 */
  memcpy(out + e, ":N:0:", 5); /* HEADER_TAIL :14 */
  e += 5;
  return e;
}

/* sum_qc(), src/lib.rs:458-474 (the range check is done by validate below) */
static void sum_qc(const uint8_t* q, size_t n, uint64_t* total, uint64_t* high) {
  uint64_t t = 0, h = 0;
  for (size_t i = 0; i < n; ++i) {
    if (q[i] >= 63) h++;
    t += q[i];
  }
  *total = t;
  *high = h;
}

/* check_read_content(), src/lib.rs:476-509.  rec = header_start..read_end */
static int check_read_content(const uint8_t* rec, size_t len, size_t seq_start, size_t plus_start,
                              size_t qc_start) {
  if (rec[0] != '@') return MGK_VAL_HEADER_AT;
  for (size_t i = seq_start; i + 1 < plus_start; ++i) {
    uint8_t n = rec[i];
    if (!(n == 'A' || n == 'C' || n == 'G' || n == 'T' || n == 'N' || n == 'a' || n == 'c' ||
          n == 'g' || n == 't' || n == 'n'))
      return MGK_VAL_BASE;
  }
  for (size_t i = qc_start; i < len; ++i)
    if (rec[i] > 73 || rec[i] < 33) return MGK_VAL_QUALITY;
  return MGK_VAL_OK;
}

static size_t memchr_all(const uint8_t* buf, size_t len, uint32_t** out) {
  size_t cap = len / 64 + 16, n = 0;
  uint32_t* p = (uint32_t*)malloc(cap * sizeof(uint32_t));
  const uint8_t* s = buf;
  const uint8_t* end = buf + len;
  while (s < end) {
    const uint8_t* q = (const uint8_t*)memchr(s, '\n', (size_t)(end - s));
    if (!q) break;
    if (n == cap) {
      cap *= 2;
      p = (uint32_t*)realloc(p, cap * sizeof(uint32_t));
    }
    p[n++] = (uint32_t)(q - buf);
    s = q + 1;
  }
  *out = p;
  return n;
}

int orc_scan_newlines(orc_ctx* ctx, const uint8_t* buf, uint64_t len, uint32_t* positions,
                      uint64_t capacity, uint64_t* count) {
  (void)ctx;
  uint32_t* p;
  size_t n = memchr_all(buf, (size_t)len, &p);
  *count = n;
  memcpy(positions, p, sizeof(uint32_t) * (n < capacity ? n : (size_t)capacity));
  free(p);
  return n <= capacity ? MGK_OK : MGK_ERR_CAPACITY;
}

int orc_match_barcodes(orc_ctx* c, const uint8_t* barcodes, uint64_t n, int32_t* sample_id,
                       int32_t* mismatches) {
  for (uint64_t i = 0; i < n; ++i) {
    match_out r = find_matching_sample(c, barcodes + i * (uint64_t)c->BL);
    if (r.panicked) {
      snprintf(c->err, sizeof c->err, "reference would panic (stale dictionary lookup) at barcode %llu",
               (unsigned long long)i);
      return MGK_ERR_FORMAT;
    }
    int s = r.sample_id;
    if (s >= c->total) s = c->total - 2; /* src/lib.rs:1020-1022 */
    sample_id[i] = s;
    mismatches[i] = r.curr_mismatch;
  }
  return MGK_OK;
}

/* ------------------------------------------------------------------------- */
/* process_buffer(), src/lib.rs:840-1348, demultiplex == true */
int orc_submit(orc_ctx* c, uint64_t seq, const uint8_t* r2, uint64_t r2_len, const uint8_t* r1,
               uint64_t r1_len, uint32_t flags) {
  (void)flags;
  struct oslot* S = NULL;
  for (int k = 0; k < c->n_slots && !S; ++k)
    if (c->slots[k].state == 0) S = &c->slots[k];
  if (!S) {
    snprintf(c->err, sizeof c->err, "orc_submit: all slots are in flight; collect and release first");
    return MGK_ERR_STATE;
  }
  const mgk_config* cfg = &c->cfg;
  const int paired = cfg->paired;
  const int total = c->total, und = total - 2, amb = total - 1;
  const size_t BL = (size_t)c->BL;
  const size_t wbl = cfg->keep_barcode ? 0 : BL; /* :883-886 */
  const size_t shift = paired ? 1 : 0;           /* :873-877 */
  const size_t L = (size_t)cfg->l_position;
  const int illumina = cfg->out_mode == MGK_OUT_ILLUMINA;
  const int full_hdr = cfg->out_mode == MGK_OUT_MGI_FULL_HEADER;

  for (int b = 0; b < total; ++b) S->bin_r2[b].n = S->bin_r1[b].n = 0;
  S->n_unassigned = 0;
  memset(&S->res, 0, sizeof S->res);

  uint32_t *nl2 = NULL, *nl1 = NULL;
  size_t n2 = memchr_all(r2, (size_t)r2_len, &nl2); /* src/lib.rs:1526 */
  size_t n1 = paired ? memchr_all(r1, (size_t)r1_len, &nl1) : 0; /* :1531 */

  size_t header_start = 0, header_start_pr = 0;
  uint64_t read_cntr = 0;
  int rc = MGK_OK;
  uint8_t* hdr = NULL; /* scratch for one rewritten header */
  size_t hdr_cap = 0;

  for (size_t i = 0;; ++i) {
    if (4 * i + 3 >= n2) break;            /* :929-956 reached_an_end */
    if (paired && 4 * i + 3 >= n1) break;  /* :959-987 */
    size_t seq_start = nl2[4 * i] + 1, plus_start = nl2[4 * i + 1] + 1, qual_start = nl2[4 * i + 2] + 1,
           read_end = nl2[4 * i + 3];
    size_t seq_start_pr = 0, plus_start_pr = 0, qual_start_pr = 0, read_end_pr = 0;
    if (paired) {
      seq_start_pr = nl1[4 * i] + 1;
      plus_start_pr = nl1[4 * i + 1] + 1;
      qual_start_pr = nl1[4 * i + 2] + 1;
      read_end_pr = nl1[4 * i + 3];
    }
    const size_t rec_start = header_start, rec_start_pr = header_start_pr;
    const size_t hl = seq_start - header_start; /* header incl. '\n' */
    /* slices the reference takes would panic / wrap on these (:1012, :1198, :1322) */
    if (plus_start - seq_start < BL + 1 || read_end - qual_start < BL || hl < 3 ||
        (illumina && hl < L + 10)) {
      snprintf(c->err, sizeof c->err, "record %llu: lines shorter than the barcode/header layout",
               (unsigned long long)i);
      rc = MGK_ERR_FORMAT;
      break;
    }
    if (hdr_cap < c->prefix_len + hl + 1024) {
      hdr_cap = c->prefix_len + hl + 1024;
      hdr = (uint8_t*)realloc(hdr, hdr_cap);
    }
    size_t sep_position = seq_start - header_start - 3; /* :1007 */
    const uint8_t* bar = r2 + plus_start - 1 - BL;      /* :1012 */
    size_t raw_shift = 0, header_shift = 0, rf_tail_offset = 0;
    match_out mo;
    if (c->reformat) { /* :1090-1147 */
      memset(&mo, 0, sizeof mo);
      mo.bar_t = mo.umi_t = -1;
      const uint8_t* sp = (const uint8_t*)memchr(r2 + header_start, ' ', hl);
      if (sp) {
        sep_position = (size_t)(sp - (r2 + header_start));
      } else { /* "You might be using MGI raw data" */
        raw_shift = 2;
        sep_position = hl - 3;
      }
      /* header_shift = seq_start - sep_position - header_start - 3 (:1129) wraps for a space in the last two
       * header bytes, and hamming_distance(..).unwrap() (:1108-1112) panics on barcodes of another length */
      if (sep_position > hl - 3 ||
          (c->rf_barcode_len && sep_position != hl - 3 && hl - 1 - (sep_position + 7) != c->rf_barcode_len) ||
          (c->rf_barcode_len && sep_position != hl - 3 && sep_position + 7 > hl - 1)) {
        snprintf(c->err, sizeof c->err, "record %llu: header does not end with ` n:N:0:<barcode of %zu bases>`",
                 (unsigned long long)i, c->rf_barcode_len);
        rc = MGK_ERR_FORMAT;
        break;
      }
      if (c->rf_barcode_len && sep_position != hl - 3) {
        const uint8_t* hb = r2 + header_start + sep_position + 7;
        for (size_t k = 0; k < c->rf_barcode_len; ++k) mo.curr_mismatch += hb[k] != (uint8_t)c->rf_barcode[k];
      }
      if (mo.curr_mismatch > 4) mo.curr_mismatch = 4; /* :1140-1141 */
      mo.sample_id = 0;                               /* :1114 */
      header_shift = hl - sep_position - 3;           /* :1129-1130 */
      rf_tail_offset = header_shift + 1;
      if (c->rf_barcode_len && sep_position == hl - 3) { /* :1132-1137 */
        header_shift = 0;
        rf_tail_offset = c->rf_barcode_len + 6;
      }
      if (paired && illumina && seq_start_pr - header_start_pr < header_shift + 2) { /* :1275 would read in front of the record */
        snprintf(c->err, sizeof c->err, "record %llu: the paired read's header is shorter than the barcode read's tail",
                 (unsigned long long)i);
        rc = MGK_ERR_FORMAT;
        break;
      }
    } else {
      mo = find_matching_sample(c, bar);
    }
    if (mo.panicked) {
      snprintf(c->err, sizeof c->err, "record %llu: reference would panic (stale dictionary lookup)",
               (unsigned long long)i);
      rc = MGK_ERR_FORMAT;
      break;
    }
    int sample_id = mo.sample_id;
    if (sample_id >= total) sample_id = und; /* :1020-1022 */
    /* curr_barcode as returned by find_matching_sample (:298-320, :387-393) */
    uint8_t barcode[2 * MGK_MAX_INDEX_LEN + 2];
    size_t barcode_len = 0;
    if (mo.bar_t >= 0) {
      const int32_t* g = c->tpl[mo.bar_t].g;
      memcpy(barcode, bar + BL - g[2], (size_t)g[1]);
      barcode_len = (size_t)g[1];
      if (c->tpl[mo.bar_t].has_i5) {
        barcode[barcode_len++] = '+';
        memcpy(barcode + barcode_len, bar + BL - g[5], (size_t)g[4]);
        barcode_len += (size_t)g[4];
      }
    }
    uint8_t umi[MGK_MAX_BARCODE_LEN + 2];
    size_t umi_len = 0;
    if (mo.umi_t >= 0) {
      const int32_t* g = c->tpl[mo.umi_t].g;
      umi[0] = ':';
      memcpy(umi + 1, bar + BL - g[8], (size_t)g[7]);
      umi_len = 1 + (size_t)g[7];
    }
    if (c->reformat && BL > 0) { /* :1117-1124: the last umi_length bases, no ':' */
      memcpy(umi, r2 + plus_start - 1 - BL, BL);
      umi_len = BL;
    }
    const size_t tail_offset = c->reformat ? rf_tail_offset : barcode_len + 6; /* :1019 */

    if (sample_id == und || sample_id == amb) {
      /* :1048-1088: the label is rebuilt on the host from this offset */
      if (cfg->report_level > 1) {
        if (S->n_unassigned == S->cap_unassigned) {
          S->cap_unassigned = S->cap_unassigned ? S->cap_unassigned * 2 : 1024;
          S->unassigned = (uint32_t*)realloc(S->unassigned, S->cap_unassigned * sizeof(uint32_t));
        }
        S->unassigned[S->n_unassigned++] =
            (uint32_t)(plus_start - 1 - BL) | (sample_id == amb ? 0x80000000u : 0u);
      }
    }

    if (cfg->validate) { /* :1154-1186 */
      int v = MGK_VAL_OK;
      if (paired) {
        v = check_read_content(r1 + header_start_pr, read_end_pr - header_start_pr,
                               seq_start_pr - header_start_pr, plus_start_pr - header_start_pr,
                               qual_start_pr - header_start_pr);
        if (!v && cfg->mgi_data) { /* :1163-1178 hamming_distance(..).unwrap() > 0 */
          size_t l1 = seq_start_pr - 2 - header_start_pr, l2 = seq_start - 2 - header_start;
          if (l1 != l2 || memcmp(r1 + header_start_pr, r2 + header_start, l1) != 0)
            v = MGK_VAL_HEADER_MISMATCH;
        }
      }
      if (!v)
        v = check_read_content(r2 + header_start, read_end - header_start, seq_start - header_start,
                               plus_start - header_start, qual_start - header_start);
      if (v) {
        S->res.validate_error = v;
        S->res.validate_record = i;
        break;
      }
    }

    if (cfg->report_level > 0) { /* :1188-1210 */
      uint64_t t, h;
      uint64_t* st = c->stats + (size_t)sample_id * MGK_N_STATS;
      if (paired) {
        sum_qc(r1 + qual_start_pr, read_end_pr - qual_start_pr, &t, &h);
        st[0] += h;
        st[6] += t;
      }
      sum_qc(r2 + read_end - BL, BL, &t, &h);
      st[2] += h;
      st[8] += t;
      sum_qc(r2 + qual_start, read_end - BL - qual_start, &t, &h);
      st[shift] += h;
      st[6 + shift] += t;
    }
    c->mism[(size_t)sample_id * (size_t)c->mism_w + (size_t)mo.curr_mismatch + 1] += 1; /* :1212 */

    const int w = c->writing[sample_id]; /* :1214 */
    if (c->reformat && !illumina) {      /* :1224: if demultiplex || illumina_format */
    } else if (sample_id >= und) {       /* :1225-1235 */
      bb_add(&S->bin_r2[sample_id], r2 + header_start, read_end + 1 - header_start);
      if (paired) bb_add(&S->bin_r1[sample_id], r1 + header_start_pr, read_end_pr + 1 - header_start_pr);
    } else {
      bytebuf* B2 = cfg->read2_has_sequence ? &S->bin_r2[w] : NULL; /* src/sample_data.rs:35-44, :683 */
      bytebuf* B1 = paired ? &S->bin_r1[w] : NULL;
      if (illumina) { /* :1240-1293 */
        size_t e = c->prefix_len;
        memcpy(hdr, c->prefix, e); /* src/sample_data.rs:256 */
        e += illumina_fields(hdr + e, r2 + header_start, hl - raw_shift, L, sep_position, umi, umi_len, c->reformat);
        if (!c->reformat) {
          memcpy(hdr + e, barcode, barcode_len); /* :1255 */
          e += barcode_len;
        } else if (sep_position == hl - 3 && c->rf_barcode_len) { /* :1256-1264 */
          hdr[e++] = ' ';
          hdr[e++] = r2[seq_start - 2];
          memcpy(hdr + e, ":N:0:", 5);
          e += 5;
          memcpy(hdr + e, c->rf_barcode, c->rf_barcode_len);
          e += c->rf_barcode_len;
        }
        if (B2) bb_add(B2, hdr, e);
        if (B1) { /* copy_header_2_paired + fix_paired_read_header, or direct write (:1265-1292) */
          if (!c->reformat) {
            hdr[e - tail_offset] = r1[seq_start_pr - 2];
          } else if (c->rf_barcode_len || sep_position < hl - 3) { /* :1269-1277 */
            hdr[e - tail_offset] = r1[seq_start_pr - 2 - header_shift];
          }
          bb_add(B1, hdr, e);
        }
        header_start = seq_start - 1; /* :1293 */
        if (paired) header_start_pr = seq_start_pr - 1;
      } else if (full_hdr) { /* :1294-1319 */
        size_t e = 0;
        hdr[e++] = ' ';
        if (umi_len > 0) {
          memcpy(hdr + e, umi + 1, umi_len - 1);
          e += umi_len - 1;
          hdr[e++] = ' ';
        }
        hdr[e++] = r2[seq_start - 2];
        memcpy(hdr + e, ":N:0:", 5);
        e += 5;
        memcpy(hdr + e, barcode, barcode_len);
        e += barcode_len;
        if (B2) {
          bb_add(B2, r2 + header_start, seq_start - 1 - header_start);
          bb_add(B2, hdr, e);
        }
        header_start = seq_start - 1;
        if (paired) {
          bb_add(B1, r1 + header_start_pr, seq_start_pr - 1 - header_start_pr);
          hdr[umi_len + 1] = r1[seq_start_pr - 2];
          bb_add(B1, hdr, e);
          header_start_pr = seq_start_pr - 1;
        }
      }
      if (B2) { /* :1321-1327 */
        bb_add(B2, r2 + header_start, plus_start - wbl - 1 - header_start);
        bb_add(B2, r2 + plus_start - 1, read_end - wbl - (plus_start - 1));
        bb_add(B2, "\n", 1);
      }
      if (B1) bb_add(B1, r1 + header_start_pr, read_end_pr + 1 - header_start_pr); /* :1328-1331 */
    }
    (void)rec_start;
    (void)rec_start_pr;
    if (paired) header_start_pr = read_end_pr + 1; /* :1340-1342 */
    header_start = read_end + 1;                   /* :1343 */
    read_cntr += 1;
  }
  free(nl2);
  free(nl1);
  free(hdr);
  if (rc != MGK_OK) return rc;

  /* lay the bins out back to back, like the GPU result */
  S->out_r2.n = S->out_r1.n = 0;
  for (int b = 0; b < total; ++b) {
    S->off2[b] = S->out_r2.n;
    S->len2[b] = S->bin_r2[b].n;
    if (S->bin_r2[b].n) bb_add(&S->out_r2, S->bin_r2[b].p, S->bin_r2[b].n);
    S->off1[b] = S->out_r1.n;
    S->len1[b] = S->bin_r1[b].n;
    if (S->bin_r1[b].n) bb_add(&S->out_r1, S->bin_r1[b].p, S->bin_r1[b].n);
  }
  S->res.seq = seq;
  S->res.n_records = read_cntr;
  S->res.consumed_r2 = header_start;
  S->res.consumed_r1 = header_start_pr;
  S->res.out_r2 = S->out_r2.p;
  S->res.out_r1 = paired ? S->out_r1.p : NULL;
  S->res.out_r2_bytes = S->out_r2.n;
  S->res.out_r1_bytes = S->out_r1.n;
  S->res.seg_off_r2 = S->off2;
  S->res.seg_len_r2 = S->len2;
  S->res.seg_off_r1 = S->off1;
  S->res.seg_len_r1 = S->len1;
  S->res.unassigned = S->unassigned;
  S->res.n_unassigned = S->n_unassigned;
  S->seq = seq;
  S->order = c->submit_clock++;
  S->state = 1;
  return MGK_OK;
}

int orc_collect(orc_ctx* c, mgk_result* res) {
  struct oslot* S = NULL;
  for (int k = 0; k < c->n_slots; ++k)
    if (c->slots[k].state == 1 && (!S || c->slots[k].order < S->order)) S = &c->slots[k];
  if (!S) {
    snprintf(c->err, sizeof c->err, "orc_collect: nothing in flight");
    return MGK_ERR_STATE;
  }
  S->state = 2;
  *res = S->res;
  return MGK_OK;
}

int orc_release(orc_ctx* c, uint64_t seq) {
  for (int k = 0; k < c->n_slots; ++k)
    if (c->slots[k].state == 2 && c->slots[k].seq == seq) {
      c->slots[k].state = 0;
      return MGK_OK;
    }
  snprintf(c->err, sizeof c->err, "orc_release: unknown chunk");
  return MGK_ERR_STATE;
}

int orc_counters(orc_ctx* c, uint64_t* stats, uint64_t* mism) {
  if (stats) memcpy(stats, c->stats, sizeof(uint64_t) * (size_t)c->total * MGK_N_STATS);
  if (mism) memcpy(mism, c->mism, sizeof(uint64_t) * (size_t)c->total * (size_t)c->mism_w);
  return MGK_OK;
}

int orc_reset_counters(orc_ctx* c) {
  memset(c->stats, 0, sizeof(uint64_t) * (size_t)c->total * MGK_N_STATS);
  memset(c->mism, 0, sizeof(uint64_t) * (size_t)c->total * (size_t)c->mism_w);
  return MGK_OK;
}

/* ReportManager::update(), src/report_manager.rs:377-416, over n worker states.  Out of place like
 * mgk_allreduce_counters: every context keeps its own share, the sum is read with orc_reduced_counters. */
int orc_allreduce_counters(orc_ctx* const* ctxs, int n) {
  if (n < 1) return MGK_ERR_ARG;
  const orc_ctx* c0 = ctxs[0];
  const size_t ns = (size_t)c0->total * MGK_N_STATS, nm = (size_t)c0->total * (size_t)c0->mism_w;
  memset(ctxs[0]->red_stats, 0, ns * sizeof(uint64_t));
  memset(ctxs[0]->red_mism, 0, nm * sizeof(uint64_t));
  for (int r = 0; r < n; ++r) {
    for (size_t i = 0; i < ns; ++i) ctxs[0]->red_stats[i] += ctxs[r]->stats[i];
    for (size_t i = 0; i < nm; ++i) ctxs[0]->red_mism[i] += ctxs[r]->mism[i];
  }
  for (int r = 1; r < n; ++r) {
    memcpy(ctxs[r]->red_stats, ctxs[0]->red_stats, ns * sizeof(uint64_t));
    memcpy(ctxs[r]->red_mism, ctxs[0]->red_mism, nm * sizeof(uint64_t));
  }
  for (int r = 0; r < n; ++r) ctxs[r]->reduced = 1;
  return MGK_OK;
}

int orc_reduced_counters(orc_ctx* c, uint64_t* stats, uint64_t* mism) {
  if (!c->reduced) {
    snprintf(c->err, sizeof c->err, "orc_reduced_counters: no counter all-reduce has run on this context");
    return MGK_ERR_STATE;
  }
  if (stats) memcpy(stats, c->red_stats, sizeof(uint64_t) * (size_t)c->total * MGK_N_STATS);
  if (mism) memcpy(mism, c->red_mism, sizeof(uint64_t) * (size_t)c->total * (size_t)c->mism_w);
  return MGK_OK;
}

void* orc_host_alloc(size_t bytes) { return malloc(bytes); }
void orc_host_free(void* p) { free(p); }

// Per-record logic of the demultiplex hot path, written lane-independently:
// every function takes the lane id as an argument and uses no warp intrinsic,
// so the very same code runs inside the sm_100a kernels (kernels.cuh) and,
// compiled as plain C++ with lanes looped, inside the CPU emulation that the
// `-m "not gpu"` tests check against the oracle (tests/emu/).
//
// What it restates (B200-first, not translated):
//   find_matching_sample      /root/reference/src/lib.rs:223-443   -> match_barcode()
//   write_illumina_header     src/sample_data.rs:569-654           -> header segments
//   trim + add_*_reads        src/lib.rs:1224-1337                 -> record_plan() / emit_*()
//   sum_qc                    src/lib.rs:458-474                   -> qc_partial()
#pragma once

#include <stdint.h>

#include "../../../include/mgikit_gpu.h"

#if defined(__CUDACC__)
#define MGK_HD __host__ __device__ __forceinline__
#define MGK_HD_NOINLINE __host__ __device__ __noinline__
#define MGK_LOOP_ROLLED _Pragma("unroll 1")
#else
#define MGK_HD inline
#define MGK_HD_NOINLINE inline
#define MGK_LOOP_ROLLED
#endif

namespace mgk {

#if defined(__CUDA_ARCH__)
MGK_HD int popc32(uint32_t x) { return __popc(x); }
MGK_HD int popc64(uint64_t x) { return __popcll(x); }
MGK_HD uint32_t ldg_u8(const uint8_t* p) { return __ldg(p); }
MGK_HD uint4 ldg_v4(const uint4* p) { return __ldg(p); }
#else
#if !defined(__CUDACC__)
struct uint4 {
  uint32_t x, y, z, w;
};
struct uint2 {
  uint32_t x, y;
};
#endif
MGK_HD int popc32(uint32_t x) { return __builtin_popcount(x); }
MGK_HD int popc64(uint64_t x) { return __builtin_popcountll(x); }
MGK_HD uint32_t ldg_u8(const uint8_t* p) { return *p; }
MGK_HD uint4 ldg_v4(const uint4* p) { return *p; }
#endif

constexpr int INF_MISMATCH = 0x3fffffff;  // usize::MAX of the reference's latest_mismatch
constexpr int PLACEHOLDER_1000 = 1000;    // src/sample_manager.rs:624

// ---------------------------------------------------------------- run tables
struct TemplateDev {
  int32_t g[10];    // parse_template geometry
  int32_t has_i5;
  int32_t n_i7, n_i5;
  int32_t off_i7, off_i5;  // first packed index in Params::codes
  int32_t off_pair;        // first entry in Params::pairs
  int32_t wide;            // any index longer than 16 bases -> 64-bit compare
  // exact-match hash tables over the raw index bytes (only when !wide): first slot in Params::hkeys / hids, slots - 1
  int32_t h7_off, h7_mask, h5_off, h5_mask;
};

struct Params {
  int32_t paired, read2_has_sequence, out_mode, keep_barcode, comprehensive, validate, report_level, mgi_data;
  int32_t m, M, L, S, total, BL, mism_w, n_tpl, prefix_len, n_codes;
  TemplateDev tpl[MGK_MAX_TEMPLATES];
  // xl7_off[d*16+t] >= 0: table translating index ids of template d into template t
  // (only consulted when the reference's dictionary iterator lags, src/lib.rs:375-376)
  int32_t xl7_off[MGK_MAX_TEMPLATES * MGK_MAX_TEMPLATES];
  int32_t xl5_off[MGK_MAX_TEMPLATES * MGK_MAX_TEMPLATES];
  const uint64_t* codes;    // 2-bit packed indexes of all templates
  const int32_t* pairs;     // pair -> sample tables of all templates
  const int32_t* xlat;      // translation tables
  const int32_t* writing;   // [total] writing_samples
  const uint8_t* consts;    // illumina prefix followed by the literals below
  const uint64_t* hkeys;    // exact-match tables: two words per slot = the index's bytes, zero padded to 16
  const int32_t* hids;      // index id of the slot, -1 = empty
};

// slot of a 16-byte key in a table of (mask + 1) slots; linear probing from there
MGK_HD uint32_t index_hash(uint64_t lo, uint64_t hi, uint32_t mask) {
  uint64_t x = lo ^ (hi * 0x9E3779B97F4A7C15ull);
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 32;
  return (uint32_t)x & mask;
}
// layout of Params::consts after the prefix
constexpr int LIT_COLON = 0;   // ":"
constexpr int LIT_SPACE = 1;   // " "
constexpr int LIT_TAIL = 2;    // ":N:0:"
constexpr int LIT_PLUS = 7;    // "+"
constexpr int LIT_NL = 8;      // "\n"
constexpr int LIT_BYTES = 16;

// ---------------------------------------------------------------- 2-bit packing
// A,C,G,T -> 0,1,3,2 ((b>>1)&3).  N packs as 0 and sets its position in nmask
// (one mismatch against anything, src/sample_manager.rs:740); any other byte
// makes the index unmatchable (it is in no dictionary key).
struct Packed {
  uint64_t code, nmask;
  int valid;
};

MGK_HD uint64_t pack_index_host(const char* s, int len) {
  uint64_t c = 0;
  for (int k = 0; k < len; ++k) c |= (uint64_t)(((uint8_t)s[k] >> 1) & 3u) << (2 * k);
  return c;
}

MGK_HD Packed pack_observed(const uint8_t* p, int len) {
  Packed r;
  r.code = 0;
  r.nmask = 0;
  r.valid = 1;
  for (int k = 0; k < len; ++k) {
    const uint32_t b = ldg_u8(p + k);
    const bool acgt = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T');
    const bool isn = b == 'N';
    if (!(acgt | isn)) r.valid = 0;
    r.code |= (uint64_t)(acgt ? ((b >> 1) & 3u) : 0u) << (2 * k);
    r.nmask |= (uint64_t)(isn ? 1u : 0u) << (2 * k);
  }
  return r;
}

struct Nearest {
  int d;      // minimum distance over the set (INF if none / invalid)
  int cnt;    // indexes at that distance
  int first;  // lowest index id at that distance
};

template <bool WIDE>
MGK_HD Nearest nearest_index(const uint64_t* codes, int n, const Packed& x) {
  Nearest r;
  r.d = INF_MISMATCH;
  r.cnt = 0;
  r.first = -1;
  if (!x.valid) return r;
  if (WIDE) {
    const uint64_t M5 = 0x5555555555555555ull;
    for (int k = 0; k < n; ++k) {
      const uint64_t v = x.code ^ codes[k];
      const int d = popc64(((v | (v >> 1)) & M5) | x.nmask);
      if (d < r.d) {
        r.d = d;
        r.cnt = 1;
        r.first = k;
      } else if (d == r.d) {
        r.cnt += 1;
      }
    }
  } else {
    const uint32_t c = (uint32_t)x.code, nm = (uint32_t)x.nmask;
    for (int k = 0; k < n; ++k) {
      const uint32_t v = c ^ (uint32_t)codes[k];
      const int d = popc32(((v | (v >> 1)) & 0x55555555u) | nm);
      if (d < r.d) {
        r.d = d;
        r.cnt = 1;
        r.first = k;
      } else if (d == r.d) {
        r.cnt += 1;
      }
    }
  }
  return r;
}

MGK_HD int dist_to(const uint64_t code, const Packed& x) {
  const uint64_t v = x.code ^ code;
  return popc64(((v | (v >> 1)) & 0x5555555555555555ull) | x.nmask);
}

// ---------------------------------------------------------------- matching
struct MatchOut {
  int sample;   // 0..total-1 (already clamped, src/lib.rs:1020-1022)
  int mism;     // curr_mismatch as returned
  int bar_t;    // template whose geometry cut curr_barcode (-1: empty)
  int umi_t;    // template whose geometry cut curr_umi (-1: empty)
  int panicked; // the reference would have panicked (stale dictionary lookup)
};

// Where the nearest-set of an observed index comes from.  The default computes it on the
// spot (XOR + popcount against every index of the dictionary); k_match_plan substitutes
// results it obtained with an exact-match hash probe or a warp-wide search.
//   di    dictionary (= template whose index set is searched; lags behind t on the reference's
//         `continue` path, src/lib.rs:375-376)
//   t     template whose geometry cuts the observed bytes
//   which 7 or 5
struct BruteNearest {
  const Params& P;
  const uint64_t* codes;
  MGK_HD Nearest get(int di, int t, int which, const uint8_t* bar) const {
    const TemplateDev& T = P.tpl[t];
    const TemplateDev& D = P.tpl[di];
    Nearest n;
    n.d = INF_MISMATCH;
    n.cnt = 0;
    n.first = -1;
    // dictionary `di` holds keys of D's index length; a lookup of another length misses
    if (which == 7) {
      if (T.g[1] == D.g[1] && D.n_i7 > 0) {
        const Packed x = pack_observed(bar + P.BL - T.g[2], T.g[1]);
        n = D.wide ? nearest_index<true>(codes + D.off_i7, D.n_i7, x) : nearest_index<false>(codes + D.off_i7, D.n_i7, x);
      }
    } else {
      if (T.g[4] == D.g[4] && D.n_i5 > 0) {
        const Packed x = pack_observed(bar + P.BL - T.g[5], T.g[4]);
        n = D.wide ? nearest_index<true>(codes + D.off_i5, D.n_i5, x) : nearest_index<false>(codes + D.off_i5, D.n_i5, x);
      }
    }
    return n;
  }
};

// find_matching_sample (src/lib.rs:223-443).  `codes` may point to a shared-memory copy of
// P.codes.  `bar` = BL observed bytes.
template <class NF>
MGK_HD MatchOut match_barcode_nf(const Params& P, const uint64_t* codes, const uint8_t* bar, const NF& nf) {
  const int total = P.total, und = total - 2, amb = total - 1;
  MatchOut r;
  r.sample = und;
  r.bar_t = r.umi_t = -1;
  r.panicked = 0;
  int latest = INF_MISMATCH, curr = INF_MISMATCH;
  int di = 0;  // dictionary iterator `template_itr` (src/lib.rs:242,437)
  for (int t = 0; t < P.n_tpl; ++t) {
    const TemplateDev& T = P.tpl[t];
    const TemplateDev& D = P.tpl[di];
    bool skip_tail = false;
    const Nearest n7 = nf.get(di, t, 7, bar);
    if (n7.d <= P.m) {  // Some(i7_matches) && .1 <= allowed  (:262-268)
      if (T.has_i5) {
        const Nearest n5 = nf.get(di, t, 5, bar);
        if (n5.d <= P.m) {  // Some(i5_matches)  (:274); dictionary entries never exceed m
          const int tot = n7.d + n5.d;
          if (!(latest < tot)) {  // :278-282
            curr = tot;           // :283
            if (curr <= P.M) {    // :284
              // candidate pairs that are samples of THIS template (:286-291)
              int npairs = 0, first_sample = -1;
              const int32_t* x7t = (di != t) ? P.xlat + P.xl7_off[di * MGK_MAX_TEMPLATES + t] : nullptr;
              const int32_t* x5t = (di != t) ? P.xlat + P.xl5_off[di * MGK_MAX_TEMPLATES + t] : nullptr;
              if (n7.cnt == 1 && n5.cnt == 1) {
                int ia = n7.first, ib = n5.first;
                if (di != t) {
                  ia = x7t[ia];
                  if (ia < 0) {
                    r.panicked = 1;
                    return r;
                  }
                  ib = x5t[ib];
                }
                const int s = ib < 0 ? -1 : P.pairs[T.off_pair + ia * T.n_i5 + ib];
                if (s >= 0) {
                  npairs = 1;
                  first_sample = s;
                }
              } else {   // several indexes at the minimum distance: rare, enumerate them
                const Packed x7 = pack_observed(bar + P.BL - T.g[2], T.g[1]);
                const Packed x5 = pack_observed(bar + P.BL - T.g[5], T.g[4]);
                for (int a = 0; a < D.n_i7 && npairs < 2; ++a) {
                  if (dist_to(codes[D.off_i7 + a], x7) != n7.d) continue;
                  int ia = a;
                  if (di != t) {
                    ia = x7t[a];
                    if (ia < 0) {
                      r.panicked = 1;
                      return r;
                    }
                  }
                  for (int b = 0; b < D.n_i5 && npairs < 2; ++b) {
                    if (dist_to(codes[D.off_i5 + b], x5) != n5.d) continue;
                    const int ib = (di != t) ? x5t[b] : b;
                    const int s = ib < 0 ? -1 : P.pairs[T.off_pair + ia * T.n_i5 + ib];
                    if (s >= 0) {
                      if (npairs == 0) first_sample = s;
                      ++npairs;
                    }
                  }
                }
              }
              if (npairs >= 1) {
                if (r.sample >= total || latest > curr) {  // :293-295
                  r.sample = first_sample;
                  latest = curr;
                  r.bar_t = t;
                  if (T.g[6] == 1) r.umi_t = t;
                  if (npairs >= 2) r.sample = amb;  // second hit takes the else arm (:360)
                } else {
                  r.sample = amb;
                }
              }
            }
          }
        }
      } else {
        if (latest < n7.d) {
          skip_tail = true;  // `continue` at :375-376: no break check, no template_itr += 1
        } else if (latest > n7.d && n7.cnt < 2) {  // :377
          curr = n7.d;
          latest = curr;
          int ia = n7.first;
          if (di != t) ia = P.xlat[P.xl7_off[di * MGK_MAX_TEMPLATES + t] + ia];
          if (ia < 0) {
            r.panicked = 1;
            return r;
          }
          r.sample = P.pairs[T.off_pair + ia];
          r.bar_t = t;
          if (T.g[6] == 1) r.umi_t = t;
        } else {
          r.sample = amb;  // :425
        }
      }
    }
    if (skip_tail) continue;
    if (r.sample != und && r.sample < total && !P.comprehensive) break;  // :433-435
    di += 1;
  }
  if (r.sample >= und) curr = 0;           // :439-441
  if (r.sample >= total) r.sample = und;   // src/lib.rs:1020-1022
  r.mism = curr;
  return r;
}

MGK_HD MatchOut match_barcode(const Params& P, const uint64_t* codes, const uint8_t* bar) {
  const BruteNearest nf{P, codes};
  return match_barcode_nf(P, codes, bar, nf);
}

// ---------------------------------------------------------------- record plan
// Byte positions of one read pair inside the chunk buffers (src/lib.rs:925-989).
struct RecGeom {
  uint32_t h2, s2, p2, q2, e2;  // header_start, seq_start, plus_start, qual_start, read_end
  uint32_t h1, s1, p1, q1, e1;  // same for the paired read
};

// Everything emit needs that is cheap to recompute from the record itself.
struct RecPlan {
  int matched;          // sample < S
  int fmt_error;        // record violates the layout the reference slices (would panic)
  uint32_t len2, len1;  // output bytes of the barcode read / paired read
  // illumina header pieces (src/sample_data.rs:584-626)
  uint32_t z;           // first non-'0' of the read-number field, == sep if none
  uint32_t f1, f2;      // offsets of the first kept digit of the two 3-digit fields
};

MGK_HD uint32_t barcode_text_len(const Params& P, int bar_t) {
  if (bar_t < 0) return 0;
  const TemplateDev& T = P.tpl[bar_t];
  return (uint32_t)T.g[1] + (T.has_i5 ? 1u + (uint32_t)T.g[4] : 0u);
}
MGK_HD uint32_t umi_text_len(const Params& P, int umi_t) { return umi_t < 0 ? 0u : 1u + (uint32_t)P.tpl[umi_t].g[7]; }

MGK_HD RecPlan record_plan(const Params& P, const uint8_t* r2, const RecGeom& G, int sample, int bar_t, int umi_t) {
  RecPlan pl;
  pl.z = pl.f1 = pl.f2 = 0;
  const uint32_t BL = (uint32_t)P.BL, L = (uint32_t)P.L;
  const uint32_t hl = G.s2 - G.h2;  // header incl. '\n'
  pl.fmt_error = (G.p2 - G.s2 < BL + 1u) || (G.e2 - G.q2 < BL) || hl < 3u ||
                 (P.out_mode == MGK_OUT_ILLUMINA && hl < L + 10u);
  pl.matched = sample < P.S;
  const uint32_t rec2 = G.e2 + 1u - G.h2, rec1 = P.paired ? G.e1 + 1u - G.h1 : 0u;
  if (pl.fmt_error) {
    pl.len2 = pl.len1 = 0;
    return pl;
  }
  if (!pl.matched) {  // src/lib.rs:1225-1235: verbatim
    pl.len2 = rec2;
    pl.len1 = rec1;
    return pl;
  }
  const uint32_t wbl = P.keep_barcode ? 0u : BL;
  const uint32_t btl = barcode_text_len(P, bar_t), utl = umi_text_len(P, umi_t);
  uint32_t hdr2, hdr1;
  if (P.out_mode == MGK_OUT_ILLUMINA) {
    const uint8_t* H = r2 + G.h2;
    const uint32_t sep = hl - 3u;
    uint32_t z = sep;
    for (uint32_t i = L + 10u; i < sep; ++i)
      if (ldg_u8(H + i) != '0') {
        z = i;
        break;
      }
    pl.z = z;
    pl.f1 = ldg_u8(H + L + 3) != '0' ? L + 3u : (ldg_u8(H + L + 4) != '0' ? L + 4u : L + 5u);
    pl.f2 = ldg_u8(H + L + 7) != '0' ? L + 7u : (ldg_u8(H + L + 8) != '0' ? L + 8u : L + 9u);
    hdr2 = (uint32_t)P.prefix_len + 2u + (sep - z) + 1u + (L + 6u - pl.f1) + 1u + (L + 10u - pl.f2) + utl + 7u + btl;
    hdr1 = hdr2;
  } else if (P.out_mode == MGK_OUT_MGI_FULL_HEADER) {
    const uint32_t tail = 1u + (utl ? utl : 0u) + 6u + btl;  // " " [umi " "] n ":N:0:" barcode
    hdr2 = (hl - 1u) + tail;
    hdr1 = P.paired ? (G.s1 - G.h1 - 1u) + tail : 0u;
  } else {
    hdr2 = hl - 1u;
    hdr1 = P.paired ? G.s1 - G.h1 - 1u : 0u;
  }
  // body: "\n" seq[..-wbl] "\n+\n" qual[..-wbl] "\n"   (src/lib.rs:1321-1327)
  const uint32_t body2 = (G.p2 - wbl - 1u - (G.s2 - 1u)) + (G.e2 - wbl - (G.p2 - 1u)) + 1u;
  pl.len2 = P.read2_has_sequence ? hdr2 + body2 : 0u;
  pl.len1 = P.paired ? hdr1 + (G.e1 + 1u - (G.s1 - 1u)) : 0u;
  return pl;
}

// ---------------------------------------------------------------- emit
// A record's rewritten header is a short list of (source pointer, length)
// segments in FIXED slots (absent pieces have length 0), so the list lives in
// registers: no dynamically indexed local array.
constexpr int MAX_SEG = 16;
struct SegList {
  const uint8_t* p[MAX_SEG];
  uint32_t n[MAX_SEG];
};

// Header segments of a matched record; which = 2 (barcode read) or 1 (paired read).
MGK_HD void header_segments(const Params& P, const uint8_t* r2, const uint8_t* r1, const RecGeom& G, const RecPlan& pl,
                            int bar_t, int umi_t, int which, SegList& s) {
  const uint8_t* lit = P.consts + P.prefix_len;
  const uint8_t* H = r2 + G.h2;
  const uint32_t hl = G.s2 - G.h2, L = (uint32_t)P.L;
  const uint8_t* bar = r2 + G.p2 - 1u - (uint32_t)P.BL;
  const uint8_t* nsrc = which == 2 ? r2 + G.s2 - 2u : r1 + G.s1 - 2u;  // src/lib.rs:1275,1288,1302,1314
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < MAX_SEG; ++j) {
    s.p[j] = lit;
    s.n[j] = 0;
  }
  const uint8_t* umi_p = lit;
  uint32_t umi_n = 0;
  if (umi_t >= 0) {
    umi_p = bar + P.BL - P.tpl[umi_t].g[8];
    umi_n = (uint32_t)P.tpl[umi_t].g[7];
  }
  if (P.out_mode == MGK_OUT_ILLUMINA) {
    const uint32_t sep = hl - 3u;
    s.p[0] = P.consts;          s.n[0] = (uint32_t)P.prefix_len;
    s.p[1] = H + L + 1;         s.n[1] = 1;
    s.p[2] = lit + LIT_COLON;   s.n[2] = 1;
    s.p[3] = H + pl.z;          s.n[3] = sep - pl.z;
    s.p[4] = lit + LIT_COLON;   s.n[4] = 1;
    s.p[5] = H + pl.f1;         s.n[5] = L + 6u - pl.f1;
    s.p[6] = lit + LIT_COLON;   s.n[6] = 1;
    s.p[7] = H + pl.f2;         s.n[7] = L + 10u - pl.f2;
    s.p[8] = lit + LIT_COLON;   s.n[8] = umi_t >= 0 ? 1u : 0u;
    s.p[9] = umi_p;             s.n[9] = umi_n;
    s.p[10] = lit + LIT_SPACE;  s.n[10] = 1;
  } else {
    if (which == 2) { s.p[0] = H;         s.n[0] = hl - 1u; }
    else            { s.p[0] = r1 + G.h1; s.n[0] = G.s1 - G.h1 - 1u; }
    if (P.out_mode != MGK_OUT_MGI_FULL_HEADER) return;
    s.p[1] = lit + LIT_SPACE;   s.n[1] = 1;
    s.p[2] = umi_p;             s.n[2] = umi_n;
    s.p[3] = lit + LIT_SPACE;   s.n[3] = umi_t >= 0 ? 1u : 0u;
  }
  s.p[11] = nsrc;               s.n[11] = 1;
  s.p[12] = lit + LIT_TAIL;     s.n[12] = 5;
  if (bar_t >= 0) {
    const TemplateDev& T = P.tpl[bar_t];
    s.p[13] = bar + P.BL - T.g[2];  s.n[13] = (uint32_t)T.g[1];
    s.p[14] = lit + LIT_PLUS;       s.n[14] = T.has_i5 ? 1u : 0u;
    s.p[15] = bar + P.BL - T.g[5];  s.n[15] = T.has_i5 ? (uint32_t)T.g[4] : 0u;
  }
}

// Lane `lane` of `nlanes` writes its share of the short segments, one byte per
// lane per step (headers are ~60 B: two steps).
MGK_HD uint32_t emit_segments(uint8_t* dst, const SegList& s, int lane, int nlanes) {
  uint32_t total = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < MAX_SEG; ++j) total += s.n[j];
  for (uint32_t k = (uint32_t)lane; k < total; k += (uint32_t)nlanes) {
    uint32_t start = 0;
    const uint8_t* src = nullptr;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < MAX_SEG; ++j) {
      if (src == nullptr && k < start + s.n[j]) src = s.p[j] + (k - start);
      start += s.n[j];
    }
    dst[k] = (uint8_t)ldg_u8(src);
  }
  return total;
}

// Long verbatim span: every lane moves 16-byte units that are ALIGNED ON THE
// DESTINATION (so the dominant traffic is full-width coalesced stores); the
// source is read as the two aligned 16-byte vectors that cover the unit and
// funnel-shifted.  Head/tail bytes that do not fill a destination unit are
// written bytewise.  `src` must stay readable up to the next 16-byte boundary
// after src+n (chunk buffers are padded).
MGK_HD void copy_span(uint8_t* dst, const uint8_t* src, uint32_t n, int lane, int nlanes) {
  const uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
  const uint32_t h = head < n ? head : n;
  for (uint32_t k = (uint32_t)lane; k < h; k += (uint32_t)nlanes) dst[k] = (uint8_t)ldg_u8(src + k);
  if (n <= h) return;
  const uint32_t units = (n - h) >> 4;
  uint8_t* d = dst + h;
  const uint8_t* s = src + h;
  const uint32_t sh = (uint32_t)((uintptr_t)s & 15u);
  const uint4* sv = (const uint4*)(s - sh);
  if (sh == 0) {
    for (uint32_t u = (uint32_t)lane; u < units; u += (uint32_t)nlanes) ((uint4*)d)[u] = ldg_v4(sv + u);
  } else {
    const uint32_t wsel = sh >> 2, bsh = (sh & 3u) * 8u;
    for (uint32_t u = (uint32_t)lane; u < units; u += (uint32_t)nlanes) {
      const uint4 a = ldg_v4(sv + u), b = ldg_v4(sv + u + 1);
      // words a.x a.y a.z a.w b.x b.y b.z b.w ; take 5 starting at wsel
      uint32_t w0, w1, w2, w3, w4;
      if (wsel == 0) { w0 = a.x; w1 = a.y; w2 = a.z; w3 = a.w; w4 = b.x; }
      else if (wsel == 1) { w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; }
      else if (wsel == 2) { w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; }
      else { w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; }
      uint4 o;
      if (bsh == 0) {
        o.x = w0; o.y = w1; o.z = w2; o.w = w3;
      } else {
#if defined(__CUDA_ARCH__)
        o.x = __funnelshift_r(w0, w1, bsh);
        o.y = __funnelshift_r(w1, w2, bsh);
        o.z = __funnelshift_r(w2, w3, bsh);
        o.w = __funnelshift_r(w3, w4, bsh);
#else
        o.x = (w0 >> bsh) | (w1 << (32u - bsh));
        o.y = (w1 >> bsh) | (w2 << (32u - bsh));
        o.z = (w2 >> bsh) | (w3 << (32u - bsh));
        o.w = (w3 >> bsh) | (w4 << (32u - bsh));
#endif
      }
      ((uint4*)d)[u] = o;
    }
  }
  const uint32_t done = h + (units << 4);
  for (uint32_t k = done + (uint32_t)lane; k < n; k += (uint32_t)nlanes) dst[k] = (uint8_t)ldg_u8(src + k);
}

// ---------------------------------------------------------------- QC sums
// sum_qc (src/lib.rs:458-474) over src[0..n): this lane's share, 16-byte
// aligned vectors with out-of-range bytes masked.  Returns (sum, count>=63).
struct Qc {
  uint32_t sum, high;
};

MGK_HD uint32_t bytes_ge63(uint32_t x) {  // per byte: 0x80 where byte >= 63, exact for 0..255
  const uint32_t low = ((x | 0x80808080u) - 0x3f3f3f3fu);
  return (low | x) & 0x80808080u;
}
MGK_HD uint32_t byte_sum(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __dp4a(x, 0x01010101u, 0u);
#else
  return (x & 0xff) + ((x >> 8) & 0xff) + ((x >> 16) & 0xff) + (x >> 24);
#endif
}
MGK_HD uint32_t keep_bytes(uint32_t w, int lo, int hi) {  // keep bytes [lo,hi) of the word, 0<=lo<=hi<=4 clipped
  if (lo < 0) lo = 0;
  if (hi > 4) hi = 4;
  if (hi <= lo) return 0u;
  uint32_t mask = hi == 4 ? 0xffffffffu : ((1u << (8 * hi)) - 1u);
  mask &= lo == 0 ? 0xffffffffu : ~((1u << (8 * lo)) - 1u);
  return w & mask;
}

MGK_HD Qc qc_partial(const uint8_t* src, uint32_t n, int lane, int nlanes) {
  Qc q;
  q.sum = q.high = 0;
  if (n == 0) return q;
  const uint32_t sh = (uint32_t)((uintptr_t)src & 15u);
  const uint4* sv = (const uint4*)(src - sh);
  const uint32_t nvec = (sh + n + 15u) >> 4;
  for (uint32_t v = (uint32_t)lane; v < nvec; v += (uint32_t)nlanes) {
    uint4 a = ldg_v4(sv + v);
    const int lo = (int)sh - (int)(v << 4);            // first valid byte inside this vector
    const int hi = (int)(sh + n) - (int)(v << 4);      // one past the last valid byte
    if (lo > 0 || hi < 16) {
      a.x = keep_bytes(a.x, lo, hi);
      a.y = keep_bytes(a.y, lo - 4, hi - 4);
      a.z = keep_bytes(a.z, lo - 8, hi - 8);
      a.w = keep_bytes(a.w, lo - 12, hi - 12);
    }
    q.sum += byte_sum(a.x) + byte_sum(a.y) + byte_sum(a.z) + byte_sum(a.w);
    q.high += (uint32_t)(popc32(bytes_ge63(a.x)) + popc32(bytes_ge63(a.y)) + popc32(bytes_ge63(a.z)) + popc32(bytes_ge63(a.w)));
  }
  return q;
}

// ---------------------------------------------------------------- validate
// check_read_content (src/lib.rs:476-509) + header equality (:1163-1178): this
// lane's share of one pair; returns the lowest failing priority (1..7) or 0.
//   1..3 paired read ('@', base, quality)  4 header mismatch  5..7 barcode read
MGK_HD int validate_partial(const Params& P, const uint8_t* r2, const uint8_t* r1, const RecGeom& G, int lane, int nlanes) {
  int worst = 0;
  auto note = [&](int p) {
    if (worst == 0 || p < worst) worst = p;
  };
  auto base_ok = [](uint32_t n) {
    return n == 'A' || n == 'C' || n == 'G' || n == 'T' || n == 'N' || n == 'a' || n == 'c' || n == 'g' || n == 't' || n == 'n';
  };
  if (P.paired) {
    if (lane == 0 && ldg_u8(r1 + G.h1) != '@') note(1);
    for (uint32_t i = G.s1 + (uint32_t)lane; i + 1u < G.p1; i += (uint32_t)nlanes)
      if (!base_ok(ldg_u8(r1 + i))) note(2);
    for (uint32_t i = G.q1 + (uint32_t)lane; i < G.e1; i += (uint32_t)nlanes) {
      const uint32_t q = ldg_u8(r1 + i);
      if (q > 73u || q < 33u) note(3);
    }
    if (P.mgi_data) {
      const uint32_t l1 = G.s1 - 2u - G.h1, l2 = G.s2 - 2u - G.h2;
      if (l1 != l2) {
        if (lane == 0) note(4);
      } else {
        for (uint32_t i = (uint32_t)lane; i < l1; i += (uint32_t)nlanes)
          if (ldg_u8(r1 + G.h1 + i) != ldg_u8(r2 + G.h2 + i)) note(4);
      }
    }
  }
  if (lane == 0 && ldg_u8(r2 + G.h2) != '@') note(5);
  for (uint32_t i = G.s2 + (uint32_t)lane; i + 1u < G.p2; i += (uint32_t)nlanes)
    if (!base_ok(ldg_u8(r2 + i))) note(6);
  for (uint32_t i = G.q2 + (uint32_t)lane; i < G.e2; i += (uint32_t)nlanes) {
    const uint32_t q = ldg_u8(r2 + i);
    if (q > 73u || q < 33u) note(7);
  }
  return worst;
}

MGK_HD int validate_kind(int priority) {
  switch (priority) {
    case 1: case 5: return MGK_VAL_HEADER_AT;
    case 2: case 6: return MGK_VAL_BASE;
    case 3: case 7: return MGK_VAL_QUALITY;
    case 4: return MGK_VAL_HEADER_MISMATCH;
    default: return MGK_VAL_OK;
  }
}

}  // namespace mgk

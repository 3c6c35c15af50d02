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
  // `reformat` (src/lib.rs:1090-1147): one sample, no matching; BL is the UMI length; the sample's barcode (rf_bc_len
  // bytes, 0: none) follows the literals in `consts`
  int32_t reformat, rf_bc_len;
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
constexpr int RF_MAX_TAIL = 255;   // reformat: bytes of the input header from its first ' ' on that one record may carry over

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
  // reformat only: bytes of the input header that follow its fields (from the first ' ' on; none for MGI raw headers),
  // whether the header is raw ("/n", no ' '), mismatches of the header's barcode against the sample's
  uint32_t rf_tcl, rf_raw, rf_mism;
};

// reformat: what the emit side needs to know about a record's header, packed into RecMeta (tpl byte + a flag bit)
MGK_HD uint32_t rf_pack(uint32_t tcl, uint32_t raw) { return tcl | (raw << 8); }
struct RfHeader {
  uint32_t sep;           // offset of the separator inside the header (src/lib.rs:1091-1102)
  uint32_t tcl;           // header bytes carried over from sep on
  uint32_t at_end;        // sep == header length - 3 (:1104, :1132, :1256)
  uint32_t header_shift;  // :1129, :1135
  uint32_t tail_offset;   // :1130, :1136
  uint32_t append;        // " n:N:0:" + the sample's barcode is appended (:1256-1264)
};
MGK_HD RfHeader rf_header(const Params& P, uint32_t hl, uint32_t rf) {   // hl: header incl. '\n'
  RfHeader h;
  const uint32_t raw = rf >> 8;
  h.tcl = rf & 0xffu;
  h.sep = hl - (raw ? 2u : 0u) - 1u - h.tcl;
  h.at_end = h.sep == hl - 3u;
  h.header_shift = hl - h.sep - 3u;
  h.tail_offset = h.header_shift + 1u;
  h.append = (h.at_end && P.rf_bc_len > 0) ? 1u : 0u;
  if (h.append) {
    h.header_shift = 0;
    h.tail_offset = (uint32_t)P.rf_bc_len + 6u;
  }
  return h;
}

MGK_HD uint32_t barcode_text_len(const Params& P, int bar_t) {
  if (bar_t < 0) return 0;
  const TemplateDev& T = P.tpl[bar_t];
  return (uint32_t)T.g[1] + (T.has_i5 ? 1u + (uint32_t)T.g[4] : 0u);
}
MGK_HD uint32_t umi_text_len(const Params& P, int umi_t) { return umi_t < 0 ? 0u : 1u + (uint32_t)P.tpl[umi_t].g[7]; }

MGK_HD RecPlan record_plan(const Params& P, const uint8_t* r2, const RecGeom& G, int sample, int bar_t, int umi_t) {
  RecPlan pl;
  pl.z = pl.f1 = pl.f2 = 0;
  pl.rf_tcl = pl.rf_raw = pl.rf_mism = 0;
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
  if (P.reformat) {   // src/lib.rs:1090-1147, :1240-1293 with demultiplex == false
    const uint8_t* H = r2 + G.h2;
    uint32_t sep = hl;
    for (uint32_t i = 0; i < hl; ++i)   // memchr(b' ', header), :1091
      if (ldg_u8(H + i) == ' ') {
        sep = i;
        break;
      }
    pl.rf_raw = sep == hl ? 1u : 0u;   // "You might be using MGI raw data": raw_shift = 2
    if (pl.rf_raw) sep = hl - 3u;
    const uint32_t bcl = (uint32_t)P.rf_bc_len;
    const bool at_end = sep == hl - 3u;
    // a ' ' in the last two header bytes wraps header_shift (:1129); a header barcode of another length panics in
    // hamming_distance(..).unwrap() (:1108-1112)
    if (sep > hl - 3u || (bcl && !at_end && (sep + 7u > hl - 1u || hl - 1u - (sep + 7u) != bcl))) {
      pl.fmt_error = 1;
      pl.len2 = pl.len1 = 0;
      return pl;
    }
    pl.rf_tcl = hl - (pl.rf_raw ? 2u : 0u) - 1u - sep;
    if (bcl && !at_end) {
      const uint8_t* bc = P.consts + P.prefix_len + LIT_BYTES;
      uint32_t d = 0;
      for (uint32_t k = 0; k < bcl; ++k) d += ldg_u8(H + sep + 7u + k) != ldg_u8(bc + k) ? 1u : 0u;
      pl.rf_mism = d > 4u ? 4u : d;   // :1140-1141
    }
    const RfHeader rh = rf_header(P, hl, rf_pack(pl.rf_tcl, pl.rf_raw));
    // the header tail must fit its byte of RecMeta; the paired read's own digit is looked up header_shift + 2 bytes
    // in front of its sequence (:1275): a shorter header would reach into the record before it
    if (pl.rf_tcl > (uint32_t)RF_MAX_TAIL || (P.paired && P.out_mode == MGK_OUT_ILLUMINA && G.s1 - G.h1 < rh.header_shift + 2u)) {
      pl.fmt_error = 1;
      pl.len2 = pl.len1 = 0;
      return pl;
    }
    if (P.out_mode != MGK_OUT_ILLUMINA) {   // :1224: nothing is written unless demultiplex || illumina_format
      pl.len2 = pl.len1 = 0;
      return pl;
    }
    uint32_t z = sep;
    for (uint32_t i = L + 10u; i < sep; ++i)
      if (ldg_u8(H + i) != '0') {
        z = i;
        break;
      }
    pl.z = z;
    pl.f1 = ldg_u8(H + L + 3) != '0' ? L + 3u : (ldg_u8(H + L + 4) != '0' ? L + 4u : L + 5u);
    pl.f2 = ldg_u8(H + L + 7) != '0' ? L + 7u : (ldg_u8(H + L + 8) != '0' ? L + 8u : L + 9u);
    const uint32_t hdr = (uint32_t)P.prefix_len + 2u + (sep - z) + 1u + (L + 6u - pl.f1) + 1u + (L + 10u - pl.f2) +
                         (BL ? 1u + BL : 0u) + pl.rf_tcl + (rh.append ? 7u + bcl : 0u);
    const uint32_t body2 = (G.p2 - BL - 1u - (G.s2 - 1u)) + (G.e2 - BL - (G.p2 - 1u)) + 1u;
    pl.len2 = hdr + body2;
    pl.len1 = P.paired ? hdr + (G.e1 + 1u - (G.s1 - 1u)) : 0u;
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

// ---------------------------------------------------------------- QC sums
// sum_qc (src/lib.rs:458-474): (sum of the quality bytes, how many are >= 63); the loops are in emit_logic.cuh
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
// ---------------------------------------------------------------- validate
// priority (validate_pair, emit_logic.cuh) -> MGK_VAL_* kind
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

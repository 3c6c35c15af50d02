// Record emission out of a STAGED TILE: header rewrite, barcode trim, Q30 /
// quality sums and --validate for one read pair whose bytes already sit in a
// fast byte array `S` (shared memory on the B200, a host array in the CPU
// emulation that the `-m "not gpu"` tests check against the oracle).
//
// Two kinds of work, two kinds of workers (k_emit):
//   * control-heavy, tiny: where a header's digits start, which template cut the barcode,
//     the quality sums.  ONE THREAD PER (record, role): build_header() front / tail,
//     qc_pair(), validate_pair() — every lane of a warp carries a different record.
//   * bulk bytes: an output record is at most three byte ranges (header | "\n" sequence |
//     "\n+\n" qualities), see plan_read() / read_piece().  On the GPU quarter-warps move them
//     (kernels.cuh: move_pair_lane / move_small_lane, PTX): 16-byte stores aligned on the
//     destination, fully coalesced, fed by unaligned reads of S (aligned words + funnel shift),
//     the < 16 bytes at either end byte by byte.  copy_piece() below states the same byte
//     movement in plain C++ (lanes looped) for the CPU emulation, which therefore checks WHAT is
//     moved WHERE (plans, pieces, headers, patch), not the PTX movers; those are checked by the
//     GPU parity tests.
// Everything else here takes no warp intrinsic, so the CPU emulation runs the very same code.
//
// What it restates (reference = /root/reference):
//   write_illumina_header        src/sample_data.rs:569-654, call site src/lib.rs:1240-1293
//   --mgi-full-header            src/lib.rs:1294-1319
//   trim + add_*_reads           src/lib.rs:1224-1337
//   sum_qc + update_stats        src/lib.rs:458-474, :1188-1210
//   check_read_content, header equality   src/lib.rs:476-509, :1163-1178
#pragma once

#include "device_logic.cuh"

namespace mgk {

// ---------------------------------------------------------------- reading S at any alignment
MGK_HD uint32_t s_u8(const uint8_t* S, uint32_t off) { return S[off]; }
MGK_HD uint32_t s_u32(const uint8_t* S, uint32_t off) { return *reinterpret_cast<const uint32_t*>(S + off); }   // off % 4 == 0
MGK_HD uint4 s_v4(const uint8_t* S, uint32_t off) { return *reinterpret_cast<const uint4*>(S + off); }          // off % 16 == 0

MGK_HD int ctz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
MGK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // sh in {0,8,16,24}
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
// 16 bytes of S at any offset as four words: five independent aligned word loads (reads S[a & ~3 .. +20))
MGK_HD uint4 ld16u(const uint8_t* S, uint32_t a) {
  const uint32_t al = a & ~3u, sh = (a & 3u) * 8u;
  const uint32_t w0 = s_u32(S, al), w1 = s_u32(S, al + 4u), w2 = s_u32(S, al + 8u), w3 = s_u32(S, al + 12u), w4 = s_u32(S, al + 16u);
  uint4 o;
  o.x = funnel_r(w0, w1, sh);
  o.y = funnel_r(w1, w2, sh);
  o.z = funnel_r(w2, w3, sh);
  o.w = funnel_r(w3, w4, sh);
  return o;
}
MGK_HD uint32_t put_byte(uint32_t w, int i, uint32_t v) {  // byte i of the word := v when 0 <= i < 4
  if (i < 0 || i > 3) return w;
  const uint32_t sh = 8u * (uint32_t)i;
  return (w & ~(0xffu << sh)) | (v << sh);
}

// ---------------------------------------------------------------- one byte range (CPU emulation of the movers)
// Lane l16 (0..15) of sixteen writes its share of dst[0 .. n) = S[src .. src+n), with output
// byte `patch_at` (>= n: none) replaced by patch_val.  Lane u takes the 16-byte units u, u+16, ...
// of the destination-aligned middle; lane l the l-th byte in front of it and the l-th after it.
MGK_HD void copy_piece(uint8_t* dst, const uint8_t* S, uint32_t src, uint32_t n, int l16, uint32_t patch_at, uint32_t patch_val) {
  uint32_t hb = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
  if (hb > n) hb = n;
  const uint32_t nfull = (n - hb) >> 4, tail0 = hb + (nfull << 4);
  for (uint32_t u = (uint32_t)l16; u < nfull; u += 16u) {
    const uint32_t x = hb + (u << 4);
    uint4 w = ld16u(S, src + x);
    if (patch_at - x < 16u) {
      const int i = (int)(patch_at - x);
      w.x = put_byte(w.x, i, patch_val);
      w.y = put_byte(w.y, i - 4, patch_val);
      w.z = put_byte(w.z, i - 8, patch_val);
      w.w = put_byte(w.w, i - 12, patch_val);
    }
    *reinterpret_cast<uint4*>(dst + x) = w;
  }
  if ((uint32_t)l16 < hb) dst[l16] = (uint8_t)((uint32_t)l16 == patch_at ? patch_val : s_u8(S, src + (uint32_t)l16));
  const uint32_t y = tail0 + (uint32_t)l16;
  if (y < n) dst[y] = (uint8_t)(y == patch_at ? patch_val : s_u8(S, src + y));
}

// ---------------------------------------------------------------- header rewrite
// Positions are offsets into S.  Illumina: where the kept digits start
// (src/sample_data.rs:584-626): leading zeros of the read number are dropped (all
// zeros: nothing is left), the two 3-digit fields keep at least their last digit.
struct HdrInfo {
  uint32_t z, f1, f2;  // first kept byte of the read number (== sep if none), of the two 3-digit fields
  uint32_t sep;        // the '/' (header length - 3, src/lib.rs:1007)
};
MGK_HD uint32_t bytes_ne(uint32_t x, uint32_t pat) {   // 0x80 in every byte of x that differs from the pattern byte, exact
  const uint32_t v = x ^ pat;
  return (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
}
MGK_HD HdrInfo hdr_info(const Params& P, const uint8_t* S, const RecGeom& G, uint32_t sep) {   // sep: S offset of the separator
  HdrInfo h;
  const uint32_t L = (uint32_t)P.L, H = G.h2;
  h.sep = sep;
  // bytes H+L+3 .. H+L+18: ccc R rrr and the first nine digits of the read number, in one go
  const uint4 v = ld16u(S, H + L + 3u);
  const uint32_t Z = 0x30303030u, C = 0x00204081u;   // 0x80 flags of a word -> 4 bits at the top of the product
  const uint32_t nz = ((bytes_ne(v.x, Z) * C) >> 28) | (((bytes_ne(v.y, Z) * C) >> 24) & 0xf0u) | (((bytes_ne(v.z, Z) * C) >> 20) & 0xf00u) |
                      (((bytes_ne(v.w, Z) * C) >> 16) & 0xf000u);   // bit k <=> byte k is not '0'
  h.f1 = (nz & 1u) ? H + L + 3u : ((nz & 2u) ? H + L + 4u : H + L + 5u);
  h.f2 = (nz & 0x10u) ? H + L + 7u : ((nz & 0x20u) ? H + L + 8u : H + L + 9u);
  uint32_t z = H + L + 10u;
  const uint32_t rest = nz >> 7;   // the read number's first nine digits
  if (rest) {
    z += (uint32_t)ctz32(rest);
  } else {
    z += 9u;
    while (z < h.sep && s_u8(S, z) == '0') ++z;   // longer runs of zeros: rare
  }
  h.z = z < h.sep ? z : h.sep;
  return h;
}

// What follows the header fields:
//   illumina : [":" umi] " " n ":N:0:" barcode            (src/sample_data.rs:627-653)
//   mgi full : " " [umi " "] n ":N:0:" barcode            (src/lib.rs:1294-1319)
struct TailSpec {
  uint32_t umi_off, umi_len;   // S offset / length of the UMI (0 length: none)
  uint32_t i7_off, i7_len, i5_off, i5_len, has_i5, has_bar;
};
MGK_HD uint32_t barcode_len(const TailSpec& t) { return t.has_bar ? t.i7_len + (t.has_i5 ? 1u + t.i5_len : 0u) : 0u; }
MGK_HD uint32_t tail_len(const TailSpec& t) {   // both flavours: one separator more than the umi, n, ":N:0:"
  return (t.umi_len ? 1u + t.umi_len : 0u) + 7u + barcode_len(t);
}
MGK_HD TailSpec make_tail(const Params& P, const RecGeom& G, int bar_t, int umi_t) {
  TailSpec t;
  const uint32_t bar = G.p2 - 1u - (uint32_t)P.BL;
  t.umi_off = t.umi_len = 0;
  if (umi_t >= 0) {
    t.umi_off = bar + (uint32_t)(P.BL - P.tpl[umi_t].g[8]);
    t.umi_len = (uint32_t)P.tpl[umi_t].g[7];
  }
  t.has_bar = bar_t >= 0 ? 1u : 0u;
  t.i7_off = t.i5_off = bar;
  t.i7_len = t.i5_len = t.has_i5 = 0;
  if (bar_t >= 0) {
    const TemplateDev& T = P.tpl[bar_t];
    t.i7_off = bar + (uint32_t)(P.BL - T.g[2]);
    t.i7_len = (uint32_t)T.g[1];
    t.has_i5 = T.has_i5 ? 1u : 0u;
    if (T.has_i5) {
      t.i5_off = bar + (uint32_t)(P.BL - T.g[5]);
      t.i5_len = (uint32_t)T.g[4];
    }
  }
  return t;
}

// Short pieces (digits, index, UMI) go through a cursor: 16 source bytes per round, fetched with
// independent word loads, then stored byte by byte (the cursor has any alignment): no load depends on
// a store.  ONE out-of-line copy: the emit code has to stay small (instruction cache).
MGK_HD_NOINLINE uint8_t* put_span(uint8_t* o, const uint8_t* S, uint32_t src, uint32_t n) {
  MGK_LOOP_ROLLED
  for (uint32_t k = 0; k < n; k += 16u) {
    const uint4 v = ld16u(S, src + k);
    const uint32_t m = n - k;   // bytes of this round that count
    uint32_t w = v.x;
    if (m > 0) o[k + 0] = (uint8_t)w;
    if (m > 1) o[k + 1] = (uint8_t)(w >> 8);
    if (m > 2) o[k + 2] = (uint8_t)(w >> 16);
    if (m > 3) o[k + 3] = (uint8_t)(w >> 24);
    w = v.y;
    if (m > 4) o[k + 4] = (uint8_t)w;
    if (m > 5) o[k + 5] = (uint8_t)(w >> 8);
    if (m > 6) o[k + 6] = (uint8_t)(w >> 16);
    if (m > 7) o[k + 7] = (uint8_t)(w >> 24);
    w = v.z;
    if (m > 8) o[k + 8] = (uint8_t)w;
    if (m > 9) o[k + 9] = (uint8_t)(w >> 8);
    if (m > 10) o[k + 10] = (uint8_t)(w >> 16);
    if (m > 11) o[k + 11] = (uint8_t)(w >> 24);
    w = v.w;
    if (m > 12) o[k + 12] = (uint8_t)w;
    if (m > 13) o[k + 13] = (uint8_t)(w >> 8);
    if (m > 14) o[k + 14] = (uint8_t)(w >> 16);
    if (m > 15) o[k + 15] = (uint8_t)(w >> 24);
  }
  return o + n;
}

// ---------------------------------------------------------------- what a read's output is made of
// Every output record, matched or not, is header | "\n" sequence | "\n+\n" qualities (+ "\n").
// which = 2: barcode read -> new header, barcode trimmed off sequence and qualities
//            (src/lib.rs:1240-1327); unmatched reads verbatim (:1225-1235)
// which = 1: paired read  -> in illumina mode the barcode read's new header with its own
//            read-number digit (src/sample_data.rs:302-328, src/lib.rs:1288), the rest of the
//            record verbatim (:1328-1331); with --mgi-full-header ITS OWN header line (:1306-1318)
// `out_len` = the record's output length as planned by k_match_plan (RecMeta::len2 / len1): the
// header length follows from it without looking at the header bytes.
enum { HDR_VERBATIM = 0, HDR_SLOT0 = 1, HDR_SLOT1 = 2 };
struct ReadPlan {
  uint32_t len;          // 0: nothing is written for this read
  uint32_t hl, n1, n2;   // bytes of the three parts (the closing '\n' of a trimmed read not counted)
  uint32_t h_src, src1, src2;   // S offsets of the header line (verbatim case), of "\n" seq, of "\n+\n" qual
  int hdr;               // HDR_VERBATIM: from the input; HDR_SLOT0/1: rewritten, from that staging slot of the record
  int trim;              // the closing '\n' is stored on its own
  uint32_t patch_at;     // illumina paired read: index of the read-number digit inside the header (else ~0)
  uint32_t patch_src;    // ... and where its own digit is in S
};
// `rf`: reformat only, rf_pack() of the record (RecMeta::tpl and flag bit 1).
MGK_HD ReadPlan plan_read(const Params& P, const RecGeom& G, int which, int sample, int bar_t, uint32_t out_len, uint32_t rf = 0) {
  ReadPlan r;
  const bool br = which == 2;
  const uint32_t h = br ? G.h2 : G.h1, s = br ? G.s2 : G.s1, p = br ? G.p2 : G.p1, e = br ? G.e2 : G.e1;
  const bool matched = sample < P.S;
  r.len = (br && matched && !P.read2_has_sequence) ? 0u : out_len;   // samples get no barcode-read file, src/sample_data.rs:35-44
  r.trim = br && matched;
  const uint32_t wbl = (r.trim && !P.keep_barcode) ? (uint32_t)P.BL : 0u;
  r.n1 = p - wbl - s;                                   // "\n" seq[..-wbl], src/lib.rs:1293,1321-1323
  r.n2 = r.trim ? e - wbl - (p - 1u) : e + 2u - p;      // "\n+\n" qual[..-wbl]; a trimmed read's closing '\n' apart (:1324-1327)
  r.hl = out_len - r.n1 - r.n2 - (r.trim ? 1u : 0u);
  r.h_src = h;
  r.src1 = s - 1u;
  r.src2 = p - 1u;
  r.hdr = HDR_VERBATIM;
  r.patch_at = 0xffffffffu;
  r.patch_src = 0;
  if (matched && P.out_mode != MGK_OUT_MGI) {
    if (P.out_mode == MGK_OUT_MGI_FULL_HEADER) {
      r.hdr = br ? HDR_SLOT0 : HDR_SLOT1;
    } else {
      r.hdr = HDR_SLOT0;
      if (!br && !P.reformat) {   // fix_paired_read_header: n is tail_offset = len(barcode) + 6 from the end, src/lib.rs:1019
        r.patch_at = r.hl - 6u - barcode_text_len(P, bar_t);
        r.patch_src = G.s1 - 2u;
      } else if (!br) {           // reformat: only when the header ends with a read number at all, src/lib.rs:1269-1277
        const RfHeader rh = rf_header(P, G.s2 - G.h2, rf);
        if (P.rf_bc_len > 0 || !rh.at_end) {
          r.patch_at = r.hl - rh.tail_offset;
          r.patch_src = G.s1 - 2u - rh.header_shift;
        }
      }
    }
  }
  return r;
}

// Rewritten header of one read into its staging slot `o` (16-byte aligned), in two independent halves:
// part 0 = everything in front of the tail, part 1 = the tail ([umi] n ":N:0:" barcode).
// which = 2: the barcode read's (illumina: also the paired read's, up to one digit); which = 1:
// the paired read's own full header (--mgi-full-header only).  `hl` = the header's length.
// reformat (`rf` = rf_pack() of the record): the tail is [":" umi] + what the input header holds from its separator on
// [+ " " n ":N:0:" + the sample's barcode when it holds nothing] (src/sample_data.rs:627-637, src/lib.rs:1256-1264).
MGK_HD void build_header(const Params& P, const uint8_t* S, uint32_t cs0, const RecGeom& G, int which, int bar_t, int umi_t,
                         uint32_t hl, uint8_t* o, int part, uint32_t rf = 0) {
  const bool full = P.out_mode == MGK_OUT_MGI_FULL_HEADER;
  const uint32_t h = which == 2 ? G.h2 : G.h1, s = which == 2 ? G.s2 : G.s1;
  RfHeader rh;
  rh.sep = G.s2 - G.h2 - 3u;   // demultiplex: the '/' (src/lib.rs:1007)
  if (P.reformat) rh = rf_header(P, G.s2 - G.h2, rf);
  if (P.reformat && part == 1) {
    const uint32_t U = (uint32_t)P.BL, bcl = (uint32_t)P.rf_bc_len;
    o += hl - ((U ? 1u + U : 0u) + rh.tcl + (rh.append ? 7u + bcl : 0u));
    if (U) {
      *o++ = ':';
      o = put_span(o, S, G.p2 - 1u - U, U);   // src/lib.rs:1117-1124
    }
    o = put_span(o, S, G.h2 + rh.sep, rh.tcl);
    if (rh.append) {
      *o++ = ' ';
      *o++ = (uint8_t)s_u8(S, G.s2 - 2u);
      *o++ = ':';
      *o++ = 'N';
      *o++ = ':';
      *o++ = '0';
      *o++ = ':';
      put_span(o, S, cs0 + (uint32_t)P.prefix_len + (uint32_t)LIT_BYTES, bcl);
    }
    return;
  }
  const TailSpec t = make_tail(P, G, bar_t, umi_t);
  const uint32_t tl = tail_len(t);
  if (part == 0) {
    if (full) {
      put_span(o, S, h, s - h - 1u);   // its own header line without the '\n'
    } else {                             // prefix lane ":" readno ":" ccc ":" rrr
      const HdrInfo hi = hdr_info(P, S, G, G.h2 + rh.sep);
      const uint32_t L = (uint32_t)P.L, H = G.h2, Pn = (uint32_t)P.prefix_len;
      if (((uintptr_t)o & 3u) == 0) {   // a staging slot: whole words (the constants are aligned too); the up to three
                                        // bytes past the prefix are overwritten by this thread right below
        MGK_LOOP_ROLLED
        for (uint32_t k = 0; k < Pn; k += 4u) *reinterpret_cast<uint32_t*>(o + k) = s_u32(S, cs0 + k);
        o += Pn;
      } else {
        o = put_span(o, S, cs0, Pn);
      }
      *o++ = (uint8_t)s_u8(S, H + L + 1u);
      *o++ = ':';
      o = put_span(o, S, hi.z, hi.sep - hi.z);
      *o++ = ':';
      o = put_span(o, S, hi.f1, H + L + 6u - hi.f1);
      *o++ = ':';
      put_span(o, S, hi.f2, H + L + 10u - hi.f2);
    }
    return;
  }
  o += hl - tl;
  const uint32_t n_digit = s_u8(S, s - 2u);   // src/lib.rs:1275,1288,1302,1314
  if (full) {
    *o++ = ' ';
    if (t.umi_len) {
      o = put_span(o, S, t.umi_off, t.umi_len);
      *o++ = ' ';
    }
  } else {
    if (t.umi_len) {
      *o++ = ':';
      o = put_span(o, S, t.umi_off, t.umi_len);
    }
    *o++ = ' ';
  }
  *o++ = (uint8_t)n_digit;
  *o++ = ':';
  *o++ = 'N';
  *o++ = ':';
  *o++ = '0';
  *o++ = ':';
  if (t.has_bar) {
    o = put_span(o, S, t.i7_off, t.i7_len);
    if (t.has_i5) {
      *o++ = '+';
      put_span(o, S, t.i5_off, t.i5_len);
    }
  }
}

constexpr uint32_t HDR_SLOT_BYTES = 80;    // staging slot of a rewritten header; longer ones are built at their destination

// The two byte ranges of a read's output that come straight from the input:
//   k = 0: [header line verbatim |] "\n" sequence
//   k = 1: "\n+\n" qualities [| "\n" unless the read is trimmed: then its closing '\n' is stored apart]
struct Piece {
  uint32_t dst_off, src, n;   // n == 0: no such piece
};
MGK_HD Piece read_piece(const ReadPlan& r, int k) {
  Piece pc;
  pc.dst_off = pc.src = pc.n = 0;
  if (r.len == 0) return pc;
  const bool staged = r.hdr != HDR_VERBATIM;
  if (k == 0) {
    pc.dst_off = staged ? r.hl : 0u;
    pc.src = staged ? r.src1 : r.h_src;
    pc.n = (staged ? 0u : r.hl) + r.n1;
  } else {
    pc.dst_off = r.hl + r.n1;
    pc.src = r.src2;
    pc.n = r.n2;
  }
  return pc;
}

// ---------------------------------------------------------------- QC sums out of S
MGK_HD uint32_t keep_range(uint32_t wd, int lo, int hi) {  // keep bytes [lo,hi) of the word
  if (lo < 0) lo = 0;
  if (hi > 4) hi = 4;
  if (hi <= lo) return 0u;
  uint32_t mask = hi == 4 ? 0xffffffffu : ((1u << (8 * hi)) - 1u);
  if (lo) mask &= ~((1u << (8 * lo)) - 1u);
  return wd & mask;
}
// One word into the running sums.  q.high counts in units of -128: the bytes >= 63 are marked 0x80 and added as signed
// bytes by the same dot-product instruction that adds the qualities (qc_line converts at the end).
MGK_HD void qc_word(Qc& q, uint32_t x) {
#if defined(__CUDA_ARCH__)
  q.sum = __dp4a(x, 0x01010101u, q.sum);
  q.high = (uint32_t)__dp4a((int)bytes_ge63(x), 0x01010101, (int)q.high);
#else
  q.sum += byte_sum(x);
  q.high -= 128u * (uint32_t)popc32(bytes_ge63(x));
#endif
}
MGK_HD void qc_add(Qc& q, const uint4& a) {
  qc_word(q, a.x);
  qc_word(q, a.y);
  qc_word(q, a.z);
  qc_word(q, a.w);
}
MGK_HD uint4 keep16(uint4 a, int lo, int hi) {
  a.x = keep_range(a.x, lo, hi);
  a.y = keep_range(a.y, lo - 4, hi - 4);
  a.z = keep_range(a.z, lo - 8, hi - 8);
  a.w = keep_range(a.w, lo - 12, hi - 12);
  return a;
}
// sum_qc (src/lib.rs:458-474) over S[start .. start+n): aligned vectors, the two edge vectors masked
MGK_HD Qc qc_line(const uint8_t* S, uint32_t start, uint32_t n) {
  Qc q;
  q.sum = q.high = 0;
  if (n == 0) return q;
  const uint32_t sh = start & 15u, al = start - sh;
  const uint32_t nvec = (sh + n + 15u) >> 4;
  const int end = (int)(sh + n);
  qc_add(q, keep16(s_v4(S, al), (int)sh, end));
  MGK_LOOP_ROLLED
  for (uint32_t v = 1; v + 1u < nvec; ++v) qc_add(q, s_v4(S, al + (v << 4)));
  if (nvec > 1u) qc_add(q, keep16(s_v4(S, al + ((nvec - 1u) << 4)), 0, end - (int)((nvec - 1u) << 4)));
  q.high = (0u - q.high) >> 7;
  return q;
}

// The six per-read numbers of src/lib.rs:1188-1210 in the order
// [paired high, paired sum, barcode high, barcode sum, rest-of-barcode-read high, rest sum].
struct QcSix {
  uint32_t v[6];
};
constexpr int QC_SLOTS = 6;
// statistics column of slot k (src/lib.rs:1190-1209; shift = 1 paired / 0 single-end, :873-877)
MGK_HD int qc_stat_column(int k, int paired) {
  const int shift = paired ? 1 : 0;
  switch (k) {
    case 0: return 0;
    case 1: return 6;
    case 2: return 2;
    case 3: return 8;
    case 4: return shift;
    default: return 6 + shift;
  }
}

// which: bit 0 the paired read's slots (0,1), bit 1 the barcode read's (2..5); the others are left 0
MGK_HD QcSix qc_pair(const Params& P, const uint8_t* S, const RecGeom& G, int which) {
  QcSix r;
  const uint32_t BL = (uint32_t)P.BL;
  Qc q1;
  q1.sum = q1.high = 0;
  if (P.paired && (which & 1)) q1 = qc_line(S, G.q1, G.e1 - G.q1);
  r.v[0] = q1.high;
  r.v[1] = q1.sum;
  r.v[2] = r.v[3] = r.v[4] = r.v[5] = 0;
  if (!(which & 2)) return r;
  // the barcode's own qualities are the last BL of the line (record_plan has refused shorter lines)
  const Qc qb = qc_line(S, G.e2 - BL, BL), qr = qc_line(S, G.q2, G.e2 - BL - G.q2);
  r.v[2] = qb.high;
  r.v[3] = qb.sum;
  r.v[4] = qr.high;
  r.v[5] = qr.sum;
  return r;
}

// ---------------------------------------------------------------- --validate out of S
// Lowest failing priority (1..7) of one pair, 0 if clean:
//   1..3 paired read ('@', base, quality)  4 header mismatch  5..7 barcode read
MGK_HD int validate_pair(const Params& P, const uint8_t* S, const RecGeom& G) {
  auto base_ok = [](uint32_t n) {
    return n == 'A' || n == 'C' || n == 'G' || n == 'T' || n == 'N' || n == 'a' || n == 'c' || n == 'g' || n == 't' || n == 'n';
  };
  if (P.paired) {
    if (s_u8(S, G.h1) != '@') return 1;
    for (uint32_t i = G.s1; i + 1u < G.p1; ++i)
      if (!base_ok(s_u8(S, i))) return 2;
    for (uint32_t i = G.q1; i < G.e1; ++i) {
      const uint32_t q = s_u8(S, i);
      if (q > 73u || q < 33u) return 3;
    }
    if (P.mgi_data) {
      const uint32_t l1 = G.s1 - 2u - G.h1, l2 = G.s2 - 2u - G.h2;
      if (l1 != l2) return 4;
      for (uint32_t i = 0; i < l1; ++i)
        if (s_u8(S, G.h1 + i) != s_u8(S, G.h2 + i)) return 4;
    }
  }
  if (s_u8(S, G.h2) != '@') return 5;
  for (uint32_t i = G.s2; i + 1u < G.p2; ++i)
    if (!base_ok(s_u8(S, i))) return 6;
  for (uint32_t i = G.q2; i < G.e2; ++i) {
    const uint32_t q = s_u8(S, i);
    if (q > 73u || q < 33u) return 7;
  }
  return 0;
}

}  // namespace mgk

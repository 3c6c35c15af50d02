// Record emission out of a STAGED TILE: header rewrite, barcode trim, Q30 /
// quality sums and --validate for one read pair whose bytes already sit in a
// fast byte array `S` (shared memory on the B200, a host array in the CPU
// emulation that the `-m "not gpu"` tests check against the oracle).
//
// ONE THREAD PER (record, role): the control-heavy part of a record (where the
// header digits start, which bytes survive the trim) is scalar work, so every
// lane of a warp carries a different record and runs the same straight-line
// code; nothing here needs a warp collective, which is also why the CPU
// emulation can call these functions as they are.
//   role A  emit_read2()   barcode read: new header, trimmed sequence and qualities
//   role B  emit_read1()   paired read: copy of the header, rest of the record verbatim
//   role C  qc_pair() / validate_pair()
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

// ---------------------------------------------------------------- 16-byte gather
MGK_HD uint32_t s_u8(const uint8_t* S, uint32_t off) { return S[off]; }
MGK_HD uint4 s_v4(const uint8_t* S, uint32_t off) { return *reinterpret_cast<const uint4*>(S + off); }

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

// 16 bytes starting at ANY offset: two aligned vectors, a word shift in two
// select stages, then a byte funnel shift.  Reads S[addr & ~15 .. +32).
MGK_HD uint4 gather16(const uint8_t* S, uint32_t addr) {
  const uint32_t al = addr & ~15u, sh = addr & 15u;
  const uint4 a = s_v4(S, al), b = s_v4(S, al + 16u);
  const bool s2 = (sh & 8u) != 0u, s1 = (sh & 4u) != 0u;
  const uint32_t t0 = s2 ? a.z : a.x, t1 = s2 ? a.w : a.y, t2 = s2 ? b.x : a.z, t3 = s2 ? b.y : a.w, t4 = s2 ? b.z : b.x,
                 t5 = s2 ? b.w : b.y;
  const uint32_t u0 = s1 ? t1 : t0, u1 = s1 ? t2 : t1, u2 = s1 ? t3 : t2, u3 = s1 ? t4 : t3, u4 = s1 ? t5 : t4;
  const uint32_t bs = (sh & 3u) * 8u;
  uint4 o;
  o.x = funnel_r(u0, u1, bs);
  o.y = funnel_r(u1, u2, bs);
  o.z = funnel_r(u2, u3, bs);
  o.w = funnel_r(u3, u4, bs);
  return o;
}

// bytes [0,k) of a, bytes [k,16) of b     (0 < k < 16)
MGK_HD uint32_t merge_word(uint32_t a, uint32_t b, int k) {  // k = bytes of this word taken from a, any int
  if (k >= 4) return a;
  if (k <= 0) return b;
  const uint32_t m = (1u << (8 * k)) - 1u;
  return (a & m) | (b & ~m);
}
MGK_HD uint4 merge16(const uint4& a, const uint4& b, int k) {
  uint4 o;
  o.x = merge_word(a.x, b.x, k);
  o.y = merge_word(a.y, b.y, k - 4);
  o.z = merge_word(a.z, b.z, k - 8);
  o.w = merge_word(a.w, b.w, k - 12);
  return o;
}
MGK_HD uint32_t put_byte(uint32_t w, int i, uint32_t v) {  // byte i of the word := v when 0 <= i < 4
  if (i < 0 || i > 3) return w;
  const uint32_t sh = 8u * (uint32_t)i;
  return (w & ~(0xffu << sh)) | (v << sh);
}
MGK_HD void put_byte16(uint4& o, int i, uint32_t v) {
  o.x = put_byte(o.x, i, v);
  o.y = put_byte(o.y, i - 4, v);
  o.z = put_byte(o.z, i - 8, v);
  o.w = put_byte(o.w, i - 12, v);
}

// ---------------------------------------------------------------- three-segment copy
// An output record is the concatenation of up to three byte ranges of S
// (rewritten header | sequence part | quality part).
struct Seg3 {
  uint32_t s0, n0, s1, n1, s2, n2;  // S offsets and lengths
};

MGK_HD void st_v4(uint8_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// lo, hi = two consecutive aligned vectors; the 16 bytes starting `sh` bytes into lo
MGK_HD uint4 shift16(const uint4& a, const uint4& b, uint32_t sh) {
  const bool s2 = (sh & 8u) != 0u, s1 = (sh & 4u) != 0u;
  const uint32_t t0 = s2 ? a.z : a.x, t1 = s2 ? a.w : a.y, t2 = s2 ? b.x : a.z, t3 = s2 ? b.y : a.w, t4 = s2 ? b.z : b.x,
                 t5 = s2 ? b.w : b.y;
  const uint32_t u0 = s1 ? t1 : t0, u1 = s1 ? t2 : t1, u2 = s1 ? t3 : t2, u3 = s1 ? t4 : t3, u4 = s1 ? t5 : t4;
  const uint32_t bs = (sh & 3u) * 8u;
  uint4 o;
  o.x = funnel_r(u0, u1, bs);
  o.y = funnel_r(u1, u2, bs);
  o.z = funnel_r(u2, u3, bs);
  o.w = funnel_r(u3, u4, bs);
  return o;
}

// 32 bytes to a 32-byte aligned global address: one full DRAM sector per lane per store
MGK_HD void st_v8(uint8_t* p, const uint4& a, const uint4& b) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
#else
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + 16) = b;
#endif
}

// the LAST n (< 32) bytes of the 32-byte value (a, b) go to [end - n, end), `end` 32-byte aligned: 16 + 8 + 4 + 2 + 1 ladder
MGK_HD void store_before(uint8_t* end, const uint4& a, const uint4& b, uint32_t n) {
  uint4 v = b;
  uint8_t* p = end;
  if (n & 16u) {
    p -= 16;
    st_v4(p, b);
    v = a;
  }
  uint32_t cur = v.w, prev = v.z;
  if (n & 8u) {
    p -= 8;
    uint2 t;
    t.x = v.z;
    t.y = v.w;
    *reinterpret_cast<uint2*>(p) = t;
    cur = v.y;
    prev = v.x;
  }
  if (n & 4u) {
    p -= 4;
    *reinterpret_cast<uint32_t*>(p) = cur;
    cur = prev;
  }
  if (n & 2u) {
    p -= 2;
    *reinterpret_cast<uint16_t*>(p) = (uint16_t)(cur >> 16);
    cur <<= 16;
  }
  if (n & 1u) {
    p -= 1;
    *p = (uint8_t)(cur >> 24);
  }
}
// the FIRST n (< 32) bytes of (a, b) go to [p, p + n), p 32-byte aligned
MGK_HD void store_after(uint8_t* p, const uint4& a, const uint4& b, uint32_t n) {
  uint4 v = a;
  if (n & 16u) {
    st_v4(p, a);
    p += 16;
    v = b;
  }
  uint32_t c0 = v.x, c1 = v.y;
  if (n & 8u) {
    uint2 t;
    t.x = v.x;
    t.y = v.y;
    *reinterpret_cast<uint2*>(p) = t;
    p += 8;
    c0 = v.z;
    c1 = v.w;
  }
  if (n & 4u) {
    *reinterpret_cast<uint32_t*>(p) = c0;
    p += 4;
    c0 = c1;
  }
  if (n & 2u) {
    *reinterpret_cast<uint16_t*>(p) = (uint16_t)(c0 & 0xffffu);
    p += 2;
    c0 >>= 16;
  }
  if (n & 1u) *p = (uint8_t)(c0 & 0xffu);
}

// A record's output may be shared by two threads: PART_HEAD takes the bytes in front of the
// first aligned unit and the first `split` units, PART_TAIL the rest.  Both compute the same
// split from the same inputs; `own0` leading output bytes (the staged header, which only the
// head thread can see) are guaranteed to stay in the head part.
enum { PART_HEAD = 0, PART_TAIL = 1, PART_ALL = 2 };

// This thread writes its part of dst[0 .. n0+n1+n2).  The bulk moves as 32-byte stores ALIGNED
// ON THE DESTINATION (one whole DRAM sector per store); inside a segment the source is read
// as a sliding window of aligned 16-byte vectors (two shared-memory loads per store).  The at
// most four "special" pieces — the < 32 bytes in front of the first aligned unit, the two
// units that straddle a segment boundary, the < 32 bytes after the last unit — are assembled
// from both sides of the boundary by ONE copy of the gather-and-merge code (instruction-cache
// footprint matters: every lane of every emit warp runs this function).
MGK_HD void copy3(uint8_t* dst, const uint8_t* S, const Seg3& g, int part, uint32_t own0) {
  const uint32_t len = g.n0 + g.n1 + g.n2;
  if (len == 0) return;
  const uint32_t o1 = g.n0, o2 = g.n0 + g.n1;
  const uint32_t b0 = g.s0, b1 = g.s1 - o1, b2 = g.s2 - o2;  // source offset of output byte y in segment j: bj + y
  uint32_t hb = (32u - (uint32_t)((uintptr_t)dst & 31u)) & 31u;
  if (hb > len) hb = len;
  const uint32_t nfull = (len - hb) >> 5;
  const uint32_t tail0 = hb + (nfull << 5), nt = len - tail0;
  uint32_t split = (nfull * 3u) >> 3;
  if (own0 > hb && ((own0 - hb + 31u) >> 5) > split) split = (own0 - hb + 31u) >> 5;
  if (split > nfull) split = nfull;
  const uint32_t x_lo = part == PART_TAIL ? hb + (split << 5) : hb;   // whole units of this part: [x_lo, x_hi)
  const uint32_t x_hi = part == PART_HEAD ? hb + (split << 5) : tail0;
  if (nfull == 0 && nt == 0) {   // the whole record sits in front of the first aligned boundary
    if (part != PART_TAIL)
      for (uint32_t y = 0; y < len; ++y) dst[y] = (uint8_t)s_u8(S, (y >= o2 ? b2 : (y >= o1 ? b1 : b0)) + y);
    return;
  }
  // ---- whole units that lie inside one segment
  MGK_LOOP_ROLLED
  for (int j = 0; j < 3; ++j) {
    const uint32_t bj = j == 0 ? b0 : (j == 1 ? b1 : b2);
    const uint32_t beg = j == 0 ? 0u : (j == 1 ? o1 : o2);
    uint32_t end = j == 0 ? o1 : (j == 1 ? o2 : len);
    if (end > x_hi) end = x_hi;
    uint32_t x = beg <= hb ? hb : hb + (((beg - hb) + 31u) & ~31u);   // first unit that starts inside the segment
    if (x < x_lo) x = x_lo;
    if (x + 32u <= end) {
      const uint32_t addr = bj + x, sh = addr & 15u;
      uint32_t al = addr - sh;
      uint4 lo = s_v4(S, al);
      do {
        const uint4 mid = s_v4(S, al + 16u), hi = s_v4(S, al + 32u);
        al += 32u;
        st_v8(dst + x, shift16(lo, mid, sh), shift16(mid, hi, sh));
        lo = hi;
        x += 32u;
      } while (x + 32u <= end);
    }
  }
  // ---- the special pieces: 0 head fragment, 1 / 2 unit across the boundary at o1 / o2, 3 tail fragment
  const int tail_owner = own0 > tail0 ? PART_HEAD : PART_TAIL;   // a staged header reaching into the tail keeps it in the head part
  uint32_t done_unit = 0xffffffffu;
  MGK_LOOP_ROLLED
  for (int k = 0; k < 4; ++k) {
    uint32_t x;
    bool on;
    if (k == 0) {
      x = hb - 32u;
      on = hb != 0u && part != PART_TAIL;
    } else if (k == 3) {
      x = tail0;
      on = nt != 0u && (part == PART_ALL || part == tail_owner);
    } else {
      const uint32_t o = k == 1 ? o1 : o2;
      x = hb + ((o - hb) & ~31u);                       // unit holding output byte o, if o is past the head fragment
      on = o > hb && o < tail0 && x != o && x >= x_lo && x + 32u <= x_hi && x != done_unit && (k == 1 ? g.n1 + g.n2 : g.n2) != 0u;
      if (on) done_unit = x;
    }
    if (!on) continue;
    // 2 x 16 output bytes starting at output offset x (may hang over either end of the record)
    uint4 wa, wb;
    wa.x = wa.y = wa.z = wa.w = 0;
    wb = wa;
    MGK_LOOP_ROLLED
    for (int h = 0; h < 2; ++h) {
      const uint32_t xh = x + 16u * (uint32_t)h, xe = xh + 16u;
      if ((int)xe <= 0 || (int)xh >= (int)len) continue;   // half entirely outside the record
      int j = ((int)xh >= (int)o2 && g.n2) ? 2 : (((int)xh >= (int)o1 && g.n1) ? 1 : 0);   // never an empty segment:
      if (j == 0 && g.n0 == 0) j = g.n1 ? 1 : 2;                                             // its base offset is meaningless
      uint4 w = gather16(S, (j == 2 ? b2 : (j == 1 ? b1 : b0)) + xh);
      if (j < 1 && (int)o1 < (int)xe && g.n1) w = merge16(w, gather16(S, b1 + xh), (int)(o1 - xh));
      if (j < 2 && (int)o2 < (int)xe && g.n2) w = merge16(w, gather16(S, b2 + xh), (int)(o2 - xh));
      if (h == 0) wa = w;
      else wb = w;
    }
    if (k == 0) store_before(dst + hb, wa, wb, hb);
    else if (k == 3) store_after(dst + tail0, wa, wb, nt);
    else st_v8(dst + x, wa, wb);
  }
}

// ---------------------------------------------------------------- header rewrite
// Positions are offsets into S.  Illumina: where the kept digits start
// (src/sample_data.rs:584-626): leading zeros of the read number are dropped (all
// zeros: nothing is left), the two 3-digit fields keep at least their last digit.
struct HdrInfo {
  uint32_t z, f1, f2;  // first kept byte of the read number (== sep if none), of the two 3-digit fields
  uint32_t sep;        // the '/' (header length - 3, src/lib.rs:1007)
};

MGK_HD HdrInfo hdr_info(const Params& P, const uint8_t* S, const RecGeom& G) {
  HdrInfo h;
  const uint32_t L = (uint32_t)P.L, H = G.h2;
  h.sep = G.s2 - 3u;
  h.f1 = s_u8(S, H + L + 3u) != '0' ? H + L + 3u : (s_u8(S, H + L + 4u) != '0' ? H + L + 4u : H + L + 5u);
  h.f2 = s_u8(S, H + L + 7u) != '0' ? H + L + 7u : (s_u8(S, H + L + 8u) != '0' ? H + L + 8u : H + L + 9u);
  uint32_t z = H + L + 10u;
  while (z < h.sep && s_u8(S, z) == '0') ++z;
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
MGK_HD uint32_t tail_len(const TailSpec& t, int full) {
  const uint32_t u = t.umi_len ? 1u + t.umi_len : 0u;
  (void)full;
  return u + 7u + barcode_len(t);   // both flavours: one separator more than the umi, n, ":N:0:"
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

// A header is written front to back through a cursor `o`: into the (record, read)'s staging
// slot in S or, for headers too long for the slot, straight to the destination (byte stores;
// correct for any length, just slower).  ONE out-of-line copy of the byte loop serves every
// piece: the emit code runs in every warp of the kernel and has to stay small.
MGK_HD_NOINLINE uint8_t* put_span(uint8_t* o, const uint8_t* S, uint32_t src, uint32_t n) {
  uint32_t k = 0;
  MGK_LOOP_ROLLED
  for (; k + 4u <= n; k += 4u) {   // loads first: the cursor may alias S as far as the compiler knows
    const uint32_t a = s_u8(S, src + k), b = s_u8(S, src + k + 1u), c = s_u8(S, src + k + 2u), d = s_u8(S, src + k + 3u);
    o[0] = (uint8_t)a;
    o[1] = (uint8_t)b;
    o[2] = (uint8_t)c;
    o[3] = (uint8_t)d;
    o += 4;
  }
  MGK_LOOP_ROLLED
  for (; k < n; ++k) *o++ = (uint8_t)s_u8(S, src + k);
  return o;
}
MGK_HD uint8_t* put_tail(uint8_t* o, const uint8_t* S, const TailSpec& t, uint32_t n_digit, int full) {
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
      o = put_span(o, S, t.i5_off, t.i5_len);
    }
  }
  return o;
}
MGK_HD uint32_t illumina_header_len(const Params& P, const RecGeom& G, const HdrInfo& hi, const TailSpec& t) {
  const uint32_t L = (uint32_t)P.L, H = G.h2;
  return (uint32_t)P.prefix_len + 2u + (hi.sep - hi.z) + 1u + (H + L + 6u - hi.f1) + 1u + (H + L + 10u - hi.f2) + tail_len(t, 0);
}
// prefix lane ":" readno ":" ccc ":" rrr tail.  `cs0` = S offset of the prefix; `vec`: the cursor is
// a 16-byte aligned staging slot, so the constant prefix goes in as whole vectors.
MGK_HD uint8_t* put_illumina_header(uint8_t* o, bool vec, const Params& P, const uint8_t* S, uint32_t cs0, const RecGeom& G,
                                    const HdrInfo& hi, const TailSpec& t, uint32_t n_digit) {
  const uint32_t L = (uint32_t)P.L, H = G.h2, Pn = (uint32_t)P.prefix_len;
  if (vec) {
    MGK_LOOP_ROLLED
    for (uint32_t k = 0; k < Pn; k += 16u) *reinterpret_cast<uint4*>(o + k) = s_v4(S, cs0 + k);
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
  o = put_span(o, S, hi.f2, H + L + 10u - hi.f2);
  return put_tail(o, S, t, n_digit, 0);
}

// ---------------------------------------------------------------- the emit roles
constexpr uint32_t HDR_SLOT_BYTES = 96;     // per (record, read) staging slot; longer headers take the byte-store path
// Every record, matched or not, is cut into the SAME three pieces (header line | sequence
// line | "+" and quality lines) so that the lanes of a warp, each on its own record, run the
// copy loops of copy3() in step whatever their records look like; and both reads go through
// the SAME function so that the four emit roles of the kernel share one set of instructions.
//
// which = 2: barcode read -> new header, barcode trimmed off sequence and qualities
//            (src/lib.rs:1240-1327); unmatched reads verbatim (:1225-1235)
// which = 1: paired read  -> in illumina mode the barcode read's new header with its own
//            read-number digit (src/sample_data.rs:302-328, src/lib.rs:1288), the rest of the
//            record verbatim (:1328-1331); with --mgi-full-header ITS OWN header line (:1306-1318)
// `hs` = this (record, read)'s staging slot (S offset, 16-byte aligned), written by the
// PART_HEAD thread only.
MGK_HD void emit_read(const Params& P, uint8_t* S, uint32_t hs, uint32_t cs0, const RecGeom& G, int which, int sample, int bar_t,
                      int umi_t, uint8_t* d, int part) {
  const bool br = which == 2;
  const uint32_t h = br ? G.h2 : G.h1, s = br ? G.s2 : G.s1, p = br ? G.p2 : G.p1, e = br ? G.e2 : G.e1;
  const bool matched = sample < P.S;
  if (br && matched && !P.read2_has_sequence) return;   // samples get no barcode-read file, src/sample_data.rs:35-44
  const bool trim = br && matched;
  const uint32_t wbl = (trim && !P.keep_barcode) ? (uint32_t)P.BL : 0u;
  Seg3 g;
  g.s1 = s - 1u;
  g.n1 = p - wbl - s;                                   // "\n" seq[..-wbl], src/lib.rs:1293,1321-1323
  g.s2 = p - 1u;
  g.n2 = trim ? e - wbl - (p - 1u) : e + 2u - p;        // "\n+\n" qual[..-wbl]; a trimmed read's closing '\n' is stored on its own
  g.s0 = h;
  g.n0 = s - 1u - h;                                    // the header line as it is
  uint32_t own0 = 0, skip = 0;
  if (matched && P.out_mode != MGK_OUT_MGI) {
    const int full = P.out_mode == MGK_OUT_MGI_FULL_HEADER;
    const TailSpec t = make_tail(P, G, bar_t, umi_t);
    const uint32_t n_digit = s_u8(S, s - 2u);           // src/lib.rs:1275,1288,1302,1314
    HdrInfo hi;
    uint32_t hl;
    if (full) {
      hi.z = hi.f1 = hi.f2 = hi.sep = 0;
      hl = s - h - 1u + tail_len(t, 1);
    } else {
      hi = hdr_info(P, S, G);
      hl = illumina_header_len(P, G, hi, t);
    }
    const bool slot = hl <= HDR_SLOT_BYTES;
    if (part != PART_TAIL) {   // too long for the slot: bytes straight to the destination
      uint8_t* o = slot ? S + hs : d;
      if (full) {
        o = put_span(o, S, h, s - h - 1u);
        put_tail(o, S, t, n_digit, 1);
      } else {
        put_illumina_header(o, slot, P, S, cs0, G, hi, t, n_digit);
      }
    }
    if (slot) {
      g.s0 = hs;
      g.n0 = hl;
      own0 = hl;
    } else {
      g.s0 = g.n0 = 0;
      skip = hl;
    }
  }
  copy3(d + skip, S, g, part, own0);
  if (trim && part != PART_HEAD) d[skip + g.n0 + g.n1 + g.n2] = '\n';   // src/lib.rs:1327
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
MGK_HD void qc_add(Qc& q, const uint4& a) {
  q.sum += byte_sum(a.x) + byte_sum(a.y) + byte_sum(a.z) + byte_sum(a.w);
  q.high += (uint32_t)(popc32(bytes_ge63(a.x)) + popc32(bytes_ge63(a.y)) + popc32(bytes_ge63(a.z)) + popc32(bytes_ge63(a.w)));
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
  const Qc q2 = qc_line(S, G.q2, G.e2 - G.q2);
  uint32_t sb = 0, hb = 0;
  MGK_LOOP_ROLLED
  for (uint32_t k = 0; k < BL; ++k) {   // the barcode's own qualities
    const uint32_t q = s_u8(S, G.e2 - BL + k);
    sb += q;
    hb += q >= 63u ? 1u : 0u;
  }
  r.v[2] = hb;
  r.v[3] = sb;
  r.v[4] = q2.high - hb;
  r.v[5] = q2.sum - sb;
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

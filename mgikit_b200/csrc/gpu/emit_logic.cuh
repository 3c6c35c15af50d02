// Record emission out of a STAGED TILE: header rewrite, barcode trim, Q30 /
// quality sums and --validate for one read pair whose bytes already sit in a
// fast byte array `S` (shared memory on the B200, a host array in the CPU
// emulation that the `-m "not gpu"` tests check against the oracle).
//
// Everything is written against a tiny warp interface `W`:
//     w.each(f)        run f(lane) for the 32 lanes          (device: this lane)
//     w.sync()         order the lanes' writes to S          (device: __syncwarp)
//     w.ballot(f)      bit l = f(l)                          (device: __ballot_sync)
//     w.sum(f)         sum of f(l) over the lanes            (device: __reduce_add_sync)
//     w.min(f)         min of f(l) over the lanes            (device: __reduce_min_sync)
//     w.lane0(f)       run f() once
// so that the byte-exact logic below is ONE piece of source for both targets.
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

// ---------------------------------------------------------------- three-segment emit
// An output record is the concatenation of up to three byte ranges of S
// (rewritten header | sequence part | quality part) with at most one byte
// overridden (the closing '\n' after a trimmed quality line, or the read-number
// digit of the paired read's copy of the header).
struct Seg3 {
  uint32_t s0, n0, s1, n1, s2, n2;  // S offsets and lengths
  uint32_t patch_at, patch_val;     // output index to override (>= len: none)
};

MGK_HD uint32_t seg3_len(const Seg3& g) { return g.n0 + g.n1 + g.n2; }

// Lane `lane` writes its share of dst[0..len).  The bulk moves as 16-byte stores
// ALIGNED ON THE DESTINATION; each takes its bytes from the segment holding its
// first byte and, where it straddles a boundary, merges in the next segment(s).
// The < 16 bytes in front of the first aligned unit go one per lane (lanes 0-15),
// those after the last unit likewise (lanes 16-31).
MGK_HD void emit3(uint8_t* dst, const uint8_t* S, const Seg3& g, int lane) {
  const uint32_t len = seg3_len(g);
  const uint32_t o1 = g.n0, o2 = g.n0 + g.n1;
  const uint32_t b0 = g.s0, b1 = g.s1 - o1, b2 = g.s2 - o2;  // source offset of output byte y in segment j: bj + y
  uint32_t hb = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
  if (hb > len) hb = len;
  const uint32_t nfull = (len - hb) >> 4;
  const uint32_t tail0 = hb + (nfull << 4);
  {
    const uint32_t y = lane < 16 ? (uint32_t)lane : tail0 + (uint32_t)(lane - 16);
    const bool on = lane < 16 ? (uint32_t)lane < hb : y < len;
    if (on) {
      const uint32_t base = y >= o2 ? b2 : (y >= o1 ? b1 : b0);
      uint32_t v = s_u8(S, base + y);
      if (y == g.patch_at) v = g.patch_val;
      dst[y] = (uint8_t)v;
    }
  }
  for (uint32_t u = (uint32_t)lane; u < nfull; u += 32u) {
    const uint32_t x = hb + (u << 4), xe = x + 16u;
    const int j = (x >= o1 ? 1 : 0) + (x >= o2 ? 1 : 0);
    uint4 w = gather16(S, (j == 2 ? b2 : (j == 1 ? b1 : b0)) + x);
    if (j < 1 && o1 < xe && g.n1) w = merge16(w, gather16(S, b1 + x), (int)(o1 - x));
    if (j < 2 && o2 < xe && g.n2) w = merge16(w, gather16(S, b2 + x), (int)(o2 - x));
    if (g.patch_at >= x && g.patch_at < xe) put_byte16(w, (int)(g.patch_at - x), g.patch_val);
    *reinterpret_cast<uint4*>(dst + x) = w;
  }
}

// ---------------------------------------------------------------- record geometry inside S
// RecGeom (device_logic.cuh) with every position an offset into S.

// Illumina header rewrite: where the kept digits start (src/sample_data.rs:584-626).
struct HdrInfo {
  uint32_t z, f1, f2;  // S offsets: first kept byte of the read number (== sep if none), of the two 3-digit fields
  uint32_t sep;        // S offset of the '/' (header length - 3, src/lib.rs:1007)
};

template <class W>
MGK_HD HdrInfo hdr_info(const W& w, const Params& P, const uint8_t* S, const RecGeom& G) {
  HdrInfo h;
  const uint32_t L = (uint32_t)P.L, H = G.h2;
  h.sep = G.s2 - 3u;
  // one ballot covers the two 3-digit fields (bits 0-2, 4-6) and the first 25 digits of the read number
  const uint32_t w0 = H + L + 3u;
  uint32_t nz = w.ballot([&](int lane) {
    const uint32_t at = w0 + (uint32_t)lane;
    return (lane < 7 || at < h.sep) && s_u8(S, at) != '0';   // the two 3-digit fields are read unconditionally
  });
  h.f1 = (nz & 1u) ? H + L + 3u : ((nz & 2u) ? H + L + 4u : H + L + 5u);
  h.f2 = (nz & 0x10u) ? H + L + 7u : ((nz & 0x20u) ? H + L + 8u : H + L + 9u);
  uint32_t rest = nz >> 7;
  uint32_t base = H + L + 10u;
  h.z = h.sep;
  for (;;) {
    if (rest) {
      h.z = base + (uint32_t)ctz32(rest);
      break;
    }
    base += (base == H + L + 10u) ? 25u : 32u;
    if (base >= h.sep) break;
    const uint32_t bb = base;
    rest = w.ballot([&](int lane) {
      const uint32_t at = bb + (uint32_t)lane;
      return at < h.sep && s_u8(S, at) != '0';
    });
  }
  return h;
}

// Source of byte k of the text that follows the original/new header fields:
//   illumina : [":" umi] " " n ":N:0:" barcode            (src/sample_data.rs:627-653)
//   mgi full : " " [umi " "] n ":N:0:" barcode            (src/lib.rs:1294-1319)
// `cs` = S offset of the literals (LIT_*), `bar` = S offset of the BL barcode bytes,
// `nsrc` = S offset of the read-number digit.
struct TailSpec {
  uint32_t cs, bar, nsrc;
  uint32_t umi_off, umi_len;   // S offset / length of the UMI (0 length: none)
  uint32_t i7_off, i7_len, i5_off, i5_len, has_i5, has_bar;
  int full;                    // 1: mgi full header flavour
};
MGK_HD uint32_t tail_len(const TailSpec& t) {
  const uint32_t u = t.umi_len ? 1u + t.umi_len : 0u;
  const uint32_t b = t.has_bar ? t.i7_len + (t.has_i5 ? 1u + t.i5_len : 0u) : 0u;
  return (t.full ? 1u : 0u) + u + (t.full ? 0u : 1u) + 1u + 5u + b;
}
MGK_HD uint32_t tail_src(const TailSpec& t, uint32_t k) {
  uint32_t u0;
  if (t.full) {  // " " [umi " "]
    if (k == 0) return t.cs + LIT_SPACE;
    u0 = 1u;
    if (t.umi_len) {
      if (k < 1u + t.umi_len) return t.umi_off + (k - 1u);
      if (k == 1u + t.umi_len) return t.cs + LIT_SPACE;
      u0 = 2u + t.umi_len;
    }
  } else {       // [":" umi] " "
    u0 = 0u;
    if (t.umi_len) {
      if (k == 0) return t.cs + LIT_COLON;
      if (k < 1u + t.umi_len) return t.umi_off + (k - 1u);
      u0 = 1u + t.umi_len;
    }
    if (k == u0) return t.cs + LIT_SPACE;
    u0 += 1u;
  }
  if (k == u0) return t.nsrc;
  if (k < u0 + 6u) return t.cs + LIT_TAIL + (k - u0 - 1u);
  const uint32_t b0 = u0 + 6u;
  if (k < b0 + t.i7_len) return t.i7_off + (k - b0);
  if (k == b0 + t.i7_len) return t.cs + LIT_PLUS;
  return t.i5_off + (k - b0 - t.i7_len - 1u);
}

MGK_HD TailSpec make_tail(const Params& P, const RecGeom& G, int bar_t, int umi_t, uint32_t cs, uint32_t nsrc, int full) {
  TailSpec t;
  t.cs = cs;
  t.bar = G.p2 - 1u - (uint32_t)P.BL;
  t.nsrc = nsrc;
  t.full = full;
  t.umi_off = t.umi_len = 0;
  if (umi_t >= 0) {
    t.umi_off = t.bar + (uint32_t)(P.BL - P.tpl[umi_t].g[8]);
    t.umi_len = (uint32_t)P.tpl[umi_t].g[7];
  }
  t.has_bar = bar_t >= 0 ? 1u : 0u;
  t.i7_off = t.i5_off = t.bar;
  t.i7_len = t.i5_len = t.has_i5 = 0;
  if (bar_t >= 0) {
    const TemplateDev& T = P.tpl[bar_t];
    t.i7_off = t.bar + (uint32_t)(P.BL - T.g[2]);
    t.i7_len = (uint32_t)T.g[1];
    t.has_i5 = T.has_i5 ? 1u : 0u;
    if (T.has_i5) {
      t.i5_off = t.bar + (uint32_t)(P.BL - T.g[5]);
      t.i5_len = (uint32_t)T.g[4];
    }
  }
  return t;
}

// Illumina header into the staging area S[hs..): prefix lane ":" readno ":" ccc ":" rrr tail.
// Returns its length.  `cs0` = S offset of the prefix (the literals follow it).
template <class W>
MGK_HD uint32_t build_illumina_header(const W& w, const Params& P, uint8_t* S, uint32_t hs, uint32_t cs0, const RecGeom& G,
                                      const HdrInfo& hi, const TailSpec& t) {
  const uint32_t L = (uint32_t)P.L, H = G.h2, Pn = (uint32_t)P.prefix_len, lit = cs0 + Pn;
  const uint32_t nr = hi.sep - hi.z, n1 = H + L + 6u - hi.f1, n2 = H + L + 10u - hi.f2;
  const uint32_t t1 = 2u + nr, t2 = t1 + 1u + n1, t3 = t2 + 1u + n2, tl = tail_len(t);
  const uint32_t total = Pn + t3 + tl;
  w.each([&](int lane) {
    for (uint32_t k = (uint32_t)lane; k < total; k += 32u) {
      uint32_t src;
      if (k < Pn) {
        src = cs0 + k;
      } else {
        const uint32_t q = k - Pn;
        if (q >= t3) src = tail_src(t, q - t3);
        else if (q == 0) src = H + L + 1u;
        else if (q == 1 || q == t1 || q == t2) src = lit + LIT_COLON;
        else if (q < t1) src = hi.z + (q - 2u);
        else if (q < t2) src = hi.f1 + (q - t1 - 1u);
        else src = hi.f2 + (q - t2 - 1u);
      }
      S[hs + k] = (uint8_t)s_u8(S, src);
    }
  });
  return total;
}

// MGI full header of one read: its own header line without the '\n', then the tail.
template <class W>
MGK_HD uint32_t build_full_header(const W& w, uint8_t* S, uint32_t hs, uint32_t h, uint32_t hl1, const TailSpec& t) {
  const uint32_t total = hl1 + tail_len(t);
  w.each([&](int lane) {
    for (uint32_t k = (uint32_t)lane; k < total; k += 32u) {
      const uint32_t src = k < hl1 ? h + k : tail_src(t, k - hl1);
      S[hs + k] = (uint8_t)s_u8(S, src);
    }
  });
  return total;
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
// this sub-lane's share of sum_qc over S[start .. start+n): aligned vectors, edges masked
MGK_HD Qc qc_span(const uint8_t* S, uint32_t start, uint32_t n, uint32_t sub, uint32_t nsub) {
  Qc q;
  q.sum = q.high = 0;
  if (n == 0) return q;
  const uint32_t sh = start & 15u, al = start - sh;
  const uint32_t nvec = (sh + n + 15u) >> 4;
  for (uint32_t v = sub; v < nvec; v += nsub) {
    uint4 a = s_v4(S, al + (v << 4));
    const int lo = (int)sh - (int)(v << 4), hi = (int)(sh + n) - (int)(v << 4);
    if (lo > 0 || hi < 16) {
      a.x = keep_range(a.x, lo, hi);
      a.y = keep_range(a.y, lo - 4, hi - 4);
      a.z = keep_range(a.z, lo - 8, hi - 8);
      a.w = keep_range(a.w, lo - 12, hi - 12);
    }
    q.sum += byte_sum(a.x) + byte_sum(a.y) + byte_sum(a.z) + byte_sum(a.w);
    q.high += (uint32_t)(popc32(bytes_ge63(a.x)) + popc32(bytes_ge63(a.y)) + popc32(bytes_ge63(a.z)) + popc32(bytes_ge63(a.w)));
  }
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

struct U4 {
  uint32_t a, b, c, d;
};

template <class W>
MGK_HD QcSix qc_pair(const W& w, const Params& P, const uint8_t* S, const RecGeom& G) {
  QcSix r;
  const uint32_t BL = (uint32_t)P.BL;
  const uint32_t n2 = G.e2 - G.q2, n1 = P.paired ? G.e1 - G.q1 : 0u;
  const int paired = P.paired;
  // whole quality lines: lanes 0-15 the paired read, lanes 16-31 the barcode read (all 32 when single-end)
  const U4 t = w.sum4([&](int lane) {
    U4 v;
    Qc q;
    if (!paired) q = qc_span(S, G.q2, n2, (uint32_t)lane, 32u);
    else if (lane < 16) q = qc_span(S, G.q1, n1, (uint32_t)lane, 16u);
    else q = qc_span(S, G.q2, n2, (uint32_t)lane - 16u, 16u);
    v.a = q.sum;
    v.b = q.high;
    v.c = (paired && lane < 16) ? q.sum : 0u;
    v.d = (paired && lane < 16) ? q.high : 0u;
    return v;
  });
  // the barcode's own qualities, one byte per lane: sum in the low half, count >= Q30 in the high half
  // (BL <= 255, so neither half overflows)
  const uint32_t bq = G.e2 - BL;
  const uint32_t pb = w.sum([&](int lane) {
    uint32_t s = 0;
    for (uint32_t k = (uint32_t)lane; k < BL; k += 32u) {
      const uint32_t q = s_u8(S, bq + k);
      s += q + (q >= 63u ? 0x10000u : 0u);
    }
    return s;
  });
  const uint32_t sum_b = pb & 0xffffu, high_b = pb >> 16;
  r.v[0] = t.d;
  r.v[1] = t.c;
  r.v[2] = high_b;
  r.v[3] = sum_b;
  r.v[4] = t.b - t.d - high_b;
  r.v[5] = t.a - t.c - sum_b;
  return r;
}

// ---------------------------------------------------------------- --validate out of S
// Lowest failing priority (1..7) of one pair, 0 if clean:
//   1..3 paired read ('@', base, quality)  4 header mismatch  5..7 barcode read
MGK_HD int validate_lane(const Params& P, const uint8_t* S, const RecGeom& G, int lane) {
  int worst = 0;
  auto note = [&](int p) {
    if (worst == 0 || p < worst) worst = p;
  };
  auto base_ok = [](uint32_t n) {
    return n == 'A' || n == 'C' || n == 'G' || n == 'T' || n == 'N' || n == 'a' || n == 'c' || n == 'g' || n == 't' || n == 'n';
  };
  if (P.paired) {
    if (lane == 0 && s_u8(S, G.h1) != '@') note(1);
    for (uint32_t i = G.s1 + (uint32_t)lane; i + 1u < G.p1; i += 32u)
      if (!base_ok(s_u8(S, i))) note(2);
    for (uint32_t i = G.q1 + (uint32_t)lane; i < G.e1; i += 32u) {
      const uint32_t q = s_u8(S, i);
      if (q > 73u || q < 33u) note(3);
    }
    if (P.mgi_data) {
      const uint32_t l1 = G.s1 - 2u - G.h1, l2 = G.s2 - 2u - G.h2;
      if (l1 != l2) {
        if (lane == 0) note(4);
      } else {
        for (uint32_t i = (uint32_t)lane; i < l1; i += 32u)
          if (s_u8(S, G.h1 + i) != s_u8(S, G.h2 + i)) note(4);
      }
    }
  }
  if (lane == 0 && s_u8(S, G.h2) != '@') note(5);
  for (uint32_t i = G.s2 + (uint32_t)lane; i + 1u < G.p2; i += 32u)
    if (!base_ok(s_u8(S, i))) note(6);
  for (uint32_t i = G.q2 + (uint32_t)lane; i < G.e2; i += 32u) {
    const uint32_t q = s_u8(S, i);
    if (q > 73u || q < 33u) note(7);
  }
  return worst;
}

// ---------------------------------------------------------------- one read pair
struct PairOut {
  int validate;   // lowest failing priority, 0 = clean (only when P.validate)
  int hdr_overflow;
  QcSix qc;       // only when P.report_level > 0
};

constexpr uint32_t HDR_STAGE_BYTES = 256;   // per warp; rewritten headers longer than this are refused
constexpr uint32_t HDR_STAGE_MAX = HDR_STAGE_BYTES - 16;

// `S` holds the tile (both reads), the constants (prefix + literals at S[cs0..))
// and this warp's header staging area S[hs .. hs+HDR_STAGE_BYTES).  `G` is in S
// offsets.  d2 / d1 = where this pair's two output records start.
template <class W>
MGK_HD PairOut emit_pair(const W& w, const Params& P, uint8_t* S, uint32_t hs, uint32_t cs0, const RecGeom& G, int sample,
                         int bar_t, int umi_t, uint8_t* d2, uint8_t* d1) {
  PairOut out;
  out.validate = 0;
  out.hdr_overflow = 0;
  const uint32_t BL = (uint32_t)P.BL, wbl = P.keep_barcode ? 0u : BL;
  const uint32_t lit = cs0 + (uint32_t)P.prefix_len;
  const bool matched = sample < P.S;
  Seg3 g2, g1;
  g2.patch_at = g1.patch_at = 0xffffffffu;
  g2.patch_val = g1.patch_val = '\n';
  g2.s2 = g2.n2 = g1.s2 = g1.n2 = 0;
  if (!matched) {  // verbatim, src/lib.rs:1225-1235
    g2.s0 = G.h2; g2.n0 = G.e2 + 1u - G.h2; g2.s1 = g2.n1 = 0;
    g1.s0 = G.h1; g1.n0 = P.paired ? G.e1 + 1u - G.h1 : 0u; g1.s1 = g1.n1 = 0;
    w.each([&](int lane) {
      emit3(d2, S, g2, lane);
      if (P.paired) emit3(d1, S, g1, lane);
    });
  } else {
    // quality part "\n+\n" qual[..-wbl] and the closing '\n' (src/lib.rs:1324-1327): one extra byte, overridden
    const uint32_t qs = G.p2 - 1u, qn = G.e2 - wbl - (G.p2 - 1u) + 1u;
    if (P.out_mode == MGK_OUT_MGI) {
      if (P.read2_has_sequence) {
        g2.s0 = G.h2; g2.n0 = G.p2 - wbl - 1u - G.h2;     // header and trimmed sequence are contiguous
        g2.s1 = qs; g2.n1 = qn;
        g2.patch_at = g2.n0 + g2.n1 - 1u;
      }
      g1.s0 = G.h1; g1.n0 = P.paired ? G.e1 + 1u - G.h1 : 0u; g1.s1 = g1.n1 = 0;
      w.each([&](int lane) {
        if (P.read2_has_sequence) emit3(d2, S, g2, lane);
        if (P.paired) emit3(d1, S, g1, lane);
      });
    } else {
      const int full = P.out_mode == MGK_OUT_MGI_FULL_HEADER;
      uint32_t hl;
      TailSpec t = make_tail(P, G, bar_t, umi_t, lit, G.s2 - 2u, full);  // n of the barcode read, src/lib.rs:1275,1302
      if (!full) {
        const HdrInfo hi = hdr_info(w, P, S, G);
        const uint32_t nr = hi.sep - hi.z;
        if ((uint32_t)P.prefix_len + nr + tail_len(t) + 16u > HDR_STAGE_MAX) {
          out.hdr_overflow = 1;
          return out;
        }
        hl = build_illumina_header(w, P, S, hs, cs0, G, hi, t);
      } else {
        const uint32_t hl1 = G.s2 - G.h2 - 1u;
        if (hl1 + tail_len(t) > HDR_STAGE_MAX || (P.paired && G.s1 - G.h1 - 1u + tail_len(t) > HDR_STAGE_MAX)) {
          out.hdr_overflow = 1;
          return out;
        }
        hl = build_full_header(w, S, hs, G.h2, hl1, t);
      }
      w.sync();
      if (P.read2_has_sequence) {
        g2.s0 = hs; g2.n0 = hl;
        g2.s1 = G.s2 - 1u; g2.n1 = G.p2 - wbl - 1u - (G.s2 - 1u);   // "\n" seq[..-wbl], src/lib.rs:1293,1321-1323
        g2.s2 = qs; g2.n2 = qn;
        g2.patch_at = g2.n0 + g2.n1 + g2.n2 - 1u;
      } else {
        g2.s0 = g2.n0 = g2.s1 = g2.n1 = 0;
      }
      if (P.paired && !full) {   // copy of the new header with the read-number digit of the paired read
        g1.s0 = hs; g1.n0 = hl;                                     // src/sample_data.rs:302-328
        g1.s1 = G.s1 - 1u; g1.n1 = G.e1 + 1u - (G.s1 - 1u);         // src/lib.rs:1328-1331
        g1.patch_at = hl - 6u - (t.has_bar ? t.i7_len + (t.has_i5 ? 1u + t.i5_len : 0u) : 0u);
        g1.patch_val = s_u8(S, G.s1 - 2u);                          // src/lib.rs:1288
      }
      w.each([&](int lane) {
        if (P.read2_has_sequence) emit3(d2, S, g2, lane);
        if (P.paired && !full) emit3(d1, S, g1, lane);
      });
      if (P.paired && full) {    // the paired read keeps ITS OWN header line (src/lib.rs:1306-1318)
        w.sync();
        t.nsrc = G.s1 - 2u;
        const uint32_t hl1 = G.s1 - G.h1 - 1u;
        const uint32_t hlp = build_full_header(w, S, hs, G.h1, hl1, t);
        w.sync();
        g1.s0 = hs; g1.n0 = hlp;
        g1.s1 = G.s1 - 1u; g1.n1 = G.e1 + 1u - (G.s1 - 1u);
        w.each([&](int lane) { emit3(d1, S, g1, lane); });
      }
      w.sync();  // the staging area is free again
    }
  }
  if (P.report_level > 0) out.qc = qc_pair(w, P, S, G);
  if (P.validate) {
    const int v = w.min([&](int lane) {
      const int x = validate_lane(P, S, G, lane);
      return x ? x : 15;
    });
    out.validate = v == 15 ? 0 : v;
  }
  return out;
}

}  // namespace mgk

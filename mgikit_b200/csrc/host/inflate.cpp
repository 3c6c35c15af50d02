#include "inflate.hpp"

#include <sys/mman.h>

#include <algorithm>
#include <cstring>
#include <new>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mgikit {

namespace {

// decode-table entry
//   bits 0-3   code bits to consume (all the codes of a literal run; subtable pointer: unused)
//   literal run (bit 7 set):   bits 4-5 number of literals (1-3), bits 8-31 the literal bytes, first one lowest
//   anything else (bit 7 clear): bit 4 end of block, bit 5 subtable pointer, bit 6 not a code (incomplete set),
//                              bits 8-11 extra bits that follow the code (subtable pointer: index bits of the subtable),
//                              bits 16-31 length base / distance base / offset of the subtable
// The primary literal/length table resolves up to three literals per lookup when their codes fit its index together:
// FASTQ out of `gzip -1` is almost only literals, and the serial chain lookup -> shift -> lookup is what bounds it.
constexpr uint32_t E_LIT = 1u << 7, E_EOB = 1u << 4, E_SUB = 1u << 5, E_BAD = 1u << 6;
constexpr uint32_t E_SPECIAL = E_EOB | E_SUB | E_BAD;

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CLEN_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}

inline void store32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }

inline unsigned reverse_bits(unsigned code, unsigned len) {
  unsigned r = 0;
  for (unsigned i = 0; i < len; ++i) {
    r = (r << 1) | (code & 1u);
    code >>= 1;
  }
  return r;
}

// Canonical Huffman code -> single-symbol decode table.  lens[sym] in 0..15; entry_of(sym) gives the entry without
// its code length.  Returns false for an over-subscribed set, and for an incomplete one unless it is a single 1-bit
// code or (allow_empty) no code at all -- the sets zlib accepts (inftrees.c).
template <typename F>
bool build_table(const uint8_t* lens, unsigned n, uint32_t* table, unsigned primary_bits, size_t table_cap, bool allow_empty, F entry_of) {
  unsigned count[16] = {0};
  for (unsigned s = 0; s < n; ++s) count[lens[s]]++;
  unsigned max_len = 15;
  while (max_len > 0 && count[max_len] == 0) --max_len;
  const size_t primary = (size_t)1 << primary_bits;
  if (max_len == 0) {
    if (!allow_empty) return false;
    for (size_t i = 0; i < primary; ++i) table[i] = E_BAD | 1u;
    return true;
  }
  int left = 1;
  for (unsigned l = 1; l <= 15; ++l) {
    left <<= 1;
    left -= (int)count[l];
    if (left < 0) return false;
  }
  if (left > 0 && max_len != 1) return false;
  const unsigned sub_bits = max_len > primary_bits ? max_len - primary_bits : 0;
  if (left > 0 || sub_bits)
    for (size_t i = 0; i < primary; ++i) table[i] = E_BAD | 1u;   // a complete short code fills every slot below
  // symbols by (length, symbol): canonical order
  unsigned offs[17], next_code[16], code = 0;
  uint16_t order[320];
  offs[1] = 0;
  for (unsigned l = 1; l <= 15; ++l) {
    offs[l + 1] = offs[l] + count[l];
    code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
    next_code[l] = code;
  }
  {
    unsigned at[16];
    for (unsigned l = 1; l <= 15; ++l) at[l] = offs[l];
    for (unsigned s = 0; s < n; ++s)
      if (lens[s]) order[at[lens[s]]++] = (uint16_t)s;
  }
  size_t sub_next = primary;
  for (unsigned l = 1; l <= max_len; ++l) {
    for (unsigned k = offs[l]; k < offs[l + 1]; ++k) {
      const unsigned s = order[k];
      const unsigned rev = reverse_bits(next_code[l]++, l);
      const uint32_t e = entry_of(s) | l;
      if (l <= primary_bits) {
        for (size_t i = rev; i < primary; i += (size_t)1 << l) table[i] = e;
      } else {
        const size_t slot = rev & (primary - 1);
        if (!(table[slot] & E_SUB)) {
          if (sub_next + ((size_t)1 << sub_bits) > table_cap) return false;
          table[slot] = E_SUB | (sub_bits << 8) | ((uint32_t)sub_next << 16);
          for (size_t i = 0; i < ((size_t)1 << sub_bits); ++i) table[sub_next + i] = E_BAD | 1u;
          sub_next += (size_t)1 << sub_bits;
        }
        uint32_t* sub = table + (table[slot] >> 16);
        for (size_t i = rev >> primary_bits; i < ((size_t)1 << sub_bits); i += (size_t)1 << (l - primary_bits)) sub[i] = e;
      }
    }
  }
  return true;
}

inline uint32_t litlen_entry(unsigned s) {
  if (s < 256) return E_LIT | (1u << 4) | (s << 8);
  if (s == 256) return E_EOB;
  if (s > 285) return E_BAD;   // 286, 287 take part in the code but never appear in valid data
  return ((uint32_t)LEN_BASE[s - 257] << 16) | ((uint32_t)LEN_EXTRA[s - 257] << 8);
}
inline uint32_t dist_entry(unsigned s) {
  if (s > 29) return E_BAD;
  return ((uint32_t)DIST_BASE[s] << 16) | ((uint32_t)DIST_EXTRA[s] << 8);
}

// Literal runs: where the bits after a literal's code spell one or two more literal codes that still fit the index,
// the slot takes them all.  `single` is the table as built above (its replication makes the entry at index y valid for
// a symbol of up to `avail` known bits exactly when the entry's length is <= avail).
void fuse_literal_runs(const uint32_t* single, uint32_t* table, unsigned bits) {
  const uint32_t n = 1u << bits;
  for (uint32_t x = 0; x < n; ++x) {
    uint32_t e = single[x];
    if (e & E_LIT) {
      const unsigned l1 = e & 15u;
      const uint32_t e2 = single[x >> l1];
      const unsigned l2 = e2 & 15u;
      if ((e2 & E_LIT) && l1 + l2 <= bits) {
        e = E_LIT | (2u << 4) | (l1 + l2) | (e & 0xff00u) | ((e2 & 0xff00u) << 8);
        const uint32_t e3 = single[x >> (l1 + l2)];
        const unsigned l3 = e3 & 15u;
        if ((e3 & E_LIT) && l1 + l2 + l3 <= bits) e = (e & ~0x3fu) | (3u << 4) | (l1 + l2 + l3) | ((e3 & 0xff00u) << 16);
      }
    }
    table[x] = e;
  }
}

}  // namespace

void Inflater::reset(const uint8_t* in, size_t n) {
  in_begin_ = in_next_ = in;
  in_end_ = in + n;
  bitbuf_ = 0;
  bitcnt_ = 0;
  state_ = ST_HEADER;
  final_ = false;
  stored_left_ = 0;
  err_ = "";
}

void Inflater::reset_at(const uint8_t* in, size_t n, uint64_t bit_offset) {
  reset(in, n);
  in_next_ = in + (bit_offset >> 3);
  const unsigned skip = (unsigned)(bit_offset & 7u);
  if (skip && in_next_ < in_end_) {   // the first byte's leading bits belong to the block before
    bitbuf_ = (uint64_t)*in_next_++ >> skip;
    bitcnt_ = 8u - skip;
  }
}

bool Inflater::build_tables(const uint8_t* lens, unsigned n_lit, unsigned n_dist) {
  if (lens[256] == 0) {
    err_ = "invalid code -- missing end-of-block";
    return false;
  }
  if (!build_table(lens, n_lit, lit_, LIT_BITS, sizeof lit_ / sizeof lit_[0], false, litlen_entry)) {
    err_ = "invalid literal/lengths set";
    return false;
  }
  memcpy(single_, lit_, sizeof single_);
  fuse_literal_runs(single_, lit_, LIT_BITS);
  if (!build_table(lens + n_lit, n_dist, dist_, DIST_BITS, sizeof dist_ / sizeof dist_[0], true, dist_entry)) {
    err_ = "invalid distances set";
    return false;
  }
  return true;
}

// Block header, byte-wise refills (once per block).
Inflater::Status Inflater::read_block_header() {
  uint64_t bb = bitbuf_;
  unsigned bc = bitcnt_;
  const uint8_t* in = in_next_;
  auto need = [&](unsigned n) {
    while (bc < n && in < in_end_) {
      bb |= (uint64_t)*in++ << bc;
      bc += 8;
    }
    return bc >= n;
  };
  auto take = [&](unsigned n) {
    const unsigned v = (unsigned)(bb & ((1ull << n) - 1));
    bb >>= n;
    bc -= n;
    return v;
  };
  auto leave = [&](Status s) {
    bitbuf_ = bb;
    bitcnt_ = bc;
    in_next_ = in;
    return s;
  };
  // the branch-free refill of the block loop leaves look-ahead bits above bitcnt: they are the bytes at in_next,
  // and OR-ing those bytes in again is harmless; drop them so that `need` starts from a clean buffer
  bb &= bc >= 64 ? ~0ull : ((1ull << bc) - 1);
  if (!need(3)) return leave(TRUNCATED);
  final_ = take(1) != 0;
  const unsigned type = take(2);
  if (type == 0) {
    take(bc & 7);                 // to the byte boundary
    in -= bc >> 3;                // whole bytes still in the bit buffer go back
    bb = 0;
    bc = 0;
    if (in_end_ - in < 4) return leave(TRUNCATED);
    const unsigned len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
    if ((len ^ 0xffffu) != nlen) {
      err_ = "invalid stored block lengths";
      return leave(BAD_DATA);
    }
    in += 4;
    stored_left_ = len;
    state_ = ST_STORED;
    return leave(DONE);
  }
  uint8_t lens[288 + 32];
  unsigned n_lit, n_dist;
  if (type == 1) {
    n_lit = 288;
    n_dist = 32;
    for (unsigned i = 0; i < 144; ++i) lens[i] = 8;
    for (unsigned i = 144; i < 256; ++i) lens[i] = 9;
    for (unsigned i = 256; i < 280; ++i) lens[i] = 7;
    for (unsigned i = 280; i < 288; ++i) lens[i] = 8;
    for (unsigned i = 0; i < 32; ++i) lens[288 + i] = 5;
  } else if (type == 2) {
    if (!need(14)) return leave(TRUNCATED);
    n_lit = take(5) + 257;
    n_dist = take(5) + 1;
    const unsigned n_clen = take(4) + 4;
    if (n_lit > 286 || n_dist > 30) {
      err_ = "too many length or distance symbols";
      return leave(BAD_DATA);
    }
    uint8_t clens[19] = {0};
    for (unsigned i = 0; i < n_clen; ++i) {
      if (!need(3)) return leave(TRUNCATED);
      clens[CLEN_ORDER[i]] = (uint8_t)take(3);
    }
    uint32_t ctab[128];
    if (!build_table(clens, 19, ctab, 7, 128, false, [](unsigned s) { return (uint32_t)s << 16; })) {
      err_ = "invalid code lengths set";
      return leave(BAD_DATA);
    }
    unsigned i = 0;
    while (i < n_lit + n_dist) {
      need(7 + 7);
      const uint32_t e = ctab[bb & 127];
      if ((e & E_BAD) || (e & 15u) > bc) {
        if (bc < 7 && in >= in_end_) return leave(TRUNCATED);
        err_ = "invalid code lengths set";
        return leave(BAD_DATA);
      }
      take(e & 15u);
      const unsigned sym = e >> 16;
      if (sym < 16) {
        lens[i++] = (uint8_t)sym;
        continue;
      }
      unsigned rep, val = 0;
      const unsigned xb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
      if (bc < xb && !need(xb)) return leave(TRUNCATED);
      if (sym == 16) {
        if (i == 0) {
          err_ = "invalid bit length repeat";
          return leave(BAD_DATA);
        }
        val = lens[i - 1];
        rep = 3 + take(2);
      } else if (sym == 17) {
        rep = 3 + take(3);
      } else {
        rep = 11 + take(7);
      }
      if (i + rep > n_lit + n_dist) {
        err_ = "invalid bit length repeat";
        return leave(BAD_DATA);
      }
      while (rep--) lens[i++] = (uint8_t)val;
    }
  } else {
    err_ = "invalid block type";
    return leave(BAD_DATA);
  }
  if (!build_tables(lens, n_lit, n_dist)) return leave(BAD_DATA);
  state_ = ST_CODES;
  return leave(DONE);
}

// The symbols of one compressed block.  FAST: at least 16 input bytes and 300 output bytes are left (checked per
// symbol group), refills are branch-free 8-byte loads, copies may overshoot by up to 15 bytes.  !FAST: byte-wise
// refills with bounds, exact copies, suspension when fewer than 258 output bytes are left.
template <bool FAST>
inline __attribute__((always_inline)) Inflater::Status Inflater::block_loop(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end) {
  uint64_t bb = bitbuf_;
  unsigned bc = bitcnt_;
  const uint8_t* in = in_next_;
  uint8_t* out = *out_next;
  const uint32_t* lit = lit_;
  const uint32_t* dst = dist_;
  Status st = DONE;
#define MGK_LEAVE(s) \
  do {               \
    st = (s);        \
    goto leave;      \
  } while (0)
  uint32_t e;
  for (;;) {
    if (FAST) {
      if (in_end_ - in < 16 || out_end - out < 300) MGK_LEAVE(NEED_OUTPUT);   // the caller goes on with the careful loop
      bb |= load64(in) << bc;
      in += (63 - bc) >> 3;
      bc |= 56;
    } else {
      if (out_end - out < 258) MGK_LEAVE(NEED_OUTPUT);
      while (bc <= 56 && in < in_end_) {
        bb |= (uint64_t)*in++ << bc;
        bc += 8;
      }
    }
    e = lit[bb & ((1u << LIT_BITS) - 1)];
  have_entry:
    if (FAST && (e & E_LIT)) {   // up to four lookups of literal runs per refill (4 x LIT_BITS <= 56 bits)
      bb >>= e & 15u;
      bc -= e & 15u;
      store32(out, e >> 8);
      out += (e >> 4) & 3u;
      e = lit[bb & ((1u << LIT_BITS) - 1)];
      if (e & E_LIT) {
        bb >>= e & 15u;
        bc -= e & 15u;
        store32(out, e >> 8);
        out += (e >> 4) & 3u;
        e = lit[bb & ((1u << LIT_BITS) - 1)];
        if (e & E_LIT) {
          bb >>= e & 15u;
          bc -= e & 15u;
          store32(out, e >> 8);
          out += (e >> 4) & 3u;
          e = lit[bb & ((1u << LIT_BITS) - 1)];
          if (e & E_LIT) {
            bb >>= e & 15u;
            bc -= e & 15u;
            store32(out, e >> 8);
            out += (e >> 4) & 3u;
            continue;
          }
        }
      }
    }
    if (!(e & E_LIT) && (e & E_SPECIAL)) {   // (a literal run keeps its count where these flags sit)
      if (e & E_SUB) e = lit[(e >> 16) + ((bb >> LIT_BITS) & ((1u << ((e >> 8) & 15u)) - 1))];
      if (!(e & E_LIT) && (e & E_BAD)) {
        if (!FAST && bc < 15 && in >= in_end_) MGK_LEAVE(TRUNCATED);
        err_ = "invalid literal/length code";
        MGK_LEAVE(BAD_DATA);
      }
    }
    if (!FAST && (e & 15u) + ((e & E_LIT) ? 0u : ((e >> 8) & 15u)) > bc) MGK_LEAVE(TRUNCATED);
    bb >>= e & 15u;
    bc -= e & 15u;
    if (e & E_LIT) {
      store32(out, e >> 8);
      out += (e >> 4) & 3u;
      continue;
    }
    if (e & E_EOB) {
      state_ = final_ ? ST_DONE : ST_HEADER;
      MGK_LEAVE(DONE);
    }
    unsigned xb = (e >> 8) & 15u;
    const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << xb) - 1));
    bb >>= xb;
    bc -= xb;
    if (FAST) {
      if (bc < 32) {   // a distance takes up to 15 + 13 bits; a match right after the refill still has them
        bb |= load64(in) << bc;
        in += (63 - bc) >> 3;
        bc |= 56;
      }
    } else {
      while (bc <= 56 && in < in_end_) {
        bb |= (uint64_t)*in++ << bc;
        bc += 8;
      }
    }
    uint32_t d = dst[bb & ((1u << DIST_BITS) - 1)];
    if (d & E_SPECIAL) {
      if (d & E_SUB) d = dst[(d >> 16) + ((bb >> DIST_BITS) & ((1u << ((d >> 8) & 15u)) - 1))];
      if (d & E_BAD) {
        if (!FAST && bc < 15 && in >= in_end_) MGK_LEAVE(TRUNCATED);
        err_ = "invalid distance code";
        MGK_LEAVE(BAD_DATA);
      }
    }
    xb = (d >> 8) & 15u;
    if (!FAST && (d & 15u) + xb > bc) MGK_LEAVE(TRUNCATED);
    bb >>= d & 15u;
    bc -= d & 15u;
    const size_t dist = (d >> 16) + (size_t)(bb & ((1u << xb) - 1));
    bb >>= xb;
    bc -= xb;
    if (dist > (size_t)(out - out_begin)) {
      err_ = "invalid distance too far back";
      MGK_LEAVE(BAD_DATA);
    }
    const uint8_t* src = out - dist;
    uint8_t* end = out + len;
    bool fetched = false;
    if (FAST && in_end_ - in >= 16 && out_end - end >= 300) {   // the next entry's load runs under the copy
      bb |= load64(in) << bc;
      in += (63 - bc) >> 3;
      bc |= 56;
      e = lit[bb & ((1u << LIT_BITS) - 1)];
      fetched = true;
    }
    if (FAST && dist >= 8) {
      memcpy(out, src, 8);
      memcpy(out + 8, src + 8, 8);
      if (len > 16) {
        out += 16;
        src += 16;
        do {
          memcpy(out, src, 8);
          out += 8;
          src += 8;
        } while (out < end);
      }
    } else if (dist == 1) {
      memset(out, *src, len);
    } else {
      while (out < end) *out++ = *src++;
    }
    out = end;
    if (FAST && fetched) goto have_entry;
  }
leave:
#undef MGK_LEAVE
  bitbuf_ = bb;
  bitcnt_ = bc;
  in_next_ = in;
  *out_next = out;
  return st;
}

// The hot loop twice: once for any x86-64, once with BMI2 (shrx / bzhi for the variable shifts and masks of the bit
// buffer), chosen at run time.
Inflater::Status Inflater::fast_generic(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end) {
  return block_loop<true>(out_begin, out_next, out_end);
}
#if defined(__x86_64__)
__attribute__((target("bmi2"))) Inflater::Status Inflater::fast_bmi2(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end) {
  return block_loop<true>(out_begin, out_next, out_end);
}
#endif

Inflater::Status Inflater::run(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end) {
#if defined(__x86_64__)
  static const bool have_bmi2 = __builtin_cpu_supports("bmi2");
#else
  const bool have_bmi2 = false;
#endif
  for (;;) {
    switch (state_) {
      case ST_DONE:
        return DONE;
      case ST_HEADER: {
        if (stop_ && stop_(bit_position())) return STOPPED;
        const Status s = read_block_header();
        if (s != DONE) return s;
        break;
      }
      case ST_STORED: {
        const size_t room = (size_t)(out_end - *out_next), have = (size_t)(in_end_ - in_next_);
        const size_t n = stored_left_ < room ? (stored_left_ < have ? stored_left_ : have) : (room < have ? room : have);
        memcpy(*out_next, in_next_, n);
        *out_next += n;
        in_next_ += n;
        stored_left_ -= n;
        if (stored_left_) return have == n && have < room ? TRUNCATED : NEED_OUTPUT;
        state_ = final_ ? ST_DONE : ST_HEADER;
        break;
      }
      case ST_CODES: {
        Status s = have_bmi2 ? fast_bmi2(out_begin, out_next, out_end) : fast_generic(out_begin, out_next, out_end);
        if (s == NEED_OUTPUT) s = block_loop<false>(out_begin, out_next, out_end);   // near an end of input or output
        if (s != DONE) return s;
        break;
      }
    }
  }
}

// ---------------------------------------------------------------- a stream taken up in its middle
// Same tables, 16-bit symbols; a match that reaches in front of the first decoded byte yields markers.  Every step
// is bounds-checked (this is the speculative side: the position may be no block at all).
Inflater::Status Inflater::run_markers(SymbolBuffer& out) {
  size_t n = out.size;   // symbols so far
  auto leave = [&](Status s) {
    out.size = n;
    return s;
  };
  for (;;) {
    if (state_ == ST_DONE) return leave(DONE);
    if (state_ == ST_HEADER) {
      out_symbols_ = n;
      if (stop_ && stop_(bit_position())) return leave(STOPPED);
      const Status s = read_block_header();
      if (s != DONE) return leave(s);
      continue;
    }
    if (state_ == ST_STORED) {
      const size_t have = (size_t)(in_end_ - in_next_);
      if (have < stored_left_) return leave(TRUNCATED);
      if (out.cap < n + stored_left_) {
        out.size = n;
        out.reserve(n + stored_left_ + (out.cap >> 1));
      }
      for (size_t i = 0; i < stored_left_; ++i) out.data[n + i] = in_next_[i];
      n += stored_left_;
      in_next_ += stored_left_;
      stored_left_ = 0;
      state_ = final_ ? ST_DONE : ST_HEADER;
      continue;
    }
    // ST_CODES
    uint64_t bb = bitbuf_;
    unsigned bc = bitcnt_;
    const uint8_t* in = in_next_;
    Status st = DONE;
    // the bulk of a block: no input bounds to watch (16 bytes ahead cover both refills of a symbol group), literal runs
    // stored as one 64-bit word, matches copied four symbols at a time
    for (;;) {
      if (out.cap - n < 640) {
        out.size = n;
        out.reserve(out.cap + (out.cap >> 1) + ((size_t)4 << 20));
      }
      if (in_end_ - in < 16) break;
      uint16_t* o = out.data;
      bb |= load64(in) << bc;
      in += (63 - bc) >> 3;
      bc |= 56;
      uint32_t e = lit_[bb & ((1u << LIT_BITS) - 1)];
      unsigned runs = 0;
      while ((e & E_LIT) && runs < 3) {   // up to three lookups of literal runs per refill, then a length still fits
        bb >>= e & 15u;
        bc -= e & 15u;
        const uint64_t v = (uint64_t)((e >> 8) & 0xffu) | ((uint64_t)((e >> 16) & 0xffu) << 16) | ((uint64_t)(e >> 24) << 32);
        memcpy(o + n, &v, 8);
        n += (e >> 4) & 3u;
        e = lit_[bb & ((1u << LIT_BITS) - 1)];
        ++runs;
      }
      if (e & E_LIT) continue;   // (fourth run: after the next refill)
      if (e & E_SPECIAL) {
        if (e & E_SUB) e = lit_[(e >> 16) + ((bb >> LIT_BITS) & ((1u << ((e >> 8) & 15u)) - 1))];
        if (e & E_LIT) {
          bb >>= e & 15u;
          bc -= e & 15u;
          o[n++] = (uint16_t)((e >> 8) & 0xffu);
          continue;
        }
        if (e & E_BAD) {
          st = BAD_DATA;
          err_ = "invalid literal/length code";
          break;
        }
        if (e & E_EOB) {
          bb >>= e & 15u;
          bc -= e & 15u;
          state_ = final_ ? ST_DONE : ST_HEADER;
          break;
        }
      }
      bb >>= e & 15u;
      bc -= e & 15u;
      unsigned xb = (e >> 8) & 15u;
      const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << xb) - 1));
      bb >>= xb;
      bc -= xb;
      if (bc < 32) {
        bb |= load64(in) << bc;
        in += (63 - bc) >> 3;
        bc |= 56;
      }
      uint32_t d = dist_[bb & ((1u << DIST_BITS) - 1)];
      if (d & E_SPECIAL) {
        if (d & E_SUB) d = dist_[(d >> 16) + ((bb >> DIST_BITS) & ((1u << ((d >> 8) & 15u)) - 1))];
        if (d & E_BAD) {
          st = BAD_DATA;
          err_ = "invalid distance code";
          break;
        }
      }
      xb = (d >> 8) & 15u;
      bb >>= d & 15u;
      bc -= d & 15u;
      const size_t dist = (d >> 16) + (size_t)(bb & ((1u << xb) - 1));
      bb >>= xb;
      bc -= xb;
      if (dist <= n) {
        const uint16_t* src = o + (n - dist);
        uint16_t* dst = o + n;
        if (dist >= 4) {
          for (unsigned i = 0; i < len; i += 4) memcpy(dst + i, src + i, 8);
        } else {
          for (unsigned i = 0; i < len; ++i) dst[i] = src[i];
        }
      } else {   // (part of) the match lies in the unknown window
        const size_t before = dist - n;
        for (unsigned i = 0; i < len; ++i) {
          const size_t back = before > i ? before - i : 0;
          o[n + i] = back ? (uint16_t)(MARKER | (INFLATE_WINDOW - back)) : o[i - before];
        }
      }
      n += len;
    }
    if (st != DONE || state_ != ST_CODES) {
      bitbuf_ = bb;
      bitcnt_ = bc;
      in_next_ = in;
      if (st != DONE) return leave(st);
      continue;
    }
    // the last bytes of the input: every step checked
    for (;;) {
      if (out.cap - n < 320) {
        out.size = n;
        out.reserve(out.cap + (out.cap >> 1) + ((size_t)4 << 20));
      }
      uint16_t* o = out.data;
      if (in_end_ - in >= 8) {
        bb |= load64(in) << bc;
        in += (63 - bc) >> 3;
        bc |= 56;
      } else {
        while (bc <= 56 && in < in_end_) {
          bb |= (uint64_t)*in++ << bc;
          bc += 8;
        }
      }
      uint32_t e = lit_[bb & ((1u << LIT_BITS) - 1)];
      if (!(e & E_LIT) && (e & E_SPECIAL)) {
        if (e & E_SUB) e = lit_[(e >> 16) + ((bb >> LIT_BITS) & ((1u << ((e >> 8) & 15u)) - 1))];
        if (!(e & E_LIT) && (e & E_BAD)) {
          st = (bc < 15 && in >= in_end_) ? TRUNCATED : BAD_DATA;
          err_ = "invalid literal/length code";
          break;
        }
      }
      if ((e & 15u) + ((e & E_LIT) ? 0u : ((e >> 8) & 15u)) > bc) {
        st = TRUNCATED;
        break;
      }
      bb >>= e & 15u;
      bc -= e & 15u;
      if (e & E_LIT) {
        const unsigned cnt = (e >> 4) & 3u;
        o[n] = (uint16_t)((e >> 8) & 0xffu);
        o[n + 1] = (uint16_t)((e >> 16) & 0xffu);
        o[n + 2] = (uint16_t)(e >> 24);
        n += cnt;
        continue;
      }
      if (e & E_EOB) {
        state_ = final_ ? ST_DONE : ST_HEADER;
        break;
      }
      unsigned xb = (e >> 8) & 15u;
      const unsigned len = (e >> 16) + (unsigned)(bb & ((1u << xb) - 1));
      bb >>= xb;
      bc -= xb;
      if (bc < 32) {
        if (in_end_ - in >= 8) {
          bb |= load64(in) << bc;
          in += (63 - bc) >> 3;
          bc |= 56;
        } else {
          while (bc <= 56 && in < in_end_) {
            bb |= (uint64_t)*in++ << bc;
            bc += 8;
          }
        }
      }
      uint32_t d = dist_[bb & ((1u << DIST_BITS) - 1)];
      if (d & E_SPECIAL) {
        if (d & E_SUB) d = dist_[(d >> 16) + ((bb >> DIST_BITS) & ((1u << ((d >> 8) & 15u)) - 1))];
        if (d & E_BAD) {
          st = (bc < 15 && in >= in_end_) ? TRUNCATED : BAD_DATA;
          err_ = "invalid distance code";
          break;
        }
      }
      xb = (d >> 8) & 15u;
      if ((d & 15u) + xb > bc) {
        st = TRUNCATED;
        break;
      }
      bb >>= d & 15u;
      bc -= d & 15u;
      const size_t dist = (d >> 16) + (size_t)(bb & ((1u << xb) - 1));
      bb >>= xb;
      bc -= xb;
      if (dist <= n) {
        const uint16_t* src = o + (n - dist);
        for (unsigned i = 0; i < len; ++i) o[n + i] = src[i];
      } else {   // (part of) the match lies in the unknown window: byte w = INFLATE_WINDOW - (distance back from the start)
        const size_t before = dist - n;   // 1 .. 32768
        for (unsigned i = 0; i < len; ++i) {
          const size_t back = before > i ? before - i : 0;
          o[n + i] = back ? (uint16_t)(MARKER | (INFLATE_WINDOW - back)) : o[i - before];
        }
      }
      n += len;
    }
    bitbuf_ = bb;
    bitcnt_ = bc;
    in_next_ = in;
    if (st != DONE) return leave(st);
  }
}

SymbolBuffer::~SymbolBuffer() {
  if (base_) munmap(base_, mapped_);
}

void SymbolBuffer::reserve(size_t n) {
  if (n <= cap) return;
  const size_t huge = (size_t)2 << 20;
  const size_t bytes = (n * sizeof(uint16_t) + huge - 1) & ~(huge - 1);
  void* p = mmap(nullptr, bytes + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) throw std::bad_alloc();
  uint16_t* d = (uint16_t*)(((uintptr_t)p + huge - 1) & ~(uintptr_t)(huge - 1));
  madvise(d, bytes, MADV_HUGEPAGE);
  if (size) memcpy(d, data, size * sizeof(uint16_t));
  if (base_) munmap(base_, mapped_);
  base_ = p;
  mapped_ = bytes + huge;
  data = d;
  cap = bytes / sizeof(uint16_t);
}

void Inflater::resolve_markers(const uint16_t* sym, size_t n, const uint8_t* window, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    const uint16_t v = sym[i];
    out[i] = (v & MARKER) ? window[v & (MARKER - 1)] : (uint8_t)v;
  }
}

// Block finder: a cheap test of the first 17 bits and of the code-length code at every bit position, the full header
// (read_block_header: both Huffman codes complete, end-of-block present) for the few that pass.
uint64_t Inflater::find_block(const uint8_t* in, size_t n, uint64_t from_bit, uint64_t to_bit) {
  if (n < 16) return ~0ull;
  const uint64_t last = (uint64_t)(n - 16) * 8u;   // 64-bit loads stay inside the buffer, headers are at most ~2300 bits
  if (to_bit > last) to_bit = last;
  Inflater probe;
  for (uint64_t b = from_bit; b < to_bit; ++b) {
    const uint64_t w = load64(in + (b >> 3)) >> (b & 7u);
    if ((w & 7u) != 4u) continue;                      // BFINAL = 0, BTYPE = 10
    if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) continue;   // HLIT, HDIST
    const unsigned n_clen = (unsigned)((w >> 13) & 15u) + 4u;
    // code-length code: 3 bits each from bit 17 on; complete <=> sum of 2^(7 - len) over the used lengths == 128
    unsigned kraft = 0, used = 0;
    uint64_t bits = w >> 17;
    unsigned have = 64u - 17u - (unsigned)(b & 7u);
    uint64_t at = b + 17u;
    for (unsigned i = 0; i < n_clen; ++i) {
      if (have < 3) {
        bits = load64(in + (at >> 3)) >> (at & 7u);
        have = 64u - (unsigned)(at & 7u);
      }
      const unsigned l = (unsigned)(bits & 7u);
      bits >>= 3;
      have -= 3;
      at += 3;
      if (l) {
        kraft += 128u >> l;
        ++used;
      }
    }
    if (kraft != 128u && !(used == 1 && kraft == 64u)) continue;
    probe.reset_at(in, n, b);
    if (probe.read_block_header() == DONE && probe.state_ == ST_CODES) return b;
  }
  return ~0ull;
}

// ---------------------------------------------------------------- gzip framing
size_t GzipMember::header_length(const uint8_t* p, size_t n) {
  if (n < 10) return (size_t)-1;
  if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xe0)) return 0;
  const unsigned flg = p[3];
  size_t at = 10;
  if (flg & 4) {   // FEXTRA
    if (n < at + 2) return (size_t)-1;
    at += 2 + (p[at] | ((size_t)p[at + 1] << 8));
    if (n < at) return (size_t)-1;
  }
  for (unsigned f = 8; f <= 16; f <<= 1) {   // FNAME, FCOMMENT: zero-terminated
    if (!(flg & f)) continue;
    while (at < n && p[at]) ++at;
    if (at >= n) return (size_t)-1;
    ++at;
  }
  if (flg & 2) at += 2;   // FHCRC
  return at <= n ? at : (size_t)-1;
}

// ---------------------------------------------------------------- CRC-32
namespace {

uint32_t g_crc_table[8][256];
struct CrcInit {
  CrcInit() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xedb88320u & (0u - (c & 1u)));
      g_crc_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xff];
  }
} g_crc_init;

uint32_t crc32_slice8(uint32_t c, const uint8_t* p, size_t n) {   // c is the running (inverted) register
  while (n >= 8) {
    const uint64_t v = load64(p) ^ c;
    c = g_crc_table[7][v & 0xff] ^ g_crc_table[6][(v >> 8) & 0xff] ^ g_crc_table[5][(v >> 16) & 0xff] ^ g_crc_table[4][(v >> 24) & 0xff] ^
        g_crc_table[3][(v >> 32) & 0xff] ^ g_crc_table[2][(v >> 40) & 0xff] ^ g_crc_table[1][(v >> 48) & 0xff] ^ g_crc_table[0][v >> 56];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xff];
  return c;
}

#if defined(__x86_64__)
// Folding by carry-less multiplication (Gopal et al., "Fast CRC computation for generic polynomials using
// PCLMULQDQ"), constants of the reflected gzip polynomial.  n >= 64, n % 16 == 0.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(uint32_t c, const uint8_t* p, size_t n) {
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
  const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124);
  const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
  __m128i x1 = _mm_loadu_si128((const __m128i*)(p + 0)), x2 = _mm_loadu_si128((const __m128i*)(p + 16));
  __m128i x3 = _mm_loadu_si128((const __m128i*)(p + 32)), x4 = _mm_loadu_si128((const __m128i*)(p + 48));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)c));
  p += 64;
  n -= 64;
  while (n >= 64) {
    const __m128i a1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), a2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
    const __m128i a3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), a4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
    x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
    x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
    x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, a1), _mm_loadu_si128((const __m128i*)(p + 0)));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, a2), _mm_loadu_si128((const __m128i*)(p + 16)));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, a3), _mm_loadu_si128((const __m128i*)(p + 32)));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, a4), _mm_loadu_si128((const __m128i*)(p + 48)));
    p += 64;
    n -= 64;
  }
#define MGK_FOLD(next)                                          \
  do {                                                          \
    const __m128i a_ = _mm_clmulepi64_si128(x1, k3k4, 0x00);    \
    x1 = _mm_clmulepi64_si128(x1, k3k4, 0x11);                  \
    x1 = _mm_xor_si128(_mm_xor_si128(x1, (next)), a_);          \
  } while (0)
  MGK_FOLD(x2);
  MGK_FOLD(x3);
  MGK_FOLD(x4);
  while (n >= 16) {
    MGK_FOLD(_mm_loadu_si128((const __m128i*)p));
    p += 16;
    n -= 16;
  }
#undef MGK_FOLD
  const __m128i mask = _mm_setr_epi32(~0, 0, ~0, 0);
  __m128i t = _mm_clmulepi64_si128(x1, k3k4, 0x10);
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
  t = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, mask);
  x1 = _mm_xor_si128(_mm_clmulepi64_si128(x1, k5k0, 0x00), t);
  t = _mm_and_si128(x1, mask);
  t = _mm_clmulepi64_si128(t, poly, 0x10);
  t = _mm_and_si128(t, mask);
  t = _mm_clmulepi64_si128(t, poly, 0x00);
  x1 = _mm_xor_si128(x1, t);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

}  // namespace

namespace {
uint32_t multmodp(uint32_t a, uint32_t b) {   // a * b mod P, reflected (zlib crc32.c)
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
  }
  return p;
}
}  // namespace

uint32_t crc32_combine_fast(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
  static uint32_t x2n[32];
  static const bool init = [] {
    uint32_t p = 1u << 30;   // x^1
    x2n[0] = p;
    for (int n = 1; n < 32; ++n) x2n[n] = p = multmodp(p, p);
    return true;
  }();
  (void)init;
  uint32_t p = 1u << 31;     // x^0
  unsigned k = 3;            // x^(len_b * 2^3)
  for (uint64_t n = len_b; n; n >>= 1, ++k)
    if (n & 1) p = multmodp(x2n[k & 31], p);
  return multmodp(p, crc_a) ^ crc_b;
}

uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n) {
  uint32_t c = ~crc;
#if defined(__x86_64__)
  static const bool have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  if (have_clmul && n >= 64) {
    const size_t body = n & ~(size_t)15;
    c = crc32_clmul(c, p, body);
    p += body;
    n -= body;
  }
#endif
  return ~crc32_slice8(c, p, n);
}

}  // namespace mgikit

// Table-driven inflate (RFC 1951) + gzip member framing (RFC 1952) for the input side of the
// pump.  Replaces flate2::MultiGzDecoder of the reference's reader threads
// (/root/reference/src/file_utils.rs:231-233, :99-164, :239-373, :542-638).
//
// Design: 64-bit bit buffer with a branch-free refill, an 11-bit primary literal/length table
// (literals resolved two or three per refill: FASTQ out of `gzip -1` is literal-heavy), an
// 8-bit primary distance table, subtables for the long codes, word-wide match copies.  A
// stream can be suspended at a symbol boundary when the output buffer is full and resumed
// into the next buffer (the caller carries the last 32 KiB over), so a member of any size
// streams through fixed-size buffers; members of a multi-member file decode in parallel.
#pragma once

#include <cstddef>
#include <cstdint>
#include <functional>
#include <utility>
#include <string>
#include <vector>

namespace mgikit {

constexpr size_t INFLATE_WINDOW = 32768;   // how far a match may reach back

// Growable array of 16-bit symbols on huge pages, never shrunk, meant to be reused: in a VM a fresh 4 KiB page costs
// microseconds and the faults of many threads serialise on the address space.
struct SymbolBuffer {
  uint16_t* data = nullptr;
  size_t size = 0, cap = 0;
  SymbolBuffer() = default;
  SymbolBuffer(const SymbolBuffer&) = delete;
  SymbolBuffer& operator=(const SymbolBuffer&) = delete;
  ~SymbolBuffer();
  void reserve(size_t n);   // keeps the first `size` symbols
  void swap(SymbolBuffer& o) {
    std::swap(data, o.data);
    std::swap(size, o.size);
    std::swap(cap, o.cap);
    std::swap(base_, o.base_);
    std::swap(mapped_, o.mapped_);
  }

 private:
  void* base_ = nullptr;
  size_t mapped_ = 0;
};

class Inflater {
 public:
  enum Status { DONE = 0, NEED_OUTPUT = 1, BAD_DATA = 2, TRUNCATED = 3, STOPPED = 4 };

  // [in, in + n) holds the raw deflate stream (it may be followed by anything); nothing outside is read
  void reset(const uint8_t* in, size_t n);
  // the same, decoding from the block that starts `bit_offset` bits into the stream
  void reset_at(const uint8_t* in, size_t n, uint64_t bit_offset);
  // bits of the stream consumed so far (exact between blocks)
  uint64_t bit_position() const { return (uint64_t)(in_next_ - in_begin_) * 8u - bitcnt_; }
  // Called in front of every block header with bit_position(); returning true ends run() / run_markers() with STOPPED
  // before the header is read (one stream decoded in parallel: a part stops where the next part begins).
  void set_block_stop(std::function<bool(uint64_t)> stop) { stop_ = std::move(stop); }

  // Decoding WITHOUT the 32 KiB that precede the starting block (a part of a stream taken up in its middle): symbols
  // are 16 bit wide, 0..255 a byte, MARKER | w the byte w of that unknown window (w = 32768 - distance back from the
  // start), copied along by every match that reaches into it; resolve_markers() turns them into bytes once the window is
  // known.  Appends to `out`; ends with DONE (final block), STOPPED (block stop) or an error.
  static constexpr uint16_t MARKER = 0x8000;
  Status run_markers(SymbolBuffer& out);
  size_t symbols_so_far() const { return out_symbols_; }   // for a block stop that bounds the output
  static void resolve_markers(const uint16_t* sym, size_t n, const uint8_t* window /* INFLATE_WINDOW bytes */, uint8_t* out);

  // First bit position in [from_bit, to_bit) where a non-final dynamic-Huffman block header parses and both of its
  // codes are complete (what a decoder taking the stream up there needs), or ~0.  Candidates, not certainties: the
  // caller confirms a position by reaching it at a block boundary from a position it trusts.
  static uint64_t find_block(const uint8_t* in, size_t n, uint64_t from_bit, uint64_t to_bit);

  // Decodes into [*out_next, out_end).  [out_begin, *out_next) must hold what the stream produced before
  // (all of it, or at least its last INFLATE_WINDOW bytes).  NEED_OUTPUT: fewer than 258 bytes are left in
  // the buffer at a symbol boundary; call again with a new buffer.
  Status run(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end);
  // bytes of input consumed up to the end of the final block (valid after DONE)
  size_t consumed() const { return (size_t)(in_next_ - in_begin_) - (bitcnt_ >> 3); }
  const char* error() const { return err_; }

 private:
  template <bool FAST>
  Status block_loop(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end);
  Status fast_generic(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end);
  Status fast_bmi2(uint8_t* out_begin, uint8_t** out_next, uint8_t* out_end);
  Status read_block_header();
  bool build_tables(const uint8_t* lens, unsigned n_lit, unsigned n_dist);

  const uint8_t *in_begin_ = nullptr, *in_next_ = nullptr, *in_end_ = nullptr;
  uint64_t bitbuf_ = 0;
  unsigned bitcnt_ = 0;
  enum { ST_HEADER, ST_STORED, ST_CODES, ST_DONE } state_ = ST_HEADER;
  bool final_ = false;
  size_t stored_left_ = 0;
  const char* err_ = "";
  std::function<bool(uint64_t)> stop_;
  size_t out_symbols_ = 0;   // run_markers: symbols decoded when the block stop is asked
  // decode tables: entry layout in inflate.cpp
  static constexpr unsigned LIT_BITS = 12, DIST_BITS = 8;
  uint32_t lit_[(1u << LIT_BITS) + 1152];    // primary + subtables (8 slots per long-code prefix, at most 144 prefixes)
  uint32_t single_[1u << LIT_BITS];          // the primary table before literal runs were fused into it
  uint32_t dist_[(1u << DIST_BITS) + 2048];  // 128 slots per long-code prefix, at most 15 prefixes
};

uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n);
// CRC-32 of A||B from the CRC-32s of A and B and the length of B (zlib's crc32_combine)
uint32_t crc32_combine_fast(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);   // gzip CRC-32 (PCLMULQDQ folding when the CPU has it)

// One gzip member: header parsing, the deflate stream, CRC-32 + ISIZE check.
struct GzipMember {
  // Parses the header at [p, p + n); returns its length, 0 if this is not a gzip (deflate) member header,
  // or (size_t)-1 if it is cut off.
  static size_t header_length(const uint8_t* p, size_t n);
  // magic + method + reserved flag bits + an XFL value RFC 1952 defines: cheap test of a candidate member start (the
  // reader does not depend on it: a member the chain arrives at is taken on header_length() alone)
  static bool plausible(const uint8_t* p, size_t n) {
    return n >= 18 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 0xe0) == 0 && (p[8] == 0 || p[8] == 2 || p[8] == 4);
  }
};

}  // namespace mgikit

// TEST INFRASTRUCTURE: C entry points over the product's host codec (mgikit_b200/csrc/host/inflate.cpp,
// gz_reader.cpp) so that pytest can hold it to zlib on fixtures and fuzzed streams.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../mgikit_b200/csrc/host/inflate.hpp"

using namespace mgikit;

extern "C" {

uint32_t hc_crc32(uint32_t crc, const uint8_t* p, size_t n) { return crc32_fast(crc, p, n); }
uint32_t hc_crc32_combine(uint32_t a, uint32_t b, uint64_t len_b) { return crc32_combine_fast(a, b, len_b); }

// Raw deflate stream -> out, through output buffers of `piece` bytes (0: one buffer), the way the reader streams a
// member.  Returns the Inflater status; *produced / *consumed as far as decoding got.
int hc_inflate_raw(const uint8_t* in, size_t n, uint8_t* out, size_t cap, size_t piece, size_t* produced, size_t* consumed) {
  Inflater inf;
  inf.reset(in, n);
  *produced = 0;
  if (piece == 0) {
    uint8_t* next = out;
    const Inflater::Status s = inf.run(out, &next, out + cap);
    *produced = (size_t)(next - out);
    *consumed = s == Inflater::DONE ? inf.consumed() : 0;
    return (int)s;
  }
  // pieces: [window | payload], the last 32 KiB carried to the front of the next buffer
  std::vector<uint8_t> buf(INFLATE_WINDOW + piece + 300);
  size_t hist = 0;
  for (;;) {
    uint8_t* begin = buf.data() + INFLATE_WINDOW - hist;
    uint8_t* next = buf.data() + INFLATE_WINDOW;
    const Inflater::Status s = inf.run(begin, &next, buf.data() + INFLATE_WINDOW + piece + 300);
    const size_t got = (size_t)(next - (buf.data() + INFLATE_WINDOW));
    if (*produced + got > cap) return -1;
    memcpy(out + *produced, buf.data() + INFLATE_WINDOW, got);
    *produced += got;
    if (s != Inflater::NEED_OUTPUT) {
      *consumed = s == Inflater::DONE ? inf.consumed() : 0;
      return (int)s;
    }
    // keep the last INFLATE_WINDOW bytes of [begin, next) in front of the payload area
    const size_t have = (size_t)(next - begin), keep = have < INFLATE_WINDOW ? have : INFLATE_WINDOW;
    memmove(buf.data() + INFLATE_WINDOW - keep, next - keep, keep);
    hist = keep;
  }
}

size_t hc_gzip_header_length(const uint8_t* p, size_t n) { return GzipMember::header_length(p, n); }
}

// ---- the member-parallel file reader (gz_reader.cpp)
#include "../../mgikit_b200/csrc/host/gz_reader.hpp"
#include "../../mgikit_b200/csrc/host/sample_sheet.hpp"

extern "C" {
// Whole file -> out (at most cap bytes).  0 = ok; 1 = the reader refused the file (message in err); 2 = cap too small.
int hc_gunzip_file2(const char* path, int threads, size_t max_buffered, size_t parallel_min, size_t part_bytes, uint8_t* out, size_t cap,
                    size_t* produced, size_t* newlines, size_t* members, char* err, size_t err_cap);
int hc_gunzip_file(const char* path, int threads, size_t max_buffered, uint8_t* out, size_t cap, size_t* produced, size_t* newlines,
                   size_t* members, char* err, size_t err_cap) {
  return hc_gunzip_file2(path, threads, max_buffered, (size_t)48 << 20, (size_t)4 << 20, out, cap, produced, newlines, members, err, err_cap);
}
int hc_gunzip_file2(const char* path, int threads, size_t max_buffered, size_t parallel_min, size_t part_bytes, uint8_t* out, size_t cap,
                    size_t* produced, size_t* newlines, size_t* members, char* err, size_t err_cap) {
  *produced = *newlines = *members = 0;
  try {
    GzReader rd(path, threads, max_buffered, parallel_min, part_bytes);
    GzPiece p;
    while (rd.next(p)) {
      if (*produced + p.len > cap) return 2;
      memcpy(out + *produced, p.data, p.len);
      *produced += p.len;
      *newlines += p.newlines;
      rd.recycle(p);
    }
    *members = rd.members_decoded();
    return 0;
  } catch (const std::exception& e) {
    snprintf(err, err_cap, "%s", e.what());
    return 1;
  }
}
}

// Streaming, member-parallel gunzip of one input file: the decoded text
// comes out in file order as pieces of a few MiB, each with its newline count, while the
// members of a multi-member file (bgzip, pigz -i, concatenated lanes, mgikit's own output)
// inflate on several threads at once.  Replaces flate2::MultiGzDecoder behind the
// reference's reader threads (/root/reference/src/file_utils.rs:231-233, :239-373, :542-638).
//
// How: the file is mapped; a scanner lists every offset that looks like a member start
// (magic, method, reserved flags: GzipMember::plausible); worker threads inflate candidates
// speculatively, in offset order, into buffers of their own; the consumer walks the chain
// member by member -- the member at offset p is taken only when its predecessor ended at
// exactly p with a matching CRC-32 and ISIZE -- so a false candidate (the magic inside
// compressed data) is never reached and costs a failed attempt of a few bytes.  A
// single-member file degenerates to one streaming task: same code, no parallelism.
#pragma once

#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mgikit {

struct GzPiece {
  const uint8_t* data = nullptr;
  size_t len = 0;
  size_t newlines = 0;     // '\n' bytes in [data, data + len)
  void* owner = nullptr;   // buffer to give back with GzReader::recycle
};

class GzReader {
 public:
  // n_threads: inflate workers of this file; max_buffered: bytes of decoded text that may wait for the consumer
  // parallel_min / part_bytes: a member whose compressed stream is at least parallel_min long is decoded in parts of
  // part_bytes by all the workers at once (gz_reader.cpp: decode_parallel)
  GzReader(const std::string& path, int n_threads, size_t max_buffered = (size_t)2 << 30, size_t parallel_min = (size_t)48 << 20,
           size_t part_bytes = (size_t)4 << 20);
  ~GzReader();
  GzReader(const GzReader&) = delete;
  GzReader& operator=(const GzReader&) = delete;

  // Next piece in stream order; false at the end of the stream.  Throws Fatal on a corrupt or truncated file.
  bool next(GzPiece& p);
  void recycle(GzPiece& p);
  size_t members_decoded() const { return members_; }

 private:
  struct Buf;
  struct Task;
  struct Progress;
  void scan();
  void work();
  Buf* take_buffer(Task* t);
  void decode(Task* t);
  void decode_serial(Task* t, size_t hl, Progress& pg);
  bool decode_parallel(Task* t, size_t hl, size_t extent, Progress& pg);
  void finish_member(Task* t, size_t used, const Progress& pg);
  void fail_task(Task* t, const std::string& why);
  std::mutex parallel_mu_;   // one member at a time is decoded in parts (its helpers are the reader's thread budget)

  std::string path_;
  int n_threads_ = 1;
  size_t PARALLEL_MIN_BYTES;   // compressed bytes from which a member is decoded in parts
  size_t PART_BYTES;
  const uint8_t* map_ = nullptr;
  size_t size_ = 0;
  int fd_ = -1;
  size_t piece_ = ((size_t)4 << 20) - 32768 - 512;   // with its window and slack a buffer is two huge pages
  size_t max_bufs_, per_task_bufs_;

  std::mutex mu_;
  std::condition_variable cv_;          // any state change
  std::deque<std::unique_ptr<Task>> tasks_;   // candidates in offset order; the front ones are dropped as the chain passes
  size_t next_claim_ = 0;               // index into tasks_ of the next candidate to start
  size_t scan_pos_ = 0;                 // the scanner has listed every candidate below this offset
  bool scan_done_ = false;
  size_t chain_pos_ = 0;                // offset of the member being consumed (or to be consumed next)
  Task* head_ = nullptr;
  std::vector<Buf*> free_;
  size_t allocated_ = 0;
  bool stop_ = false;
  size_t members_ = 0;
  std::vector<std::thread> threads_;
};

}  // namespace mgikit

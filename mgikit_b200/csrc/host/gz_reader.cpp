#include "gz_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "inflate.hpp"
#include "sample_sheet.hpp"   // Fatal

namespace mgikit {

namespace {

constexpr size_t RESERVE_BUFS = 2;          // buffers only the member being consumed may take
constexpr size_t SCAN_AHEAD = 8192;         // candidates listed ahead of the workers
constexpr size_t TAIL_SLACK = 320;          // the fast loop of the inflater may overshoot a match copy

size_t count_newlines(const uint8_t* p, size_t n) {
  size_t cnt = 0, i = 0;
#if defined(__x86_64__)
  const __m128i nl = _mm_set1_epi8('\n');
  for (; i + 64 <= n; i += 64) {
    const unsigned a = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i)), nl));
    const unsigned b = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 16)), nl));
    const unsigned c = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 32)), nl));
    const unsigned d = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 48)), nl));
    cnt += (size_t)__builtin_popcountll((uint64_t)a | ((uint64_t)b << 16) | ((uint64_t)c << 32) | ((uint64_t)d << 48));
  }
#endif
  for (; i < n; ++i) cnt += p[i] == '\n';
  return cnt;
}

}  // namespace

struct GzReader::Buf {
  uint8_t* mem = nullptr;   // [INFLATE_WINDOW history | piece_ payload | slack], huge-page backed where the kernel allows
  size_t bytes = 0;
  bool pooled = false;      // one of the reader's circulating piece buffers (else: a part's own text, freed when recycled)
  explicit Buf(size_t n) {
    bytes = (n + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    void* p = mmap(nullptr, bytes + ((size_t)2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) throw Fatal("out of memory for the gunzip buffers");
    base_ = p;
    mem = (uint8_t*)(((uintptr_t)p + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1));
    madvise(mem, bytes, MADV_HUGEPAGE);   // 2 MiB pages: a thousand times fewer faults on first touch and frees at exit
  }
  ~Buf() { munmap(base_, bytes + ((size_t)2 << 20)); }
  Buf(const Buf&) = delete;
  Buf& operator=(const Buf&) = delete;

 private:
  void* base_ = nullptr;
};

struct GzReader::Task {
  size_t start = 0;
  enum State { PENDING, RUNNING, OK, FAILED } state = PENDING;
  size_t end = 0;                 // offset after the member's trailer (OK)
  std::string error;
  std::deque<GzPiece> ready;      // decoded, not yet consumed
  bool cancel = false;            // the chain has passed this offset: a false candidate
};

GzReader::GzReader(const std::string& path, int n_threads, size_t max_buffered, size_t parallel_min, size_t part_bytes)
    : path_(path), n_threads_(n_threads < 1 ? 1 : n_threads), PARALLEL_MIN_BYTES(parallel_min), PART_BYTES(part_bytes < 65536 ? 65536 : part_bytes) {
  fd_ = open(path.c_str(), O_RDONLY);
  if (fd_ < 0) throw Fatal("Could not open the file");
  struct stat st;
  if (fstat(fd_, &st) != 0) {
    close(fd_);
    throw Fatal("Could not open the file");
  }
  size_ = (size_t)st.st_size;
  if (size_) {
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) {
      close(fd_);
      throw Fatal("Could not map the file: " + path);
    }
    map_ = (const uint8_t*)m;
    madvise(m, size_, MADV_SEQUENTIAL);
  }
  if (n_threads < 1) n_threads = 1;
  max_bufs_ = max_buffered / piece_;
  if (max_bufs_ < (size_t)n_threads * 2 + RESERVE_BUFS + 2) max_bufs_ = (size_t)n_threads * 2 + RESERVE_BUFS + 2;
  per_task_bufs_ = (max_bufs_ - RESERVE_BUFS) / (size_t)n_threads;
  if (size_) {
    threads_.emplace_back([this] { scan(); });
    for (int i = 0; i < n_threads; ++i) threads_.emplace_back([this] { work(); });
  } else {
    scan_done_ = true;
  }
}

GzReader::~GzReader() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& t : threads_) t.join();
  for (Buf* b : free_) delete b;
  for (auto& t : tasks_)
    for (auto& p : t->ready) delete (Buf*)p.owner;
  if (map_) munmap((void*)map_, size_);
  if (fd_ >= 0) close(fd_);
}

// Lists every plausible member start, in offset order, a bounded distance ahead of the workers.
void GzReader::scan() {
  size_t pos = 0;
  std::vector<size_t> found;
  while (pos < size_) {
    const size_t win_end = std::min(size_, pos + ((size_t)8 << 20));
    found.clear();
    const uint8_t* p = map_ + pos;
    const uint8_t* e = map_ + win_end;
    while (p < e) {
      p = (const uint8_t*)memchr(p, 0x1f, (size_t)(e - p));
      if (!p) break;
      if (GzipMember::plausible(p, (size_t)(map_ + size_ - p))) found.push_back((size_t)(p - map_));
      ++p;
    }
    std::unique_lock<std::mutex> lk(mu_);
    for (size_t s : found) {
      std::unique_ptr<Task> t(new Task());
      t->start = s;
      tasks_.push_back(std::move(t));
    }
    scan_pos_ = win_end;
    pos = win_end;
    cv_.notify_all();
    cv_.wait(lk, [this] { return stop_ || tasks_.size() - std::min(next_claim_, tasks_.size()) < SCAN_AHEAD; });
    if (stop_) return;
  }
  std::lock_guard<std::mutex> lk(mu_);
  scan_done_ = true;
  cv_.notify_all();
}

namespace {
// Inflating is bulk work that starts before the device contexts exist: it yields to the thread that sets those up
// (driver initialisation is CPU-bound and takes two to three times longer on a machine whose cores are all busy).
void lower_thread_priority() {
  setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10);
}
}  // namespace

void GzReader::work() {
  lower_thread_priority();
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_.wait(lk, [this] { return stop_ || next_claim_ < tasks_.size(); });
    if (stop_) return;
    Task* t = tasks_[next_claim_++].get();
    if (t->state != Task::PENDING) continue;
    if (t->start < chain_pos_ || t->cancel) {   // the chain is already past it
      t->state = Task::FAILED;
      continue;
    }
    t->state = Task::RUNNING;
    lk.unlock();
    decode(t);
    lk.lock();
    cv_.notify_all();
  }
}

// A buffer for task t, or null when the task is cancelled / the reader stops.  The member being consumed always gets
// one (the consumer gives them back as it goes); a speculative member only while it stays inside its share.
GzReader::Buf* GzReader::take_buffer(Task* t) {
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    if (stop_ || t->cancel) return nullptr;
    const size_t avail = free_.size() + (max_bufs_ - allocated_);
    const bool head = t->start == chain_pos_;
    if (avail > 0 && (head || (t->ready.size() < per_task_bufs_ && avail > RESERVE_BUFS))) break;
    cv_.wait(lk);
  }
  if (!free_.empty()) {
    Buf* b = free_.back();
    free_.pop_back();
    return b;
  }
  ++allocated_;
  lk.unlock();
  Buf* b = new Buf(INFLATE_WINDOW + piece_ + TAIL_SLACK);
  b->pooled = true;
  return b;
}

// Where a member's decoding stands: everything in front of `bit` is decoded, checked and queued.
struct GzReader::Progress {
  uint64_t bit = 0;                      // position in the deflate stream (a block boundary)
  uint32_t crc = 0;
  uint64_t total = 0;                    // bytes decoded so far
  size_t hist = 0;                       // valid bytes at the end of `window`
  std::unique_ptr<uint8_t[]> window;     // the last INFLATE_WINDOW bytes decoded (right-aligned)
  Progress() : window(new uint8_t[INFLATE_WINDOW]) {}
  void slide(const uint8_t* data, size_t n) {   // the window after n more bytes
    if (n >= INFLATE_WINDOW) {
      memcpy(window.get(), data + n - INFLATE_WINDOW, INFLATE_WINDOW);
      hist = INFLATE_WINDOW;
    } else {
      const size_t keep = std::min(hist, INFLATE_WINDOW - n);
      memmove(window.get() + INFLATE_WINDOW - n - keep, window.get() + INFLATE_WINDOW - keep, keep);
      memcpy(window.get() + INFLATE_WINDOW - n, data, n);
      hist = keep + n;
    }
  }
};

void GzReader::fail_task(Task* t, const std::string& why) {
  std::lock_guard<std::mutex> lk(mu_);
  t->state = Task::FAILED;
  t->error = why;
}

// Trailer of the member whose deflate stream ended `used` bytes after the member's start.
void GzReader::finish_member(Task* t, size_t used, const Progress& pg) {
  const uint8_t* in = map_ + t->start;
  const size_t left = size_ - t->start;
  std::string why;
  if (left < used + 8) {
    why = "unexpected end of file";
  } else {
    const uint8_t* tr = in + used;
    const uint32_t want_crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
    const uint32_t want_len = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24);
    if (want_crc != pg.crc) why = "incorrect data check";
    else if (want_len != (uint32_t)pg.total) why = "incorrect length check";
  }
  std::lock_guard<std::mutex> lk(mu_);
  if (!why.empty()) {
    t->state = Task::FAILED;
    t->error = why;
    return;
  }
  t->end = t->start + used + 8;
  t->state = Task::OK;
}

void GzReader::decode(Task* t) {
  const uint8_t* in = map_ + t->start;
  const size_t left = size_ - t->start;
  const size_t hl = GzipMember::header_length(in, left);
  if (hl == 0 || hl == (size_t)-1) return fail_task(t, hl ? "unexpected end of file" : "not a gzip member");
  Progress pg;
  // a member that runs on for a long way is taken up at several places at once
  if (n_threads_ > 1 && left - hl >= PARALLEL_MIN_BYTES) {
    size_t extent_end = size_;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || t->cancel || scan_done_ || scan_pos_ >= t->start + hl + PARALLEL_MIN_BYTES; });
      if (stop_ || t->cancel) {
        t->state = Task::FAILED;
        t->error = "cancelled";
        return;
      }
      for (auto& other : tasks_)
        if (other->start > t->start) {
          extent_end = other->start;
          break;
        }
    }
    // (the next listed start decides whether this member is long; the parts themselves run to the end of the file:
    // a listed start may be no member, and a member that does end there ends its parts with its final block)
    if (extent_end - (t->start + hl) >= PARALLEL_MIN_BYTES && parallel_mu_.try_lock()) {
      const bool finished = decode_parallel(t, hl, left - hl, pg);
      parallel_mu_.unlock();
      if (finished) return;   // queued to its end (or failed for good)
    }
  }
  decode_serial(t, hl, pg);
}

// One thread, block after block from pg.bit on, through buffers of piece_ bytes with the window carried along.
void GzReader::decode_serial(Task* t, size_t hl, Progress& pg) {
  const uint8_t* in = map_ + t->start;
  const size_t left = size_ - t->start;
  Inflater inf;
  inf.reset_at(in + hl, left - hl, pg.bit);
  for (;;) {
    Buf* b = take_buffer(t);
    if (!b) return fail_task(t, "cancelled");
    uint8_t* payload = b->mem + INFLATE_WINDOW;
    if (pg.hist) memcpy(payload - pg.hist, pg.window.get() + INFLATE_WINDOW - pg.hist, pg.hist);
    uint8_t* next = payload;
    const Inflater::Status st = inf.run(payload - pg.hist, &next, payload + piece_ + TAIL_SLACK);
    const size_t got = (size_t)(next - payload);
    if (st == Inflater::BAD_DATA || st == Inflater::TRUNCATED) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(b);
      }
      return fail_task(t, st == Inflater::TRUNCATED ? "unexpected end of file" : inf.error());
    }
    pg.crc = crc32_fast(pg.crc, payload, got);
    pg.total += got;
    if (st == Inflater::NEED_OUTPUT) pg.slide(payload, got);   // the last 32 KiB go along into the next buffer
    GzPiece piece;
    piece.data = payload;
    piece.len = got;
    piece.newlines = count_newlines(payload, got);
    piece.owner = b;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (got) t->ready.push_back(piece);
      else free_.push_back(b);
    }
    cv_.notify_all();
    if (st == Inflater::DONE) return finish_member(t, hl + inf.consumed(), pg);
  }
}

// One member on several threads.  The deflate stream is cut into parts of PART_BYTES; a helper takes a part, finds the
// first block that starts in it (Inflater::find_block) and decodes from there WITHOUT the 32 KiB in front of it: bytes
// it cannot know yet are carried as markers (Inflater::run_markers) until it reaches, at a block boundary, the exact
// bit where a later part was taken up.  A part counts once its predecessor has ended exactly on its first bit (a wrong
// guess of the finder is never reached and costs nothing but its helper's time): the predecessor then hands it the
// window -- only its own last 32 KiB have to be bytes for that -- and the helper turns its markers into bytes, CRC and
// newline count included, while the parts behind it do the same.  This thread walks the chain in order, joins the
// CRCs (crc32_combine) and queues the text.  Returns true when the member is finished (t->state set); false leaves
// `pg` at a confirmed block boundary for decode_serial.
bool GzReader::decode_parallel(Task* t, size_t hl, size_t extent, Progress& pg) {
  const uint8_t* deflate = map_ + t->start + hl;
  const size_t avail = size_ - t->start - hl;
  const size_t n_parts = (extent + PART_BYTES - 1) / PART_BYTES;
  const uint64_t NONE = ~0ull;
  const size_t NO_PART = (size_t)-1;
  struct Part {
    uint64_t bit = ~0ull;        // first bit of the block it was taken up at (~0: none found)
    bool placed = false;         // `bit` is final
    bool decoded = false;
    Inflater::Status st = Inflater::BAD_DATA;
    uint64_t end_bit = 0;        // where its decoding stopped
    size_t next = (size_t)-1;    // the part that was taken up there
    size_t consumed = 0;         // bytes of the stream up to the end of the final block (st == DONE)
    SymbolBuffer sym;
    // set by the predecessor (part 0: by the walker)
    bool confirmed = false;
    std::unique_ptr<uint8_t[]> window;   // the INFLATE_WINDOW bytes in front of it, right-aligned
    size_t window_valid = 0;
    // set by its helper once the window is there
    bool resolved = false, bad_marker = false;
    Buf* text = nullptr;         // the part's bytes (a buffer of its own, outside the pool)
    size_t text_len = 0, newlines = 0;
    uint32_t crc = 0;
  };
  std::vector<Part> parts(n_parts);
  static const bool trace = getenv("MGK_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  std::mutex mu;
  std::condition_variable cv;
  size_t next_find = 0, next_part = 0, walker_at = 0;   // parts placed / decoding started up to here; the walker's part
  bool quit = false;
  std::vector<std::unique_ptr<SymbolBuffer>> spare;
  const size_t n_helpers = (size_t)n_threads_;
  const size_t ahead = 2 * n_helpers + 2;

  // Two kinds of work for a helper, the cheap one first: placing parts (the finder, well under a millisecond each) runs
  // twice as far ahead as decoding, so that a decoder always finds the parts behind its own placed when it gets there.
  auto helper = [&] {
    lower_thread_priority();
    for (;;) {
      size_t i;
      bool find;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] {
          return quit || (next_find < n_parts && next_find < walker_at + 2 * ahead) ||
                 (next_part < next_find && next_part < walker_at + ahead && parts[next_part].placed);
        });
        if (quit) return;
        find = next_find < n_parts && next_find < walker_at + 2 * ahead;
        i = find ? next_find++ : next_part++;
      }
      Part& P = parts[i];
      const auto tf = std::chrono::steady_clock::now();
      if (find) {
        const uint64_t b = i == 0 ? pg.bit : Inflater::find_block(deflate, avail, (uint64_t)i * PART_BYTES * 8u, (uint64_t)(i + 1) * PART_BYTES * 8u);
        {
          std::lock_guard<std::mutex> lk(mu);
          P.bit = b;
          P.placed = true;
        }
        cv.notify_all();
        continue;
      }
      const uint64_t bit = P.bit;
      if (bit == NONE) continue;
      {   // symbol buffers go round (allocating and unmapping tens of megabytes per part from many threads at once
          // costs more than the decoding: page faults and TLB shoot-downs serialise on the address space)
        std::lock_guard<std::mutex> lk(mu);
        if (!spare.empty()) {
          P.sym.swap(*spare.back());
          spare.pop_back();
        }
      }
      P.sym.size = 0;
      P.sym.reserve(PART_BYTES * 3);
      Inflater inf;
      inf.reset_at(deflate, avail, bit);
      Inflater* self = &inf;
      size_t landed = NO_PART;
      inf.set_block_stop([&, i, self](uint64_t pos) {
        // stop on the first bit of any later part that was taken up exactly here
        const size_t j = std::min<size_t>((size_t)(pos / (PART_BYTES * 8u)), n_parts - 1);
        std::unique_lock<std::mutex> lk(mu);
        if (quit) return true;
        for (size_t k = i + 1; k <= j; ++k) {
          if (k >= next_find) break;   // nobody has looked at that part yet (this one ran far ahead through parts
                                       // without a block to hold on to): it is simply run through
          cv.wait(lk, [&] { return quit || parts[k].placed; });
          if (quit) return true;
          if (parts[k].bit == pos) {
            landed = k;
            return true;
          }
        }
        return self->symbols_so_far() > ((size_t)1 << 30);   // a helper on a false start must not run away
      });
      const auto td = std::chrono::steady_clock::now();
      const Inflater::Status st = inf.run_markers(P.sym);
      if (trace)
        fprintf(stderr, "[gz trace] part %zu: found +%.1f ms at %.3f s, decoded %zu symbols in %.1f ms (status %d)\n", i,
                std::chrono::duration<double, std::milli>(td - tf).count(), std::chrono::duration<double>(td - t_begin).count(), P.sym.size,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td).count(), (int)st);
      {
        std::unique_lock<std::mutex> lk(mu);
        P.st = st;
        P.end_bit = inf.bit_position();
        P.next = landed;
        if (st == Inflater::DONE) P.consumed = inf.consumed();
        P.decoded = true;
        cv.notify_all();
        // the window in front of this part: from the part that ends on its first bit
        cv.wait(lk, [&] { return quit || P.confirmed || walker_at > i; });
        if (quit) return;
        if (!P.confirmed) {   // the chain went past it: a false start of the finder
          spare.emplace_back(new SymbolBuffer());
          spare.back()->swap(P.sym);
          continue;
        }
      }
      const bool usable = (st == Inflater::STOPPED && landed != NO_PART) || st == Inflater::DONE;
      if (usable) {
        // its own last 32 KiB first: that is all the next part waits for
        const size_t n = P.sym.size, tail = std::min(n, INFLATE_WINDOW);
        if (landed != NO_PART) {
          Part& N = parts[landed];
          std::unique_ptr<uint8_t[]> w(new uint8_t[INFLATE_WINDOW]);
          size_t valid = std::min(INFLATE_WINDOW, P.window_valid + n);
          if (tail < INFLATE_WINDOW) memcpy(w.get(), P.window.get() + tail, INFLATE_WINDOW - tail);   // what stays of the old window
          Inflater::resolve_markers(P.sym.data + n - tail, tail, P.window.get(), w.get() + INFLATE_WINDOW - tail);
          std::lock_guard<std::mutex> lk(mu);
          N.window = std::move(w);
          N.window_valid = valid;
          N.confirmed = true;
          cv.notify_all();
        }
        // then all of it, into a buffer of its own
        if (P.window_valid < INFLATE_WINDOW)   // a marker in front of the stream's first byte: "distance too far back"
          for (size_t q = 0; q < n; ++q) {
            const uint16_t v = P.sym.data[q];
            if ((v & Inflater::MARKER) && (size_t)(v & (Inflater::MARKER - 1)) < INFLATE_WINDOW - P.window_valid) {
              P.bad_marker = true;
              break;
            }
          }
        if (!P.bad_marker) {
          P.text = new Buf(n + 64);
          Inflater::resolve_markers(P.sym.data, n, P.window.get(), P.text->mem);
          P.text_len = n;
          P.crc = crc32_fast(0, P.text->mem, n);
          P.newlines = count_newlines(P.text->mem, n);
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        spare.emplace_back(new SymbolBuffer());
        spare.back()->swap(P.sym);
        P.resolved = true;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> helpers;
  for (size_t h = 0; h < n_helpers; ++h) helpers.emplace_back(helper);
  auto stop_helpers = [&] {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
    }
    cv.notify_all();
    for (auto& th : helpers) th.join();
    for (auto& P : parts)
      if (P.text) delete P.text;
  };
  {   // part 0 starts where the member stands, with what is known of the window
    std::lock_guard<std::mutex> lk(mu);
    parts[0].window.reset(new uint8_t[INFLATE_WINDOW]);
    memcpy(parts[0].window.get(), pg.window.get(), INFLATE_WINDOW);
    parts[0].window_valid = pg.hist;
    parts[0].confirmed = true;
  }

  bool cancelled = false;
  size_t cur = 0;
  for (;;) {
    Part& P = parts[cur];
    {
      std::unique_lock<std::mutex> lk(mu);
      walker_at = cur;
      cv.notify_all();
      cv.wait(lk, [&] { return P.resolved; });
    }
    if (trace) fprintf(stderr, "[gz trace] walker has part %zu at %.3f s\n", cur, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    const bool usable = ((P.st == Inflater::STOPPED && P.next != NO_PART) || P.st == Inflater::DONE) && !P.bad_marker;
    if (!usable) {   // serial from this part's first bit on says what is wrong with the stream
      memcpy(pg.window.get(), P.window.get(), INFLATE_WINDOW);
      pg.hist = P.window_valid;
      pg.bit = P.bit;
      break;
    }
    {   // the consumer's budget holds for this member too: wait while it is behind
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return stop_ || t->cancel || t->start == chain_pos_ || t->ready.size() < per_task_bufs_; });
      if (stop_ || t->cancel) cancelled = true;
    }
    if (cancelled) break;
    pg.crc = crc32_combine_fast(pg.crc, P.crc, P.text_len);
    pg.total += P.text_len;
    pg.bit = P.end_bit;
    if (P.text_len) {
      GzPiece piece;
      piece.data = P.text->mem;
      piece.len = P.text_len;
      piece.newlines = P.newlines;
      piece.owner = P.text;
      P.text = nullptr;
      std::lock_guard<std::mutex> lk(mu_);
      t->ready.push_back(piece);
    } else if (P.text) {
      delete P.text;
      P.text = nullptr;
    }
    cv_.notify_all();
    if (P.st == Inflater::DONE) {
      stop_helpers();
      finish_member(t, hl + P.consumed, pg);
      return true;
    }
    {   // parts in between were run through: nobody needs to look at them any more
      std::lock_guard<std::mutex> lk(mu);
      if (next_part < P.next) next_part = P.next;
      if (next_find < P.next) next_find = P.next;
    }
    cur = P.next;
  }
  stop_helpers();
  if (cancelled) {
    fail_task(t, "cancelled");
    return true;
  }
  return false;
}

bool GzReader::next(GzPiece& p) {
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    if (!head_) {
      if (chain_pos_ >= size_) return false;
      cv_.wait(lk, [this] { return scan_done_ || scan_pos_ > chain_pos_; });
      // candidates the chain has passed were false: drop them (a running one is told to stop and kept until it has)
      size_t dropped = 0;
      while (!tasks_.empty() && tasks_.front()->start < chain_pos_) {
        Task* t = tasks_.front().get();
        if (t->state == Task::RUNNING) {
          t->cancel = true;
          cv_.notify_all();
          cv_.wait(lk, [t] { return t->state != Task::RUNNING; });
        }
        for (auto& q : t->ready) {
          if (((Buf*)q.owner)->pooled) free_.push_back((Buf*)q.owner);
          else delete (Buf*)q.owner;
        }
        tasks_.pop_front();
        ++dropped;
      }
      next_claim_ = std::max(next_claim_, dropped) - dropped;
      if ((tasks_.empty() || tasks_.front()->start != chain_pos_) && chain_pos_ < size_) {
        // the scanner lists likely member starts only; what the chain arrives at is a member if its header says so
        const size_t hl = GzipMember::header_length(map_ + chain_pos_, size_ - chain_pos_);
        if (hl != 0 && hl != (size_t)-1) {
          std::unique_ptr<Task> nt(new Task());
          nt->start = chain_pos_;
          tasks_.push_front(std::move(nt));
          next_claim_ = 0;   // (tasks already taken are skipped by their state)
          cv_.notify_all();
        }
      }
      if (tasks_.empty() || tasks_.front()->start != chain_pos_) {
        if (chain_pos_ == 0) throw Fatal("corrupt gzip input: " + path_ + " (not a gzip file)");
        for (size_t i = chain_pos_; i < size_; ++i)   // zero padding after the last member is tolerated
          if (map_[i]) throw Fatal("corrupt gzip input: " + path_ + " (bytes that are no gzip member after a member)");
        chain_pos_ = size_;
        return false;
      }
      head_ = tasks_.front().get();
      cv_.notify_all();   // the head may take the reserved buffers
    }
    if (!head_->ready.empty()) {
      p = head_->ready.front();
      head_->ready.pop_front();
      cv_.notify_all();
      return true;
    }
    if (head_->state == Task::OK) {
      chain_pos_ = head_->end;
      ++members_;
      head_ = nullptr;
      tasks_.pop_front();
      next_claim_ = std::max<size_t>(next_claim_, 1) - 1;
      cv_.notify_all();
      continue;
    }
    if (head_->state == Task::FAILED) throw Fatal("corrupt gzip input: " + path_ + " (" + head_->error + ")");
    cv_.wait(lk);
  }
}

void GzReader::recycle(GzPiece& p) {
  if (!p.owner) return;
  Buf* b = (Buf*)p.owner;
  if (b->pooled) {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back(b);
  } else {
    delete b;
  }
  p.owner = nullptr;
  cv_.notify_all();
}

}  // namespace mgikit

#include "gz_reader.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

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

GzReader::GzReader(const std::string& path, int n_threads, size_t max_buffered) : path_(path) {
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

void GzReader::work() {
  std::unique_lock<std::mutex> lk(mu_);
  for (;;) {
    cv_.wait(lk, [this] { return stop_ || next_claim_ < tasks_.size(); });
    if (stop_) return;
    Task* t = tasks_[next_claim_++].get();
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
  return new Buf(INFLATE_WINDOW + piece_ + TAIL_SLACK);
}

void GzReader::decode(Task* t) {
  auto finish = [&](Task::State s, const std::string& why) {
    std::lock_guard<std::mutex> lk(mu_);
    t->state = s;
    t->error = why;
  };
  const uint8_t* in = map_ + t->start;
  const size_t left = size_ - t->start;
  const size_t hl = GzipMember::header_length(in, left);
  if (hl == 0 || hl == (size_t)-1) return finish(Task::FAILED, hl ? "unexpected end of file" : "not a gzip member");
  Inflater inf;
  inf.reset(in + hl, left - hl);
  uint32_t crc = 0;
  uint64_t total = 0;
  std::unique_ptr<uint8_t[]> carry(new uint8_t[INFLATE_WINDOW]);
  size_t hist = 0;
  for (;;) {
    Buf* b = take_buffer(t);
    if (!b) return finish(Task::FAILED, "cancelled");
    uint8_t* payload = b->mem + INFLATE_WINDOW;
    if (hist) memcpy(payload - hist, carry.get(), hist);
    uint8_t* next = payload;
    const Inflater::Status st = inf.run(payload - hist, &next, payload + piece_ + TAIL_SLACK);
    const size_t got = (size_t)(next - payload);
    if (st == Inflater::BAD_DATA || st == Inflater::TRUNCATED) {
      std::lock_guard<std::mutex> lk(mu_);
      free_.push_back(b);
      t->state = Task::FAILED;
      t->error = st == Inflater::TRUNCATED ? "unexpected end of file" : inf.error();
      return;
    }
    crc = crc32_fast(crc, payload, got);
    total += got;
    if (st == Inflater::NEED_OUTPUT) {   // the last 32 KiB go along into the next buffer
      const size_t have = hist + got, keep = have < INFLATE_WINDOW ? have : INFLATE_WINDOW;
      memmove(carry.get(), next - keep, keep);
      hist = keep;
    }
    GzPiece piece;
    piece.data = payload;
    piece.len = got;
    piece.newlines = count_newlines(payload, got);
    piece.owner = b;
    if (st == Inflater::DONE) {
      const size_t used = hl + inf.consumed();
      std::string why;
      if (left < used + 8) {
        why = "unexpected end of file";
      } else {
        const uint8_t* tr = in + used;
        const uint32_t want_crc = tr[0] | (tr[1] << 8) | (tr[2] << 16) | ((uint32_t)tr[3] << 24);
        const uint32_t want_len = tr[4] | (tr[5] << 8) | (tr[6] << 16) | ((uint32_t)tr[7] << 24);
        if (want_crc != crc) why = "incorrect data check";
        else if (want_len != (uint32_t)total) why = "incorrect length check";
      }
      std::lock_guard<std::mutex> lk(mu_);
      if (!why.empty()) {
        free_.push_back(b);
        t->state = Task::FAILED;
        t->error = why;
        return;
      }
      if (got) t->ready.push_back(piece);
      else free_.push_back(b);
      t->end = t->start + used + 8;
      t->state = Task::OK;
      return;
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      t->ready.push_back(piece);
    }
    cv_.notify_all();
  }
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
        for (auto& q : t->ready) free_.push_back((Buf*)q.owner);
        tasks_.pop_front();
        ++dropped;
      }
      next_claim_ = std::max(next_claim_, dropped) - dropped;
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
  {
    std::lock_guard<std::mutex> lk(mu_);
    free_.push_back((Buf*)p.owner);
  }
  p.owner = nullptr;
  cv_.notify_all();
}

}  // namespace mgikit

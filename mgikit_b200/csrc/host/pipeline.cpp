#include "pipeline.hpp"

#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <unordered_map>
#include <memory>
#include <mutex>
#include <thread>

#include "gz_reader.hpp"
#include "log.hpp"

namespace mgikit {

namespace {

// ---------------------------------------------------------------- thread pool
// Plain task queue: the codec tasks of several chunks are in flight at once, nobody waits for a chunk's tasks as a
// group (the last task of a chunk hands the chunk on).
class ThreadPool {
 public:
  explicit ThreadPool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void submit(std::function<void()> fn) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(std::move(fn));
    }
    cv_.notify_one();
  }

 private:
  void loop() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
      if (q_.empty()) return;   // stop_ and drained
      std::function<void()> fn = std::move(q_.front());
      q_.pop_front();
      lk.unlock();
      fn();   // tasks catch their own exceptions
      lk.lock();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  bool stop_ = false;
};

// ------------------------------------------------------- record-aligned blocks
struct Block {
  uint8_t* data = nullptr;
  size_t len = 0;        // bytes of whole records
  size_t n_records = 0;
  bool last = false;     // nothing follows
  std::string error;
};

// One assembler thread per input file on top of the member-parallel gunzip (gz_reader.hpp): decoded pieces are
// copied into pinned buffers and every block is cut after exactly `records_per_block` records (fill_send_buffers,
// src/file_utils.rs:99-164: truncate to batch_size*4 lines, carry the tail).  The pieces arrive with their newline
// counts, so only the piece a cut falls into is scanned.
class BlockReader {
 public:
  // `gz` is already decoding (it was started before the device contexts were made); the pinned chunk buffers are
  // allocated by the assembler thread itself, the first block is on its way while the others are still being pinned
  BlockReader(std::unique_ptr<GzReader> gz, size_t capacity, size_t records_per_block, int n_buffers, const Backend& be, ThreadPool& pool)
      : gz_(std::move(gz)), cap_(capacity), rpb_(records_per_block), n_buffers_(n_buffers), be_(be), pool_(pool) {
    th_ = std::thread([this] { run(); });
  }
  ~BlockReader() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    for (auto* p : all_) be_.host_free(p);
  }
  Block next() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !ready_.empty(); });
    Block b = ready_.front();
    ready_.pop_front();
    return b;
  }
  void give_back(uint8_t* p) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      free_.push_back(p);
    }
    cv_.notify_all();
  }

 private:
  void push(const Block& b) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      ready_.push_back(b);
    }
    cv_.notify_all();
  }
  uint8_t* take() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (stop_) return nullptr;
      if (free_.empty() && (int)all_.size() < n_buffers_) {   // not all pinned yet: one more
        lk.unlock();
        auto* p = (uint8_t*)be_.host_alloc(cap_);
        if (!p) throw Fatal("can not allocate chunk buffers");
        lk.lock();
        all_.push_back(p);
        return p;
      }
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return stop_ || !free_.empty(); });
    if (stop_) return nullptr;
    uint8_t* p = free_.front();
    free_.pop_front();
    return p;
  }
  // decoded pieces -> pinned chunk buffer: the copies of a block run on the pool (one thread moves ~3 GB/s into
  // pinned memory, a lane needs more)
  struct Copy {
    uint8_t* dst;
    const uint8_t* src;
    size_t n;
  };
  static constexpr size_t COPY_GRAIN = (size_t)1 << 20;
  void run_copies(const std::vector<Copy>& copies) {
    if (copies.size() <= 1) {
      for (const auto& c : copies) memcpy(c.dst, c.src, c.n);
      return;
    }
    std::atomic<size_t> next{0};
    const size_t lanes = std::min<size_t>(copies.size(), 8);
    std::mutex mu;
    std::condition_variable cv;
    size_t left = lanes;
    for (size_t l = 0; l < lanes; ++l)
      pool_.submit([&] {
        for (size_t i = next.fetch_add(1); i < copies.size(); i = next.fetch_add(1)) memcpy(copies[i].dst, copies[i].src, copies[i].n);
        std::lock_guard<std::mutex> lk(mu);
        if (--left == 0) cv.notify_all();
      });
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return left == 0; });
  }
  void fail(const std::string& why) {
    Block b;
    b.last = true;
    b.error = why;
    push(b);
  }
  void run() {
    try {
      GzReader& gz = *gz_;
      GzPiece piece;           // the piece being copied out, from `at` on
      size_t at = 0, piece_nl = 0;
      bool have_piece = false, eof = false;
      std::vector<uint8_t> carry;
      size_t carry_nl = 0;
      const size_t target = 4 * rpb_;
      std::vector<Copy> copies;
      std::vector<GzPiece> spent;
      while (!eof || !carry.empty() || have_piece) {
        uint8_t* buf = take();
        if (!buf) break;
        size_t have = carry.size(), lines = carry_nl;
        if (have) memcpy(buf, carry.data(), have);
        carry.clear();
        carry_nl = 0;
        while (lines < target && have < cap_) {
          if (!have_piece) {
            if (eof || !gz.next(piece)) {
              eof = true;
              break;
            }
            have_piece = true;
            at = 0;
            piece_nl = piece.newlines;
          }
          const size_t left = piece.len - at;
          size_t take_n = left, take_nl = piece_nl;
          if (lines + piece_nl > target) {   // the cut falls into this piece: up to the newline that completes the block
            const uint8_t* s = piece.data + at;
            size_t need = target - lines;
            take_nl = need;
            while (need) {
              s = (const uint8_t*)memchr(s, '\n', (size_t)(piece.data + piece.len - s)) + 1;
              --need;
            }
            take_n = (size_t)(s - (piece.data + at));
          }
          if (take_n > cap_ - have) {   // buffer full first (records longer than the first one): whatever fits
            take_n = cap_ - have;
            take_nl = 0;
            for (const uint8_t *s = piece.data + at, *e = s + take_n; (s = (const uint8_t*)memchr(s, '\n', (size_t)(e - s))) != nullptr; ++s) ++take_nl;
          }
          for (size_t o = 0; o < take_n; o += COPY_GRAIN) copies.push_back({buf + have + o, piece.data + at + o, std::min(COPY_GRAIN, take_n - o)});
          have += take_n;
          lines += take_nl;
          at += take_n;
          piece_nl -= take_nl;
          if (at == piece.len) {
            spent.push_back(piece);   // given back once its bytes have been copied out
            have_piece = false;
          }
        }
        run_copies(copies);
        copies.clear();
        for (auto& sp : spent) gz.recycle(sp);
        spent.clear();
        if (eof && have && buf[have - 1] != '\n' && have < cap_) {
          buf[have++] = '\n';
          ++lines;
        }
        Block b;
        b.data = buf;
        size_t cut = have;
        if (lines == target || (eof && (lines & 3) == 0)) {
          b.n_records = lines / 4;
        } else {   // full buffer or ragged end of file: back to the last newline that closes a record
          size_t n = 0, last4 = 0;
          for (const uint8_t *s = buf, *e = buf + have; (s = (const uint8_t*)memchr(s, '\n', (size_t)(e - s))) != nullptr; ++s)
            if ((++n & 3) == 0) last4 = (size_t)(s - buf) + 1;
          cut = last4;
          b.n_records = n / 4;
          carry_nl = n - 4 * b.n_records;
        }
        b.len = cut;
        if (cut < have) carry.assign(buf + cut, buf + have);
        if (b.n_records == 0) {
          if (!(eof && !have_piece)) {
            fail("a read record does not fit into one chunk buffer; increase the chunk size");
            return;
          }
          carry.clear();  // trailing partial record at end of file: nothing more to pair
          carry_nl = 0;
        }
        b.last = eof && !have_piece && carry.empty();
        push(b);
        if (b.last) return;
      }
      Block b;
      b.last = true;
      push(b);
    } catch (const std::exception& e) {
      fail(e.what());
    }
  }
  std::unique_ptr<GzReader> gz_;
  size_t cap_, rpb_;
  int n_buffers_;
  const Backend& be_;
  ThreadPool& pool_;
  std::vector<uint8_t*> all_;
  std::deque<uint8_t*> free_;
  std::deque<Block> ready_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread th_;
  bool stop_ = false;
};

// --------------------------------------------------------------- output files
// create_output_file_name(), src/sample_data.rs:536-567
void output_file_names(const std::string& name, const std::string& lane, size_t sample_index, bool illumina,
                       bool paired, std::string& r1, std::string& r2) {
  const char* br = paired ? "R2" : "R1";
  if (illumina) {
    if (sample_index == (size_t)-1) {
      r1 = name + "_" + lane + "_R1_001.fastq.gz";
      r2 = name + "_" + lane + "_" + br + "_001.fastq.gz";
    } else {
      r1 = name + "_S" + std::to_string(sample_index) + "_" + lane + "_R1_001.fastq.gz";
      r2 = name + "_S" + std::to_string(sample_index) + "_" + lane + "_" + br + "_001.fastq.gz";
    }
  } else {
    r1 = name + "_" + lane + "_R1.fastq.gz";
    r2 = name + "_" + lane + "_" + br + ".fastq.gz";
  }
}

struct SampleFiles {
  std::string barcode_path, paired_path;  // "" = this sample has no such file
};

// one gzip member per flush — compress_buffer(), src/sample_data.rs:472-490
void gzip_member(const uint8_t* src, size_t n, int level, std::vector<uint8_t>& out) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level > 9 ? 9 : level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
    throw Fatal("deflateInit2 failed");
  out.resize(deflateBound(&zs, (uLong)n) + 32);
  zs.next_in = const_cast<Bytef*>(src);
  zs.avail_in = (uInt)n;
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  const int rc = deflate(&zs, Z_FINISH);
  if (rc != Z_STREAM_END) {
    deflateEnd(&zs);
    throw Fatal("deflate failed");
  }
  out.resize(zs.total_out);
  deflateEnd(&zs);
}

// write_data(), src/sample_data.rs:492-503: open(append)/write/close
void append_file(const std::string& path, const uint8_t* data, size_t n) {
  FILE* f = fopen(path.c_str(), "ab");
  if (!f) throw Fatal("couldn't create output");
  const size_t w = fwrite(data, 1, n, f);
  fclose(f);
  if (w != n) throw Fatal("couldn't flush output");
}

void compress_segment(const uint8_t* src, size_t n, int level, std::vector<uint8_t>& out) {
  const size_t piece = (size_t)1 << 30;  // zlib counts in 32 bit
  out.clear();
  std::vector<uint8_t> z;
  for (size_t off = 0; off < n; off += piece) {
    gzip_member(src + off, std::min(piece, n - off), level, z);
    if (off == 0 && n <= piece) {
      out.swap(z);
      return;
    }
    out.insert(out.end(), z.begin(), z.end());
  }
}

// One chunk on its way through the host side: submitted -> collected -> codec tasks -> committed -> released.
struct ChunkJob {
  uint64_t seq = 0;
  int ctx = 0;
  Block b2, b1;
  mgk_result res{};
  struct Part {
    const uint8_t* p;
    size_t n;
    const std::string* path;
    std::vector<uint8_t> z;   // gzip member(s) made on the host; empty when the device already made them
  };
  std::vector<Part> parts;
  std::unordered_map<std::string, uint64_t> und, amb;   // this chunk's share of the barcode histograms
  std::atomic<int> pending{0};
  std::mutex err_mu;
  std::string error;
};

template <typename T>
class Channel {
 public:
  void push(T v) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(std::move(v));
    }
    cv_.notify_all();
  }
  bool try_pop(T& v) {
    std::lock_guard<std::mutex> lk(mu_);
    if (q_.empty()) return false;
    v = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  T pop() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return !q_.empty(); });
    T v = std::move(q_.front());
    q_.pop_front();
    return v;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
};

}  // namespace

SampleSheet dummy_sample_sheet(const std::string& label) {
  SampleSheet sh;
  sh.info.push_back({label, "", "", "", "", "", "."});
  sh.info.push_back({"Undetermined", "", "", "", "", "", "."});
  sh.info.push_back({"Ambiguous", "", "", "", "", "", "."});
  sh.writing_samples = {0, 1, 2};
  sh.unique_samples_ids = {0, 1, 2};
  sh.project_samples["."] = {};
  return sh;
}

void run_demultiplex(const Backend& be, const SampleSheet& sheet, RunInfo& run, const DemuxOptions& opt) {
  const auto t_start = std::chrono::steady_clock::now();
  const bool paired = run.paired_read_input();
  const bool illumina = run.illumina_format;
  const ReformatSample* rf = opt.reformat;
  const int total = sheet.total(), und = total - 2;
  const int BL = rf ? rf->umi_length : sheet.barcode_length();   // src/lib.rs:878-882
  const int m = rf ? 5 : opt.mismatches;                          // analyse_fastq(.., 5, ..), src/lib.rs:2431

  if (!rf) log_info(std::string("Comprehensive scan mode: ") + (run.comprehensive_scan ? "true" : "false"));
  if (!rf) {
    log_info("Allowed mismatches when finding the index are: " + std::to_string(m));
    if (!opt.per_index_error)
      log_info("The allowed mismatches will be compared to the total mismatches for all indicies combined.");
    else
      log_info("The allowed mismatches will be considered per index.");
    if (m > 2)
      log_warn("It is not recommended to allow more than 2 mismatches per index! This might decrease the accuracy of the demultipexing!");
    log_info("Reads that match with multiple samples will be saved in ambiguous_read1/2.fastq file");
    log_info("The paired read files are assumed to contain the paired reads in the same order!");
  }

  // ---- output files: create_sample_data_list(), src/sample_data.rs:656-726
  std::vector<SampleFiles> files((size_t)total);
  for (int i = 0; i < total; ++i) {
    if (sheet.writing_samples[(size_t)i] != i) continue;
    std::string r1, r2;
    const bool pseudo = i >= und;
    if (rf) {   // one sample, named by --sample-index (src/lib.rs:1386-1400); SampleData::minimal_sample() has no files
      if (i != 0) continue;
      output_file_names(rf->label, run.lane, rf->sample_index, illumina, paired, r1, r2);
      files[0].barcode_path = path_join(run.output_dir, r2);
      if (paired) files[0].paired_path = path_join(run.output_dir, r1);
      continue;
    }
    output_file_names(sheet.info[(size_t)i][SAMPLE_COLUMN], run.lane,
                      (pseudo && illumina) ? (size_t)-1 : (size_t)sheet.unique_samples_ids[(size_t)i] + 1, illumina, paired,
                      r1, r2);
    if (run.read2_has_sequence || pseudo) files[(size_t)i].barcode_path = path_join(run.output_dir, r2);
    if (paired) files[(size_t)i].paired_path = path_join(run.output_dir, r1);
  }
  if (run.force)  // clean_output_directory(), src/sample_data.rs:728-787
    for (const auto& f : files) {
      if (!f.barcode_path.empty()) remove(f.barcode_path.c_str());
      if (!f.paired_path.empty()) remove(f.paired_path.c_str());
    }

  // ---- the plan handed to the hot path
  RunFlags fl;
  fl.paired = paired;
  fl.read2_has_sequence = run.read2_has_sequence;
  fl.keep_barcode = run.keep_barcode;
  fl.comprehensive_scan = run.comprehensive_scan;
  fl.validate = run.check_content;
  fl.mgi_data = run.mgi_data;
  fl.per_index_error = opt.per_index_error;
  fl.out_mode = illumina ? MGK_OUT_ILLUMINA : (run.mgi_full_header ? MGK_OUT_MGI_FULL_HEADER : MGK_OUT_MGI);
  fl.report_level = opt.report_level;
  fl.mismatches = m;
  fl.l_position = (int)run.l_position();
  if (illumina) fl.illumina_prefix = run.create_illumina_header_prefix();
  GpuPlan plan;
  build_plan(sheet, fl, plan);
  if (rf) {
    plan.cfg.reformat = 1;
    plan.cfg.reformat_umi_length = rf->umi_length;
    plan.cfg.reformat_barcode = rf->barcode.c_str();
  }

  // records per chunk: batch_size, src/lib.rs:600-605
  const size_t rec_len = std::max(run.barcode_read_info.read_length, run.paired_read_info.read_length);
  if (rec_len == 0) throw Fatal("Something wrong in the input files!");
  size_t chunk_bytes = std::max(opt.chunk_bytes, rec_len * 8);
  if (chunk_bytes >= ((size_t)1 << 31)) chunk_bytes = ((size_t)1 << 31) - 4096;
  const size_t rpb = std::max<size_t>(1, (size_t)(0.95 * (double)chunk_bytes / (double)rec_len));
  plan.cfg.max_chunk_bytes = chunk_bytes;
  plan.cfg.max_chunk_records = rpb + 16;
  plan.cfg.n_slots = std::max(2, std::min(8, opt.slots));

  int n_threads = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  const int decode_threads = std::max(1, opt.reader_threads > 0 ? opt.reader_threads : n_threads / (paired ? 2 : 1));
  // the inputs start inflating now: device contexts and pinned buffers take a second to set up
  std::unique_ptr<GzReader> gz2(new GzReader(run.barcode_reads, decode_threads));
  std::unique_ptr<GzReader> gz1;
  if (paired) gz1.reset(new GzReader(run.paired_reads, decode_threads));

  const int G = std::max(1, opt.gpus);
  std::vector<void*> ctxs((size_t)G, nullptr);
  auto destroy_all = [&] {
    for (auto*& c : ctxs)
      if (c) {
        be.destroy(c);
        c = nullptr;
      }
  };
  for (int g = 0; g < G; ++g) {
    plan.cfg.device = g;
    if (be.create(&plan.cfg, &ctxs[(size_t)g]) != MGK_OK) {
      const std::string msg = be.last_error(nullptr);
      destroy_all();
      throw Fatal("can not start the " + std::string(be.name) + " hot path: " + msg);
    }
  }

  {
    char msg[96];
    snprintf(msg, sizeof msg, "hot path contexts ready after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    log_info(msg);
  }
  // the device makes the gzip members itself at the default level; any other level is zlib on the host threads
  const bool device_gzip = be.gzip_flag != 0 && opt.compression_level == 1 && !opt.host_gzip;
  const uint32_t submit_flags = device_gzip ? be.gzip_flag : 0u;
  log_info("Hot path: " + std::string(be.name) + " on " + std::to_string(G) + " device(s); chunk = " + std::to_string(rpb) +
           " reads; host threads: " + std::to_string(decode_threads) + " inflate per input file, " + std::to_string(n_threads) +
           (device_gzip ? " for the output (gzip members are made on the device)" : " deflate"));

  ReportData report(total, m);
  try {
    const int slots = plan.cfg.n_slots;
    const int max_in_flight = G * slots;
    const int n_buffers = max_in_flight + 2;
    ThreadPool pool(n_threads);
    std::unique_ptr<BlockReader> rd2_owner(new BlockReader(std::move(gz2), chunk_bytes, rpb, n_buffers, be, pool));
    BlockReader& rd2 = *rd2_owner;
    std::unique_ptr<BlockReader> rd1;
    if (paired) rd1.reset(new BlockReader(std::move(gz1), chunk_bytes, rpb, n_buffers, be, pool));

    const bool single_template = sheet.templates.size() == 1;
    static const std::array<int, 10> no_geometry{};
    const auto& g0 = sheet.templates.empty() ? no_geometry : sheet.templates[0].geometry;
    Channel<ChunkJob*> to_commit, finished;
    std::atomic<bool> failed{false};

    // ---- committer: appends the chunks' members to the sample files in chunk order (per-sample order == input
    // order for any number of devices) and merges the barcode histograms; then hands the chunk back to the pump
    {
      char msg[96];
      snprintf(msg, sizeof msg, "input readers started after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
      log_info(msg);
    }
    double commit_seconds = 0, input_wait = 0, device_wait = 0, done_wait = 0, submit_seconds = 0;
    std::thread committer([&] {
      std::map<uint64_t, ChunkJob*> waiting;
      uint64_t next_seq = 0;
      for (;;) {
        ChunkJob* j = to_commit.pop();
        if (!j) break;
        waiting[j->seq] = j;
        while (!waiting.empty() && waiting.begin()->first == next_seq) {
          ChunkJob* c = waiting.begin()->second;
          waiting.erase(waiting.begin());
          ++next_seq;
          if (c->error.empty() && !failed.load()) {
            try {
              // write_data(), src/sample_data.rs:492-503: open(append)/write/close.  The files of one chunk are all
              // different, so its appends run side by side; the next chunk starts when this one is on disk.
              const auto tc = std::chrono::steady_clock::now();
              const size_t n_parts = c->parts.size(), lanes = std::min<size_t>(n_parts, (size_t)std::max(1, n_threads));
              std::atomic<size_t> next_part{0};
              std::atomic<int> left{(int)lanes};
              std::mutex done_mu;
              std::condition_variable done_cv;
              std::string write_error;
              for (size_t l = 0; l < lanes; ++l)
                pool.submit([&] {
                  try {
                    for (size_t i = next_part.fetch_add(1); i < n_parts; i = next_part.fetch_add(1)) {
                      auto& part = c->parts[i];
                      append_file(*part.path, part.z.empty() ? part.p : part.z.data(), part.z.empty() ? part.n : part.z.size());
                    }
                  } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> lk(done_mu);
                    if (write_error.empty()) write_error = e.what();
                  }
                  std::lock_guard<std::mutex> lk(done_mu);
                  if (--left == 0) done_cv.notify_all();
                });
              {
                std::unique_lock<std::mutex> lk(done_mu);
                done_cv.wait(lk, [&] { return left.load() == 0; });
              }
              if (!write_error.empty()) throw Fatal(write_error);
              commit_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tc).count();
              for (auto& kv : c->und) report.undetermined_barcodes[kv.first] += kv.second;
              for (auto& kv : c->amb) report.ambiguous_barcodes[kv.first] += kv.second;
            } catch (const std::exception& e) {
              c->error = e.what();
            }
          }
          if (!c->error.empty()) failed.store(true);
          finished.push(c);
        }
      }
    });

    auto last_task_done = [&](ChunkJob* j) {
      if (j->pending.fetch_sub(1) == 1) to_commit.push(j);
    };
    auto task_error = [&](ChunkJob* j, const std::string& why) {
      std::lock_guard<std::mutex> lk(j->err_mu);
      if (j->error.empty()) j->error = why;
    };

    std::deque<ChunkJob*> fifo;                    // submitted, not collected (oldest first)
    std::vector<int> busy((size_t)G, 0);           // slots of a context that are not free (submitted .. released)
    std::vector<int> uncollected((size_t)G, 0);
    size_t held = 0;                               // collected, not yet released

    auto collect_oldest = [&] {
      ChunkJob* j = fifo.front();
      fifo.pop_front();
      --uncollected[(size_t)j->ctx];
      void* ctx = ctxs[(size_t)j->ctx];
      ++held;
      mgk_result& res = j->res;
      if (be.collect(ctx, &res) != MGK_OK) throw Fatal(std::string("hot path failed: ") + be.last_error(ctx));
      if (res.validate_error != MGK_VAL_OK) {
        static const char* what[] = {"", "Read header must starts with '@'!",
                                     "Seqeunce bases need to be in [A, C, G, T, a, c, g, t, N, n]!",
                                     "Quality score must be between [0 and 40]!", "Headers seem to be in differnet orders"};
        throw Fatal(std::string(what[res.validate_error]) + " (read " + std::to_string(res.validate_record) +
                    " of chunk " + std::to_string(j->seq) + ")");
      }
      if (res.n_records != j->b2.n_records || res.consumed_r2 != j->b2.len || (paired && res.consumed_r1 != j->b1.len))
        throw Fatal("Something wrong in the input files!");
      // per-sample segments -> gzip members (compress_and_write, src/sample_data.rs:180-207), one task each
      for (int b = 0; b < total; ++b) {
        if (res.seg_len_r2[b]) {
          if (files[(size_t)b].barcode_path.empty()) throw Fatal("Read must attached to a specific sample!");
          j->parts.push_back({res.out_r2 + res.seg_off_r2[b], (size_t)res.seg_len_r2[b], &files[(size_t)b].barcode_path, {}});
        }
        if (paired && res.seg_len_r1[b]) {
          if (files[(size_t)b].paired_path.empty()) throw Fatal("Read must attached to a specific sample!");
          j->parts.push_back({res.out_r1 + res.seg_off_r1[b], (size_t)res.seg_len_r1[b], &files[(size_t)b].paired_path, {}});
        }
      }
      const bool histogram = res.n_unassigned > 0;
      const size_t n_codec = device_gzip ? 0 : j->parts.size();
      j->pending.store((int)n_codec + (histogram ? 1 : 0) + 1);
      if (histogram)   // undetermined / ambiguous barcode histograms, src/lib.rs:1048-1088
        pool.submit([&, j] {
          try {
            const mgk_result& r = j->res;
            std::string label;
            for (uint64_t k = 0; k < r.n_unassigned; ++k) {
              const uint32_t v = r.unassigned[k];
              const uint8_t* bar = j->b2.data + (v & 0x7fffffffu);
              if (single_template) {
                label.assign((const char*)bar + BL - g0[2], (size_t)g0[1]);
                if (g0[3] == 1) {
                  label.push_back('+');
                  label.append((const char*)bar + BL - g0[5], (size_t)g0[4]);
                }
              } else {
                label.assign((const char*)bar, (size_t)BL);
              }
              ((v & 0x80000000u) ? j->amb : j->und)[label] += 1;
            }
          } catch (const std::exception& e) {
            task_error(j, e.what());
          }
          last_task_done(j);
        });
      for (size_t i = 0; i < n_codec; ++i)
        pool.submit([&, j, i] {
          try {
            if (!failed.load()) compress_segment(j->parts[i].p, j->parts[i].n, opt.compression_level, j->parts[i].z);
          } catch (const std::exception& e) {
            task_error(j, e.what());
          }
          last_task_done(j);
        });
      last_task_done(j);   // the chunk's own count: it moves on even when it has no task at all
    };
    auto release = [&](ChunkJob* j) {
      void* ctx = ctxs[(size_t)j->ctx];
      const std::string err = j->error;
      const int rc = be.release(ctx, j->seq);
      rd2.give_back(j->b2.data);
      if (paired) rd1->give_back(j->b1.data);
      --busy[(size_t)j->ctx];
      --held;
      delete j;
      if (!err.empty()) throw Fatal(err);
      if (rc != MGK_OK) throw Fatal(std::string("hot path failed: ") + be.last_error(ctx));
    };

    std::string pump_error;
    const auto t_pump = std::chrono::steady_clock::now();
    uint64_t n_chunks = 0;
    try {
      uint64_t seq = 0;
      bool input_done = false;
      while (!input_done || !fifo.empty() || held) {
        ChunkJob* done = nullptr;
        while (finished.try_pop(done)) release(done);
        const int g = (int)(seq % (uint64_t)G);
        if (!input_done && busy[(size_t)g] < slots && uncollected[(size_t)g] < 2) {
          const auto ti = std::chrono::steady_clock::now();
          Block b2 = rd2.next();
          Block b1;
          if (paired) b1 = rd1->next();
          input_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - ti).count();
          if (!b2.error.empty()) throw Fatal(b2.error);
          if (paired && !b1.error.empty()) throw Fatal(b1.error);
          if (paired && b2.n_records != b1.n_records) {   // src/lib.rs:1490-1494
            // a chunk is cut after `rpb` records, sized from the FIRST record of either file with 5 % of slack; one file
            // falling short means its records grew beyond that (or the files are not in step at all)
            const bool full = !b2.last && !b1.last && (b2.n_records < rpb || b1.n_records < rpb);
            throw Fatal(full ? "Something wrong in the input files! (" + std::to_string(rpb) + " records of one file do not fit a chunk of " +
                                   std::to_string(chunk_bytes) + " bytes: the records are longer than the first one of the file; raise --chunk-size)"
                             : std::string("Something wrong in the input files!"));
          }
          if (b2.n_records > 0) {
            const auto ts = std::chrono::steady_clock::now();
            if (be.submit(ctxs[(size_t)g], seq, b2.data, b2.len, paired ? b1.data : nullptr, paired ? b1.len : 0, submit_flags) != MGK_OK)
              throw Fatal(std::string("hot path failed: ") + be.last_error(ctxs[(size_t)g]));
            submit_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts).count();
            ChunkJob* j = new ChunkJob();
            j->seq = seq;
            j->ctx = g;
            j->b2 = b2;
            j->b1 = b1;
            fifo.push_back(j);
            ++busy[(size_t)g];
            ++uncollected[(size_t)g];
            ++seq;
            ++n_chunks;
          } else {
            if (b2.data) rd2.give_back(b2.data);
            if (paired && b1.data) rd1->give_back(b1.data);
          }
          if (b2.last || (paired && b1.last)) {
            if (paired && b2.last != b1.last) throw Fatal("Something wrong in the input files!");
            input_done = true;
          }
        } else if (!fifo.empty()) {
          const auto tg = std::chrono::steady_clock::now();
          collect_oldest();
          device_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tg).count();
        } else if (held) {
          const auto tw = std::chrono::steady_clock::now();
          release(finished.pop());
          done_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw).count();
        }
      }
    } catch (const std::exception& e) {
      pump_error = e.what();
      failed.store(true);
    }
    // drain: chunks still with the workers come back through the committer; chunks never collected are dropped
    while (held) {
      ChunkJob* j = finished.pop();
      try {
        release(j);
      } catch (const std::exception& e) {
        if (pump_error.empty()) pump_error = e.what();
      }
    }
    {
      char msg[96];
      snprintf(msg, sizeof msg, "last chunk released after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
      log_info(msg);
    }
    to_commit.push(nullptr);
    committer.join();
    for (ChunkJob* j : fifo) delete j;
    if (!pump_error.empty()) throw Fatal(pump_error);
    {
      char msg[256];
      snprintf(msg, sizeof msg, "pump phase: %llu chunks in %.3f s (first submit to last release)", (unsigned long long)n_chunks,
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pump).count());
      log_info(msg);
      snprintf(msg, sizeof msg, "pump waited %.2f s for input, %.2f s for the device, %.2f s for the output side, spent %.2f s submitting; appends took %.2f s",
               input_wait, device_wait, done_wait, submit_seconds, commit_seconds);
      log_info(msg);
    }

    // ReportManager::update across workers (src/lib.rs:819-827) == sum over devices
    std::vector<uint64_t> stats((size_t)total * MGK_N_STATS), mism((size_t)total * (size_t)(2 * m + 2));
    if (G > 1) {
      if (be.allreduce(ctxs.data(), G) != MGK_OK) throw Fatal(std::string("counter all-reduce failed: ") + be.last_error(ctxs[0]));
      if (be.reduced_counters(ctxs[0], stats.data(), mism.data()) != MGK_OK)
        throw Fatal(std::string("hot path failed: ") + be.last_error(ctxs[0]));
    } else if (be.counters(ctxs[0], stats.data(), mism.data()) != MGK_OK) {
      throw Fatal(std::string("hot path failed: ") + be.last_error(ctxs[0]));
    }
    report.add_counters(stats.data(), mism.data());
    if (opt.leave_teardown_to_exit) {   // the process ends right after the reports: unpinning and freeing gigabytes one
      rd2_owner.release();              // allocation at a time would only delay them (the OS takes everything back at once)
      rd1.release();
      for (auto*& c : ctxs) c = nullptr;
    }
    {
      char msg[96];
      snprintf(msg, sizeof msg, "counters read after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
      log_info(msg);
    }
  } catch (...) {
    destroy_all();
    throw;
  }
  destroy_all();

  {
    char msg[96];
    snprintf(msg, sizeof msg, "hot path closed after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    log_info(msg);
  }
  report.set_sample_total_reads();  // src/lib.rs:1602
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  log_info(std::to_string(report.get_total_reads()) + " reads were processed in " + std::to_string((long)secs) + " secs.");

  // src/lib.rs:1827-1853
  if (rf) {   // src/lib.rs:2445-2457
    if (opt.report_level > 0) report.prepare_final_data(5, paired ? 1 : 0, run.barcode_read_info, run.paired_read_info, BL);
    report.write_sample_report(run, sheet, opt.report_level);
  } else {
    const int max_mismatches = !opt.per_index_error ? m + 1 : 2 * m + 1;
    report.prepare_final_data(max_mismatches, paired ? 1 : 0, run.barcode_read_info, run.paired_read_info, BL);
    report.write_reports(run, sheet, opt.report_level, opt.report_limit, max_mismatches);
  }
  {
    char msg[96];
    snprintf(msg, sizeof msg, "reports written after %.2f s", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    log_info(msg);
  }
}

}  // namespace mgikit

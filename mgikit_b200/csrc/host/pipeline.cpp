#include "pipeline.hpp"

#include <zlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "log.hpp"

namespace mgikit {

namespace {

// ---------------------------------------------------------------- thread pool
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
  // run fn(i) for i in [0,n), return when all are done; first exception is rethrown
  void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn;
    next_ = 0;
    total_ = n;
    done_ = 0;
    error_ = nullptr;
    cv_.notify_all();
    done_cv_.wait(lk, [this] { return done_ == total_; });
    fn_ = nullptr;
    if (error_) std::rethrow_exception(error_);
  }

 private:
  void loop() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      cv_.wait(lk, [this] { return stop_ || (fn_ && next_ < total_); });
      if (stop_) return;
      const size_t i = next_++;
      const auto* fn = fn_;
      lk.unlock();
      std::exception_ptr err;
      try {
        (*fn)(i);
      } catch (...) {
        err = std::current_exception();
      }
      lk.lock();
      if (err && !error_) error_ = err;
      if (++done_ == total_) done_cv_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t next_ = 0, total_ = 0, done_ = 0;
  std::exception_ptr error_;
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

// One reader thread per input file: gunzip into pinned buffers, cut every
// block after exactly `records_per_block` records (fill_send_buffers,
// src/file_utils.rs:99-164: truncate to batch_size*4 lines, carry the tail).
class BlockReader {
 public:
  BlockReader(const std::string& path, size_t capacity, size_t records_per_block, int n_buffers, const Backend& be)
      : path_(path), cap_(capacity), rpb_(records_per_block), be_(be) {
    for (int i = 0; i < n_buffers; ++i) {
      auto* p = (uint8_t*)be_.host_alloc(cap_);
      if (!p) throw Fatal("can not allocate chunk buffers");
      all_.push_back(p);
      free_.push_back(p);
    }
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
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return stop_ || !free_.empty(); });
    if (stop_) return nullptr;
    uint8_t* p = free_.front();
    free_.pop_front();
    return p;
  }
  void run() {
    gzFile g = gzopen(path_.c_str(), "rb");
    if (!g) {
      Block b;
      b.last = true;
      b.error = "Could not open the file";
      push(b);
      return;
    }
    gzbuffer(g, 1u << 20);
    std::vector<uint8_t> carry;
    bool eof = false;
    while (!eof || !carry.empty()) {
      uint8_t* buf = take();
      if (!buf) break;
      size_t have = carry.size();
      if (have) memcpy(buf, carry.data(), have);
      carry.clear();
      while (!eof && have < cap_) {
        const size_t want = std::min(cap_ - have, (size_t)1 << 30);
        int got = gzread(g, buf + have, (unsigned)want);
        if (got < 0) {
          Block b;
          b.last = true;
          b.error = "corrupt gzip input: " + path_;
          push(b);
          gzclose(g);
          return;
        }
        if (got == 0) {
          eof = true;
          if (have && buf[have - 1] != '\n' && have < cap_) buf[have++] = '\n';
          break;
        }
        have += (size_t)got;
      }
      // cut after the (4*rpb_)-th newline, or after the last 4k-th one
      size_t lines = 0, cut = 0;
      const uint8_t* s = buf;
      const uint8_t* end = buf + have;
      while (s < end && lines < 4 * rpb_) {
        const uint8_t* q = (const uint8_t*)memchr(s, '\n', (size_t)(end - s));
        if (!q) break;
        ++lines;
        s = q + 1;
        if ((lines & 3) == 0) cut = (size_t)(s - buf);
      }
      Block b;
      b.data = buf;
      b.len = cut;
      b.n_records = lines / 4;
      if (cut < have) carry.assign(buf + cut, buf + have);
      if (b.n_records == 0) {
        if (!eof) {
          b.error = "a read record does not fit into one chunk buffer; increase the chunk size";
          b.last = true;
          push(b);
          break;
        }
        carry.clear();  // trailing partial record at end of file: nothing more to pair
      }
      b.last = eof && carry.empty();
      push(b);
      if (b.last) {
        gzclose(g);
        return;
      }
    }
    gzclose(g);
    Block b;
    b.last = true;
    push(b);
  }
  std::string path_;
  size_t cap_, rpb_;
  const Backend& be_;
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
void append_file(const std::string& path, const std::vector<uint8_t>& data) {
  FILE* f = fopen(path.c_str(), "ab");
  if (!f) throw Fatal("couldn't create output");
  const size_t w = fwrite(data.data(), 1, data.size(), f);
  fclose(f);
  if (w != data.size()) throw Fatal("couldn't flush output");
}

void compress_append(const uint8_t* src, size_t n, int level, const std::string& path) {
  const size_t piece = (size_t)1 << 30;  // zlib counts in 32 bit
  std::vector<uint8_t> z;
  for (size_t off = 0; off < n; off += piece) {
    gzip_member(src + off, std::min(piece, n - off), level, z);
    append_file(path, z);
  }
}

struct InFlight {
  uint64_t seq;
  int ctx;
  Block b2, b1;
};

}  // namespace

void run_demultiplex(const Backend& be, const SampleSheet& sheet, RunInfo& run, const DemuxOptions& opt) {
  const auto t_start = std::chrono::steady_clock::now();
  const bool paired = run.paired_read_input();
  const bool illumina = run.illumina_format;
  const int total = sheet.total(), und = total - 2;
  const int BL = sheet.barcode_length();
  const int m = opt.mismatches;

  log_info(std::string("Comprehensive scan mode: ") + (run.comprehensive_scan ? "true" : "false"));
  log_info("Allowed mismatches when finding the index are: " + std::to_string(m));
  if (!opt.per_index_error)
    log_info("The allowed mismatches will be compared to the total mismatches for all indicies combined.");
  else
    log_info("The allowed mismatches will be considered per index.");
  if (m > 2)
    log_warn("It is not recommended to allow more than 2 mismatches per index! This might decrease the accuracy of the demultipexing!");
  log_info("Reads that match with multiple samples will be saved in ambiguous_read1/2.fastq file");
  log_info("The paired read files are assumed to contain the paired reads in the same order!");

  // ---- output files: create_sample_data_list(), src/sample_data.rs:656-726
  std::vector<SampleFiles> files((size_t)total);
  for (int i = 0; i < total; ++i) {
    if (sheet.writing_samples[(size_t)i] != i) continue;
    std::string r1, r2;
    const bool pseudo = i >= und;
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

  // records per chunk: batch_size, src/lib.rs:600-605
  const size_t rec_len = std::max(run.barcode_read_info.read_length, run.paired_read_info.read_length);
  if (rec_len == 0) throw Fatal("Something wrong in the input files!");
  size_t chunk_bytes = std::max(opt.chunk_bytes, rec_len * 8);
  if (chunk_bytes >= ((size_t)1 << 31)) chunk_bytes = ((size_t)1 << 31) - 4096;
  const size_t rpb = std::max<size_t>(1, (size_t)(0.95 * (double)chunk_bytes / (double)rec_len));
  plan.cfg.max_chunk_bytes = chunk_bytes;
  plan.cfg.max_chunk_records = rpb + 16;
  plan.cfg.n_slots = 2;

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

  int n_threads = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  log_info("Hot path: " + std::string(be.name) + " on " + std::to_string(G) + " device(s); chunk = " +
           std::to_string(rpb) + " reads; host codec threads = " + std::to_string(n_threads));

  ReportData report(total, m);
  try {
    const int max_in_flight = G * plan.cfg.n_slots;
    const int n_buffers = max_in_flight + 2;
    BlockReader rd2(run.barcode_reads, chunk_bytes, rpb, n_buffers, be);
    std::unique_ptr<BlockReader> rd1;
    if (paired) rd1.reset(new BlockReader(run.paired_reads, chunk_bytes, rpb, n_buffers, be));
    ThreadPool pool(n_threads);

    std::deque<InFlight> fifo;
    const bool single_template = sheet.templates.size() == 1;
    const auto& g0 = sheet.templates[0].geometry;

    auto finish_oldest = [&] {
      InFlight f = fifo.front();
      fifo.pop_front();
      void* ctx = ctxs[(size_t)f.ctx];
      mgk_result res;
      if (be.collect(ctx, &res) != MGK_OK) throw Fatal(std::string("hot path failed: ") + be.last_error(ctx));
      if (res.validate_error != MGK_VAL_OK) {
        static const char* what[] = {"", "Read header must starts with '@'!",
                                     "Seqeunce bases need to be in [A, C, G, T, a, c, g, t, N, n]!",
                                     "Quality score must be between [0 and 40]!", "Headers seem to be in differnet orders"};
        throw Fatal(std::string(what[res.validate_error]) + " (read " + std::to_string(res.validate_record) +
                    " of chunk " + std::to_string(f.seq) + ")");
      }
      if (res.n_records != f.b2.n_records || res.consumed_r2 != f.b2.len || (paired && res.consumed_r1 != f.b1.len))
        throw Fatal("Something wrong in the input files!");
      // undetermined / ambiguous barcode histograms, src/lib.rs:1048-1088
      for (uint64_t k = 0; k < res.n_unassigned; ++k) {
        const uint32_t v = res.unassigned[k];
        const uint8_t* bar = f.b2.data + (v & 0x7fffffffu);
        std::string label;
        if (single_template) {
          label.assign((const char*)bar + BL - g0[2], (size_t)g0[1]);
          if (g0[3] == 1) {
            label.push_back('+');
            label.append((const char*)bar + BL - g0[5], (size_t)g0[4]);
          }
        } else {
          label.assign((const char*)bar, (size_t)BL);
        }
        if (v & 0x80000000u) report.ambiguous_barcodes[label] += 1;
        else report.undetermined_barcodes[label] += 1;
      }
      // per-sample segments -> gzip members -> append (compress_and_write, src/sample_data.rs:180-207)
      struct Task {
        const uint8_t* p;
        size_t n;
        const std::string* path;
      };
      std::vector<Task> tasks;
      for (int b = 0; b < total; ++b) {
        if (res.seg_len_r2[b]) {
          if (files[(size_t)b].barcode_path.empty()) throw Fatal("Read must attached to a specific sample!");
          tasks.push_back({res.out_r2 + res.seg_off_r2[b], (size_t)res.seg_len_r2[b], &files[(size_t)b].barcode_path});
        }
        if (paired && res.seg_len_r1[b]) {
          if (files[(size_t)b].paired_path.empty()) throw Fatal("Read must attached to a specific sample!");
          tasks.push_back({res.out_r1 + res.seg_off_r1[b], (size_t)res.seg_len_r1[b], &files[(size_t)b].paired_path});
        }
      }
      pool.parallel_for(tasks.size(), [&](size_t i) {
        compress_append(tasks[i].p, tasks[i].n, opt.compression_level, *tasks[i].path);
      });
      if (be.release(ctx, f.seq) != MGK_OK) throw Fatal(std::string("hot path failed: ") + be.last_error(ctx));
      rd2.give_back(f.b2.data);
      if (paired) rd1->give_back(f.b1.data);
    };

    uint64_t seq = 0;
    for (;;) {
      Block b2 = rd2.next();
      Block b1;
      if (paired) b1 = rd1->next();
      if (!b2.error.empty()) throw Fatal(b2.error);
      if (paired && !b1.error.empty()) throw Fatal(b1.error);
      if (paired && b2.n_records != b1.n_records) throw Fatal("Something wrong in the input files!");  // src/lib.rs:1490-1494
      if (b2.n_records > 0) {
        if ((int)fifo.size() == max_in_flight) finish_oldest();
        const int g = (int)(seq % (uint64_t)G);
        if (be.submit(ctxs[(size_t)g], seq, b2.data, b2.len, paired ? b1.data : nullptr, paired ? b1.len : 0, 0) != MGK_OK)
          throw Fatal(std::string("hot path failed: ") + be.last_error(ctxs[(size_t)g]));
        fifo.push_back({seq, g, b2, b1});
        ++seq;
      } else {
        if (b2.data) rd2.give_back(b2.data);
        if (paired && b1.data) rd1->give_back(b1.data);
      }
      if (b2.last || (paired && b1.last)) {
        if (paired && b2.last != b1.last) throw Fatal("Something wrong in the input files!");
        break;
      }
    }
    while (!fifo.empty()) finish_oldest();

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
  } catch (...) {
    destroy_all();
    throw;
  }
  destroy_all();

  report.set_sample_total_reads();  // src/lib.rs:1602
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  log_info(std::to_string(report.get_total_reads()) + " reads were processed in " + std::to_string((long)secs) + " secs.");

  // src/lib.rs:1827-1853
  const int max_mismatches = !opt.per_index_error ? m + 1 : 2 * m + 1;
  report.prepare_final_data(max_mismatches, paired ? 1 : 0, run.barcode_read_info, run.paired_read_info, BL);
  report.write_reports(run, sheet, opt.report_level, opt.report_limit, max_mismatches);
}

}  // namespace mgikit

// The demultiplex run: gunzip -> record-aligned chunk pairs -> hot path
// (through the C ABI of include/mgikit_gpu.h) -> per-sample gzip members ->
// reports.  Replaces demultiplex()/analyse_fastq() of the reference
// (/root/reference/src/lib.rs:511-838, :1350-1604) and its reader threads
// (src/file_utils.rs:99-164, :239-373).
#pragma once

#include <cstdint>
#include <string>

#include "../../../include/mgikit_gpu.h"
#include "report.hpp"
#include "run_info.hpp"
#include "sample_sheet.hpp"

namespace mgikit {

// The hot path as the host sees it: exactly the entry points of mgikit_gpu.h.
// The product binds these to libmgikit_gpu.so; the test-only oracle CLI binds
// them to oracle/ — the host code does not know which.
struct Backend {
  const char* name;
  int (*create)(const mgk_config*, void**);
  void (*destroy)(void*);
  const char* (*last_error)(const void*);
  int (*submit)(void*, uint64_t, const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint32_t);
  int (*collect)(void*, mgk_result*);
  int (*release)(void*, uint64_t);
  int (*counters)(void*, uint64_t*, uint64_t*);
  int (*allreduce)(void* const*, int);
  int (*reduced_counters)(void*, uint64_t*, uint64_t*);
  void* (*host_alloc)(size_t);
  void (*host_free)(void*);
  uint32_t gzip_flag;   // mgk_submit flag that makes the device return gzip members (0: the backend has none)
  bool exits_after_command;   // the binary leaves with _exit: contexts and pinned buffers need no orderly teardown
};

// `mgikit reformat` (src/lib.rs:2266-2464, src/formater.rs): the run of ONE sample whose reads are already apart
struct ReformatSample {
  std::string label;        // --sample-label, or the third '_' part of the barcode read's file name
  size_t sample_index = 1;  // --sample-index: the S<n> of the Illumina file names
  int umi_length = 0;       // --umi-length
  std::string barcode;      // --barcode
};

struct DemuxOptions {
  int mismatches = 1;
  bool per_index_error = false;
  int report_level = 2;
  size_t report_limit = 20;
  int compression_level = 1;
  int threads = 0;          // -t: host codec threads (0 = all cores)
  int gpus = 1;             // --gpus (extension): chunk k goes to GPU k % gpus
  size_t chunk_bytes = 64u << 20;  // decompressed bytes per read file per chunk
  int slots = 4;            // chunks per GPU between submit and release (>= 2)
  int reader_threads = 0;   // --reader-threads: inflate threads per input file (0 = share of -t)
  bool leave_teardown_to_exit = false;   // the product binary: it calls _exit right after the command
  bool host_gzip = false;   // --host-gzip (extension): gzip on the host threads even at the default level
  const ReformatSample* reformat = nullptr;   // set: the `reformat` command (the sheet is dummy_sample_sheet())
};

// SampleManager::dummy_sample(), src/sample_manager.rs:133-185: the sample, Undetermined, Ambiguous
SampleSheet dummy_sample_sheet(const std::string& label);

// Runs the whole command; throws Fatal on any error the reference would panic on.
void run_demultiplex(const Backend& be, const SampleSheet& sheet, RunInfo& run, const DemuxOptions& opt);

}  // namespace mgikit

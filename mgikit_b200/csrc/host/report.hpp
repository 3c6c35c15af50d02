// Counters -> report files.  Host-side mirror of ReportManager
// (/root/reference/src/report_manager.rs); text must be byte-identical.
#pragma once

#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "run_info.hpp"
#include "sample_sheet.hpp"

namespace mgikit {

struct ReportData {
  int total_samples = 0;
  std::vector<std::vector<uint64_t>> sample_mismatches;  // [total][2m+2]
  std::vector<std::vector<uint64_t>> sample_statistics;  // [total][10 (+ mismatch columns)]
  std::unordered_map<std::string, uint64_t> undetermined_barcodes, ambiguous_barcodes;

  ReportData(int total, int allowed_mismatches);                 // ReportManager::new, :264-295
  // tables as returned by mgk_counters(): stats [total][10], mism [total][2m+2]
  void add_counters(const uint64_t* stats, const uint64_t* mism);  // == update(), :377-416
  void set_sample_total_reads();                                 // :367-375
  uint64_t get_total_reads() const;                              // :357-365
  void prepare_final_data(int max_mismatches, int shift, const ReadInfo& barcode_read,
                          const ReadInfo& paired_read, int barcode_length);  // :437-463
  void write_reports(const RunInfo& run, const SampleSheet& sheet, int reporting_level, size_t report_limit,
                     int max_mismatches) const;                  // :465-676 (individual_sample == usize::MAX)
  // the same with individual_sample == 0 (`reformat`, src/lib.rs:2457): only <sample>.mgikit.sample_stats, the row of
  // sample 0, mismatch columns cut after the last one that is not zero (:504-512)
  void write_sample_report(const RunInfo& run, const SampleSheet& sheet, int reporting_level) const;
};

std::string fmt_f64(double v);    // Rust `{}`
std::string fmt_f64_3(double v);  // Rust `{:.3?}`

}  // namespace mgikit

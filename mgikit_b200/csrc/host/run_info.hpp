// Run description: inputs, output/report directories, first-record facts.
// Host-side mirror of RunManager / ReadInfo (/root/reference/src/run_manager.rs).
#pragma once

#include <string>

#include "sample_sheet.hpp"

namespace mgikit {

struct ReadInfo {  // src/run_manager.rs:321-345
  size_t sequence_length = 0;
  size_t read_length = 0;  // bytes of the whole 4-line record
  size_t l_position = 0;
  std::string flowcell;
  std::string lane;
};

struct RunArgs {  // the RunManager::new argument list, src/run_manager.rs:55-75
  std::string input_dir, read1, read2, output_dir, report_dir, lane, instrument, run;
  bool mgi_data = true, force = false;
  std::string r1_suffix = "_read_1.fq.gz", r2_suffix = "_read_2.fq.gz", info_file;
  bool illumina_format = true, keep_barcode = false, comprehensive_scan = false,
       ignore_undetermined = false, check_content = false, mgi_full_header = false;
};

struct RunInfo {
  std::string barcode_reads, paired_reads;  // paths ("" = none)
  std::string output_dir, report_dir, lane, instrument, run;
  bool mgi_data = true, force = false, illumina_format = true, keep_barcode = false,
       comprehensive_scan = false, ignore_undetermined = false, check_content = false,
       mgi_full_header = false, read2_has_sequence = true;
  ReadInfo barcode_read_info, paired_read_info;

  bool paired_read_input() const { return paired_read_info.read_length > 0; }  // :192
  const std::string& flowcell() const { return barcode_read_info.flowcell; }
  size_t l_position() const { return barcode_read_info.l_position; }

  void get_read_information();                          // :220-318
  void confirm_format() const;                          // :145-164
  std::string create_illumina_header_prefix() const;    // :166-186
};

RunInfo make_run_info(const RunArgs& a);  // RunManager::new, :55-143

bool path_is_file(const std::string& p);
bool path_is_dir(const std::string& p);
void create_dir_all(const std::string& p);
std::string path_join(const std::string& dir, const std::string& name);

}  // namespace mgikit

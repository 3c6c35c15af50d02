// `mgikit template`: detect_template(), /root/reference/src/lib.rs:1948-2264;
// flags as in src/main.rs:400-511.
#pragma once

#include <cstddef>
#include <string>

namespace mgikit {

struct TemplateArgs {
  std::string input_dir, read1, read2, sample_sheet, info_file, output;
  std::string r1_suffix = "_read_1.fq.gz", r2_suffix = "_read_2.fq.gz";
  size_t barcode_length = 0;     // 0: sequence length of R2 minus R1
  size_t testing_reads = 5000;
  size_t max_umi_length = 10;
  bool popular_template = false;
  bool add_umi = true;           // !--no-umi
};

// Writes <output>_template.tsv and <output>_details.tsv; throws Fatal where the reference panics.
void run_template(const TemplateArgs& a);

}  // namespace mgikit

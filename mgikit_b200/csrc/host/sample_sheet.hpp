// Sample sheet, barcode templates and the index tables handed to the GPU.
// Host-side mirror of SampleManager (/root/reference/src/sample_manager.rs).
#pragma once

#include <array>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/mgikit_gpu.h"

namespace mgikit {

// column order of SampleManager::sample_information rows (src/variables.rs:2-8)
enum { SAMPLE_COLUMN = 0, I7_COLUMN = 1, I5_COLUMN = 2, TEMPLATE_COLUMN = 3, I7_RC_COLUMN = 4,
       I5_RC_COLUMN = 5, PROJECT_ID_COLUMN = 6 };

struct Fatal : std::runtime_error {
  using std::runtime_error::runtime_error;
};

std::string reverse_complement(const std::string& seq);          // src/sample_manager.rs:852-871
std::array<int, 10> parse_template(const std::string& tmpl);      // src/sample_manager.rs:249-288

// One entry of all_template_data (src/sample_manager.rs:24-34), with the string
// sets flattened into "unique index list + pair->sample matrix".
struct TemplateData {
  unsigned count = 0;                 // .0 samples using the template
  std::string tmpl;                   // .3
  bool has_i5 = false;                // .5
  std::array<int, 10> geometry{};     // .6
  std::vector<std::string> i7s, i5s;  // .1/.2 unique, first-appearance order
  std::vector<int> pair_sample;       // .4: has_i5 ? [n_i7*n_i5] : [n_i7]; -1 = none
  // rows kept while building
  std::vector<std::array<int, 3>> rows;  // (i7 idx, i5 idx or -1, sample)
};

struct SampleSheet {
  std::vector<std::array<std::string, 7>> info;   // + Undetermined, Ambiguous rows
  std::vector<int> writing_samples;                // src/sample_manager.rs:692-730
  std::vector<int> unique_samples_ids;
  std::map<std::string, std::vector<int>> project_samples;  // src/sample_manager.rs:458-504
  std::vector<TemplateData> templates;             // matching order
  int n_samples() const { return (int)info.size() - 2; }
  int total() const { return (int)info.size(); }
  int barcode_length() const { return templates[0].geometry[9]; }
};

// load_sample_sheet (src/sample_manager.rs:290-456): the rows as written, no pseudo samples
std::vector<std::array<std::string, 7>> load_sample_sheet(const std::string& path);

// SampleManager::new (src/sample_manager.rs:38-113)
SampleSheet load_sample_manager(const std::string& path, const std::string& tmpl, bool i7_rc,
                                bool i5_rc, const std::string& undetermined_label,
                                const std::string& ambiguous_label);

// Owns the storage a mgk_config points into.
struct GpuPlan {
  mgk_config cfg{};
  std::vector<mgk_template> tpls;
  std::vector<std::string> i7_flat, i5_flat;
  std::vector<std::vector<int32_t>> pairs;
  std::vector<int32_t> writing;
  std::string prefix;
};

struct RunFlags {
  bool paired = true, read2_has_sequence = true, keep_barcode = false, comprehensive_scan = false,
       validate = false, mgi_data = true, per_index_error = false;
  int out_mode = MGK_OUT_ILLUMINA;
  int report_level = 2;
  int mismatches = 1;
  int l_position = 0;
  std::string illumina_prefix;
};

void build_plan(const SampleSheet& sheet, const RunFlags& flags, GpuPlan& plan);

}  // namespace mgikit

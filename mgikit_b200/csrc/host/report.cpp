#include "report.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <set>

#include "log.hpp"

namespace mgikit {

// Rust `format!("{}", f64)`: shortest round-trip digits, never scientific, no ".0".
std::string fmt_f64(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

// Rust `format!("{:.3?}", f64)`
std::string fmt_f64_3(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  char buf[512];
  int n = std::snprintf(buf, sizeof buf, "%.3f", v);
  return std::string(buf, (size_t)n);
}

namespace {

void write_file(const std::string& path, const std::string& content) {
  std::ofstream f(path, std::ios::binary | std::ios::trunc);
  if (!f) throw Fatal("couldn't create output");
  f.write(content.data(), (std::streamsize)content.size());
}

bool contains(const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

std::string lower(std::string s) {
  for (auto& c : s) c = (char)tolower((unsigned char)c);
  return s;
}

bool all_ids_unique(const SampleSheet& sheet) {
  std::set<std::string> u;
  for (const auto& r : sheet.info) u.insert(r[SAMPLE_COLUMN]);
  return u.size() == sheet.info.size();
}

// write_index_info_report(), src/report_manager.rs:195-252
void write_index_info_report(const SampleSheet& sheet, const std::vector<std::vector<uint64_t>>& mism,
                             const std::vector<int>& kept, int max_mismatches, const std::string& out_file,
                             const std::vector<int>& excluded) {
  std::string rep = "sample";
  for (int i = 0; i < max_mismatches; ++i) rep += "\t" + std::to_string(i) + "-mismatches";
  rep += "\n";
  auto copy = mism;
  for (size_t i = 0; i < copy.size(); ++i) copy[i].push_back(i);
  if (all_ids_unique(sheet))
    std::stable_sort(copy.begin(), copy.end(), [](const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
      if (a[0] != b[0]) return a[0] > b[0];
      return a.back() < b.back();
    });
  for (size_t itr = 0; itr < copy.size(); ++itr) {
    const int sample_index = (int)copy[itr].back();
    if (!kept.empty() && !contains(kept, sample_index)) continue;
    if (contains(excluded, (int)itr)) continue;  // :234 — tested against the sorted POSITION
    rep += sheet.info[(size_t)sample_index][SAMPLE_COLUMN];
    for (int cnt = 1; cnt < max_mismatches + 1; ++cnt) {
      if ((size_t)cnt < copy[itr].size()) rep += "\t" + std::to_string(copy[itr][(size_t)cnt]);
      else rep += "\t0";
    }
    rep += "\n";
  }
  write_file(out_file, rep);
}

// write_general_info_report(), src/report_manager.rs:12-193
void write_general_info_report(const SampleSheet& sheet, const std::vector<std::vector<uint64_t>>& stats,
                               const std::vector<int>& kept, const std::string& run, const std::string& lane,
                               const std::string& out_file, const std::vector<int>& excluded) {
  std::string out =
      "#sample general info\nSample ID\tM Clusters\tMb Yield \xE2\x89\xA5 Q30\t% R1 Yield \xE2\x89\xA5 Q30\t% R2 Yield \xE2\x89\xA5 "
      "Q30\t% R3 Yield \xE2\x89\xA5 Q30\t% Perfect Index\n";
  double lane_st[5] = {0, 0, 0, 0, 0};
  auto copy = stats;
  for (size_t i = 0; i < copy.size(); ++i) copy[i].push_back(i);
  double excluded_reads = 0.0;
  if (all_ids_unique(sheet))
    std::stable_sort(copy.begin(), copy.end(), [](const std::vector<uint64_t>& a, const std::vector<uint64_t>& b) {
      if (a[9] != b[9]) return a[9] > b[9];
      return a.back() < b.back();
    });
  for (size_t itr = 0; itr < copy.size(); ++itr) {
    const int sample_id = (int)copy[itr].back();
    if (!kept.empty() && !contains(kept, sample_id)) continue;
    if (contains(excluded, sample_id)) continue;
    std::vector<double> st(copy[itr].begin(), copy[itr].end());
    std::string id = sheet.info[(size_t)sample_id][SAMPLE_COLUMN];
    out += id + "\t";
    out += fmt_f64(st[9] / 1000000.0) + "\t";
    const double all_q30 = st[0] + st[1];
    lane_st[0] += st[3] + st[4];
    lane_st[1] += st[9];
    lane_st[2] += all_q30;
    out += fmt_f64(all_q30 / 1000000.0) + "\t";
    out += fmt_f64_3((st[3] == 0.0 ? 0.0 : st[0] / st[3]) * 100.0) + "\t";
    out += fmt_f64_3((st[4] == 0.0 ? 0.0 : st[1] / st[4]) * 100.0) + "\t";
    out += fmt_f64_3((st[5] == 0.0 ? 0.0 : st[2] / st[5]) * 100.0) + "\t";
    out += fmt_f64_3((st[9] == 0.0 ? 0.0 : st[10] / st[9]) * 100.0);
    lane_st[3] += st[6] + st[7];
    id = lower(id);
    const bool non_sample = id == "undetermined" || id == "ambiguous";
    if (non_sample) excluded_reads += st[9];
    if (!non_sample) lane_st[4] += st[10];
    out += "\n";
  }
  std::string fin =
      "#Lane statistics\nRun ID-Lane\tMb Total Yield\tM Total Clusters\t% bases \xE2\x89\xA5 Q30\tMean Quality\t% Perfect "
      "Index\n";
  fin += run + "-" + lane + "\t";
  fin += fmt_f64(lane_st[0] / 1000000.0) + "\t";
  fin += fmt_f64(lane_st[1] / 1000000.0) + "\t";
  fin += fmt_f64_3((lane_st[2] / lane_st[0]) * 100.0) + "\t";
  fin += fmt_f64_3(lane_st[3] / lane_st[0]) + "\t";
  fin += fmt_f64_3((lane_st[4] / (lane_st[1] - excluded_reads)) * 100.0) + "\n";
  fin += out;
  write_file(out_file, fin);
}

}  // namespace

ReportData::ReportData(int total, int allowed_mismatches) : total_samples(total) {
  sample_mismatches.assign((size_t)total, std::vector<uint64_t>((size_t)(2 * allowed_mismatches + 2), 0));
  sample_statistics.assign((size_t)total, std::vector<uint64_t>(10, 0));
}

void ReportData::add_counters(const uint64_t* stats, const uint64_t* mism) {
  for (int s = 0; s < total_samples; ++s) {
    for (size_t j = 0; j < 10; ++j) sample_statistics[(size_t)s][j] += stats[(size_t)s * 10 + j];
    const size_t w = sample_mismatches[(size_t)s].size();
    for (size_t j = 0; j < w; ++j) sample_mismatches[(size_t)s][j] += mism[(size_t)s * w + j];
  }
}

void ReportData::set_sample_total_reads() {
  for (auto& row : sample_mismatches) {
    uint64_t reads = 0;
    for (size_t j = 1; j < row.size(); ++j) reads += row[j];
    row[0] = reads;
  }
}

uint64_t ReportData::get_total_reads() const {
  uint64_t reads = 0;
  for (const auto& row : sample_mismatches)
    for (size_t j = 1; j < row.size(); ++j) reads += row[j];
  return reads;
}

void ReportData::prepare_final_data(int max_mismatches, int shift, const ReadInfo& br, const ReadInfo& pr, int BL) {
  for (int s = 0; s < total_samples; ++s) {
    auto& st = sample_statistics[(size_t)s];
    const auto& mm = sample_mismatches[(size_t)s];
    st[3] = (uint64_t)pr.sequence_length * mm[0];
    st[6] -= st[3] * 33;  // wrapping like the reference's release build
    st[(size_t)shift + 3] = (uint64_t)(br.sequence_length - (size_t)BL) * mm[0];
    st[5] = (uint64_t)BL * mm[0];
    st[6 + (size_t)shift] -= st[(size_t)shift + 3] * 33;
    st[8] -= st[5] * 33;
    st[9] = mm[0];
    for (int cnt = 1; cnt < max_mismatches + 1; ++cnt) st.push_back(mm[(size_t)cnt]);
  }
}

void ReportData::write_reports(const RunInfo& run, const SampleSheet& sheet, int reporting_level, size_t report_limit,
                               int max_mismatches) const {
  const size_t sample_stats_width = sample_statistics[0].size();
  std::vector<int> excluded;
  if (sample_mismatches[(size_t)total_samples - 2][0] == 0) excluded.push_back(total_samples - 2);
  if (sample_mismatches[(size_t)total_samples - 1][0] == 0) excluded.push_back(total_samples - 1);
  const std::string main = run.flowcell() + "." + run.lane + ".mgikit.";

  for (const auto& kv : sheet.project_samples) {
    std::string extra;
    if (kv.first != ".") {
      log_info("generating report for job_number: " + kv.first + " with " + std::to_string(kv.second.size()) + " samples.");
      extra = kv.first + "_";
    } else {
      log_info("generating report for the whole run with " + std::to_string(sample_mismatches.size()) + " samples.");
    }
    write_index_info_report(sheet, sample_mismatches, kv.second, max_mismatches,
                            path_join(run.report_dir, extra + main + "info"), excluded);
    if (reporting_level > 0)
      write_general_info_report(sheet, sample_statistics, kv.second, run.flowcell(), run.lane,
                                path_join(run.report_dir, extra + main + "general"), excluded);
  }
  if (reporting_level > 0) {
    std::string out =
        "job_number\tsample_id\tr1_qc_30\tr2_qc_30\tr3_qc_30\tr1_bases\tr2_bases\tr3_bases\tr1_qc\tr2_qc\tr3_qc\tall_reads";
    for (int cnt = 0; cnt < max_mismatches; ++cnt) out += "\t" + std::to_string(cnt) + "-mismatches";
    out += "\n";
    for (size_t s = 0; s < sample_statistics.size(); ++s) {
      if (contains(excluded, (int)s)) continue;
      out += sheet.info[s][PROJECT_ID_COLUMN] + "\t" + sheet.info[s][SAMPLE_COLUMN];
      for (size_t c = 0; c < sample_stats_width; ++c) out += "\t" + std::to_string(sample_statistics[s][c]);
      out += "\n";
    }
    write_file(path_join(run.report_dir, main + "sample_stats"), out);
  }
  if (reporting_level > 1) {
    using Item = std::pair<std::string, uint64_t>;
    auto sorted = [](const std::unordered_map<std::string, uint64_t>& m) {
      std::vector<Item> v(m.begin(), m.end());
      std::sort(v.begin(), v.end(), [](const Item& a, const Item& b) {  // (count, barcode) descending, :611,:646
        if (a.second != b.second) return a.second > b.second;
        return a.first > b.first;
      });
      return v;
    };
    size_t rep_itr = 0;
    if (!ambiguous_barcodes.empty()) {
      const auto v = sorted(ambiguous_barcodes);
      std::string s;
      rep_itr = 0;
      for (const auto& it : v) {
        s += it.first + "\t" + std::to_string(it.second) + "\n";
        if (++rep_itr == report_limit) break;
      }
      write_file(path_join(run.report_dir, main + "ambiguous_barcode"), s);
      s.clear();
      for (const auto& it : v) s += it.first + "\t" + std::to_string(it.second) + "\n";
      write_file(path_join(run.report_dir, main + "ambiguous_barcode.complete"), s);
    }
    if (!undetermined_barcodes.empty()) {
      const auto v = sorted(undetermined_barcodes);
      std::string s;
      for (const auto& it : v) {  // rep_itr is NOT reset here (:647-655)
        s += it.first + "\t" + std::to_string(it.second) + "\n";
        if (++rep_itr == report_limit) break;
      }
      write_file(path_join(run.report_dir, main + "undetermined_barcode"), s);
      s.clear();
      for (const auto& it : v) s += it.first + "\t" + std::to_string(it.second) + "\n";
      write_file(path_join(run.report_dir, main + "undetermined_barcode.complete"), s);
    }
  }
}

void ReportData::write_sample_report(const RunInfo& run, const SampleSheet& sheet, int reporting_level) const {
  if (reporting_level < 1) return;   // :567: nothing else is written for one sample (the barcode histograms are empty)
  const auto& row = sample_statistics[0];
  size_t width = row.size();
  for (size_t i = row.size(); i-- > 11;) {   // :504-512
    if (row[i] != 0) break;
    width = i;
  }
  const int max_mismatches = (int)width - 10;
  std::string out =
      "job_number\tsample_id\tr1_qc_30\tr2_qc_30\tr3_qc_30\tr1_bases\tr2_bases\tr3_bases\tr1_qc\tr2_qc\tr3_qc\tall_reads";
  for (int cnt = 0; cnt < max_mismatches; ++cnt) out += "\t" + std::to_string(cnt) + "-mismatches";
  out += "\n";
  out += sheet.info[0][PROJECT_ID_COLUMN] + "\t" + sheet.info[0][SAMPLE_COLUMN];
  for (size_t c = 0; c < width; ++c) out += "\t" + std::to_string(row[c]);
  out += "\n";
  const std::string path = path_join(run.report_dir, sheet.info[0][SAMPLE_COLUMN] + ".mgikit.sample_stats");
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw Fatal("couldn't create output");
  fwrite(out.data(), 1, out.size(), f);
  fclose(f);
}

}  // namespace mgikit

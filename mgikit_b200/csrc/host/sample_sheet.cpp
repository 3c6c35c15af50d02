#include "sample_sheet.hpp"

#include <algorithm>
#include <fstream>
#include <sstream>
#include <sys/stat.h>

namespace mgikit {

namespace {

std::string trim(const std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

std::string lower(std::string s) {
  for (auto& c : s) c = (char)tolower((unsigned char)c);
  return s;
}

std::vector<std::string> split_trim(const std::string& line, char delim) {
  std::vector<std::string> out;
  size_t start = 0;
  for (;;) {
    size_t p = line.find(delim, start);
    if (p == std::string::npos) {
      out.push_back(trim(line.substr(start)));
      break;
    }
    out.push_back(trim(line.substr(start, p - start)));
    start = p + 1;
  }
  return out;
}

bool all_acgt(const std::string& s) {
  for (char c : s)
    if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
  return true;
}

bool is_file(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

const std::string& col(const std::vector<std::string>& vals, size_t i) {
  if (i >= vals.size()) throw Fatal("Sample sheet row has fewer columns than its header!");
  return vals[i];
}

}  // namespace

// load_sample_sheet(), src/sample_manager.rs:290-456
std::vector<std::array<std::string, 7>> load_sample_sheet(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Fatal("Sample sheet file is invalid!");
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string content = ss.str();

  std::vector<std::array<std::string, 7>> rows;
  std::vector<std::string> header;
  const size_t NONE = (size_t)-1;
  size_t c_id = NONE, c_tpl = NONE, c_i7 = NONE, c_i5 = NONE, c_i7rc = NONE, c_i5rc = NONE, c_prj = NONE;
  char delim = '\t';

  size_t pos = 0;
  while (pos < content.size()) {
    size_t nl = content.find('\n', pos);
    std::string line = content.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
    pos = nl == std::string::npos ? content.size() : nl + 1;
    if (!line.empty() && line.back() == '\r') line.pop_back();  // str::lines()
    if (trim(line).size() < 5) continue;
    if (header.empty()) {
      header = split_trim(lower(line), '\t');
      if (header.size() < 2) {
        header = split_trim(lower(line), ',');
        delim = ',';
        if (header.size() < 2) throw Fatal("Sample sheet columns should be separated by ',' or '\t'!");
      }
      for (size_t i = 0; i < header.size(); ++i) {
        if (header[i] == "sample_id") c_id = i;
        else if (header[i] == "template") c_tpl = i;
        else if (header[i] == "i7") c_i7 = i;
        else if (header[i] == "i5") c_i5 = i;
        else if (header[i] == "i7_rc") c_i7rc = i;
        else if (header[i] == "i5_rc") c_i5rc = i;
        else if (header[i] == "job_number") c_prj = i;
      }
      continue;
    }
    const auto vals = split_trim(line, delim);
    std::array<std::string, 7> r;
    if (c_id == NONE || col(vals, c_id).empty())
      throw Fatal("sample_id column is mandatory in the samplesheet and must not be an empty string!");
    r[SAMPLE_COLUMN] = col(vals, c_id);
    if (c_i7 == NONE) {
      r[I7_COLUMN] = ".";
    } else {
      if (!all_acgt(col(vals, c_i7)))
        throw Fatal("Index must only contain A, C, G and T letters! found '" + col(vals, c_i7) + "'");
      r[I7_COLUMN] = col(vals, c_i7);
    }
    r[I5_COLUMN] = c_i5 == NONE ? "." : col(vals, c_i5);
    r[TEMPLATE_COLUMN] = c_tpl == NONE ? "." : col(vals, c_tpl);
    auto rc_col = [&](size_t c, const char* name) -> std::string {
      if (c == NONE) return ".";
      const std::string& v = col(vals, c);
      if (v != "." && v != "0" && v != "1")
        throw Fatal(std::string(name) + " must be either '.', '0' or '1' when reverse complementary! found '" + v + "'");
      return v;
    };
    r[I7_RC_COLUMN] = rc_col(c_i7rc, "i7_rc");
    r[I5_RC_COLUMN] = rc_col(c_i5rc, "i5_rc");
    r[PROJECT_ID_COLUMN] = c_prj == NONE ? "." : col(vals, c_prj);
    if (r[I7_COLUMN] == "." || r[I7_COLUMN].size() < 3)
      throw Fatal("i7 (" + r[I7_COLUMN] + ") should be longer than 3 chars!");
    if (r[I5_COLUMN] != ".") {
      if (r[I5_COLUMN].size() < 3) throw Fatal("i5 (" + r[I5_COLUMN] + ") should be longer than 3 chars!");
      if (!all_acgt(r[I5_COLUMN]))
        throw Fatal("Index must only contain A, C, G and T letters! found '" + r[I5_COLUMN] + "'");
    }
    rows.push_back(r);
  }
  if (rows.empty()) throw Fatal("Sample sheet seems to be empty! No sample is found!");
  return rows;
}

namespace {

int find_or_add(std::vector<std::string>& v, const std::string& s) {
  for (size_t i = 0; i < v.size(); ++i)
    if (v[i] == s) return (int)i;
  v.push_back(s);
  return (int)v.size() - 1;
}

// extract_templates_information(), src/sample_manager.rs:506-690
std::vector<TemplateData> extract_templates(const std::vector<std::array<std::string, 7>>& info,
                                            const std::string& tmpl, bool i7_rc, bool i5_rc) {
  std::vector<TemplateData> ls;  // first-appearance order (the reference uses a HashMap)
  int barcode_length = 0;
  for (size_t s = 0; s < info.size(); ++s) {
    const auto& r = info[s];
    const std::string curr_template = tmpl.empty() ? r[TEMPLATE_COLUMN] : tmpl;
    if ((r[I5_COLUMN] == "." || r[I5_COLUMN].size() < 3) && curr_template.find("i5") != std::string::npos)
      throw Fatal("i5 (" + r[I5_COLUMN] + ") should be longer than 3 chars! or the template should not contains i5");
    const bool check_i5 = r[I5_COLUMN].size() > 1;
    const std::string i7 = ((!tmpl.empty() && i7_rc) || (tmpl.empty() && r[I7_RC_COLUMN] == "1"))
                               ? reverse_complement(r[I7_COLUMN]) : r[I7_COLUMN];
    std::string i5;
    if (check_i5)
      i5 = ((!tmpl.empty() && i5_rc) || (tmpl.empty() && r[I5_RC_COLUMN] == "1"))
               ? reverse_complement(r[I5_COLUMN]) : r[I5_COLUMN];

    TemplateData* T = nullptr;
    for (auto& t : ls)
      if (t.tmpl == curr_template) T = &t;
    if (!T) {
      TemplateData t;
      t.tmpl = curr_template;
      t.has_i5 = check_i5;
      t.geometry = parse_template(curr_template);
      if (!ls.empty()) {
        if (t.geometry[9] != barcode_length) throw Fatal("The barcode length should be the same for all samples!");
      } else {
        barcode_length = t.geometry[9];
      }
      ls.push_back(t);
      T = &ls.back();
    }
    T->count += 1;
    const int a = find_or_add(T->i7s, i7);
    const int b = check_i5 ? find_or_add(T->i5s, i5) : -1;
    for (const auto& row : T->rows) {
      if (row[0] != a) continue;
      if (check_i5) {
        if (row[1] == b) throw Fatal("Two samples having the same indexes! i7: " + i7 + " and i5: " + i5);
      } else {
        throw Fatal("Two samples having the same i7 indexes! i7: " + i7);
      }
    }
    T->rows.push_back({a, b, (int)s});
  }
  for (auto& T : ls) {
    // An index whose length differs from the template's can never equal the
    // bytes cut from a read (the reference's dictionary lookup simply misses),
    // so it is left out of the tables; its sample just receives no reads.
    auto compact = [&](std::vector<std::string>& v, int want, int col) {
      std::vector<int> remap(v.size(), -1);
      std::vector<std::string> kept;
      for (size_t i = 0; i < v.size(); ++i)
        if ((int)v[i].size() == want) {
          remap[i] = (int)kept.size();
          kept.push_back(v[i]);
        }
      v.swap(kept);
      std::vector<std::array<int, 3>> rows;
      for (auto row : T.rows) {
        if (row[col] >= 0) {
          row[col] = remap[(size_t)row[col]];
          if (row[col] < 0) continue;
        }
        rows.push_back(row);
      }
      T.rows.swap(rows);
    };
    compact(T.i7s, T.geometry[1], 0);
    if (T.has_i5) compact(T.i5s, T.geometry[4], 1);
    const size_t n7 = T.i7s.size(), n5 = T.i5s.size();
    if (T.has_i5) {
      T.pair_sample.assign(n7 * n5, -1);
      for (const auto& row : T.rows)
        if (row[1] >= 0) T.pair_sample[(size_t)row[0] * n5 + (size_t)row[1]] = row[2];
    } else {
      // (sample_itr, {}) or the placeholder (1000, {i5: sample}) — src/sample_manager.rs:621-627
      T.pair_sample.assign(n7, -1);
      for (const auto& row : T.rows)
        if (T.pair_sample[(size_t)row[0]] < 0) T.pair_sample[(size_t)row[0]] = row[1] >= 0 ? 1000 : row[2];
      T.i5s.clear();
    }
  }
  // src/sample_manager.rs:673-687: by sample count, descending.  Ties are in
  // HashMap order in the reference (unspecified); we keep first appearance.
  std::stable_sort(ls.begin(), ls.end(), [](const TemplateData& a, const TemplateData& b) { return a.count > b.count; });
  return ls;
}

}  // namespace

std::string reverse_complement(const std::string& seq) {
  std::string out;
  out.reserve(seq.size());
  for (auto it = seq.rbegin(); it != seq.rend(); ++it) {
    switch (*it) {
      case 'n': case 'N': out.push_back('N'); break;
      case 'A': case 'a': out.push_back('T'); break;
      case 'T': case 't': out.push_back('A'); break;
      case 'C': case 'c': out.push_back('G'); break;
      case 'G': case 'g': out.push_back('C'); break;
      default: throw Fatal("Wrong neucltide!");
    }
  }
  return out;
}

std::array<int, 10> parse_template(const std::string& tmpl) {
  std::array<int, 10> g{};
  const std::string msg = "template (" + tmpl + ") does not match the expected format as explianed in mgikit documenation!";
  int shift = 0;
  size_t end = tmpl.size();
  for (;;) {  // rsplit(':')
    size_t p = end == 0 ? std::string::npos : tmpl.rfind(':', end - 1);
    const std::string item = tmpl.substr(p == std::string::npos ? 0 : p + 1, end - (p == std::string::npos ? 0 : p + 1));
    if (item.size() < 2) throw Fatal(msg);
    const std::string kind = item.substr(0, 2);
    if (kind != "i7" && kind != "i5" && kind != "um" && kind != "--") throw Fatal(msg);
    const std::string num = item.substr(2);
    if (num.empty() || num.find_first_not_of("0123456789") != std::string::npos || num.size() > 6) throw Fatal(msg);
    const int length = std::stoi(num);
    shift += length;
    if (kind == "i7") { g[0] = 1; g[1] = length; g[2] = shift; }
    else if (kind == "i5") { g[3] = 1; g[4] = length; g[5] = shift; }
    else if (kind == "um") { g[6] = 1; g[7] = length; g[8] = shift; }
    if (p == std::string::npos) break;
    end = p;
  }
  g[9] = shift;
  return g;
}

SampleSheet load_sample_manager(const std::string& path, const std::string& tmpl, bool i7_rc, bool i5_rc,
                                const std::string& undetermined_label, const std::string& ambiguous_label) {
  if (!is_file(path)) throw Fatal("Sample sheet file is invalid!");
  SampleSheet sh;
  sh.info = load_sample_sheet(path);

  {  // get_writing_unique_samples(), src/sample_manager.rs:692-730
    std::map<std::string, int> first;
    int dup_ids = 0;
    for (size_t i = 0; i < sh.info.size(); ++i) {
      auto it = first.find(sh.info[i][SAMPLE_COLUMN]);
      if (it != first.end()) {
        sh.writing_samples.push_back(it->second);
        sh.unique_samples_ids.push_back(sh.unique_samples_ids[(size_t)it->second]);
      } else {
        first.emplace(sh.info[i][SAMPLE_COLUMN], (int)i);
        sh.writing_samples.push_back((int)i);
        sh.unique_samples_ids.push_back(dup_ids++);
      }
    }
    for (int k = 0; k < 2; ++k) {
      sh.writing_samples.push_back((int)sh.writing_samples.size());
      sh.unique_samples_ids.push_back((int)sh.unique_samples_ids.size());
    }
  }
  sh.templates = extract_templates(sh.info, tmpl, i7_rc, i5_rc);
  sh.info.push_back({undetermined_label, "", "", "", "", "", "."});
  sh.info.push_back({ambiguous_label, "", "", "", "", "", "."});
  // extract_project_samples(), src/sample_manager.rs:458-504
  sh.project_samples["."] = {};
  for (size_t i = 0; i < sh.info.size(); ++i) sh.project_samples[sh.info[i][PROJECT_ID_COLUMN]].push_back((int)i);
  sh.project_samples["."].clear();
  return sh;
}

void build_plan(const SampleSheet& sheet, const RunFlags& f, GpuPlan& plan) {
  const size_t nt = sheet.templates.size();
  if (nt > MGK_MAX_TEMPLATES) throw Fatal("Too many barcode templates in the sample sheet for the GPU path!");
  plan.tpls.assign(nt, mgk_template{});
  plan.i7_flat.assign(nt, "");
  plan.i5_flat.assign(nt, "");
  plan.pairs.assign(nt, {});
  for (size_t t = 0; t < nt; ++t) {
    const TemplateData& T = sheet.templates[t];
    if (T.geometry[1] > MGK_MAX_INDEX_LEN || T.geometry[4] > MGK_MAX_INDEX_LEN || T.geometry[9] > MGK_MAX_BARCODE_LEN)
      throw Fatal("Index longer than 32 bases or barcode longer than 255 bases is not supported by the GPU path!");
    for (const auto& s : T.i7s) plan.i7_flat[t] += s;
    for (const auto& s : T.i5s) plan.i5_flat[t] += s;
    plan.pairs[t].assign(T.pair_sample.begin(), T.pair_sample.end());
    mgk_template& o = plan.tpls[t];
    for (int k = 0; k < 10; ++k) o.geometry[k] = T.geometry[k];
    o.has_i5 = T.has_i5;
    o.n_i7 = (int)T.i7s.size();
    o.n_i5 = (int)T.i5s.size();
  }
  for (size_t t = 0; t < nt; ++t) {  // pointers after all storage is final
    plan.tpls[t].i7 = plan.i7_flat[t].data();
    plan.tpls[t].i5 = plan.i5_flat[t].data();
    plan.tpls[t].pair_sample = plan.pairs[t].data();
  }
  plan.writing.assign(sheet.writing_samples.begin(), sheet.writing_samples.end());
  plan.prefix = f.illumina_prefix;
  mgk_config& c = plan.cfg;
  c = mgk_config{};
  c.abi_version = MGK_ABI_VERSION;
  c.paired = f.paired;
  c.read2_has_sequence = f.read2_has_sequence;
  c.out_mode = f.out_mode;
  c.keep_barcode = f.keep_barcode;
  c.comprehensive_scan = f.comprehensive_scan;
  c.validate = f.validate;
  c.report_level = f.report_level;
  c.mgi_data = f.mgi_data;
  c.allowed_mismatches = f.mismatches;
  c.total_allowed_mismatches = f.per_index_error ? 2 * f.mismatches : f.mismatches;  // src/lib.rs:1500-1504, :1813-1815
  c.l_position = f.l_position;
  c.illumina_prefix = plan.prefix.c_str();
  c.n_samples = sheet.n_samples();
  c.writing_sample = plan.writing.data();
  c.n_templates = (int)nt;
  c.templates = plan.tpls.data();
  c.n_slots = 2;
}

}  // namespace mgikit

#include "run_info.hpp"

#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <cstdio>
#include <ctime>
#include <fstream>
#include <vector>

#include "log.hpp"

namespace mgikit {

bool path_is_file(const std::string& p) {
  struct stat st;
  return !p.empty() && stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool path_is_dir(const std::string& p) {
  struct stat st;
  return !p.empty() && stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
void create_dir_all(const std::string& p) {
  log_info("A directory is created: " + p);
  std::string cur;
  for (size_t i = 0; i <= p.size(); ++i) {
    if (i == p.size() || p[i] == '/') {
      if (!cur.empty() && !path_is_dir(cur)) mkdir(cur.c_str(), 0777);
    }
    if (i < p.size()) cur.push_back(p[i]);
  }
  if (!path_is_dir(p)) throw Fatal("can not create directory " + p);
}
std::string path_join(const std::string& dir, const std::string& name) {
  if (dir.empty()) return name;
  if (dir.back() == '/') return dir + name;
  return dir + "/" + name;
}

namespace {

std::string file_name_of(const std::string& p) {
  size_t s = p.find_last_of('/');
  return s == std::string::npos ? p : p.substr(s + 1);
}

std::string now_fmt(const char* fmt) {
  char buf[64];
  time_t t = time(nullptr);
  struct tm tmv;
  localtime_r(&t, &tmv);
  strftime(buf, sizeof buf, fmt, &tmv);
  return buf;
}

void check_file(const std::string& p) {
  if (!path_is_file(p)) throw Fatal("File is not accessible: " + p);
}

std::string trim(const std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  return s.substr(b, e - b);
}

std::string replace_all(std::string s, const std::string& what, const std::string& with) {
  size_t p = 0;
  while ((p = s.find(what, p)) != std::string::npos) {
    s.replace(p, what.size(), with);
    p += with.size();
  }
  return s;
}

bool starts_with(const std::string& s, const char* pre) { return s.rfind(pre, 0) == 0; }

// get_lane_from_file(), src/run_manager.rs:347-356
std::string get_lane_from_file(const std::string& path, size_t loc, char sep) {
  const std::string name = file_name_of(path);
  std::vector<std::string> parts;
  size_t start = 0;
  for (;;) {
    size_t p = name.find(sep, start);
    parts.push_back(name.substr(start, p == std::string::npos ? std::string::npos : p - start));
    if (p == std::string::npos) break;
    start = p + 1;
  }
  if (parts.size() > loc && starts_with(parts[loc], "L0")) return parts[loc];
  return "";
}

// validate_and_assigne_input_reads(), src/run_manager.rs:358-385
void assign_input_reads(const std::string& r1, const std::string& r2, std::string& paired, std::string& barcode) {
  paired.clear();
  if (r2.empty()) {
    barcode = r1;
    log_info("Single end read input was detected!");
    log_info("Read with Barcode or R1: " + barcode);
  } else {
    paired = r1;
    barcode = r2;
    log_info("Paired ended read input was detected!");
    log_info("Paired read or R1: " + paired);
    log_info("Read with Barcode or R2: " + barcode);
    if (paired.empty()) throw Fatal("Input reads are invalid! For single end fastq, use -f or -i parameters!");
    check_file(paired);
  }
  if (barcode.empty()) throw Fatal("Input reads are invalid! check the path " + barcode);
  check_file(barcode);
}

// get_read_files_from_input_dir(), src/run_manager.rs:420-459
void read_files_from_dir(const std::string& dir, const std::string& r1_suf, const std::string& r2_suf,
                         std::string& paired, std::string& barcode) {
  DIR* d = opendir(dir.c_str());
  if (!d) throw Fatal("can not read directory content!");
  std::string r1, r2;
  int found = 0;
  auto ends_with = [](const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
  };
  while (struct dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    if (name == "." || name == "..") continue;
    if (ends_with(name, r1_suf)) {
      r1 = path_join(dir, name);
      ++found;
    } else if (ends_with(name, r2_suf)) {
      r2 = path_join(dir, name);
      ++found;
    }
    if (found > 1) break;
  }
  closedir(d);
  if (found == 0)
    throw Fatal("Can not find files that ends with " + dir + " or " + r1_suf + " under the directory " + r2_suf);
  assign_input_reads(r1, r2, paired, barcode);
}

// find_info_file(), src/run_manager.rs:461-483 (input_dir argument is always empty at the call site :85-89)
std::string find_info_file(const std::string& info_file_arg, const std::string& r_file) {
  if (path_is_file(info_file_arg)) return info_file_arg;
  if (path_is_file(r_file)) {
    size_t s = r_file.find_last_of('/');
    const std::string cand = (s == std::string::npos ? std::string() : r_file.substr(0, s + 1)) + "BioInfo.csv";
    struct stat st;
    if (stat(cand.c_str(), &st) == 0) return cand;
  }
  return "";
}

// parse_info_file(), src/run_manager.rs:485-513
void parse_info_file(const std::string& path, std::string& instrument, std::string& run) {
  instrument.clear();
  run.clear();
  if (!path_is_file(path)) return;
  std::ifstream f(path);
  std::string line;
  auto second = [](const std::string& s) {
    size_t a = s.find(',');
    if (a == std::string::npos) throw Fatal("BioInfo.csv line without a value: " + s);
    size_t b = s.find(',', a + 1);
    return trim(s.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1));
  };
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (starts_with(line, "Machine ID,")) {
      instrument = second(line);
    } else if (starts_with(line, "Sequence Date") || starts_with(line, "Sequence Start Date")) {
      run = replace_all(second(line), "-", "");
    } else if (starts_with(line, "Sequence Time") || starts_with(line, "Sequence Start Time")) {
      run += replace_all(second(line), ":", "");
    }
  }
}

// prepare_output_report_dir(), src/run_manager.rs:515-571
void prepare_dirs(const std::string& out_arg, const std::string& rep_arg, bool force, std::string& out_dir,
                  std::string& rep_dir) {
  out_dir = out_arg.empty() ? now_fmt("mgiKit_%Y%m%dT%H%M%S") : out_arg;
  bool same;
  if (rep_arg.empty()) {
    log_info("The same output directory will be used for reports.");
    same = true;
    rep_dir = out_dir;
  } else {
    same = false;
    rep_dir = rep_arg;
  }
  if (path_is_dir(out_dir)) {
    if (!force) throw Fatal("Output directly exists. Use --froce to overwrite their data: " + out_dir);
    log_info("Output directory exists. Data will be overwritten at: " + out_dir + ".");
  } else {
    create_dir_all(out_dir);
  }
  if (path_is_dir(rep_dir)) {
    if (!same) {
      if (!force) throw Fatal("Report directly exists. Use --froce to overwrite their data: " + rep_dir);
      log_info("Report directory exists. Data will be overwritten at: " + rep_dir + ".");
    }
  } else {
    create_dir_all(rep_dir);
  }
}

// get_read_parts(), src/run_manager.rs:387-398 — first four lines, newline kept
void first_record(const std::string& path, std::string parts[4]) {
  gzFile g = gzopen(path.c_str(), "rb");
  if (!g) throw Fatal("Could not open the file");
  char buf[1 << 16];
  for (int i = 0; i < 4; ++i) {
    parts[i].clear();
    for (;;) {
      if (!gzgets(g, buf, sizeof buf)) break;
      parts[i] += buf;
      if (!parts[i].empty() && parts[i].back() == '\n') break;
    }
  }
  gzclose(g);
}

}  // namespace

RunInfo make_run_info(const RunArgs& a) {
  RunInfo r;
  if (!a.input_dir.empty()) {
    log_info("Input directory: " + a.input_dir);
    read_files_from_dir(a.input_dir, a.r1_suffix, a.r2_suffix, r.paired_reads, r.barcode_reads);
  } else {
    assign_input_reads(a.read1, a.read2, r.paired_reads, r.barcode_reads);
  }
  std::string fin_instrument, fin_run;
  if (a.instrument.empty() || a.run.empty())
    parse_info_file(find_info_file(a.info_file, r.barcode_reads), fin_instrument, fin_run);
  if (!a.instrument.empty()) fin_instrument = a.instrument;
  if (!a.run.empty()) fin_run = a.run;
  r.lane = a.lane.empty() ? get_lane_from_file(r.barcode_reads, 1, '_') : a.lane;
  prepare_dirs(a.output_dir, a.report_dir, a.force, r.output_dir, r.report_dir);
  log_info("Output directory: " + r.output_dir);
  log_info("Reports directory: " + r.report_dir);
  log_info(std::string("MGI input fastq files:  ") + (a.mgi_data ? "true" : "false"));
  log_info(std::string("Validate fastq content: ") + (a.check_content ? "true" : "false"));
  log_info(std::string("Trim Barcode: ") + (!a.keep_barcode ? "true" : "false"));
  if (a.illumina_format && !a.mgi_data)
    throw Fatal("mgikit does not refomat output files in Illumina foramt unless the input fastq files are in MGI format! Disable `--not-mgi` or enable `--disable-illumina-format`");
  r.instrument = fin_instrument;
  r.run = fin_run;
  r.mgi_data = a.mgi_data;
  r.force = a.force;
  r.illumina_format = a.illumina_format;
  r.keep_barcode = a.keep_barcode;
  r.comprehensive_scan = a.comprehensive_scan;
  r.ignore_undetermined = a.ignore_undetermined;
  r.check_content = a.check_content;
  r.mgi_full_header = a.mgi_full_header;
  return r;
}

void RunInfo::get_read_information() {
  std::string p[4];
  first_record(barcode_reads, p);
  const size_t whole = p[0].size() + p[1].size() + p[2].size() + p[3].size();
  ReadInfo b;
  std::string header_lane;
  if (!mgi_data) {
    b.flowcell = now_fmt("%Y%m%dT%H%M%S");
    b.l_position = 0;
  } else {
    // get_flowcell_lane_info(), src/run_manager.rs:400-418: last 'L' of the header line
    const std::string& h = p[0];
    const size_t l = h.rfind('L');
    if (l == std::string::npos || l == 0) throw Fatal("Can not find the flowcell id in this header " + h + "!");
    b.flowcell = h.substr(1, l - 1);
    header_lane = h.substr(l + 1, 1);
    b.l_position = l;
    log_info("Detected flowcell from the header of the first read is " + b.flowcell + ".");
    log_info("Detected lane from the header of the first read is " + header_lane + ".");
    if (std::string("1234").find(header_lane) == std::string::npos)
      log_warn("The detected lane ( = " + header_lane + ") is not recognised! Expected 1, 2, 3 or 4!");
  }
  if (p[1].empty()) throw Fatal("can not read sequence!");
  b.sequence_length = p[1].size() - 1;
  b.read_length = whole;
  b.lane = "L0" + header_lane;
  if (p[2] != "+\n")
    throw Fatal("Expected read format is not satisified. You can try rerunning using --flexible parameter.");
  ReadInfo pr;
  if (path_is_file(paired_reads)) {  // self.paired_reads().exists()
    std::string q[4];
    first_record(paired_reads, q);
    if (q[1].empty()) throw Fatal("can not read sequence!");
    pr.read_length = q[0].size() + q[1].size() + q[2].size() + q[3].size();
    pr.sequence_length = q[1].size() - 1;
    if (q[2] != "+\n")
      throw Fatal("Expected read format is not satisified. You can try running --flexible command.");
  }
  paired_read_info = pr;
  barcode_read_info = b;
  log_info("The length of the read with barcode is: " + std::to_string(b.sequence_length));
  log_info("The length of the paired read is: " + std::to_string(pr.sequence_length));
}

void RunInfo::confirm_format() const {
  if (illumina_format) {
    log_info("Output format is Illumina");
    if (run.empty() || instrument.empty() || flowcell().empty()) {
      log_error("Missing information for Illumina header! instrument: " + instrument + ", lane: " + lane + ", run" + run +
                ", flowcell: " + flowcell());
      throw Fatal("Missing mandatory information!");
    }
  } else {
    log_info("Output format is MGI");
    if (mgi_full_header) log_info("Sample barcode and UMI if available will be written into the read header.");
  }
}

std::string RunInfo::create_illumina_header_prefix() const {
  if (instrument.empty() || lane.empty() || flowcell().empty() || run.empty()) {
    log_error("Missing information for Illumina header! instrument: " + instrument + ", lane: " + lane + ", run" + run +
              ", flowcell: " + flowcell());
    throw Fatal("Missing mandatory information!");
  }
  return "@" + instrument + ":" + run + ":" + flowcell() + ":";
}

}  // namespace mgikit

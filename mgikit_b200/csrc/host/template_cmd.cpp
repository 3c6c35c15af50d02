// `mgikit template`: barcode-template detection on the first reads of a lane, the
// pre-step of BASELINE config 4.  Host port of detect_template()
// (/root/reference/src/lib.rs:1948-2264), find_matches() (:72-138) and
// get_mgikit_template() (:140-221); goldens tests/integration_test.rs:73-160.
// 5-10 k reads: milliseconds on one core, so this stays on the host.
#include "template_cmd.hpp"

#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>

#include "log.hpp"
#include "run_info.hpp"
#include "sample_sheet.hpp"

namespace mgikit {

namespace {

// str::match_indices: left to right, the search resumes AFTER a match (no overlaps)
std::vector<size_t> match_indices(const std::string& hay, const std::string& needle) {
  std::vector<size_t> out;
  if (needle.empty()) {   // Rust yields every char boundary for ""; never reached: indexes are >= 3 long
    for (size_t i = 0; i <= hay.size(); ++i) out.push_back(i);
    return out;
  }
  for (size_t p = hay.find(needle); p != std::string::npos; p = hay.find(needle, p + needle.size())) out.push_back(p);
  return out;
}

// histogram: template key -> matches per sample; keys keep their first-seen rank so that ties are broken the
// same way on every run (the reference iterates a HashMap: its order among ties is arbitrary)
struct MatchStats {
  std::map<std::string, size_t> index;
  std::vector<std::string> keys;
  std::vector<std::vector<size_t>> counts;
  size_t n_samples;
  void add(const std::string& key, size_t sample) {   // add_match(), src/lib.rs:54-70
    auto it = index.find(key);
    if (it == index.end()) {
      it = index.emplace(key, keys.size()).first;
      keys.push_back(key);
      counts.emplace_back(n_samples, 0);
    }
    counts[it->second][sample] += 1;
  }
};

// find_matches(), src/lib.rs:72-138
void find_matches(size_t sample, MatchStats& st, const std::string& bar, const std::array<std::string, 4>& idx) {
  const size_t n_idx = idx[2].size() < 2 ? 2 : 4;
  for (size_t i = 0; i < 2; ++i) {
    const std::vector<size_t> first = match_indices(bar, idx[i]);
    if (first.empty()) continue;
    for (size_t j = 2; j < n_idx; ++j) {
      const std::vector<size_t> second = match_indices(bar, idx[j]);
      for (size_t s : second)
        for (size_t f : first)
          if (s >= f + idx[0].size() || f >= s + idx[2].size())
            st.add(std::to_string(f) + ":" + std::to_string(s) + "_" + std::to_string(i) + ":" + std::to_string(j - 2), sample);
    }
    if (n_idx == 2)
      for (size_t f : first) st.add(std::to_string(f) + "_" + std::to_string(i) + ":.", sample);
  }
}

struct TemplateInfo {
  std::string tmpl, i7_rc, i5_rc;
};

std::vector<std::string> split(const std::string& s, char c) {
  std::vector<std::string> out;
  size_t p = 0;
  for (;;) {
    const size_t q = s.find(c, p);
    out.push_back(s.substr(p, q == std::string::npos ? std::string::npos : q - p));
    if (q == std::string::npos) break;
    p = q + 1;
  }
  return out;
}

// get_mgikit_template(), src/lib.rs:140-221: "pos7[:pos5]_rc7:rc5" -> "--a:i7n:--b:i5m:--c" with the longest gap as UMI
TemplateInfo get_mgikit_template(const std::string& initial, bool umi, size_t barcode_length, size_t i7_len, size_t i5_len) {
  const auto elements = split(initial, '_');
  const auto rc = split(elements[1], ':');
  std::vector<size_t> pos;
  for (const auto& p : split(elements[0], ':')) pos.push_back((size_t)std::stoull(p));
  std::vector<std::string> items;
  size_t shift = 0;
  const bool no_swap = pos.size() == 1 || pos[0] < pos[1];
  std::sort(pos.begin(), pos.end());
  size_t umi_loc = 0, umi_size = 0, added = 0;
  if (pos[0] > 0) {
    items.push_back("--" + std::to_string(pos[0]));
    shift += pos[0];
    umi_size = pos[0];
    ++added;
  }
  if (no_swap) {
    items.push_back("i7" + std::to_string(i7_len));
    shift += i7_len;
  } else {
    items.push_back("i5" + std::to_string(i5_len));
    shift += i5_len;
  }
  ++added;
  if (pos.size() == 2) {
    if (shift < pos[1]) {
      items.push_back("--" + std::to_string(pos[1] - shift));
      if (pos[1] - shift > umi_size) {
        umi_loc = added;
        umi_size = pos[1] - shift;
      }
      shift = pos[1];
      ++added;
    }
    if (no_swap) {
      items.push_back("i5" + std::to_string(i5_len));
      shift += i5_len;
    } else {
      items.push_back("i7" + std::to_string(i7_len));
      shift += i7_len;
    }
    ++added;
  }
  if (shift < barcode_length) {
    items.push_back("--" + std::to_string(barcode_length - shift));
    if (barcode_length - shift > umi_size) umi_loc = added;
  }
  if (umi) {   // String::replace("--", "um") on that item
    std::string& it = items[umi_loc];
    for (size_t p = it.find("--"); p != std::string::npos; p = it.find("--", p + 2)) it.replace(p, 2, "um");
  }
  TemplateInfo out;
  for (size_t i = 0; i < items.size(); ++i) out.tmpl += (i ? ":" : "") + items[i];
  out.i7_rc = rc[0];
  out.i5_rc = rc.size() > 1 ? rc[1] : "";
  return out;
}

void write_file(const std::string& path, const std::string& content) {   // src/file_utils.rs:411-416
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw Fatal("couldn't create output");
  const size_t w = fwrite(content.data(), 1, content.size(), f);
  fclose(f);
  if (w != content.size()) throw Fatal("couldn't create output");
}

// BufRead::read_line over the (multi-member) gzip stream; false at end of file
bool read_line(gzFile g, std::string& line) {
  line.clear();
  char buf[1 << 14];
  for (;;) {
    if (!gzgets(g, buf, sizeof buf)) return !line.empty();
    line += buf;
    if (line.back() == '\n') return true;
  }
}

}  // namespace

void run_template(const TemplateArgs& a) {
  const auto t0 = std::chrono::steady_clock::now();
  // SampleManager::from_simple_sheet(), src/sample_manager.rs:115-131
  if (!path_is_file(a.sample_sheet)) throw Fatal("Sample sheet file is invalid!");
  const auto info = load_sample_sheet(a.sample_sheet);
  log_info(std::to_string(info.size()) + " Samples were found in the input sample sheet.");

  // Path::file_name / Path::parent of the -o prefix (src/lib.rs:1958-1966)
  std::string prefix = a.output;
  while (prefix.size() > 1 && prefix.back() == '/') prefix.pop_back();
  const size_t slash = prefix.find_last_of('/');
  const std::string out_name = slash == std::string::npos ? prefix : prefix.substr(slash + 1);
  const std::string parent = slash == std::string::npos ? "" : (slash == 0 ? "/" : prefix.substr(0, slash));

  RunArgs ra;   // RunManager::new(..., true, true, ..., true, false, true, true, true, false), src/lib.rs:1968-2003
  ra.input_dir = a.input_dir;
  ra.read1 = a.read1;
  ra.read2 = a.read2;
  ra.output_dir = parent;
  ra.mgi_data = true;
  ra.force = true;
  ra.r1_suffix = a.r1_suffix;
  ra.r2_suffix = a.r2_suffix;
  ra.info_file = a.info_file;
  ra.illumina_format = true;
  ra.keep_barcode = false;
  ra.comprehensive_scan = true;
  ra.ignore_undetermined = true;
  ra.check_content = true;
  ra.mgi_full_header = false;
  RunInfo run = make_run_info(ra);
  run.get_read_information();

  // get_samples_indices(), src/sample_manager.rs:224-246
  std::vector<std::array<std::string, 4>> idx;
  for (const auto& r : info) {
    if (r[I5_COLUMN] != ".")
      idx.push_back({r[I7_COLUMN], reverse_complement(r[I7_COLUMN]), r[I5_COLUMN], reverse_complement(r[I5_COLUMN])});
    else
      idx.push_back({r[I7_COLUMN], reverse_complement(r[I7_COLUMN]), "", ""});
  }

  size_t barcode_length;
  if (a.barcode_length > 0) {
    barcode_length = a.barcode_length;
  } else {
    if (run.barcode_read_info.sequence_length < run.paired_read_info.sequence_length)
      throw Fatal("attempt to subtract with overflow");   // usize underflow in the reference (src/lib.rs:2025)
    barcode_length = run.barcode_read_info.sequence_length - run.paired_read_info.sequence_length;
    log_info("Barcode length is calculated as the difference between R2 length and R1.");
    size_t max_len = 0;
    for (const auto& it : idx) max_len = std::max(max_len, it[0].size() + it[2].size());
    if (barcode_length > max_len + a.max_umi_length)
      throw Fatal("The difference in read length is " + std::to_string(barcode_length) +
                  ". It is greater than the the length of indexes and possible UMI " + std::to_string(max_len + a.max_umi_length) +
                  ". You need to prvide barcode length for this run!");
  }
  log_info("Barcode length: " + std::to_string(barcode_length));

  MatchStats st;
  st.n_samples = idx.size();
  size_t reads = 0;
  {
    gzFile g = gzopen(run.barcode_reads.c_str(), "rb");
    if (!g) throw Fatal("Could not open the file");
    gzbuffer(g, 1u << 18);
    std::string line, seq;
    while (reads < a.testing_reads) {
      if (!read_line(g, line)) break;
      read_line(g, seq);
      if (seq.size() < barcode_length + 1) {
        gzclose(g);
        throw Fatal("a read is shorter than the barcode length");   // slice panic in the reference (src/lib.rs:2069-2071)
      }
      const std::string bar = seq.substr(seq.size() - barcode_length - 1, barcode_length);
      for (size_t s = 0; s < idx.size(); ++s) find_matches(s, st, bar, idx[s]);
      read_line(g, line);
      read_line(g, line);
      ++reads;
    }
    gzclose(g);
  }

  // per sample: its templates by matches, descending (stable: first-seen order among ties)
  std::vector<std::vector<std::pair<std::string, size_t>>> per_sample(info.size());
  for (size_t s = 0; s < info.size(); ++s) {
    for (size_t k = 0; k < st.keys.size(); ++k)
      if (st.counts[k][s] > 0) per_sample[s].push_back({st.keys[k], st.counts[k][s]});
    if (per_sample[s].empty()) {
      log_warn("Sample (" + info[s][SAMPLE_COLUMN] + ") has no matches with any possible template! might be an issue with its indexes!");
      log_warn("Consder increasing the '--testing-reads' parameter if the indexes are correct!");
    }
    std::stable_sort(per_sample[s].begin(), per_sample[s].end(), [](const auto& x, const auto& y) { return x.second > y.second; });
  }
  std::string popular;
  size_t popular_cnt = 0;
  for (size_t k = 0; k < st.keys.size(); ++k) {
    size_t cnt = 0;
    for (size_t v : st.counts[k]) cnt += v;
    if (cnt > popular_cnt) {
      popular = st.keys[k];
      popular_cnt = cnt;
    }
  }
  if (popular_cnt == 0) throw Fatal("Something wrong, there is no matches with any sample!");
  if (a.popular_template) {
    const TemplateInfo t = get_mgikit_template(popular, a.add_umi, barcode_length, idx[0][0].size(), idx[0][2].size());
    log_info("The most frequent template is " + t.tmpl + " i7_rc is " + t.i7_rc + " and i5_rc is " + t.i5_rc + " - Appeared " +
             std::to_string(popular_cnt) + " times!");
  }

  std::string out = "sample_id\ti7\ti5\tjob_number\ttemplate\ti7_rc\ti5_rc\n";
  std::string full = "sample_id\ti7\ti5\tall-matches\n";
  for (size_t s = 0; s < info.size(); ++s) {
    const auto& r = info[s];
    out += r[SAMPLE_COLUMN] + "\t" + r[I7_COLUMN] + "\t" + r[I5_COLUMN] + "\t" + r[PROJECT_ID_COLUMN] + "\t";
    TemplateInfo t;
    if (a.popular_template) {
      if (!per_sample[s].empty() && popular != per_sample[s][0].first)
        log_info("Using popular template for sample " + r[SAMPLE_COLUMN] + " instead of its detected template " + per_sample[s][0].first + "!");
      else if (per_sample[s].empty())
        log_info("Using popular template for sample " + r[SAMPLE_COLUMN] + " as no matches were found!");
      // lengths of the sheet's own strings: a single-index row carries "." here (src/lib.rs:2185-2191)
      t = get_mgikit_template(popular, a.add_umi, barcode_length, r[I7_COLUMN].size(), r[I5_COLUMN].size());
    } else if (!per_sample[s].empty()) {
      t = get_mgikit_template(per_sample[s][0].first, a.add_umi, barcode_length, idx[s][0].size(), idx[s][2].size());
    } else {
      t = {"No matches with this sample", ".", "."};
    }
    out += t.tmpl + "\t" + t.i7_rc + "\t" + t.i5_rc + "\n";
    full += r[SAMPLE_COLUMN] + "\t" + r[I7_COLUMN] + "\t" + r[I5_COLUMN] + "\t";
    if (!per_sample[s].empty()) {
      for (size_t k = 0; k < per_sample[s].size(); ++k)
        full += (k ? "\t" : "") + per_sample[s][k].first + " (" + std::to_string(per_sample[s][k].second) + ")";
    } else {
      full += "No matches found!";
    }
    full += "\n";
  }
  write_file(run.output_dir + "/" + out_name + "_template.tsv", out);
  write_file(run.output_dir + "/" + out_name + "_details.tsv", full);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  log_info(std::to_string(reads) + " reads were processed in " + std::to_string((long)secs) + " secs.");
}

}  // namespace mgikit

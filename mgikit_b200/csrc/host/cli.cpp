#include "cli.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "log.hpp"
#include "template_cmd.hpp"

namespace mgikit {

namespace {

const char* VERSION = "2.2.0-b200";

struct Opt {
  const char* long_name;
  char short_name;
  bool takes_value;
  const char* def;
  const char* alias;
};

// src/main.rs:93-399 (+ --gpus / --chunk-size, extensions of this build)
const Opt OPTS[] = {
    {"input", 'i', true, "", nullptr},
    {"read2", 'r', true, "", "2"},
    {"read1", 'f', true, "", "1"},
    {"sample-sheet", 's', true, nullptr, nullptr},
    {"output", 'o', true, "", nullptr},
    {"reports", 0, true, "", nullptr},
    {"mismatches", 'm', true, "1", nullptr},
    {"template", 0, true, "", nullptr},
    {"i7-rc", 0, false, "false", nullptr},
    {"i5-rc", 0, false, "false", nullptr},
    {"disable-illumina", 0, false, "false", nullptr},
    {"keep-barcode", 0, false, "false", nullptr},
    {"writing-buffer-size", 0, true, "67108864", nullptr},
    {"lane", 0, true, "", nullptr},
    {"instrument", 0, true, "", nullptr},
    {"run", 0, true, "", nullptr},
    {"undetermined-label", 0, true, "Undetermined", nullptr},
    {"ambiguous-label", 0, true, "Ambiguous", nullptr},
    {"comprehensive-scan", 0, false, "false", nullptr},
    {"force", 0, false, "false", nullptr},
    {"report-limit", 0, true, "20", nullptr},
    {"in-r1-file-suf", 0, true, "_read_1.fq.gz", nullptr},
    {"in-r2-file-suf", 0, true, "_read_2.fq.gz", nullptr},
    {"info-file", 0, true, "", nullptr},
    {"report-level", 0, true, "2", nullptr},
    {"compression-level", 0, true, "1", nullptr},
    {"compression-buffer-size", 0, true, "131072", nullptr},
    {"ignore-undetermined", 0, false, "false", nullptr},
    {"log", 0, true, "", nullptr},
    {"mgi-full-header", 0, false, "false", nullptr},
    {"per-index-error", 0, false, "false", nullptr},
    {"memory", 0, true, "0", nullptr},
    {"not-mgi", 0, false, "false", nullptr},
    {"threads", 't', true, "0", nullptr},
    {"validate", 0, false, "false", nullptr},
    {"reader-threads", 0, true, "0", nullptr},
    {"writer-threads", 0, true, "0", nullptr},
    {"log-level", 'l', true, "info", nullptr},
    {"gpus", 0, true, "1", nullptr},
    {"chunk-size", 0, true, "67108864", nullptr},
    {"slots", 0, true, "4", nullptr},
    {"host-gzip", 0, false, "false", nullptr},
};

// `template` subcommand, src/main.rs:400-511
const Opt TEMPLATE_OPTS[] = {
    {"input", 'i', true, "", nullptr},
    {"read2", 'r', true, "", "2"},
    {"read1", 'f', true, "", "1"},
    {"sample-sheet", 's', true, nullptr, nullptr},
    {"info-file", 0, true, "", nullptr},
    {"output", 'o', true, "", nullptr},
    {"barcode-length", 0, true, "0", nullptr},
    {"popular-template", 0, false, "false", nullptr},
    {"no-umi", 0, false, "false", nullptr},
    {"testing-reads", 0, true, "5000", nullptr},
    {"max-umi-len", 0, true, "10", nullptr},
    {"in-r1-file-suf", 0, true, "_read_1.fq.gz", nullptr},
    {"in-r2-file-suf", 0, true, "_read_2.fq.gz", nullptr},
    {"log-level", 'l', true, "info", nullptr},
};

// `reformat` subcommand, src/main.rs:541-712 (+ the same extensions as demultiplex)
const Opt REFORMAT_OPTS[] = {
    {"read2", 'r', true, "", "2"},
    {"read1", 'f', true, "", "1"},
    {"output", 'o', true, "", nullptr},
    {"reports", 0, true, "", nullptr},
    {"disable-illumina", 0, false, "false", nullptr},
    {"writing-buffer-size", 0, true, "67108864", nullptr},
    {"lane", 0, true, "", nullptr},
    {"instrument", 0, true, "", nullptr},
    {"run", 0, true, "", nullptr},
    {"force", 0, false, "false", nullptr},
    {"info-file", 0, true, "", nullptr},
    {"report-level", 0, true, "2", nullptr},
    {"compression-level", 0, true, "1", nullptr},
    {"compression-buffer-size", 0, true, "131072", nullptr},
    {"memory", 0, true, "0", nullptr},
    {"validate", 0, false, "false", nullptr},
    {"umi-length", 0, true, "0", nullptr},
    {"sample-index", 0, true, "1", nullptr},
    {"barcode", 0, true, "", nullptr},
    {"sample-label", 0, true, "", nullptr},
    {"log-level", 'l', true, "info", nullptr},
    {"threads", 't', true, "0", nullptr},
    {"reader-threads", 0, true, "0", nullptr},
    {"gpus", 0, true, "1", nullptr},
    {"chunk-size", 0, true, "67108864", nullptr},
    {"slots", 0, true, "4", nullptr},
    {"host-gzip", 0, false, "false", nullptr},
};

// parse_sb_file_name(), src/formater.rs:44-61: Flowcell_Lane_SampleLabel_1/2 -> (label, flowcell, lane, rest)
void parse_sb_file_name(const std::string& stem, std::string& label, std::string& lane) {
  std::vector<std::string> parts;
  size_t a = 0;
  for (;;) {
    const size_t b = stem.find('_', a);
    parts.push_back(stem.substr(a, b == std::string::npos ? std::string::npos : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  label.clear();
  lane.clear();
  if (parts.size() < 3) {
    log_warn("Input file name is expected to have the following format 'Flowcell_Lane_SampleLabel_1/2.fq/fastq.gz'");
    log_warn("Ignore this warning if you provided Sample label and lane parameters.");
    return;
  }
  for (size_t k = 2; k + 1 < parts.size(); ++k) label += (k > 2 ? "_" : "") + parts[k];
  lane = parts[1];
}

void print_logo() {
  std::printf("\n");
  std::printf("        MGIKIT  -  B200 hot path\n");
  std::printf("        (v%s) demultiplex on sm_100a\n\n", VERSION);
}

void usage() {
  std::printf("MGIKIT - MGI data demultipexing kit (B200 hot path).\n\nUsage: mgikit demultiplex [OPTIONS] --sample-sheet <FILE>\n"
              "       mgikit template [OPTIONS] --sample-sheet <FILE>   (barcode template detection)\n"
              "       mgikit reformat -f <R1> -r <R2> [OPTIONS]         (one sample: Illumina headers + quality report)\n\nOptions of demultiplex:\n");
  for (const auto& o : OPTS) {
    if (o.short_name) std::printf("  -%c, --%s", o.short_name, o.long_name);
    else std::printf("      --%s", o.long_name);
    if (o.takes_value) std::printf(" <VALUE>%s%s%s", o.def ? " [default: " : "", o.def ? o.def : "", o.def ? "]" : "");
    std::printf("\n");
  }
}

size_t parse_usize(const std::string& name, const std::string& v) {
  if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos) {
    std::fprintf(stderr, "error: invalid value '%s' for '--%s <VALUE>': invalid digit found in string\n", v.c_str(), name.c_str());
    std::exit(2);
  }
  return (size_t)std::strtoull(v.c_str(), nullptr, 10);
}

}  // namespace

int cli_main(int argc, char** argv, const Backend& be) {
  // the subcommand decides the option table: it is the first word that is neither an option nor the value of the
  // global -l/--log-level in front of it
  bool is_template = false, is_reformat = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-l" || a == "--log-level") {
      ++i;
      continue;
    }
    if (!a.empty() && a[0] == '-') continue;
    is_template = a == "template";
    is_reformat = a == "reformat";
    break;
  }
  const Opt* table = is_template ? TEMPLATE_OPTS : (is_reformat ? REFORMAT_OPTS : OPTS);
  const size_t n_table = is_template ? sizeof(TEMPLATE_OPTS) / sizeof(Opt)
                                     : (is_reformat ? sizeof(REFORMAT_OPTS) / sizeof(Opt) : sizeof(OPTS) / sizeof(Opt));
  std::map<std::string, std::string> val;
  for (size_t k = 0; k < n_table; ++k)
    if (table[k].def) val[table[k].long_name] = table[k].def;
  std::string command;
  auto find_long = [&](const std::string& n) -> const Opt* {
    for (size_t k = 0; k < n_table; ++k)
      if (n == table[k].long_name || (table[k].alias && n == table[k].alias)) return &table[k];
    return nullptr;
  };
  auto find_short = [&](char c) -> const Opt* {
    for (size_t k = 0; k < n_table; ++k)
      if (table[k].short_name == c) return &table[k];
    return nullptr;
  };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    const Opt* o = nullptr;
    std::string inline_val;
    bool has_inline = false;
    if (a == "--help" || a == "-h") {
      usage();
      return 0;
    }
    if (a == "--version" || a == "-V") {
      std::printf("MGIKIT - MGI data demultipexing kit. %s\n", VERSION);
      return 0;
    }
    if (a.rfind("--", 0) == 0) {
      std::string n = a.substr(2);
      size_t eq = n.find('=');
      if (eq != std::string::npos) {
        inline_val = n.substr(eq + 1);
        n = n.substr(0, eq);
        has_inline = true;
      }
      o = find_long(n);
    } else if (a.size() >= 2 && a[0] == '-') {
      o = find_short(a[1]);
      if (o && a.size() > 2) {
        inline_val = a.substr(a[2] == '=' ? 3 : 2);
        has_inline = true;
      }
    } else if (command.empty()) {
      command = a;
      continue;
    }
    if (!o) {
      std::fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str());
      return 2;
    }
    if (!o->takes_value) {
      val[o->long_name] = "true";
    } else if (has_inline) {
      val[o->long_name] = inline_val;
    } else {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "error: a value is required for '--%s <VALUE>' but none was supplied\n", o->long_name);
        return 2;
      }
      val[o->long_name] = argv[++i];
    }
  }
  print_logo();
  if (command.empty()) {
    usage();
    return 2;
  }
  if (command != "demultiplex" && command != "template" && command != "reformat") {
    std::fprintf(stderr, "error: this build provides the `demultiplex`, `template` and `reformat` commands (got '%s'); "
                         "`report` panics in the reference itself (v2.2.0, index out of bounds in write_reports)\n", command.c_str());
    return 2;
  }
  if (command != "reformat" && !val.count("sample-sheet")) {
    std::fprintf(stderr, "error: the following required arguments were not provided:\n  --sample-sheet <arg_sample_sheet_file_path>\n");
    return 2;
  }
  set_log_level(val["log-level"]);
  auto flag = [&](const char* n) { return val[n] == "true"; };
  const auto t0 = std::chrono::steady_clock::now();
  if (command == "template") {
    try {
      TemplateArgs ta;
      ta.input_dir = val["input"];
      ta.read1 = val["read1"];
      ta.read2 = val["read2"];
      ta.sample_sheet = val["sample-sheet"];
      ta.info_file = val["info-file"];
      ta.output = val["output"];
      ta.r1_suffix = val["in-r1-file-suf"];
      ta.r2_suffix = val["in-r2-file-suf"];
      ta.barcode_length = parse_usize("barcode-length", val["barcode-length"]);
      ta.testing_reads = parse_usize("testing-reads", val["testing-reads"]);
      ta.max_umi_length = parse_usize("max-umi-len", val["max-umi-len"]);
      ta.popular_template = flag("popular-template");
      ta.add_umi = !flag("no-umi");
      run_template(ta);
    } catch (const std::exception& e) {
      std::fflush(stdout);
      std::fprintf(stderr, "mgikit error: %s\n", e.what());
      return 101;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    log_info("Exection time is " + std::to_string((long)secs) + " seconds.");
    return 0;
  }
  if (command == "reformat") {
    try {
      // reformat(), src/lib.rs:2266-2464
      DemuxOptions opt;
      opt.report_level = (int)parse_usize("report-level", val["report-level"]);
      RunArgs ra;
      ra.read1 = val["read1"];
      ra.read2 = val["read2"];
      ra.output_dir = val["output"];
      ra.report_dir = val["reports"];
      ra.lane = val["lane"];
      ra.instrument = val["instrument"];
      ra.run = val["run"];
      ra.mgi_data = true;
      ra.force = flag("force");
      ra.r1_suffix = "";
      ra.r2_suffix = "";
      ra.info_file = val["info-file"];
      ra.illumina_format = !flag("disable-illumina");
      ra.check_content = flag("validate");
      RunInfo run = make_run_info(ra);
      run.get_read_information();
      if (run.lane.empty()) {
        log_info("lane detected in the read header will be used for this run!");
        run.lane = run.barcode_read_info.lane;
      }
      run.confirm_format();
      const size_t cbs = parse_usize("compression-buffer-size", val["compression-buffer-size"]);
      const size_t wbs = parse_usize("writing-buffer-size", val["writing-buffer-size"]);
      std::string stem = run.barcode_reads;   // Path::file_stem(): the name without its last extension
      if (stem.find_last_of('/') != std::string::npos) stem = stem.substr(stem.find_last_of('/') + 1);
      if (stem.find_last_of('.') != std::string::npos && stem.find_last_of('.') > 0) stem = stem.substr(0, stem.find_last_of('.'));
      std::string label_sb, lane_sb;
      parse_sb_file_name(stem, label_sb, lane_sb);
      ReformatSample rf;
      rf.label = val["sample-label"];
      if (rf.label.empty()) {
        if (label_sb.empty())
          throw Fatal("Sample label needs to be passed either through '--sample-label' parameter or the third part of the file name after separation by '_'");
        rf.label = label_sb;
      }
      if (lane_sb != run.lane)
        log_warn("The lane extracted from the file name (" + lane_sb + ") is not the same as the provided lane (" + run.lane + ")");
      rf.sample_index = parse_usize("sample-index", val["sample-index"]);
      if (rf.sample_index < 1)   // ReformatedSample::new, src/formater.rs:23-25
        throw Fatal("Sample index (val: " + std::to_string(rf.sample_index) + ") needs to be greater than 0!");
      rf.umi_length = (int)parse_usize("umi-length", val["umi-length"]);
      rf.barcode = val["barcode"];
      log_info("Sample's id, extracted from the file name is: " + rf.label);
      log_info("Sample's index to be used in file naming is: " + std::to_string(rf.sample_index));
      log_info("Sample's barcode: " + rf.barcode);
      if (rf.barcode.empty())
        log_warn("No mismatches will be considered as sample's barcode is not provided to compare against!");
      log_info("Umi length to be extracted from the tail of R2: " + std::to_string(rf.umi_length));
      // BufferInfo::new(), src/sample_data.rs:413-455 — kept for flag compatibility
      if (cbs > wbs)
        throw Fatal("Compression buffer size '--compression-buffer-size' should be less than Writing buffer size ('--writing-buffer-size').");
      if (wbs < 65536)
        throw Fatal("Writing buffer size '--writing-buffer-size' will be increased to minimal allowed value (65536).");
      if (wbs > 536870912)
        throw Fatal("Writing buffer size '--writing-buffer-size' will be reduced to the maximum allowed value (536870912).");
      const double mem = std::atof(val["memory"].c_str());
      if (0.0 < mem && mem <= 0.5) throw Fatal("Requested memory should be greater than 0.5 GB!");
      if ((size_t)rf.umi_length > run.barcode_read_info.sequence_length)
        throw Fatal("The UMI is longer than the read that carries it!");
      opt.compression_level = (int)parse_usize("compression-level", val["compression-level"]);
      if (opt.compression_level > 12) throw Fatal("compression level must be in [0, 12]");
      opt.threads = (int)parse_usize("threads", val["threads"]);
      opt.gpus = (int)parse_usize("gpus", val["gpus"]);
      opt.chunk_bytes = parse_usize("chunk-size", val["chunk-size"]);
      opt.slots = (int)parse_usize("slots", val["slots"]);
      opt.reader_threads = (int)parse_usize("reader-threads", val["reader-threads"]);
      opt.host_gzip = flag("host-gzip");
      opt.leave_teardown_to_exit = be.exits_after_command;
      opt.reformat = &rf;
      const SampleSheet sheet = dummy_sample_sheet(rf.label);
      log_info("Reporting level is: " + std::to_string(opt.report_level));
      run_demultiplex(be, sheet, run, opt);
    } catch (const std::exception& e) {
      std::fflush(stdout);
      std::fprintf(stderr, "mgikit error: %s\n", e.what());
      return 101;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    log_info("Exection time is " + std::to_string((long)secs) + " seconds.");
    return 0;
  }
  try {
    // initiate_demultiplexing(), src/lib.rs:1606-1856
    SampleSheet sheet = load_sample_manager(val["sample-sheet"], val["template"], flag("i7-rc"), flag("i5-rc"),
                                            val["undetermined-label"], val["ambiguous-label"]);
    log_info(std::to_string(sheet.n_samples()) + " Samples were found in the input sample sheet.");
    if (sheet.templates.size() > 1) log_info("Mixed library is detected! different barcode templates for some samples!");
    else log_info("Same barcode template is used for all samples!");
    RunArgs ra;
    ra.input_dir = val["input"];
    ra.read1 = val["read1"];
    ra.read2 = val["read2"];
    ra.output_dir = val["output"];
    ra.report_dir = val["reports"];
    ra.lane = val["lane"];
    ra.instrument = val["instrument"];
    ra.run = val["run"];
    ra.mgi_data = !flag("not-mgi");
    ra.force = flag("force");
    ra.r1_suffix = val["in-r1-file-suf"];
    ra.r2_suffix = val["in-r2-file-suf"];
    ra.info_file = val["info-file"];
    ra.illumina_format = !flag("disable-illumina");
    ra.keep_barcode = flag("keep-barcode");
    ra.comprehensive_scan = flag("comprehensive-scan");
    ra.ignore_undetermined = flag("ignore-undetermined");
    ra.check_content = flag("validate");
    ra.mgi_full_header = flag("mgi-full-header");
    RunInfo run = make_run_info(ra);
    run.get_read_information();
    if (run.lane.empty()) {
      log_info("lane detected in the read header will be used for this run!");
      run.lane = run.barcode_read_info.lane;
    }
    const int BL = sheet.barcode_length();
    if ((size_t)BL == run.barcode_read_info.sequence_length) {
      log_info("It is assumed that read 2 contains barcode only without read sequence!");
      run.read2_has_sequence = false;
    }
    if ((size_t)BL > run.barcode_read_info.sequence_length)
      throw Fatal("The barcode described by the template is longer than the read that carries it!");
    run.confirm_format();
    // BufferInfo::new(), src/sample_data.rs:413-455 — kept for flag compatibility
    const size_t wbs = parse_usize("writing-buffer-size", val["writing-buffer-size"]);
    const size_t cbs = parse_usize("compression-buffer-size", val["compression-buffer-size"]);
    if (cbs > wbs)
      throw Fatal("Compression buffer size '--compression-buffer-size' should be less than Writing buffer size ('--writing-buffer-size').");
    if (wbs < 65536)
      throw Fatal("Writing buffer size '--writing-buffer-size' will be increased to minimal allowed value (65536).");
    if (wbs > 536870912)
      throw Fatal("Writing buffer size '--writing-buffer-size' will be reduced to the maximum allowed value (536870912).");
    const double mem = std::atof(val["memory"].c_str());
    if (0.0 < mem && mem <= 0.5) throw Fatal("Requested memory should be greater than 0.5 GB!");

    DemuxOptions opt;
    opt.mismatches = (int)parse_usize("mismatches", val["mismatches"]);
    opt.per_index_error = flag("per-index-error");
    opt.report_level = (int)parse_usize("report-level", val["report-level"]);
    opt.report_limit = parse_usize("report-limit", val["report-limit"]);
    opt.compression_level = (int)parse_usize("compression-level", val["compression-level"]);
    if (opt.compression_level > 12) throw Fatal("compression level must be in [0, 12]");
    opt.threads = (int)parse_usize("threads", val["threads"]);
    opt.gpus = (int)parse_usize("gpus", val["gpus"]);
    opt.chunk_bytes = parse_usize("chunk-size", val["chunk-size"]);
    opt.slots = (int)parse_usize("slots", val["slots"]);
    opt.reader_threads = (int)parse_usize("reader-threads", val["reader-threads"]);
    opt.host_gzip = flag("host-gzip");
    opt.leave_teardown_to_exit = be.exits_after_command;
    log_info("Reporting level is: " + std::to_string(opt.report_level));
    run_demultiplex(be, sheet, run, opt);
  } catch (const std::exception& e) {
    std::fflush(stdout);
    std::fprintf(stderr, "mgikit error: %s\n", e.what());
    return 101;
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  log_info("Exection time is " + std::to_string((long)secs) + " seconds.");
  return 0;
}

}  // namespace mgikit

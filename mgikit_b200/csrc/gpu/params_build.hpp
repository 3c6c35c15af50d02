// mgk_config -> Params + flat tables (host side).  Shared by libmgikit_gpu.so
// (which then uploads the tables) and by the CPU emulation used in tests.
#pragma once

#include <cstring>
#include <string>
#include <vector>

#include "device_logic.cuh"

namespace mgk {

struct HostTables {
  std::vector<uint64_t> codes;   // 2-bit packed unique indexes, all templates
  std::vector<int32_t> pairs;    // pair -> sample tables
  std::vector<int32_t> xlat;     // index-id translation between templates
  std::vector<int32_t> writing;
  std::string consts;            // illumina prefix + literal bytes
  std::vector<uint64_t> hkeys;   // exact-match hash tables (two words per slot)
  std::vector<int32_t> hids;
};

// open-addressing table over n index strings of `len` (<= 16) bytes; returns the first slot, sets mask
inline int build_index_hash(HostTables& tb, const char* seqs, int n, int len, int32_t& mask) {
  uint32_t slots = 16;
  while (slots < 2u * (uint32_t)n) slots <<= 1;
  mask = (int32_t)(slots - 1);
  const int off = (int)tb.hids.size();
  tb.hids.resize((size_t)off + slots, -1);
  tb.hkeys.resize(2 * ((size_t)off + slots), 0);
  for (int i = 0; i < n; ++i) {
    uint64_t k[2] = {0, 0};
    memcpy(k, seqs + (size_t)i * len, (size_t)len);
    uint32_t s = index_hash(k[0], k[1], (uint32_t)mask);
    while (tb.hids[(size_t)off + s] >= 0) s = (s + 1) & (uint32_t)mask;
    tb.hids[(size_t)off + s] = i;
    tb.hkeys[2 * ((size_t)off + s)] = k[0];
    tb.hkeys[2 * ((size_t)off + s) + 1] = k[1];
  }
  return off;
}

// Fills P (pointer members left null) and tb.  Returns false with a reason.
inline bool build_params_host(const mgk_config* cfg, Params& P, HostTables& tb, std::string& why) {
  memset(&P, 0, sizeof P);
  P.paired = cfg->paired;
  P.read2_has_sequence = cfg->read2_has_sequence;
  P.out_mode = cfg->out_mode;
  P.keep_barcode = cfg->keep_barcode;
  P.comprehensive = cfg->comprehensive_scan;
  P.validate = cfg->validate;
  P.report_level = cfg->report_level;
  P.mgi_data = cfg->mgi_data;
  P.m = cfg->allowed_mismatches;
  P.M = cfg->total_allowed_mismatches;
  P.L = cfg->l_position;
  P.S = cfg->n_samples;
  P.total = cfg->n_samples + 2;
  P.reformat = cfg->reformat ? 1 : 0;
  const std::string rf_barcode = (cfg->reformat && cfg->reformat_barcode) ? cfg->reformat_barcode : "";
  P.rf_bc_len = (int)rf_barcode.size();
  if (P.reformat) {   // src/lib.rs:2302-2322, :2422-2441
    if (cfg->n_samples != 1 || cfg->n_templates != 0) { why = "reformat: one sample, no barcode templates"; return false; }
    if (!cfg->read2_has_sequence || cfg->out_mode == MGK_OUT_MGI_FULL_HEADER || cfg->keep_barcode) { why = "reformat: flags the command does not have"; return false; }
    if (cfg->reformat_umi_length < 0 || cfg->reformat_umi_length > MGK_MAX_BARCODE_LEN) { why = "umi length outside [0,255]"; return false; }
    if (P.rf_bc_len > MGK_MAX_BARCODE_LEN) { why = "barcode longer than 255 bases"; return false; }
    if (cfg->allowed_mismatches < 4) { why = "reformat counts up to 4 mismatches: allowed_mismatches must be >= 4 (the reference passes 5)"; return false; }
  } else if (cfg->n_templates < 1 || !cfg->templates) {
    why = "no barcode template";
    return false;
  }
  P.BL = P.reformat ? cfg->reformat_umi_length : cfg->templates[0].geometry[9];
  P.mism_w = 2 * cfg->allowed_mismatches + 2;
  P.n_tpl = cfg->n_templates;
  const std::string prefix = cfg->illumina_prefix ? cfg->illumina_prefix : "";
  P.prefix_len = (int)prefix.size();
  if (P.out_mode < 0 || P.out_mode > 2) { why = "unknown out_mode"; return false; }
  if (P.out_mode == MGK_OUT_ILLUMINA && prefix.empty()) { why = "illumina output needs the header prefix"; return false; }
  if (P.BL < (P.reformat ? 0 : 1) || P.BL > MGK_MAX_BARCODE_LEN) { why = "barcode length outside [1,255]"; return false; }
  if (P.L < 0) { why = "negative l_position"; return false; }
  for (int t = 0; t < P.n_tpl; ++t) {
    const mgk_template& s = cfg->templates[t];
    TemplateDev& T = P.tpl[t];
    for (int k = 0; k < 10; ++k) T.g[k] = s.geometry[k];
    if (T.g[9] != P.BL) { why = "all templates must describe the same barcode length"; return false; }
    if (T.g[1] < 1 || T.g[1] > MGK_MAX_INDEX_LEN || T.g[4] < 0 || T.g[4] > MGK_MAX_INDEX_LEN) { why = "index lengths must be in [1,32]"; return false; }
    if (T.g[2] > P.BL || T.g[2] < T.g[1] || (s.has_i5 && (T.g[5] > P.BL || T.g[5] < T.g[4])) || (T.g[6] && (T.g[8] > P.BL || T.g[8] < T.g[7]))) {
      why = "a template places an item outside the barcode";
      return false;
    }
    T.has_i5 = s.has_i5 ? 1 : 0;
    T.n_i7 = s.n_i7;
    T.n_i5 = s.has_i5 ? s.n_i5 : 0;
    if (T.n_i7 < 0 || T.n_i5 < 0 || !s.pair_sample) { why = "a template has no index table"; return false; }
    // the 2-bit codes and the raw-byte hash agree only on [ACGT] (the sample sheet's own alphabet,
    // src/sample_manager.rs:362-372): anything else would silently pack as another base
    auto acgt = [](const char* p, size_t n) {
      for (size_t i = 0; i < n; ++i)
        if (p[i] != 'A' && p[i] != 'C' && p[i] != 'G' && p[i] != 'T') return false;
      return true;
    };
    if ((T.n_i7 && !s.i7) || (T.n_i5 && !s.i5)) { why = "a template has no index table"; return false; }
    if (!acgt(s.i7, (size_t)T.n_i7 * T.g[1]) || !acgt(s.i5, (size_t)T.n_i5 * T.g[4])) { why = "index tables must hold [ACGT] only"; return false; }
    T.wide = (T.g[1] > 16 || (T.has_i5 && T.g[4] > 16)) ? 1 : 0;
    T.off_i7 = (int)tb.codes.size();
    for (int i = 0; i < T.n_i7; ++i) tb.codes.push_back(pack_index_host(s.i7 + (size_t)i * T.g[1], T.g[1]));
    T.off_i5 = (int)tb.codes.size();
    for (int i = 0; i < T.n_i5; ++i) tb.codes.push_back(pack_index_host(s.i5 + (size_t)i * T.g[4], T.g[4]));
    T.h7_off = T.h5_off = T.h7_mask = T.h5_mask = 0;
    if (!T.wide) {
      T.h7_off = build_index_hash(tb, s.i7, T.n_i7, T.g[1], T.h7_mask);
      if (T.has_i5) T.h5_off = build_index_hash(tb, s.i5, T.n_i5, T.g[4], T.h5_mask);
    }
    T.off_pair = (int)tb.pairs.size();
    const size_t np = T.has_i5 ? (size_t)T.n_i7 * T.n_i5 : (size_t)T.n_i7;
    tb.pairs.insert(tb.pairs.end(), s.pair_sample, s.pair_sample + np);
  }
  // index-id translation between templates (stale dictionary iterator of the reference)
  for (int d = 0; d < P.n_tpl; ++d)
    for (int t = 0; t < P.n_tpl; ++t)
      for (int which = 0; which < 2; ++which) {
        int32_t* off = which == 0 ? P.xl7_off : P.xl5_off;
        off[d * MGK_MAX_TEMPLATES + t] = (int)tb.xlat.size();
        const mgk_template &sd = cfg->templates[d], &st = cfg->templates[t];
        const int nd = which == 0 ? P.tpl[d].n_i7 : P.tpl[d].n_i5, nt = which == 0 ? P.tpl[t].n_i7 : P.tpl[t].n_i5;
        const int ld = which == 0 ? P.tpl[d].g[1] : P.tpl[d].g[4], lt = which == 0 ? P.tpl[t].g[1] : P.tpl[t].g[4];
        const char* seqd = which == 0 ? sd.i7 : sd.i5;
        const char* seqt = which == 0 ? st.i7 : st.i5;
        for (int i = 0; i < nd; ++i) {
          int found = -1;
          if (d == t) found = i;
          else if (ld == lt)
            for (int j = 0; j < nt && found < 0; ++j)
              if (memcmp(seqd + (size_t)i * ld, seqt + (size_t)j * lt, (size_t)ld) == 0) found = j;
          tb.xlat.push_back(found);
        }
        if (nd == 0) tb.xlat.push_back(-1);
      }
  P.n_codes = (int)tb.codes.size();
  tb.consts = prefix;
  char lit[LIT_BYTES];
  memset(lit, 0, sizeof lit);
  lit[LIT_COLON] = ':';
  lit[LIT_SPACE] = ' ';
  memcpy(lit + LIT_TAIL, ":N:0:", 5);
  lit[LIT_PLUS] = '+';
  lit[LIT_NL] = '\n';
  tb.consts.append(lit, LIT_BYTES);
  tb.consts += rf_barcode;
  tb.writing.assign(cfg->writing_sample, cfg->writing_sample + P.total);
  for (int i = 0; i < P.total; ++i)
    if (tb.writing[(size_t)i] < 0 || tb.writing[(size_t)i] >= P.total) { why = "writing_sample out of range"; return false; }
  return true;
}

}  // namespace mgk

// TEST INFRASTRUCTURE ONLY.
// CPU emulation of the CUDA path's per-record logic: the very functions of
// mgikit_b200/csrc/gpu/device_logic.cuh, compiled by g++ with the 32 lanes of a
// warp looped.  It lets the `-m "not gpu"` tests check matching, output-length
// planning, header rewrite, destination-aligned span copies and QC sums against
// the oracle before any GPU time is spent.  The warp/block collectives (scan,
// look-back, in-tile ranking) are NOT covered here — those only run on the GPU.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../mgikit_b200/csrc/gpu/params_build.hpp"

using namespace mgk;

struct emu_ctx {
  Params P;
  HostTables tb;
  std::vector<uint64_t> stats, mism;
  std::vector<uint8_t> out2, out1;
  std::vector<uint64_t> off2, len2, off1, len1;
  std::vector<uint32_t> unassigned;
  mgk_result res;
  bool have = false;
  std::string err;
};

extern "C" {

const char* emu_last_error(const emu_ctx* c) { return c ? c->err.c_str() : "emu_create failed"; }

int emu_create(const mgk_config* cfg, emu_ctx** out) {
  emu_ctx* c = new emu_ctx();
  std::string why;
  if (!build_params_host(cfg, c->P, c->tb, why)) {
    delete c;
    return MGK_ERR_ARG;
  }
  c->P.codes = c->tb.codes.data();
  c->P.pairs = c->tb.pairs.data();
  c->P.xlat = c->tb.xlat.data();
  c->P.writing = c->tb.writing.data();
  c->P.consts = (const uint8_t*)c->tb.consts.data();
  c->stats.assign((size_t)c->P.total * MGK_N_STATS, 0);
  c->mism.assign((size_t)c->P.total * c->P.mism_w, 0);
  *out = c;
  return MGK_OK;
}

void emu_destroy(emu_ctx* c) { delete c; }

static std::vector<uint32_t> newlines(const uint8_t* b, size_t n) {
  std::vector<uint32_t> v;
  for (size_t i = 0; i < n; ++i)
    if (b[i] == '\n') v.push_back((uint32_t)i);
  return v;
}

int emu_submit(emu_ctx* c, uint64_t seq, const uint8_t* r2_in, uint64_t r2_len, const uint8_t* r1_in, uint64_t r1_len, uint32_t) {
  const Params& P = c->P;
  // 16-byte aligned, padded private copies like the device buffers
  uint8_t* r2 = (uint8_t*)aligned_alloc(64, ((size_t)r2_len + 128 + 63) / 64 * 64);
  uint8_t* r1 = (uint8_t*)aligned_alloc(64, ((size_t)r1_len + 128 + 63) / 64 * 64);
  memset(r2 + r2_len, 0, 64);
  memset(r1 + r1_len, 0, 64);
  memcpy(r2, r2_in, r2_len);
  if (r1_len) memcpy(r1, r1_in, r1_len);
  const auto nl2 = newlines(r2, r2_len);
  const auto nl1 = P.paired ? newlines(r1, r1_len) : std::vector<uint32_t>();
  uint32_t n_rec = (uint32_t)(nl2.size() / 4);
  if (P.paired) n_rec = std::min<uint32_t>(n_rec, (uint32_t)(nl1.size() / 4));
  const int total = P.total;
  struct Rec {
    RecGeom G;
    MatchOut mo;
    RecPlan pl;
    int bin;
    uint64_t d2, d1;
  };
  std::vector<Rec> recs(n_rec);
  std::vector<uint64_t> bin2((size_t)total, 0), bin1((size_t)total, 0);
  c->unassigned.clear();
  int rc = MGK_OK;
  for (uint32_t i = 0; i < n_rec; ++i) {
    Rec& R = recs[i];
    RecGeom& G = R.G;
    G.h2 = i ? nl2[4 * i - 1] + 1 : 0;
    G.s2 = nl2[4 * i] + 1;
    G.p2 = nl2[4 * i + 1] + 1;
    G.q2 = nl2[4 * i + 2] + 1;
    G.e2 = nl2[4 * i + 3];
    G.h1 = G.s1 = G.p1 = G.q1 = G.e1 = 0;
    if (P.paired) {
      G.h1 = i ? nl1[4 * i - 1] + 1 : 0;
      G.s1 = nl1[4 * i] + 1;
      G.p1 = nl1[4 * i + 1] + 1;
      G.q1 = nl1[4 * i + 2] + 1;
      G.e1 = nl1[4 * i + 3];
    }
    R.mo.sample = total - 2;
    R.mo.mism = 0;
    R.mo.bar_t = R.mo.umi_t = -1;
    R.mo.panicked = 0;
    if (G.p2 - G.s2 >= (uint32_t)P.BL + 1u) R.mo = match_barcode(P, P.codes, r2 + G.p2 - 1 - P.BL);
    if (R.mo.panicked) { rc = MGK_ERR_FORMAT; c->err = "panic"; break; }
    R.pl = record_plan(P, r2, G, R.mo.sample, R.mo.bar_t, R.mo.umi_t);
    if (R.pl.fmt_error) { rc = MGK_ERR_FORMAT; c->err = "format"; break; }
    R.bin = R.pl.matched ? P.writing[R.mo.sample] : R.mo.sample;
    R.d2 = bin2[(size_t)R.bin];
    R.d1 = bin1[(size_t)R.bin];
    bin2[(size_t)R.bin] += R.pl.len2;
    bin1[(size_t)R.bin] += R.pl.len1;
    c->mism[(size_t)R.mo.sample * P.mism_w + R.mo.mism + 1] += 1;
    if (!R.pl.matched && P.report_level > 1)
      c->unassigned.push_back((G.p2 - 1 - (uint32_t)P.BL) | (R.mo.sample == total - 1 ? 0x80000000u : 0u));
  }
  if (rc != MGK_OK) { free(r2); free(r1); return rc; }
  c->off2.assign((size_t)total, 0); c->len2 = bin2; c->off1.assign((size_t)total, 0); c->len1 = bin1;
  uint64_t t2 = 0, t1 = 0;
  for (int b = 0; b < total; ++b) { c->off2[(size_t)b] = t2; t2 += bin2[(size_t)b]; c->off1[(size_t)b] = t1; t1 += bin1[(size_t)b]; }
  // 16-byte aligned outputs (vector::data of a fresh allocation is) with slack
  c->out2.assign((size_t)t2 + 64, 0xEE);
  c->out1.assign((size_t)t1 + 64, 0xEE);
  uint8_t* o2 = c->out2.data();
  uint8_t* o1 = c->out1.data();
  memset(&c->res, 0, sizeof c->res);
  const uint32_t BL = (uint32_t)P.BL, wbl = P.keep_barcode ? 0u : BL;
  unsigned long long vkey = ~0ull;
  for (uint32_t i = 0; i < n_rec; ++i) {
    const Rec& R = recs[i];
    const RecGeom& G = R.G;
    const RecPlan pl = record_plan(P, r2, G, R.mo.sample, R.mo.bar_t, R.mo.umi_t);
    uint8_t* d2 = o2 + c->off2[(size_t)R.bin] + R.d2;
    uint8_t* d1 = o1 + c->off1[(size_t)R.bin] + R.d1;
    Qc qa{0, 0}, qb{0, 0}, qc{0, 0};
    int vworst = 0;
    for (int lane = 0; lane < 32; ++lane) {  // one warp, lanes looped
      if (!pl.matched) {
        copy_span(d2, r2 + G.h2, G.e2 + 1u - G.h2, lane, 32);
        if (P.paired) copy_span(d1, r1 + G.h1, G.e1 + 1u - G.h1, lane, 32);
      } else {
        const bool rewrite = P.out_mode != MGK_OUT_MGI;
        if (P.read2_has_sequence) {
          uint32_t n = 0;
          if (rewrite) {
            SegList s;
            header_segments(P, r2, r1, G, pl, R.mo.bar_t, R.mo.umi_t, 2, s);
            n = emit_segments(d2, s, lane, 32);
          }
          const uint32_t a0 = rewrite ? G.s2 - 1u : G.h2;
          const uint32_t na = G.p2 - wbl - 1u - a0;
          copy_span(d2 + n, r2 + a0, na, lane, 32);
          const uint32_t nb = G.e2 - wbl - (G.p2 - 1u);
          copy_span(d2 + n + na, r2 + G.p2 - 1u, nb, lane, 32);
          if (lane == 0) d2[n + na + nb] = '\n';
        }
        if (P.paired) {
          uint32_t n = 0;
          if (rewrite) {
            SegList s;
            header_segments(P, r2, r1, G, pl, R.mo.bar_t, R.mo.umi_t, 1, s);
            n = emit_segments(d1, s, lane, 32);
          }
          const uint32_t a0 = rewrite ? G.s1 - 1u : G.h1;
          copy_span(d1 + n, r1 + a0, G.e1 + 1u - a0, lane, 32);
        }
      }
      if (P.report_level > 0) {
        if (P.paired) { Qc q = qc_partial(r1 + G.q1, G.e1 - G.q1, lane, 32); qa.sum += q.sum; qa.high += q.high; }
        { Qc q = qc_partial(r2 + G.e2 - BL, BL, lane, 32); qb.sum += q.sum; qb.high += q.high; }
        { Qc q = qc_partial(r2 + G.q2, G.e2 - BL - G.q2, lane, 32); qc.sum += q.sum; qc.high += q.high; }
      }
      if (P.validate) {
        const int v = validate_partial(P, r2, r1, G, lane, 32);
        if (v && (vworst == 0 || v < vworst)) vworst = v;
      }
    }
    if (P.report_level > 0) {
      uint64_t* st = c->stats.data() + (size_t)R.mo.sample * MGK_N_STATS;
      const int shift = P.paired ? 1 : 0;
      if (P.paired) { st[0] += qa.high; st[6] += qa.sum; }
      st[2] += qb.high; st[8] += qb.sum;
      st[shift] += qc.high; st[6 + shift] += qc.sum;
    }
    if (vworst) {
      const unsigned long long k = ((unsigned long long)i << 4) | (unsigned)vworst;
      if (k < vkey) vkey = k;
    }
  }
  free(r2);
  free(r1);
  c->res.seq = seq;
  c->res.n_records = n_rec;
  c->res.consumed_r2 = n_rec ? nl2[4 * n_rec - 1] + 1 : 0;
  c->res.consumed_r1 = (P.paired && n_rec) ? nl1[4 * n_rec - 1] + 1 : 0;
  c->res.out_r2 = o2;
  c->res.out_r1 = P.paired ? o1 : nullptr;
  c->res.out_r2_bytes = t2;
  c->res.out_r1_bytes = t1;
  c->res.seg_off_r2 = c->off2.data(); c->res.seg_len_r2 = c->len2.data();
  c->res.seg_off_r1 = c->off1.data(); c->res.seg_len_r1 = c->len1.data();
  c->res.unassigned = c->unassigned.data();
  c->res.n_unassigned = c->unassigned.size();
  if (vkey != ~0ull) { c->res.validate_error = validate_kind((int)(vkey & 15)); c->res.validate_record = vkey >> 4; }
  c->have = true;
  return MGK_OK;
}

int emu_collect(emu_ctx* c, mgk_result* r) {
  if (!c->have) return MGK_ERR_STATE;
  *r = c->res;
  return MGK_OK;
}
int emu_release(emu_ctx* c, uint64_t) { c->have = false; return MGK_OK; }
int emu_counters(emu_ctx* c, uint64_t* stats, uint64_t* mism) {
  if (stats) memcpy(stats, c->stats.data(), c->stats.size() * 8);
  if (mism) memcpy(mism, c->mism.data(), c->mism.size() * 8);
  return MGK_OK;
}
int emu_reset_counters(emu_ctx* c) {
  std::fill(c->stats.begin(), c->stats.end(), 0);
  std::fill(c->mism.begin(), c->mism.end(), 0);
  return MGK_OK;
}
int emu_match_barcodes(emu_ctx* c, const uint8_t* bars, uint64_t n, int32_t* sample, int32_t* mism) {
  for (uint64_t i = 0; i < n; ++i) {
    MatchOut mo = match_barcode(c->P, c->P.codes, bars + i * (uint64_t)c->P.BL);
    if (mo.panicked) return MGK_ERR_FORMAT;
    sample[i] = mo.sample;
    mism[i] = mo.mism;
  }
  return MGK_OK;
}
}

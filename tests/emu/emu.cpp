// TEST INFRASTRUCTURE ONLY.
// CPU emulation of the CUDA path's per-record logic: the very functions of
// mgikit_b200/csrc/gpu/device_logic.cuh, compiled by g++ with the 32 lanes of a
// warp looped.  It lets the `-m "not gpu"` tests check matching, output-length
// planning, header rewrite, destination-aligned span copies and QC sums against
// the oracle before any GPU time is spent.  The warp/block collectives (scan,
// look-back, in-tile ranking) are NOT covered here — those only run on the GPU.
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../mgikit_b200/csrc/gpu/emit_logic.cuh"
#include "../../mgikit_b200/csrc/gpu/params_build.hpp"

using namespace mgk;

struct emu_chunk {   // what one submitted chunk leaves behind until it is released
  std::vector<uint8_t> out2, out1;
  std::vector<uint64_t> off2, len2, off1, len1;
  std::vector<uint32_t> unassigned;
  mgk_result res;
};

struct emu_ctx {
  Params P;
  HostTables tb;
  std::vector<uint64_t> stats, mism;
  std::deque<std::unique_ptr<emu_chunk>> submitted, collected;   // FIFO like the device contexts: any number in flight
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
  c->P.hkeys = c->tb.hkeys.data();
  c->P.hids = c->tb.hids.data();
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
  std::unique_ptr<emu_chunk> chunk(new emu_chunk());
  emu_chunk& K = *chunk;
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
  K.unassigned.clear();
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
    if (P.reformat) R.mo.sample = 0;   // src/lib.rs:1114
    else if (G.p2 - G.s2 >= (uint32_t)P.BL + 1u) R.mo = match_barcode(P, P.codes, r2 + G.p2 - 1 - P.BL);
    if (R.mo.panicked) { rc = MGK_ERR_FORMAT; c->err = "panic"; break; }
    R.pl = record_plan(P, r2, G, R.mo.sample, R.mo.bar_t, R.mo.umi_t);
    if (R.pl.fmt_error) { rc = MGK_ERR_FORMAT; c->err = P.reformat ? "record does not have the header / line layout `reformat` slices" : "format"; break; }
    if (P.reformat) R.mo.mism = (int)R.pl.rf_mism;
    R.bin = R.pl.matched ? P.writing[R.mo.sample] : R.mo.sample;
    R.d2 = bin2[(size_t)R.bin];
    R.d1 = bin1[(size_t)R.bin];
    bin2[(size_t)R.bin] += R.pl.len2;
    bin1[(size_t)R.bin] += R.pl.len1;
    c->mism[(size_t)R.mo.sample * P.mism_w + R.mo.mism + 1] += 1;
    if (!R.pl.matched && P.report_level > 1)
      K.unassigned.push_back((G.p2 - 1 - (uint32_t)P.BL) | (R.mo.sample == total - 1 ? 0x80000000u : 0u));
  }
  if (rc != MGK_OK) { free(r2); free(r1); return rc; }
  K.off2.assign((size_t)total, 0); K.len2 = bin2; K.off1.assign((size_t)total, 0); K.len1 = bin1;
  uint64_t t2 = 0, t1 = 0;
  for (int b = 0; b < total; ++b) { K.off2[(size_t)b] = t2; t2 += bin2[(size_t)b]; K.off1[(size_t)b] = t1; t1 += bin1[(size_t)b]; }
  // 16-byte aligned outputs (vector::data of a fresh allocation is) with slack
  K.out2.assign((size_t)t2 + 64, 0xEE);
  K.out1.assign((size_t)t1 + 64, 0xEE);
  uint8_t* o2 = K.out2.data();
  uint8_t* o1 = K.out1.data();
  memset(&K.res, 0, sizeof K.res);
  // the staged tile: the whole chunk as ONE tile, laid out like the kernel's shared memory
  //   [slack | prefix + literals | two header staging slots | barcode read bytes | paired read bytes | slack]
  const uint32_t cs0 = 64;
  const uint32_t hs = (cs0 + (uint32_t)c->tb.consts.size() + 63u) / 16u * 16u + 16u;
  const uint32_t b2 = hs + 2u * HDR_SLOT_BYTES + 32u;
  const uint32_t b1 = (b2 + (uint32_t)r2_len + 15u) / 16u * 16u + 16u + (uint32_t)(seq % 3) * 16u;
  const size_t s_bytes = (size_t)b1 + r1_len + 128;
  uint8_t* S = (uint8_t*)aligned_alloc(64, (s_bytes + 63) / 64 * 64);
  memset(S, 0x5a, s_bytes);
  memcpy(S + cs0, c->tb.consts.data(), c->tb.consts.size());
  memcpy(S + b2, r2, r2_len);
  if (r1_len) memcpy(S + b1, r1, r1_len);
  unsigned long long vkey = ~0ull;
  for (uint32_t i = 0; i < n_rec; ++i) {
    const Rec& R = recs[i];
    RecGeom G = R.G;   // into S offsets
    G.h2 += b2; G.s2 += b2; G.p2 += b2; G.q2 += b2; G.e2 += b2;
    G.h1 += b1; G.s1 += b1; G.p1 += b1; G.q1 += b1; G.e1 += b1;
    uint8_t* dst[3] = {nullptr, o1 + K.off1[(size_t)R.bin] + R.d1, o2 + K.off2[(size_t)R.bin] + R.d2};
    // what k_emit does for one pair: header halves by scalar roles, byte ranges by copy_piece (lanes looped; the kernel's PTX movers do the same bytes)
    ReadPlan plan[3];
    const uint32_t rf = rf_pack(R.pl.rf_tcl, R.pl.rf_raw);   // what the kernel carries in RecMeta
    plan[2] = plan_read(P, G, 2, R.mo.sample, R.mo.bar_t, R.pl.len2, rf);
    plan[1] = plan_read(P, G, 1, R.mo.sample, R.mo.bar_t, P.paired ? R.pl.len1 : 0u, rf);
    for (int which = 2; which >= 1; --which) {   // scalar roles: the rewritten headers
      const ReadPlan& r = plan[which];
      const int slot = which == 2 ? HDR_SLOT0 : HDR_SLOT1;
      const bool own = r.hdr == slot && r.len != 0, lent = which == 2 && plan[1].hdr == HDR_SLOT0 && plan[1].len != 0;
      if (!own && !lent) continue;
      const uint32_t hl = own ? r.hl : plan[1].hl;   // illumina: both reads carry the same header
      if (hl <= HDR_SLOT_BYTES) {
        uint8_t* o = S + hs + (uint32_t)(slot - 1) * HDR_SLOT_BYTES;
        build_header(P, S, cs0, G, which, R.mo.bar_t, R.mo.umi_t, hl, o, (i & 1) ? 1 : 0, rf);   // the two halves, either order
        build_header(P, S, cs0, G, which, R.mo.bar_t, R.mo.umi_t, hl, o, (i & 1) ? 0 : 1, rf);
      } else {   // too long for a slot: at its destination(s)
        for (int to = 2; to >= 1; --to) {
          if (plan[to].hdr != slot || plan[to].len == 0) continue;
          build_header(P, S, cs0, G, which, R.mo.bar_t, R.mo.umi_t, hl, dst[to], 0, rf);
          build_header(P, S, cs0, G, which, R.mo.bar_t, R.mo.umi_t, hl, dst[to], 1, rf);
          if (plan[to].patch_at != 0xffffffffu) dst[to][plan[to].patch_at] = S[plan[to].patch_src];
        }
      }
    }
    for (int which = 2; which >= 1; --which) {   // byte ranges
      const ReadPlan& r = plan[which];
      if (r.len == 0) continue;
      for (int k = 0; k < 2; ++k) {
        const Piece pc = read_piece(r, k);
        for (int l = 0; l < 16; ++l)
          if (pc.n) copy_piece(dst[which] + pc.dst_off, S, pc.src, pc.n, l, 0xffffffffu, 0u);
      }
      if (r.trim) dst[which][r.len - 1] = '\n';
      if (r.hdr != HDR_VERBATIM && r.hl <= HDR_SLOT_BYTES)
        for (int l = 0; l < 16; ++l)
          copy_piece(dst[which], S, hs + (uint32_t)(r.hdr - 1) * HDR_SLOT_BYTES, r.hl, l, r.patch_at,
                     r.patch_at != 0xffffffffu ? S[r.patch_src] : 0u);
    }
    if (P.report_level > 0) {
      QcSix q = qc_pair(P, S, G, 1);
      const QcSix qb = qc_pair(P, S, G, 2);
      for (int k = 2; k < QC_SLOTS; ++k) q.v[k] = qb.v[k];
      uint64_t* st = c->stats.data() + (size_t)R.mo.sample * MGK_N_STATS;
      for (int k = P.paired ? 0 : 2; k < QC_SLOTS; ++k) st[qc_stat_column(k, P.paired)] += q.v[k];
    }
    if (P.validate) {
      const int v = validate_pair(P, S, G);
      if (v) {
        const unsigned long long k = ((unsigned long long)i << 4) | (unsigned)v;
        if (k < vkey) vkey = k;
      }
    }
  }
  free(S);
  if (rc != MGK_OK) { free(r2); free(r1); return rc; }
  free(r2);
  free(r1);
  K.res.seq = seq;
  K.res.n_records = n_rec;
  K.res.consumed_r2 = n_rec ? nl2[4 * n_rec - 1] + 1 : 0;
  K.res.consumed_r1 = (P.paired && n_rec) ? nl1[4 * n_rec - 1] + 1 : 0;
  K.res.out_r2 = o2;
  K.res.out_r1 = P.paired ? o1 : nullptr;
  K.res.out_r2_bytes = t2;
  K.res.out_r1_bytes = t1;
  K.res.seg_off_r2 = K.off2.data(); K.res.seg_len_r2 = K.len2.data();
  K.res.seg_off_r1 = K.off1.data(); K.res.seg_len_r1 = K.len1.data();
  K.res.unassigned = K.unassigned.data();
  K.res.n_unassigned = K.unassigned.size();
  if (vkey != ~0ull) { K.res.validate_error = validate_kind((int)(vkey & 15)); K.res.validate_record = vkey >> 4; }
  c->submitted.push_back(std::move(chunk));
  return MGK_OK;
}

int emu_collect(emu_ctx* c, mgk_result* r) {
  if (c->submitted.empty()) return MGK_ERR_STATE;
  c->collected.push_back(std::move(c->submitted.front()));
  c->submitted.pop_front();
  *r = c->collected.back()->res;
  return MGK_OK;
}
int emu_release(emu_ctx* c, uint64_t seq) {
  for (auto it = c->collected.begin(); it != c->collected.end(); ++it)
    if ((*it)->res.seq == seq) {
      c->collected.erase(it);
      return MGK_OK;
    }
  return MGK_ERR_STATE;
}
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

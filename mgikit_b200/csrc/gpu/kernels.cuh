// sm_100a kernels of the demultiplex hot path.
//
//   K1 k_scan_newlines   record-boundary scan of both read streams: 16-byte vector
//                        loads, SIMD-in-word '\n' detection, block scan, single-pass
//                        decoupled look-back -> sorted newline positions
//                        (replaces memchr_iter, /root/reference/src/file_utils.rs:95)
//   K2 k_match_plan      one thread per read pair: barcode extract + 2-bit pack +
//                        XOR/popc nearest-set match against the index table staged in
//                        shared memory, output lengths, stable in-tile bin offsets
//                        (find_matching_sample src/lib.rs:223-443, update_mismatches :1212)
//   K3 k_scan_*          per-bin exclusive scan over tiles -> order-preserving
//                        destination of every record (multisplit)
//   K4 k_scatter         one warp per read pair: header rewrite, barcode trim,
//                        destination-aligned 16-byte stores, fused Q30/quality sums
//                        (src/lib.rs:1188-1337, src/sample_data.rs:569-654)
//
// All four are HBM-bound integer/byte kernels; no tensor cores on purpose.
#pragma once

#include <cuda_runtime.h>

#include "device_logic.cuh"

namespace mgk {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_BYTES_PER_THREAD = 64;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_BYTES_PER_THREAD;  // 16 KiB of text per CTA
constexpr int TILE_RECORDS = 256;                                // read pairs per K2/K4 tile
constexpr int GROUP_TILES = 64;                                  // tiles per scan group

enum : uint32_t {
  ERR_NL_CAPACITY = 1u,    // more newlines than the record table holds
  ERR_REC_CAPACITY = 2u,   // more records than max_chunk_records
  ERR_OUT_CAPACITY = 4u,   // output larger than the output buffer
  ERR_PANIC = 8u           // a read hit the reference's stale-dictionary panic
};

struct __align__(16) RecMeta {
  uint32_t off2, off1;   // byte offset of the record inside its (tile, bin) run
  uint16_t sample, bin;
  int8_t bar_t, umi_t;
  uint8_t mism, flags;   // flags bit0: format error
};

struct Status {
  uint32_t n_records, consumed2, consumed1, error;
  unsigned long long fmt_record;     // lowest record index violating the layout (~0 = none)
  unsigned long long validate_key;   // lowest (record << 4 | priority)       (~0 = none)
  uint32_t n_unassigned, pad;
};

struct ChunkDev {
  const uint8_t* r2;
  const uint8_t* r1;
  uint32_t len2, len1;
  uint32_t* nl2;
  uint32_t* nl1;
  uint32_t nl_cap, max_records;
  unsigned long long* look2;   // decoupled look-back state per scan tile
  unsigned long long* look1;
  uint32_t* tickets;           // [2] dynamic tile ids
  uint32_t* counts;            // [2] newline totals
  RecMeta* meta;
  uint32_t* tile_bins;         // [tiles][2*total]: bytes, then absolute offsets
  uint32_t* group_sums;        // [groups][2*total]
  unsigned long long* seg;     // [4][total]: off2, len2, off1, len1
  uint8_t* out2;
  uint8_t* out1;
  unsigned long long out_cap2, out_cap1;
  uint32_t* unassigned;
  Status* status;
  unsigned long long* stats;   // [total][10]   running totals of the context
  unsigned long long* mism;    // [total][2m+2]
};

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {  // 0x80 in every byte that is '\n', exact
  const uint32_t x = w ^ 0x0a0a0a0au;
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t device_n_records(const ChunkDev& c, const Params& P) {
  const uint32_t a = c.counts[0] >> 2;
  const uint32_t n = P.paired ? min(a, c.counts[1] >> 2) : a;
  return min(n, c.max_records);
}

template <typename T>
__device__ __forceinline__ void copy_params_to_smem(T* dst, const T* src) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (uint32_t i = threadIdx.x; i < sizeof(T) / 4; i += blockDim.x) d[i] = s[i];
}

// ------------------------------------------------------------------ K1
// grid = (tiles of the longer stream, 2 streams).  Tile ids are taken from a
// ticket so that look-back only ever waits on CTAs that are already running.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_newlines(ChunkDev c, int n_files) {
  const int file = blockIdx.y;
  if (file >= n_files) return;
  const uint8_t* __restrict__ buf = file == 0 ? c.r2 : c.r1;
  const uint32_t len = file == 0 ? c.len2 : c.len1;
  uint32_t* __restrict__ nl = file == 0 ? c.nl2 : c.nl1;
  unsigned long long* look = file == 0 ? c.look2 : c.look1;
  const uint32_t ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
  if (blockIdx.x >= ntiles) return;

  __shared__ uint32_t s_tile, s_excl;
  __shared__ uint32_t s_warp[SCAN_THREADS / 32];
  if (threadIdx.x == 0) s_tile = atomicAdd(&c.tickets[file], 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_BYTES_PER_THREAD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  uint32_t f[16];
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t off = base + 16u * j;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (off < len) v = __ldg(reinterpret_cast<const uint4*>(buf + off));
    uint32_t w[4] = {nl_flags(v.x), nl_flags(v.y), nl_flags(v.z), nl_flags(v.w)};
    if (off >= len) {
      w[0] = w[1] = w[2] = w[3] = 0;
    } else if (off + 16u > len) {  // last, partial vector of the stream
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int nb = (int)(len - off) - 4 * k;
        if (nb <= 0) w[k] = 0;
        else if (nb < 4) w[k] &= (1u << (8 * nb)) - 1u;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[4 * j + k] = w[k];
      cnt += __popc(w[k]);
    }
  }
  // block exclusive scan of cnt
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0, agg = 0;
#pragma unroll
  for (int wi = 0; wi < SCAN_THREADS / 32; ++wi) {
    const uint32_t t = s_warp[wi];
    if (wi < warp) warp_off += t;
    agg += t;
  }
  // decoupled look-back: status word = flag << 32 | value; flag 1 aggregate, 2 inclusive prefix
  if (warp == 0) {
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_relaxed(&look[0], (2ull << 32) | agg);
    } else {
      if (lane == 0) st_relaxed(&look[tile], (1ull << 32) | agg);
      int look_at = (int)tile - 1;
      for (;;) {
        const int idx = look_at - lane;
        unsigned long long st = (2ull << 32);  // before the first tile: prefix 0
        if (idx >= 0) {
          do {
            st = ld_relaxed(&look[idx]);
          } while ((st >> 32) == 0);
        }
        const uint32_t flag = (uint32_t)(st >> 32), val = (uint32_t)st;
        const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2u);
        const int stop = pmask ? __ffs(pmask) - 1 : 31;  // nearest tile that already has its prefix
        uint32_t contrib = lane <= stop ? val : 0u;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
        excl += contrib;
        if (pmask) break;
        look_at -= 32;
      }
      if (lane == 0) st_relaxed(&look[tile], (2ull << 32) | (excl + agg));
    }
    if (lane == 0) {
      s_excl = excl;
      if (tile == ntiles - 1) c.counts[file] = excl + agg;
    }
  }
  __syncthreads();
  uint32_t idx = s_excl + warp_off + incl - cnt;
  if (cnt) {
    bool overflow = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      uint32_t w = f[k];
      while (w) {
        const int b = __ffs(w) - 1;
        w &= w - 1;
        if (idx < c.nl_cap) nl[idx] = base + 4u * k + (uint32_t)(b >> 3);
        else overflow = true;
        ++idx;
      }
    }
    if (overflow) atomicOr(&c.status->error, ERR_NL_CAPACITY);
  }
}

// ------------------------------------------------------------------ K2
__device__ __forceinline__ RecGeom load_geom(const ChunkDev& c, const Params& P, uint32_t rec) {
  RecGeom G;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(c.nl2) + rec);
  G.h2 = rec ? __ldg(c.nl2 + 4u * rec - 1u) + 1u : 0u;
  G.s2 = a.x + 1u;
  G.p2 = a.y + 1u;
  G.q2 = a.z + 1u;
  G.e2 = a.w;
  G.h1 = G.s1 = G.p1 = G.q1 = G.e1 = 0;
  if (P.paired) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(c.nl1) + rec);
    G.h1 = rec ? __ldg(c.nl1 + 4u * rec - 1u) + 1u : 0u;
    G.s1 = b.x + 1u;
    G.p1 = b.y + 1u;
    G.q1 = b.z + 1u;
    G.e1 = b.w;
  }
  return G;
}

// dynamic shared memory: uint64 codes[n_codes] | uint32 bin_acc[2*total]
__global__ void __launch_bounds__(TILE_RECORDS) k_match_plan(const Params* __restrict__ Pg, ChunkDev c) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Params P;
  copy_params_to_smem(&P, Pg);
  __syncthreads();
  uint64_t* s_codes = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* bin_acc = reinterpret_cast<uint32_t*>(s_codes + P.n_codes);
  for (int i = threadIdx.x; i < P.n_codes; i += blockDim.x) s_codes[i] = P.codes[i];
  const int total = P.total, B2 = 2 * total;
  const uint32_t raw2 = c.counts[0] >> 2, raw1 = P.paired ? c.counts[1] >> 2 : raw2;
  const uint32_t n_rec = device_n_records(c, P);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.status->n_records = n_rec;
    c.status->consumed2 = n_rec ? c.nl2[4u * n_rec - 1u] + 1u : 0u;
    c.status->consumed1 = (P.paired && n_rec) ? c.nl1[4u * n_rec - 1u] + 1u : 0u;
    if (min(raw2, raw1) > c.max_records) atomicOr(&c.status->error, ERR_REC_CAPACITY);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_tiles = (n_rec + TILE_RECORDS - 1) / TILE_RECORDS;

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < B2; i += blockDim.x) bin_acc[i] = 0;
    __syncthreads();
    const uint32_t rec = tile * TILE_RECORDS + threadIdx.x;
    const bool active = rec < n_rec;
    int bin = 0;
    uint32_t l2 = 0, l1 = 0;
    RecMeta mt;
    mt.off2 = mt.off1 = 0;
    mt.sample = mt.bin = 0;
    mt.bar_t = mt.umi_t = -1;
    mt.mism = mt.flags = 0;
    bool unassigned = false, is_amb = false;
    uint32_t bar_off = 0;
    if (active) {
      const RecGeom G = load_geom(c, P, rec);
      const uint32_t BL = (uint32_t)P.BL;
      const bool layout_ok = (G.p2 - G.s2 >= BL + 1u);
      MatchOut mo;
      mo.sample = total - 2;
      mo.mism = 0;
      mo.bar_t = mo.umi_t = -1;
      mo.panicked = 0;
      if (layout_ok) {
        bar_off = G.p2 - 1u - BL;
        mo = match_barcode(P, s_codes, c.r2 + bar_off);
      }
      if (mo.panicked) atomicOr(&c.status->error, ERR_PANIC);
      const RecPlan pl = record_plan(P, c.r2, G, mo.sample, mo.bar_t, mo.umi_t);
      if (pl.fmt_error) {
        atomicMin(&c.status->fmt_record, (unsigned long long)rec);
        mt.flags = 1;
      }
      bin = pl.matched ? P.writing[mo.sample] : mo.sample;
      l2 = pl.len2;
      l1 = pl.len1;
      mt.sample = (uint16_t)mo.sample;
      mt.bin = (uint16_t)bin;
      mt.bar_t = (int8_t)mo.bar_t;
      mt.umi_t = (int8_t)mo.umi_t;
      mt.mism = (uint8_t)mo.mism;
      if (!pl.fmt_error) {
        // update_mismatches(sample_id, curr_mismatch + 1, 1), src/lib.rs:1212
        atomicAdd(&c.mism[(size_t)mo.sample * P.mism_w + mo.mism + 1], 1ull);
        unassigned = !pl.matched && P.report_level > 1;  // src/lib.rs:1082-1088
        is_amb = mo.sample == total - 1;
      }
    }
    // Undetermined / Ambiguous reads -> compact list (one atomic per warp)
    {
      const uint32_t um = __ballot_sync(0xffffffffu, unassigned);
      if (um) {
        uint32_t basei = 0;
        if (lane == __ffs(um) - 1) basei = atomicAdd(&c.status->n_unassigned, (uint32_t)__popc(um));
        basei = __shfl_sync(0xffffffffu, basei, __ffs(um) - 1);
        if (unassigned) c.unassigned[basei + __popc(um & ((1u << lane) - 1u))] = bar_off | (is_amb ? 0x80000000u : 0u);
      }
    }
    // stable offsets inside the tile, per bin: lanes of a warp first ...
    const uint32_t act = __ballot_sync(0xffffffffu, active);
    uint32_t pre2 = 0, pre1 = 0, tot2 = 0, tot1 = 0;
    bool leader = false;
    uint32_t peers = 0;
    if (active) {
      peers = __match_any_sync(act, bin);
      leader = lane == __ffs(peers) - 1;
      uint32_t rem = peers;
      while (rem) {
        const int l = __ffs(rem) - 1;
        rem &= rem - 1;
        const uint32_t v2 = __shfl_sync(peers, l2, l), v1 = __shfl_sync(peers, l1, l);
        if (l < lane) {
          pre2 += v2;
          pre1 += v1;
        }
        tot2 += v2;
        tot1 += v1;
      }
    }
    // ... then warps in order through the shared per-bin accumulators
    for (int w = 0; w < TILE_RECORDS / 32; ++w) {
      if (warp == w && active) {
        const uint32_t b2 = bin_acc[bin], b1 = bin_acc[total + bin];
        __syncwarp(act);
        if (leader) {
          bin_acc[bin] = b2 + tot2;
          bin_acc[total + bin] = b1 + tot1;
        }
        mt.off2 = b2 + pre2;
        mt.off1 = b1 + pre1;
      }
      __syncthreads();
    }
    if (active) c.meta[rec] = mt;
    uint32_t* row = c.tile_bins + (size_t)tile * B2;
    for (int i = threadIdx.x; i < B2; i += blockDim.x) row[i] = bin_acc[i];
  }
}

// ------------------------------------------------------------------ K3
// (a) bytes per (group of 64 tiles, bin)
__global__ void k_scan_groups(const Params* __restrict__ Pg, ChunkDev c) {
  const int B2 = 2 * Pg->total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t n_tiles = (n_rec + TILE_RECORDS - 1) / TILE_RECORDS;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = id / B2, b = id % B2;
  if (g >= n_groups) return;
  const uint32_t t0 = g * GROUP_TILES, t1 = min(n_tiles, t0 + GROUP_TILES);
  uint32_t sum = 0;
#pragma unroll 8
  for (uint32_t t = t0; t < t1; ++t) sum += c.tile_bins[(size_t)t * B2 + b];
  c.group_sums[(size_t)g * B2 + b] = sum;
}

// (b) one CTA: exclusive scan over groups per bin, then over bins -> segment table
__global__ void __launch_bounds__(1024) k_scan_bins(const Params* __restrict__ Pg, ChunkDev c) {
  const int total = Pg->total, B2 = 2 * total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t n_tiles = (n_rec + TILE_RECORDS - 1) / TILE_RECORDS;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  unsigned long long* seg = c.seg;  // [off2 | len2 | off1 | len1] x total
  for (int b = threadIdx.x; b < B2; b += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t g = 0; g < n_groups; ++g) {
      const uint32_t v = c.group_sums[(size_t)g * B2 + b];
      c.group_sums[(size_t)g * B2 + b] = run;
      run += v;
    }
    seg[(b < total ? 1 : 3) * total + (b % total)] = run;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < 2) {  // warp 0: barcode-read bins, warp 1: paired-read bins
    unsigned long long* len = seg + (warp == 0 ? 1 : 3) * total;
    unsigned long long* off = seg + (warp == 0 ? 0 : 2) * total;
    unsigned long long carry = 0;
    for (int b0 = 0; b0 < total; b0 += 32) {
      const int b = b0 + lane;
      const unsigned long long v = b < total ? len[b] : 0ull;
      unsigned long long incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (b < total) off[b] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0 && carry > (warp == 0 ? c.out_cap2 : c.out_cap1)) atomicOr(&c.status->error, ERR_OUT_CAPACITY);
  }
}

// (c) absolute destination of every (tile, bin) run
__global__ void k_scan_tiles(const Params* __restrict__ Pg, ChunkDev c) {
  const int total = Pg->total, B2 = 2 * total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t n_tiles = (n_rec + TILE_RECORDS - 1) / TILE_RECORDS;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = id / B2, b = id % B2;
  if (g >= n_groups) return;
  const uint32_t t0 = g * GROUP_TILES, t1 = min(n_tiles, t0 + GROUP_TILES);
  uint32_t run = c.group_sums[(size_t)g * B2 + b] + (uint32_t)c.seg[(b < (uint32_t)total ? 0 : 2) * total + (b % total)];
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t v = c.tile_bins[(size_t)t * B2 + b];
    c.tile_bins[(size_t)t * B2 + b] = run;
    run += v;
  }
}

// ------------------------------------------------------------------ K4
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

__global__ void __launch_bounds__(TILE_RECORDS) k_scatter(const Params* __restrict__ Pg, ChunkDev c) {
  __shared__ Params P;
  copy_params_to_smem(&P, Pg);
  __syncthreads();
  const int total = P.total, B2 = 2 * total;
  const uint32_t n_rec = device_n_records(c, P);
  const uint32_t n_tiles = (n_rec + TILE_RECORDS - 1) / TILE_RECORDS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t BL = (uint32_t)P.BL, wbl = P.keep_barcode ? 0u : BL;
  if (c.status->error & (ERR_OUT_CAPACITY | ERR_NL_CAPACITY)) return;  // nothing safe to write

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t* tb = c.tile_bins + (size_t)tile * B2;
    const uint32_t rec0 = tile * TILE_RECORDS + warp * 32;
    for (uint32_t r = 0; r < 32; ++r) {
      const uint32_t rec = rec0 + r;
      if (rec >= n_rec) break;
      const RecMeta mt = c.meta[rec];
      if (mt.flags & 1) continue;  // layout error: reported through status, nothing emitted
      const RecGeom G = load_geom(c, P, rec);
      const int sample = mt.sample, bin = mt.bin;
      const RecPlan pl = record_plan(P, c.r2, G, sample, mt.bar_t, mt.umi_t);
      uint8_t* d2 = c.out2 + tb[bin] + mt.off2;
      uint8_t* d1 = P.paired ? c.out1 + tb[total + bin] + mt.off1 : nullptr;
      if (!pl.matched) {
        copy_span(d2, c.r2 + G.h2, G.e2 + 1u - G.h2, lane, 32);
        if (P.paired) copy_span(d1, c.r1 + G.h1, G.e1 + 1u - G.h1, lane, 32);
      } else {
        const bool rewrite = P.out_mode != MGK_OUT_MGI;
        if (P.read2_has_sequence) {
          uint32_t n = 0;
          if (rewrite) {
            SegList s;
            header_segments(P, c.r2, c.r1, G, pl, mt.bar_t, mt.umi_t, 2, s);
            n = emit_segments(d2, s, lane, 32);
          }
          const uint32_t a0 = rewrite ? G.s2 - 1u : G.h2;         // src/lib.rs:1293 / unchanged header_start
          const uint32_t na = G.p2 - wbl - 1u - a0;               // :1321-1323
          copy_span(d2 + n, c.r2 + a0, na, lane, 32);
          const uint32_t nb = G.e2 - wbl - (G.p2 - 1u);           // :1324-1326
          copy_span(d2 + n + na, c.r2 + G.p2 - 1u, nb, lane, 32);
          if (lane == 0) d2[n + na + nb] = '\n';                  // :1327
        }
        if (P.paired) {
          uint32_t n = 0;
          if (rewrite) {
            SegList s;
            header_segments(P, c.r2, c.r1, G, pl, mt.bar_t, mt.umi_t, 1, s);
            n = emit_segments(d1, s, lane, 32);
          }
          const uint32_t a0 = rewrite ? G.s1 - 1u : G.h1;
          copy_span(d1 + n, c.r1 + a0, G.e1 + 1u - a0, lane, 32);  // :1328-1331
        }
      }
      if (P.report_level > 0) {  // sum_qc x3 + update_stats x6, src/lib.rs:1188-1210
        Qc qa, qb, qc;
        qa.sum = qa.high = 0;
        if (P.paired) qa = qc_partial(c.r1 + G.q1, G.e1 - G.q1, lane, 32);
        qb = qc_partial(c.r2 + G.e2 - BL, BL, lane, 32);
        qc = qc_partial(c.r2 + G.q2, G.e2 - BL - G.q2, lane, 32);
        const uint32_t v0 = warp_sum(qa.high), v1 = warp_sum(qa.sum), v2 = warp_sum(qb.high), v3 = warp_sum(qb.sum),
                       v4 = warp_sum(qc.high), v5 = warp_sum(qc.sum);
        const int shift = P.paired ? 1 : 0;
        unsigned long long* st = c.stats + (size_t)sample * MGK_N_STATS;
        if (lane == 0 && P.paired) atomicAdd(st + 0, (unsigned long long)v0);
        if (lane == 1 && P.paired) atomicAdd(st + 6, (unsigned long long)v1);
        if (lane == 2) atomicAdd(st + 2, (unsigned long long)v2);
        if (lane == 3) atomicAdd(st + 8, (unsigned long long)v3);
        if (lane == 4) atomicAdd(st + shift, (unsigned long long)v4);
        if (lane == 5) atomicAdd(st + 6 + shift, (unsigned long long)v5);
      }
      if (P.validate) {
        int v = validate_partial(P, c.r2, c.r1, G, lane, 32);
        v = v ? v : 15;
        v = __reduce_min_sync(0xffffffffu, v);
        if (v != 15 && lane == 0) atomicMin(&c.status->validate_key, ((unsigned long long)rec << 4) | (unsigned)v);
      }
    }
  }
}

// ------------------------------------------------------------------ stage kernels (parity tests)
__global__ void k_match_only(const Params* __restrict__ Pg, const uint8_t* __restrict__ bars, uint32_t n, int32_t* sample,
                             int32_t* mism, uint32_t* error) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Params P;
  copy_params_to_smem(&P, Pg);
  __syncthreads();
  uint64_t* s_codes = reinterpret_cast<uint64_t*>(smem_raw);
  for (int i = threadIdx.x; i < P.n_codes; i += blockDim.x) s_codes[i] = P.codes[i];
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const MatchOut mo = match_barcode(P, s_codes, bars + (size_t)i * P.BL);
    if (mo.panicked) atomicOr(error, ERR_PANIC);
    sample[i] = mo.sample;
    mism[i] = mo.mism;
  }
}

}  // namespace mgk

// sm_100a kernels of the demultiplex hot path.
//
//   K1 k_scan_newlines   record-boundary scan of both read streams: every warp streams its own
//      k_gather_newlines range through a TMA-fed shared-memory ring, SIMD-in-word '\n' detection,
//                        warp scan -> sorted newline positions
//                        (replaces memchr_iter, /root/reference/src/file_utils.rs:95)
//   K2 k_match_plan      one thread per read pair: exact-match hash probe of the raw index bytes,
//                        misses by a whole warp (2-bit pack + XOR/popc nearest-set match against the
//                        index table staged in shared memory), output lengths, stable in-tile bin offsets
//                        (find_matching_sample src/lib.rs:223-443, update_mismatches :1212)
//   K3 k_scan_*          per-bin exclusive scan over tiles -> order-preserving
//                        destination of every record (multisplit)
//   K4 k_emit            persistent: TMA-staged tiles, scalar roles (header rewrite, Q30/quality sums,
//                        --validate), quarter-warp movers with destination-aligned 16-byte stores
//                        (src/lib.rs:1188-1337, src/sample_data.rs:569-654)
//
// All of them are HBM-bound integer/byte kernels; no tensor cores on purpose.
#pragma once

#include <cuda_runtime.h>

#include "emit_logic.cuh"

namespace mgk {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ROWS = 2;                                      // 32-byte pieces per thread and tile
constexpr int SCAN_TILE = SCAN_THREADS * 32 * SCAN_ROWS * 2;     // 32 KiB of text per CTA and stage (SCAN_HALVES pieces per warp)
constexpr int SCAN_STAGES = 2;                                    // bulk copies in flight per warp (3 CTAs of 8 warps per SM)
constexpr int SCAN_HALVES = 2;                                    // 2 KiB pieces per stage: one bulk copy, one wait per 4 KiB
constexpr int TILE_RECORDS = 256;                                // read pairs per K2/K4 tile
constexpr int GROUP_TILES = 64;                                  // tiles per scan group

enum : uint32_t {
  ERR_NL_CAPACITY = 1u,    // more newlines than the record table holds
  ERR_REC_CAPACITY = 2u,   // more records than max_chunk_records
  ERR_OUT_CAPACITY = 4u,   // output larger than the output buffer
  ERR_PANIC = 8u,          // a read hit the reference's stale-dictionary panic
  ERR_REC_TOO_LONG = 16u   // one read pair does not fit a shared-memory stage
};

struct __align__(16) RecMeta {
  uint32_t off2, off1;   // byte offset of the record inside its (tile, bin) run, barcode read / paired read
  uint16_t sample;       // 0..S+1; the bin is writing[sample] for a matched read, the sample itself otherwise
  uint16_t len2, len1;   // output bytes of the two records
  uint8_t tpl;           // template that cut the barcode text (low nibble) and the UMI (high nibble); 15 = none
                         // (reformat: bytes of the input header carried over, rf_pack())
  uint8_t flags;         // bit0: format error; bit1 (reformat): MGI raw header
};
static_assert(sizeof(RecMeta) == 16, "RecMeta is one 16-byte row");

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
  uint32_t* stg2;              // newline positions as found, one run per scan CTA (scan_cap entries each)
  uint32_t* stg1;
  uint32_t* range_counts;      // [2][scan_grid * SCAN_WARPS] newlines found by each scan warp in its range
  uint32_t scan_grid, scan_cap;
  uint32_t* counts;            // [2] newline totals
  RecMeta* meta;
  uint32_t* tile_bins;         // [tiles][bins_stride]: bytes, then absolute offsets (rows 16-byte aligned)
  uint32_t bins_stride, emit_cap;   // emit_cap: bytes of text one shared-memory stage of k_emit holds
  uint32_t* group_sums;        // [groups][2*total]
  unsigned long long* seg;     // [4][total]: off2, len2, off1, len1
  uint8_t* out2;
  uint8_t* out1;
  unsigned long long out_cap2, out_cap1;
  uint32_t* unassigned;
  Status* status;
  unsigned long long* stats;   // [total][10]   running totals of the context
  unsigned long long* mism;    // [total][2m+2]
  unsigned long long* prof;    // optional [grid][warps][6] cycle counters of k_emit (MGK_PROFILE=1), else null
};

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ uint32_t device_n_records(const ChunkDev& c, const Params& P) {
  const uint32_t a = c.counts[0] >> 2;
  const uint32_t n = P.paired ? min(a, c.counts[1] >> 2) : a;
  return min(n, c.max_records);
}

// Records per plan tile (the unit of the multisplit: one row of tile_bins, one CTA pass of k_match_plan).  256 when
// half of it -- what k_emit stages at a time -- fits a shared-memory stage at the chunk's mean record size; with many
// samples or long records less fits, and the plan tile shrinks to twice what does (two staged tiles per plan tile
// instead of three uneven ones).  Every kernel derives it from the same device-resident numbers.
__device__ __forceinline__ uint32_t device_tile_records(const ChunkDev& c, const Params& P, uint32_t n_rec) {
  if (n_rec == 0 || c.emit_cap == 0) return (uint32_t)TILE_RECORDS;
  const unsigned long long bytes = (unsigned long long)c.nl2[4u * n_rec - 1u] + (P.paired ? c.nl1[4u * n_rec - 1u] : 0u) + 2ull;
  const uint32_t mean = (uint32_t)(bytes / n_rec) + 1u;
  const uint32_t fit = (c.emit_cap - 256u) / mean;
  if (fit >= (uint32_t)TILE_RECORDS / 2u || fit < (uint32_t)TILE_RECORDS / 4u) return (uint32_t)TILE_RECORDS;
  return 2u * fit;
}

template <typename T>
__device__ __forceinline__ void copy_params_to_smem(T* dst, const T* src) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (uint32_t i = threadIdx.x; i < sizeof(T) / 4; i += blockDim.x) d[i] = s[i];
}

// ------------------------------------------------------------------ TMA / mbarrier helpers
// shared-memory loads by 32-bit shared-window address (the movers of k_emit: no generic-address arithmetic, predicates
// evaluated once for a group of loads; registers of a predicated-off load keep their value)
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// One lane's share (of a quarter-warp's) of a destination range g[0, N) fed by two source runs: byte j comes from
// shared-window address sa + j if j < n0, else sb + j.  hb bytes lie in front of the first 16-byte boundary of g,
// nfull whole units follow; the lane takes units l8, l8 + 8, l8 + 16 (unit u < ubf lies wholly in the first run, the
// others in the second, unit ubs -- if any -- straddles the two and goes byte by byte), the bytes l8 and l8 + 8 of
// the head, of the straddling unit and of the tail (z = first tail byte + l8), and lane 0 a closing '\n' if nl.
// Stores carry their predicates; loads do not (a predicated load would make its register live across the whole
// loop): they may read up to 512 bytes past the range, which the shared-memory layout leaves room for.
__device__ __forceinline__ void move_pair_lane(uint8_t* g, uint32_t l8, uint32_t hb, uint32_t nfull, uint32_t N, uint32_t n0, uint32_t sa,
                                               uint32_t sb, uint32_t ubf, uint32_t ubs, uint32_t z, uint32_t nl) {
  asm volatile(
      "{\n"
      ".reg .pred p0, p1, p2, c, v0, v1, v2, v3, v4, v5, pn;\n"
      ".reg .u32 u1, u2, x0, a0, a1, a2, b0, b1, b2, s0, s1, s2, t, j0, j1, j2, j3, j4, j5, e0, e1, e2, e3, e4, e5, k;\n"
      ".reg .u32 w<15>;\n"
      ".reg .u64 gu, gh, gt, gb, gn, o;\n"
      "add.u32 u1, %1, 8;\n"
      "add.u32 u2, %1, 16;\n"
      "setp.lt.u32 p0, %1, %3;\n"
      "setp.ne.and.u32 p0, %1, %9, p0;\n"
      "setp.lt.u32 p1, u1, %3;\n"
      "setp.ne.and.u32 p1, u1, %9, p1;\n"
      "setp.lt.u32 p2, u2, %3;\n"
      "setp.ne.and.u32 p2, u2, %9, p2;\n"
      "shl.b32 x0, %1, 4;\n"
      "add.u32 x0, x0, %2;\n"
      "setp.lt.u32 c, %1, %8;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 a0, t, x0;\n"
      "setp.lt.u32 c, u1, %8;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 a1, t, x0;\n"
      "setp.lt.u32 c, u2, %8;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 a2, t, x0;\n"
      "shl.b32 s0, a0, 3;\n"
      "shl.b32 s1, a1, 3;\n"
      "shl.b32 s2, a2, 3;\n"
      "and.b32 b0, a0, 0xfffffffc;\n"
      "and.b32 b1, a1, 0xfffffffc;\n"
      "and.b32 b2, a2, 0xfffffffc;\n"
      "ld.shared.u32 w0, [b0];\n"
      "ld.shared.u32 w1, [b0+4];\n"
      "ld.shared.u32 w2, [b0+8];\n"
      "ld.shared.u32 w3, [b0+12];\n"
      "ld.shared.u32 w4, [b0+16];\n"
      "ld.shared.u32 w5, [b1+128];\n"
      "ld.shared.u32 w6, [b1+132];\n"
      "ld.shared.u32 w7, [b1+136];\n"
      "ld.shared.u32 w8, [b1+140];\n"
      "ld.shared.u32 w9, [b1+144];\n"
      "ld.shared.u32 w10, [b2+256];\n"
      "ld.shared.u32 w11, [b2+260];\n"
      "ld.shared.u32 w12, [b2+264];\n"
      "ld.shared.u32 w13, [b2+268];\n"
      "ld.shared.u32 w14, [b2+272];\n"
      // the bytes: head (j0, j1), straddling unit (j2, j3), tail (j4, j5)
      "mov.u32 j0, %1;\n"
      "add.u32 j1, %1, 8;\n"
      "shl.b32 k, %9, 4;\n"
      "add.u32 k, k, %2;\n"
      "add.u32 j2, k, %1;\n"
      "add.u32 j3, j2, 8;\n"
      "mov.u32 j4, %10;\n"
      "add.u32 j5, %10, 8;\n"
      "setp.lt.u32 v0, j0, %2;\n"
      "setp.lt.u32 v1, j1, %2;\n"
      "setp.ne.u32 v2, %9, 0xffffffff;\n"
      "setp.ne.u32 v3, %9, 0xffffffff;\n"
      "setp.lt.u32 v4, j4, %4;\n"
      "setp.lt.u32 v5, j5, %4;\n"
      "setp.lt.u32 c, j0, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j0;\n"
      "ld.shared.u8 e0, [t];\n"
      "setp.lt.u32 c, j1, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j1;\n"
      "ld.shared.u8 e1, [t];\n"
      "setp.lt.u32 c, j2, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j2;\n"
      "ld.shared.u8 e2, [t];\n"
      "setp.lt.u32 c, j3, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j3;\n"
      "ld.shared.u8 e3, [t];\n"
      "setp.lt.u32 c, j4, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j4;\n"
      "ld.shared.u8 e4, [t];\n"
      "setp.lt.u32 c, j5, %5;\n"
      "selp.u32 t, %6, %7, c;\n"
      "add.u32 t, t, j5;\n"
      "ld.shared.u8 e5, [t];\n"
      // realign (shf takes the shift amount modulo 32: 8 * (address & 3))
      "shf.r.wrap.b32 w0, w0, w1, s0;\n"
      "shf.r.wrap.b32 w1, w1, w2, s0;\n"
      "shf.r.wrap.b32 w2, w2, w3, s0;\n"
      "shf.r.wrap.b32 w3, w3, w4, s0;\n"
      "shf.r.wrap.b32 w5, w5, w6, s1;\n"
      "shf.r.wrap.b32 w6, w6, w7, s1;\n"
      "shf.r.wrap.b32 w7, w7, w8, s1;\n"
      "shf.r.wrap.b32 w8, w8, w9, s1;\n"
      "shf.r.wrap.b32 w10, w10, w11, s2;\n"
      "shf.r.wrap.b32 w11, w11, w12, s2;\n"
      "shf.r.wrap.b32 w12, w12, w13, s2;\n"
      "shf.r.wrap.b32 w13, w13, w14, s2;\n"
      "cvt.u64.u32 o, x0;\n"
      "add.u64 gu, %0, o;\n"
      "@p0 st.global.v4.u32 [gu], {w0, w1, w2, w3};\n"
      "@p1 st.global.v4.u32 [gu+128], {w5, w6, w7, w8};\n"
      "@p2 st.global.v4.u32 [gu+256], {w10, w11, w12, w13};\n"
      "cvt.u64.u32 o, j0;\n"
      "add.u64 gh, %0, o;\n"
      "@v0 st.global.u8 [gh], e0;\n"
      "@v1 st.global.u8 [gh+8], e1;\n"
      "cvt.u64.u32 o, j2;\n"
      "add.u64 gb, %0, o;\n"
      "@v2 st.global.u8 [gb], e2;\n"
      "@v3 st.global.u8 [gb+8], e3;\n"
      "cvt.u64.u32 o, j4;\n"
      "add.u64 gt, %0, o;\n"
      "@v4 st.global.u8 [gt], e4;\n"
      "@v5 st.global.u8 [gt+8], e5;\n"
      "setp.ne.u32 pn, %11, 0;\n"
      "setp.eq.and.u32 pn, %1, 0, pn;\n"
      "cvt.u64.u32 o, %4;\n"
      "add.u64 gn, %0, o;\n"
      "mov.u32 k, 10;\n"
      "@pn st.global.u8 [gn], k;\n"
      "}\n" ::"l"(g),
      "r"(l8), "r"(hb), "r"(nfull), "r"(N), "r"(n0), "r"(sa), "r"(sb), "r"(ubf), "r"(ubs), "r"(z), "r"(nl)
      : "memory");
}
// The same for a short range out of one source run (a staged header): unit l8, bytes l8 and l8 + 8 of head and tail.
__device__ __forceinline__ void move_small_lane(uint8_t* g, uint32_t l8, uint32_t hb, uint32_t nfull, uint32_t n, uint32_t sa, uint32_t z) {
  asm volatile(
      "{\n"
      ".reg .pred p0, v0, v1, v4, v5;\n"
      ".reg .u32 x0, a0, b0, s0, t, j1, j5, e0, e1, e4, e5;\n"
      ".reg .u32 w<5>;\n"
      ".reg .u64 gu, gh, gt, o;\n"
      "setp.lt.u32 p0, %1, %3;\n"
      "shl.b32 x0, %1, 4;\n"
      "add.u32 x0, x0, %2;\n"
      "add.u32 a0, %5, x0;\n"
      "shl.b32 s0, a0, 3;\n"
      "and.b32 b0, a0, 0xfffffffc;\n"
      "ld.shared.u32 w0, [b0];\n"
      "ld.shared.u32 w1, [b0+4];\n"
      "ld.shared.u32 w2, [b0+8];\n"
      "ld.shared.u32 w3, [b0+12];\n"
      "ld.shared.u32 w4, [b0+16];\n"
      "add.u32 j1, %1, 8;\n"
      "add.u32 j5, %6, 8;\n"
      "setp.lt.u32 v0, %1, %2;\n"
      "setp.lt.u32 v1, j1, %2;\n"
      "setp.lt.u32 v4, %6, %4;\n"
      "setp.lt.u32 v5, j5, %4;\n"
      "add.u32 t, %5, %1;\n"
      "ld.shared.u8 e0, [t];\n"
      "ld.shared.u8 e1, [t+8];\n"
      "add.u32 t, %5, %6;\n"
      "ld.shared.u8 e4, [t];\n"
      "ld.shared.u8 e5, [t+8];\n"
      "shf.r.wrap.b32 w0, w0, w1, s0;\n"
      "shf.r.wrap.b32 w1, w1, w2, s0;\n"
      "shf.r.wrap.b32 w2, w2, w3, s0;\n"
      "shf.r.wrap.b32 w3, w3, w4, s0;\n"
      "cvt.u64.u32 o, x0;\n"
      "add.u64 gu, %0, o;\n"
      "@p0 st.global.v4.u32 [gu], {w0, w1, w2, w3};\n"
      "cvt.u64.u32 o, %1;\n"
      "add.u64 gh, %0, o;\n"
      "@v0 st.global.u8 [gh], e0;\n"
      "@v1 st.global.u8 [gh+8], e1;\n"
      "cvt.u64.u32 o, %6;\n"
      "add.u64 gt, %0, o;\n"
      "@v4 st.global.u8 [gt], e4;\n"
      "@v5 st.global.u8 [gt+8], e5;\n"
      "}\n" ::"l"(g),
      "r"(l8), "r"(hb), "r"(nfull), "r"(n), "r"(sa), "r"(z)
      : "memory");
}
__device__ __forceinline__ long long tick(bool on) {   // cycle counter, ordered with memory operations (profiling aid)
  long long t = 0;
  if (on) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map): 16-byte aligned both sides, size % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


// ------------------------------------------------------------------ K1
// Newline scan without any dependency between warps.  grid = (scan_grid, files); every WARP owns
// ONE contiguous range of its stream (scan_grid * SCAN_WARPS ranges per stream) and walks it
// through its own ring of shared-memory stages that the TMA engine keeps full (cp.async.bulk +
// mbarrier, issued by the warp's lane 0), so the bytes in flight per SM are set by the ring depth,
// not by registers, and no CTA-wide barrier is ever taken.  The index of a newline inside its
// range needs only the warp's own running count; the positions go to a per-range run of a staging
// array and k_gather_newlines then packs the runs into the final table (32 B per read pair,
// L2-resident) once the per-range totals are known.  Order inside a warp's 2 KiB: (row, lane),
// lane l = bytes 32l..32l+31 of the row.
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int SCAN_PIECE = 32 * 32 * SCAN_ROWS;          // bytes a warp scans at a time
constexpr int SCAN_WTILE = SCAN_PIECE * SCAN_HALVES;     // bytes per warp and stage
static_assert(SCAN_TILE == SCAN_WARPS * SCAN_WTILE, "stage size");
// '\n' flags of the words a (earlier bytes) and b as 8 bits in byte order at the top of the product: the flags sit at
// bits 3,11,19,27 (a) and 7,15,23,31 (b), the four partial products of 0x00204081 put them at 24..27 and 28..31, and no
// two of the 32 partial-product bits coincide (so no carries).  k0a / k7f: 0x0a0a0a0a / 0x7f7f7f7f held in registers
// (lop3 takes one immediate at most).  Exact for every byte value.
__device__ __forceinline__ uint32_t nl_raw(uint32_t w, uint32_t k0a, uint32_t k7f) {   // bit 7 of a byte CLEAR <=> it is '\n'
  uint32_t x7, r;
  asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(x7) : "r"(w), "r"(k0a), "r"(k7f));            // (w ^ k0a) & k7f
  const uint32_t t = x7 + 0x7f7f7f7fu;
  asm("lop3.b32 %0, %1, %2, %3, 0xf6;" : "=r"(r) : "r"(t), "r"(w), "r"(k0a));               // t | (w ^ k0a)
  return r;
}
__device__ __forceinline__ uint32_t nl_pair(uint32_t a, uint32_t b, uint32_t k0a, uint32_t k7f) {
  const uint32_t za = nl_raw(a, k0a, k7f) >> 4, zb = nl_raw(b, k0a, k7f);
  uint32_t x;
  asm("lop3.b32 %0, %1, %2, 0x08080808, 0xe4;" : "=r"(x) : "r"(za), "r"(zb));                // bits 3,11,.. of za, the others of zb
  return (~x & 0x88888888u) * 0x00204081u;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_newlines(ChunkDev c, int n_files) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_full[SCAN_WARPS][SCAN_STAGES];
  const int file = blockIdx.y;
  if (file >= n_files) return;
  const uint8_t* __restrict__ buf = file == 0 ? c.r2 : c.r1;
  const uint32_t len = file == 0 ? c.len2 : c.len1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_ranges = gridDim.x * SCAN_WARPS, range = blockIdx.x * SCAN_WARPS + (uint32_t)warp;
  uint32_t* __restrict__ out = (file == 0 ? c.stg2 : c.stg1) + (size_t)range * c.scan_cap;
  const uint32_t ntiles = (len + SCAN_WTILE - 1) / SCAN_WTILE;
  const uint32_t per = (ntiles + n_ranges - 1) / n_ranges;
  const uint32_t t0 = min(range * per, ntiles), t1 = min(t0 + per, ntiles);
  unsigned char* ring = smem + (size_t)warp * (SCAN_STAGES * SCAN_WTILE);
  uint64_t* full = s_full[warp];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < SCAN_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](uint32_t t) {   // lane 0: tile t -> stage (t - t0) % SCAN_STAGES of the warp's ring
    const int s = (int)((t - t0) % SCAN_STAGES);
    const uint32_t off = t * SCAN_WTILE;
    const uint32_t bytes = (min(len - off, (uint32_t)SCAN_WTILE) + 15u) & ~15u;   // chunk buffers are padded
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(ring + (size_t)s * SCAN_WTILE, buf + off, bytes, &full[s]);
  };
  if (lane == 0)
    for (uint32_t t = t0; t < t1 && t < t0 + SCAN_STAGES; ++t) issue(t);
  uint32_t k0a, k7f;
  asm volatile("mov.u32 %0, 0x0a0a0a0a;" : "=r"(k0a));   // opaque: kept in registers
  asm volatile("mov.u32 %0, 0x7f7f7f7f;" : "=r"(k7f));

  uint32_t run = 0;   // newlines of this range so far
  bool overflow = false;
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t it = t - t0;
    const int s = (int)(it % SCAN_STAGES);
    mbar_wait(&full[s], (it / SCAN_STAGES) & 1u);
#pragma unroll 1
    for (int h = 0; h < SCAN_HALVES; ++h) {
      const unsigned char* st = ring + (size_t)s * SCAN_WTILE + (size_t)h * SCAN_PIECE + 32u * (uint32_t)lane;
      const uint32_t wbase = t * SCAN_WTILE + (uint32_t)h * SCAN_PIECE + 32u * (uint32_t)lane;
      if (t * SCAN_WTILE + (uint32_t)h * SCAN_PIECE >= len) break;                          // warp-uniform
      const bool last_piece = t * SCAN_WTILE + (uint32_t)(h + 1) * SCAN_PIECE > len;      // only the stream's last piece holds bytes beyond len
      uint32_t m[SCAN_ROWS];   // bit p set <=> byte p of the lane's 32 bytes is a newline
      uint32_t cp = 0;         // newline counts of the rows, 16 bits each (a row holds at most 1024)
#pragma unroll
      for (int j = 0; j < SCAN_ROWS; ++j) {
        const uint32_t off = wbase + 1024u * j;
        const uint4 va = *reinterpret_cast<const uint4*>(st + 1024u * j);
        const uint4 vb = *reinterpret_cast<const uint4*>(st + 1024u * j + 16u);
        const uint32_t lo = __byte_perm(nl_pair(va.x, va.y, k0a, k7f), nl_pair(va.z, va.w, k0a, k7f), 0x0073u);
        const uint32_t hi = __byte_perm(nl_pair(vb.x, vb.y, k0a, k7f), nl_pair(vb.z, vb.w, k0a, k7f), 0x0073u);
        uint32_t mm = __byte_perm(lo, hi, 0x5410u);
        if (last_piece && off + 32u > len) mm = off >= len ? 0u : mm & ((1u << (len - off)) - 1u);   // bytes beyond the stream
        m[j] = mm;
        cp += (uint32_t)__popc(mm) << (16 * j);
      }
      uint32_t ip = cp;   // inclusive scan over the lanes
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, ip, d);
        if (lane >= d) ip += x;
      }
      const uint32_t tot = __shfl_sync(0xffffffffu, ip, 31);
      const uint32_t row0 = tot & 0xffffu, wtotal = row0 + (tot >> 16);
      if (run + wtotal > c.scan_cap) {   // warp-uniform
        overflow = true;
      } else {
        uint32_t* orun = out + run;
#pragma unroll
        for (int j = 0; j < SCAN_ROWS; ++j) {
          uint32_t mm = m[j];
          if (mm == 0) continue;
          const uint32_t cj = (cp >> (16 * j)) & 0xffffu, ij = (ip >> (16 * j)) & 0xffffu;
          uint32_t* o = orun + ((j ? row0 : 0u) + ij - cj);
          const uint32_t off = wbase + 1024u * j;
          o[0] = off + (uint32_t)(__ffs(mm) - 1);          // a record's "\n+\n" makes two per lane common: straight-line
          mm &= mm - 1;
          if (mm) {
            o[1] = off + (uint32_t)(__ffs(mm) - 1);
            mm &= mm - 1;
            for (uint32_t q = 2; mm; ++q) {
              o[q] = off + (uint32_t)(__ffs(mm) - 1);
              mm &= mm - 1;
            }
          }
        }
      }
      run += wtotal;
    }
    __syncwarp();   // every lane has read its part of stage s
    if (lane == 0 && t + SCAN_STAGES < t1) issue(t + SCAN_STAGES);
  }
  if (overflow && lane == 0) atomicOr(&c.status->error, ERR_NL_CAPACITY);
  if (lane == 0) c.range_counts[(size_t)file * n_ranges + range] = run;
}

// Packs the per-range runs into the final newline table and leaves the total in counts[file].
// grid = (ranges, files).
__global__ void __launch_bounds__(256) k_gather_newlines(ChunkDev c, int n_files) {
  __shared__ uint32_t s_part[8];
  const int file = blockIdx.y;
  if (file >= n_files) return;
  const uint32_t range = blockIdx.x, n_ranges = gridDim.x;
  const uint32_t* rc = c.range_counts + (size_t)file * n_ranges;
  {   // newlines in front of this range: all threads, independent loads
    uint32_t sum = 0;
    for (uint32_t i = threadIdx.x; i < range; i += blockDim.x) sum += rc[i];
    sum = __reduce_add_sync(0xffffffffu, sum);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = sum;
  }
  __syncthreads();
  uint32_t prefix = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) prefix += s_part[w];
  const uint32_t n = min(rc[range], c.scan_cap);
  const uint32_t* __restrict__ src = (file == 0 ? c.stg2 : c.stg1) + (size_t)range * c.scan_cap;
  uint32_t* __restrict__ nl = file == 0 ? c.nl2 : c.nl1;
  bool overflow = false;
  for (uint32_t i0 = 0; i0 < n; i0 += 4u * blockDim.x) {   // four loads in flight per thread
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i = i0 + k * blockDim.x + threadIdx.x;
      v[k] = i < n ? src[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i = i0 + k * blockDim.x + threadIdx.x;
      if (i < n) {
        if (prefix + i < c.nl_cap) nl[prefix + i] = v[k];
        else overflow = true;
      }
    }
  }
  if (overflow) atomicOr(&c.status->error, ERR_NL_CAPACITY);
  if (range == n_ranges - 1 && threadIdx.x == 0) c.counts[file] = prefix + rc[range];   // every newline, stored or not
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

// Matching of one barcode per thread, CTA-wide.  With a single template of indexes up to 16
// bases (every BASELINE configuration) the nearest-set of an observed index is found in two
// steps that keep the warps converged:
//   1. every thread probes the exact-match hash table with the raw bytes (an index found
//      there is at distance 0 and alone at that distance);
//   2. the misses (sequencing errors, N, random barcodes: ~1 read in 8) are queued in shared
//      memory and each is searched by a WHOLE warp: lane k packs base k, then the lanes split
//      the index table, XOR + popcount, min / count / first by warp reductions.
// The decision logic on top (match_barcode_nf) is the one the CPU emulation runs.
struct MatchShared {
  const uint8_t* bar[TILE_RECORDS];
  uint2 res[2][TILE_RECORDS];        // x = distance, y = count << 16 | first
  uint32_t queue[2 * TILE_RECORDS];  // thread | which << 16
  uint32_t nq;
};

struct StoredNearest {
  Nearest n7, n5;
  __device__ __forceinline__ Nearest get(int, int, int which, const uint8_t*) const { return which == 7 ? n7 : n5; }
};

__device__ __forceinline__ bool probe_exact(const Params& P, const uint8_t* p, int len, int off, int mask, Nearest& n) {
  // 16 bytes at any alignment: five aligned words, funnel-shifted
  const uint32_t sh = (uint32_t)((uintptr_t)p & 3u) * 8u;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p - ((uintptr_t)p & 3u));
  const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);
  uint4 v;
  v.x = __funnelshift_r(w0, w1, sh);
  v.y = __funnelshift_r(w1, w2, sh);
  v.z = __funnelshift_r(w2, w3, sh);
  v.w = __funnelshift_r(w3, w4, sh);
  uint64_t lo = (uint64_t)v.x | ((uint64_t)v.y << 32), hi = (uint64_t)v.z | ((uint64_t)v.w << 32);
  if (len < 8) {
    lo &= (1ull << (8 * len)) - 1ull;
    hi = 0;
  } else if (len < 16) {
    hi &= len == 8 ? 0ull : (1ull << (8 * (len - 8))) - 1ull;
  }
  uint32_t s = index_hash(lo, hi, (uint32_t)mask);
  for (;;) {
    const int id = __ldg(P.hids + off + s);
    if (id < 0) return false;
    if (__ldg(P.hkeys + 2 * (off + s)) == lo && __ldg(P.hkeys + 2 * (off + s) + 1) == hi) {
      n.d = 0;
      n.cnt = 1;
      n.first = id;
      return true;
    }
    s = (s + 1u) & (uint32_t)mask;
  }
}

// all threads of the CTA call this; bar == nullptr: nothing to match for this thread
__device__ __forceinline__ MatchOut cta_match(const Params& P, const uint64_t* s_codes, const uint8_t* bar, MatchShared& ms) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  MatchOut mo;
  mo.sample = P.total - 2;
  mo.mism = 0;
  mo.bar_t = mo.umi_t = -1;
  mo.panicked = 0;
  if (P.n_tpl != 1 || P.tpl[0].wide) {   // general path: every thread on its own
    if (bar) mo = match_barcode(P, s_codes, bar);
    return mo;
  }
  const TemplateDev& T = P.tpl[0];
  StoredNearest nf;
  nf.n7.d = nf.n5.d = INF_MISMATCH;
  nf.n7.cnt = nf.n5.cnt = 0;
  nf.n7.first = nf.n5.first = -1;
  if (tid == 0) ms.nq = 0;
  ms.bar[tid] = bar;
  __syncthreads();
  bool slow7 = false, slow5 = false;
  if (bar) {
    slow7 = T.n_i7 > 0 && !probe_exact(P, bar + P.BL - T.g[2], T.g[1], T.h7_off, T.h7_mask, nf.n7);
    slow5 = T.has_i5 && T.n_i5 > 0 && !probe_exact(P, bar + P.BL - T.g[5], T.g[4], T.h5_off, T.h5_mask, nf.n5);
    if (slow7) ms.queue[atomicAdd(&ms.nq, 1u)] = (uint32_t)tid;
    if (slow5) ms.queue[atomicAdd(&ms.nq, 1u)] = (uint32_t)tid | 0x10000u;
  }
  __syncthreads();
  const uint32_t nq = ms.nq;
  for (uint32_t q = (uint32_t)warp; q < nq; q += blockDim.x / 32) {
    const uint32_t item = ms.queue[q], who = item & 0xffffu, which = item >> 16;
    const int len = which ? T.g[4] : T.g[1], n = which ? T.n_i5 : T.n_i7;
    const uint8_t* p = ms.bar[who] + P.BL - (which ? T.g[5] : T.g[2]);
    const uint64_t* codes = s_codes + (which ? T.off_i5 : T.off_i7);
    const uint32_t b = lane < len ? __ldg(p + lane) : (uint32_t)'A';
    const bool acgt = (b == 'A') | (b == 'C') | (b == 'G') | (b == 'T'), isn = b == 'N';
    const bool invalid = __any_sync(0xffffffffu, !(acgt | isn));
    const uint32_t code = __reduce_or_sync(0xffffffffu, (acgt && lane < len) ? ((b >> 1) & 3u) << (2 * (lane & 15)) : 0u);
    const uint32_t nm = __reduce_or_sync(0xffffffffu, isn ? 1u << (2 * (lane & 15)) : 0u);
    int best = INF_MISMATCH, cnt = 0, first = 0x7fffffff;
    for (int k = lane; k < n; k += 32) {
      const uint32_t v = code ^ (uint32_t)codes[k];
      const int d = __popc(((v | (v >> 1)) & 0x55555555u) | nm);
      if (d < best) {
        best = d;
        cnt = 1;
        first = k;
      } else if (d == best) {
        ++cnt;
      }
    }
    const int dmin = __reduce_min_sync(0xffffffffu, best);
    const int ctot = __reduce_add_sync(0xffffffffu, best == dmin ? cnt : 0);
    const int fmin = __reduce_min_sync(0xffffffffu, best == dmin ? first : 0x7fffffff);
    if (lane == 0) ms.res[which][who] = invalid ? make_uint2((uint32_t)INF_MISMATCH, 0u) : make_uint2((uint32_t)dmin, ((uint32_t)ctot << 16) | (uint32_t)fmin);
  }
  __syncthreads();
  if (!bar) return mo;
  if (slow7) {
    const uint2 r = ms.res[0][tid];
    nf.n7.d = (int)r.x;
    nf.n7.cnt = (int)(r.y >> 16);
    nf.n7.first = r.x == (uint32_t)INF_MISMATCH ? -1 : (int)(r.y & 0xffffu);
  }
  if (slow5) {
    const uint2 r = ms.res[1][tid];
    nf.n5.d = (int)r.x;
    nf.n5.cnt = (int)(r.y >> 16);
    nf.n5.first = r.x == (uint32_t)INF_MISMATCH ? -1 : (int)(r.y & 0xffffu);
  }
  return match_barcode_nf(P, s_codes, bar, nf);
}

// dynamic shared memory: uint64 codes[n_codes] | uint32 bin_acc[2*total]
__global__ void __launch_bounds__(TILE_RECORDS, 5) k_match_plan(const Params* __restrict__ Pg, ChunkDev c, int mism_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Params P;
  __shared__ MatchShared ms;
  copy_params_to_smem(&P, Pg);
  __syncthreads();
  uint64_t* s_codes = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* bin_acc = reinterpret_cast<uint32_t*>(s_codes + P.n_codes);
  for (int i = threadIdx.x; i < P.n_codes; i += blockDim.x) s_codes[i] = P.codes[i];
  const int total = P.total, B2 = 2 * total;
  // per-CTA mismatch counters in shared memory when the table is small enough (mism_smem of the launch)
  uint32_t* mism_acc = mism_smem ? bin_acc + B2 : nullptr;
  if (mism_acc)
    for (int i = threadIdx.x; i < total * P.mism_w; i += blockDim.x) mism_acc[i] = 0;
  // a scan range that overflowed its staging run leaves holes in the newline table: nothing below may build
  // addresses from it (the error was raised by an earlier launch on this stream, so the read is uniform)
  if (c.status->error & ERR_NL_CAPACITY) return;
  const uint32_t raw2 = c.counts[0] >> 2, raw1 = P.paired ? c.counts[1] >> 2 : raw2;
  const uint32_t n_rec = device_n_records(c, P);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.status->n_records = n_rec;
    c.status->consumed2 = n_rec ? c.nl2[4u * n_rec - 1u] + 1u : 0u;
    c.status->consumed1 = (P.paired && n_rec) ? c.nl1[4u * n_rec - 1u] + 1u : 0u;
    if (min(raw2, raw1) > c.max_records) atomicOr(&c.status->error, ERR_REC_CAPACITY);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tr = device_tile_records(c, P, n_rec);
  const uint32_t n_tiles = (n_rec + tr - 1) / tr;

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < B2; i += blockDim.x) bin_acc[i] = 0;
    __syncthreads();
    const uint32_t rec = tile * tr + threadIdx.x;
    const bool active = threadIdx.x < tr && rec < n_rec;
    int bin = 0;
    uint32_t l2 = 0, l1 = 0;
    RecMeta mt;
    mt.off2 = mt.off1 = 0;
    mt.sample = 0;
    mt.len2 = mt.len1 = 0;
    mt.tpl = 0xff;
    mt.flags = 0;
    bool unassigned = false, is_amb = false;
    uint32_t bar_off = 0;
    RecGeom G;
    G.h2 = G.s2 = G.p2 = G.q2 = G.e2 = G.h1 = G.s1 = G.p1 = G.q1 = G.e1 = 0;
    const uint8_t* bar = nullptr;
    if (active) {
      G = load_geom(c, P, rec);
      if (G.p2 - G.s2 >= (uint32_t)P.BL + 1u) {   // the sequence line holds a whole barcode
        bar_off = G.p2 - 1u - (uint32_t)P.BL;
        bar = c.r2 + bar_off;
      }
    }
    MatchOut mo;
    if (P.reformat) {   // one sample, nothing to match (src/lib.rs:1114); the mismatches come with the record's plan
      mo.sample = 0;
      mo.mism = 0;
      mo.bar_t = mo.umi_t = -1;
      mo.panicked = 0;
    } else {
      mo = cta_match(P, s_codes, bar, ms);   // CTA-wide: every thread calls it
    }
    if (active) {
      if (mo.panicked) atomicOr(&c.status->error, ERR_PANIC);
      const RecPlan pl = record_plan(P, c.r2, G, mo.sample, mo.bar_t, mo.umi_t);
      if (P.reformat) mo.mism = (int)pl.rf_mism;
      if (pl.fmt_error) {
        atomicMin(&c.status->fmt_record, (unsigned long long)rec);
        mt.flags = 1;
      }
      bin = pl.matched ? P.writing[mo.sample] : mo.sample;
      l2 = pl.len2;
      l1 = pl.len1;
      mt.sample = (uint16_t)mo.sample;
      mt.len2 = (uint16_t)l2;
      mt.len1 = (uint16_t)l1;
      mt.tpl = (uint8_t)((mo.bar_t & 15) | ((mo.umi_t & 15) << 4));
      if (P.reformat) {
        mt.tpl = (uint8_t)pl.rf_tcl;
        mt.flags |= (uint8_t)(pl.rf_raw << 1);
      }
      if (l2 > 0xffffu || l1 > 0xffffu) atomicOr(&c.status->error, ERR_REC_TOO_LONG);
      if (!pl.fmt_error) {
        // update_mismatches(sample_id, curr_mismatch + 1, 1), src/lib.rs:1212
        if (mism_acc) atomicAdd(&mism_acc[mo.sample * P.mism_w + mo.mism + 1], 1u);   // flushed once per CTA
        else atomicAdd(&c.mism[(size_t)mo.sample * P.mism_w + mo.mism + 1], 1ull);
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
    uint32_t* row = c.tile_bins + (size_t)tile * c.bins_stride;
    for (int i = threadIdx.x; i < B2; i += blockDim.x) row[i] = bin_acc[i];
  }
  if (mism_acc) {   // a CTA sees fewer than 2^32 records
    __syncthreads();
    for (int i = threadIdx.x; i < total * P.mism_w; i += blockDim.x) {
      const uint32_t v = mism_acc[i];
      if (v) atomicAdd(&c.mism[i], (unsigned long long)v);
    }
  }
}

// ------------------------------------------------------------------ K3
// (a) bytes per (group of 64 tiles, bin)
__global__ void k_scan_groups(const Params* __restrict__ Pg, ChunkDev c) {
  if (c.status->error & ERR_NL_CAPACITY) return;   // no plan was made (k_match_plan)
  const int B2 = 2 * Pg->total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t tr = device_tile_records(c, *Pg, n_rec);
  const uint32_t n_tiles = (n_rec + tr - 1) / tr;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = id / B2, b = id % B2;
  if (g >= n_groups) return;
  const uint32_t t0 = g * GROUP_TILES, t1 = min(n_tiles, t0 + GROUP_TILES);
  uint32_t sum = 0;
#pragma unroll 8
  for (uint32_t t = t0; t < t1; ++t) sum += c.tile_bins[(size_t)t * c.bins_stride + b];
  c.group_sums[(size_t)g * B2 + b] = sum;
}

// (b) one CTA: exclusive scan over groups per bin, then over bins -> segment table
__global__ void __launch_bounds__(1024) k_scan_bins(const Params* __restrict__ Pg, ChunkDev c) {
  if (c.status->error & ERR_NL_CAPACITY) return;
  const int total = Pg->total, B2 = 2 * total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t tr = device_tile_records(c, *Pg, n_rec);
  const uint32_t n_tiles = (n_rec + tr - 1) / tr;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  unsigned long long* seg = c.seg;  // [off2 | len2 | off1 | len1] x total
  for (int b = threadIdx.x; b < B2; b += blockDim.x) {
    uint32_t run = 0;
    for (uint32_t g0 = 0; g0 < n_groups; g0 += 8) {   // loads of a batch first: reads and writes alias for the compiler
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = g0 + k < n_groups ? c.group_sums[(size_t)(g0 + k) * B2 + b] : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (g0 + k < n_groups) c.group_sums[(size_t)(g0 + k) * B2 + b] = run;
        run += v[k];
      }
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
  if (c.status->error & ERR_NL_CAPACITY) return;   // no plan was made (k_match_plan)
  const int total = Pg->total, B2 = 2 * total;
  const uint32_t n_rec = device_n_records(c, *Pg);
  const uint32_t tr = device_tile_records(c, *Pg, n_rec);
  const uint32_t n_tiles = (n_rec + tr - 1) / tr;
  const uint32_t n_groups = (n_tiles + GROUP_TILES - 1) / GROUP_TILES;
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = id / B2, b = id % B2;
  if (g >= n_groups) return;
  const uint32_t t0 = g * GROUP_TILES, t1 = min(n_tiles, t0 + GROUP_TILES);
  uint32_t run = c.group_sums[(size_t)g * B2 + b] + (uint32_t)c.seg[(b < (uint32_t)total ? 0 : 2) * total + (b % total)];
  for (uint32_t tb = t0; tb < t1; tb += 8) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = tb + k < t1 ? c.tile_bins[(size_t)(tb + k) * c.bins_stride + b] : 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (tb + k < t1) c.tile_bins[(size_t)(tb + k) * c.bins_stride + b] = run;
      run += v[k];
    }
  }
}

// ------------------------------------------------------------------ K4
// TMA in, scalar threads for the control-heavy crumbs, quarter-warps for the bytes.  Persistent,
// one CTA per SM.  A producer lane keeps two shared-memory stages full with bulk copies
// (cp.async.bulk + mbarrier) of the next 128 records: text of both reads, newline rows, RecMeta
// rows, bin-offset row.  On staged tile k, between two CTA-wide barriers:
//   * 16 warps, one thread per (record, role): role 0 / 1 build the front / the tail of the
//     rewritten header in a staging slot, role 2 / 3 the QC sums of the paired / barcode read
//     (and --validate): scalar work with data-dependent branches, a different record per lane;
//   * 8 of the other warps, one thread per (record, read): the descriptors of tile k + 1 (where every
//     byte range of the record comes from and goes to), into the table of the other parity;
//   * every warp, when that is done, draws rounds of four bodies from a shared cursor: a
//     quarter-warp moves "\n" sequence | "\n+\n" qualities of one read as one destination range in
//     16-byte units aligned on the destination, adjacent lanes adjacent units (move_pair_lane);
// then, after a barrier, all warps move the staged headers out the same way (move_small_lane; the
// paired read's copy with its own read-number digit patched in).  Every input byte crosses HBM->SM
// once (bulk copies), every output byte SM->HBM once, in coalesced stores.
constexpr int EMIT_NMAX = 128;               // records per staged tile (<= TILE_RECORDS, divides it)
constexpr int EMIT_SROLES = 4;
constexpr int EMIT_SWARPS = EMIT_SROLES * EMIT_NMAX / 32;
constexpr int EMIT_WWARPS = 11;
constexpr int EMIT_PRODUCER = EMIT_SWARPS + EMIT_WWARPS;
constexpr int EMIT_THREADS = (EMIT_PRODUCER + 1) * 32;
static_assert(EMIT_WWARPS >= 2 * (EMIT_NMAX / 32), "the descriptor warps (one thread per (record, read)) are worker warps without a role");
static_assert(EMIT_THREADS <= 1024 && EMIT_THREADS % 128 == 0, "one CTA per SM, whole warp-allocation granules");
constexpr uint32_t EMIT_FLUSH_TILES = 48;    // QC accumulators (u32) are flushed this often
// staging slot k (0 barcode read's header, 1 paired read's own full header) of record i; inside a warp's 32 records,
// rows of 8 are skewed by one word so that the 32 lanes of a warp, writing the same byte of their headers, hit 32 banks
// (the slot stride is a multiple of 4 words)
constexpr uint32_t HDR_WARP_BYTES = 32u * HDR_SLOT_BYTES + 64u;
constexpr uint32_t HDR_KIND_BYTES = (EMIT_NMAX / 32) * HDR_WARP_BYTES;
__host__ __device__ inline uint32_t hdr_slot(uint32_t k, uint32_t i) {
  return k * HDR_KIND_BYTES + (i >> 5) * HDR_WARP_BYTES + (i & 31u) * HDR_SLOT_BYTES + 4u * ((i >> 3) & 3u);
}

struct EmitLayout {       // byte offsets inside the dynamic shared memory, all 16-byte aligned
  uint32_t consts;        // prefix + literals (+ slack)
  uint32_t stats;         // [total][QC_SLOTS] u32
  uint32_t writing;       // [total] writing_samples (bin of a matched sample)
  uint32_t hdr;           // header staging slots, hdr_slot()
  uint32_t desc;          // [EMIT_NMAX][EMIT_PIECES] PieceDesc of the tile in flight
  uint32_t stage0, stage_stride;   // two stages, each: nl2 | nl1 | meta | bins | bytes
  uint32_t nl1, meta, bins, bytes;   // offsets inside a stage (nl2 at 0)
  uint32_t cap;           // bytes of text one stage holds (both reads)
  uint32_t total_bytes;
};

// What the movers work from, four 16-byte descriptors per record:
//   kind 0 / 1: the body of the barcode / paired read, "\n" sequence | "\n+\n" qualities [| "\n"], two runs of the
//               staged input that follow each other in the output ({dst, src0 | buffer << 31, src1 | newline << 31,
//               n0 | n1 << 16}; newline: a trimmed read's closing '\n' is stored on its own);
//   kind 2 / 3: the staged header of the barcode / paired read ({dst, src, n, aux}).
constexpr int EMIT_PIECES = 4;
struct __align__(16) PieceDesc {
  uint32_t dst;   // offset in the output buffer of its read
  uint32_t src;   // S offset
  uint32_t n;     // bytes (0: nothing)
  uint32_t aux;   // bit 31: paired-read buffer; bits 0-15: 1 + index of a byte to replace (0: none), bits 16-23: its value
};
static_assert(HDR_SLOT_BYTES <= 8 * 16, "a staged header is at most one unit per lane of a quarter-warp");

struct StageDesc {
  uint32_t n;             // records in the stage (0: no more tiles)
  uint32_t r0;            // first record
  uint32_t base2, base1;  // chunk offset of the first staged byte of either read (16-byte aligned)
  uint32_t off1;          // where the paired read starts inside the stage's byte area
  uint32_t pad[3];
};

__global__ void __launch_bounds__(EMIT_THREADS, 1) k_emit(const Params* __restrict__ Pg, ChunkDev c, EmitLayout lay) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Params P;
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ StageDesc s_desc[2];
  __shared__ uint32_t s_claim[2][2];   // cursors of a tile's two work lists (bodies, staged headers), by tile parity
  __shared__ uint32_t s_headers_built;  // header-builder warps that have finished their tile, counted over the whole launch
  __shared__ uint32_t s_abort;          // errors of the earlier kernels, read once for the whole CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  copy_params_to_smem(&P, Pg);
  if (tid == 0) {
    s_claim[0][0] = s_claim[0][1] = s_claim[1][0] = s_claim[1][1] = 0;
    s_headers_built = 0;
    s_abort = c.status->error & (ERR_OUT_CAPACITY | ERR_NL_CAPACITY | ERR_REC_TOO_LONG);
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = P.total, paired = P.paired;
  {  // constants, QC accumulators, bin of every sample
    const uint32_t nconst = (uint32_t)P.prefix_len + LIT_BYTES + (uint32_t)P.rf_bc_len;
    for (uint32_t i = tid; i < nconst; i += EMIT_THREADS) smem[lay.consts + i] = P.consts[i];
    uint32_t* st = reinterpret_cast<uint32_t*>(smem + lay.stats);
    for (int i = tid; i < total * QC_SLOTS; i += EMIT_THREADS) st[i] = 0;
    int32_t* wr = reinterpret_cast<int32_t*>(smem + lay.writing);
    for (int i = tid; i < total; i += EMIT_THREADS) wr[i] = P.writing[i];
  }
  const int32_t* s_writing = reinterpret_cast<const int32_t*>(smem + lay.writing);
  const uint32_t n_rec = device_n_records(c, P);
  if (s_abort) return;  // nothing safe to write (one read: another CTA may be raising ERR_REC_TOO_LONG right now)
  if (n_rec == 0) return;
  const uint32_t bins_stride = (2u * (uint32_t)total + 3u) & ~3u;

  // records per tile from the chunk's mean pair size (identical in every CTA; oversized tiles are shrunk below), and
  // the unit of ownership: EMIT_NMAX records when a whole tile of them fits a stage, else a plan tile cut into equal
  // parts (many samples or long records leave less room: three tiles of 86 are better than a tile of 117 and a sliver)
  const uint32_t plan_tile = device_tile_records(c, P, n_rec);
  uint32_t n_tile, unit;
  {
    const unsigned long long bytes = (unsigned long long)c.nl2[4u * n_rec - 1u] + (paired ? c.nl1[4u * n_rec - 1u] : 0u) + 2ull;
    const uint32_t mean = (uint32_t)(bytes / n_rec) + 1u;
    n_tile = max(1u, min((lay.cap - 256u) / mean, (uint32_t)EMIT_NMAX));
    unit = n_tile < plan_tile / 2u ? plan_tile : plan_tile / 2u;   // half a plan tile when it fits (device_tile_records sees to that)
    const uint32_t parts = (unit + n_tile - 1u) / n_tile;
    n_tile = (unit + parts - 1u) / parts;
  }
  // this CTA's units: blockIdx.x, + gridDim.x, ...  Interleaved: at any moment the CTAs of the grid work on
  // neighbouring records, so the pages they touch (input and every bin's output region) are few whatever the chunk size.
  if (blockIdx.x >= (n_rec + unit - 1u) / unit) return;

  // producer state (lane 0 of the producer warp)
  uint32_t next_rec = blockIdx.x * unit, start2 = 0, start1 = 0;
  bool fresh = true;   // next_rec opens a plan tile: its byte offsets come from the newline table
  auto produce = [&](int s) {   // lane 0 of the producer warp
    StageDesc& d = s_desc[s];
    if (next_rec >= n_rec) {   // no more tiles: an empty stage (the waiters read d.n after the barrier)
      d.n = 0;
      mbar_arrive(&s_bar[s]);
      return;
    }
    const uint32_t r0 = next_rec;
    if (fresh) {
      start2 = c.nl2[(long long)4 * r0 - 1] + 1u;   // c.nl2[-1] is a sentinel -1: record 0 starts at byte 0
      start1 = paired ? c.nl1[(long long)4 * r0 - 1] + 1u : 0u;
    }
    const uint32_t pt_end = min((r0 / unit + 1u) * unit, n_rec);
    uint32_t lim = min(r0 + n_tile, pt_end);
    const uint32_t base2 = start2 & ~15u, base1 = start1 & ~15u;
    uint32_t e2, e1, len2, len1, off1;
    for (;;) {
      e2 = c.nl2[4u * lim - 1u] + 1u;
      e1 = paired ? c.nl1[4u * lim - 1u] + 1u : 0u;
      len2 = ((e2 + 15u) & ~15u) - base2;
      len1 = paired ? ((e1 + 15u) & ~15u) - base1 : 0u;
      off1 = (len2 + 127u) & ~127u;
      if (off1 + len1 + 64u <= lay.cap || lim == r0 + 1u) break;
      lim = r0 + max(1u, (lim - r0) * 15u / 16u);   // rare: records much longer than the chunk's mean
    }
    if (off1 + len1 + 64u > lay.cap) {   // one read pair larger than a stage
      atomicOr(&c.status->error, ERR_REC_TOO_LONG);
      d.n = 0;
      next_rec = n_rec;
      mbar_arrive(&s_bar[s]);
      return;
    }
    const uint32_t n = lim - r0;
    d.n = n;
    d.r0 = r0;
    d.base2 = base2;
    d.base1 = base1;
    d.off1 = off1;
    unsigned char* st = smem + (lay.stage0 + (uint32_t)s * lay.stage_stride);
    const uint32_t nl_bytes = 16u * (n + 1u), meta_bytes = 16u * n, bins_bytes = 4u * bins_stride;
    mbar_expect_tx(&s_bar[s], len2 + len1 + (paired ? 2u : 1u) * nl_bytes + meta_bytes + bins_bytes);
    bulk_g2s(st + lay.bytes, c.r2 + base2, len2, &s_bar[s]);
    if (paired) bulk_g2s(st + lay.bytes + off1, c.r1 + base1, len1, &s_bar[s]);
    bulk_g2s(st, c.nl2 + ((long long)4 * r0 - 4), nl_bytes, &s_bar[s]);
    if (paired) bulk_g2s(st + lay.nl1, c.nl1 + ((long long)4 * r0 - 4), nl_bytes, &s_bar[s]);
    bulk_g2s(st + lay.meta, c.meta + r0, meta_bytes, &s_bar[s]);
    bulk_g2s(st + lay.bins, c.tile_bins + (size_t)(r0 / plan_tile) * bins_stride, bins_bytes, &s_bar[s]);
    fresh = lim == pt_end;
    next_rec = fresh ? (r0 / unit + gridDim.x) * unit : lim;
    start2 = e2;
    start1 = e1;
  };
  if (warp == EMIT_PRODUCER && lane == 0) produce(0);
  __syncthreads();

  uint32_t* s_stats = reinterpret_cast<uint32_t*>(smem + lay.stats);
  const uint32_t smem_s = smem_u32(smem), desc_s = smem_s + lay.desc;   // shared-window addresses
  auto flush_stats = [&]() {   // all threads; s_stats quiescent
    for (int i = tid; i < total * QC_SLOTS; i += EMIT_THREADS) {
      const uint32_t v = s_stats[i];
      if (v) {
        const int k = i % QC_SLOTS;
        atomicAdd(c.stats + (size_t)(i / QC_SLOTS) * MGK_N_STATS + qc_stat_column(k, paired), (unsigned long long)v);
        s_stats[i] = 0;
      }
    }
  };

  // One staged tile as the worker threads see it
  struct TileView {
    const uint32_t *nl2s, *nl1s, *bins;
    const RecMeta* metas;
    uint32_t o2, o1;   // S offsets of chunk byte 0 of either read
    uint32_t n, r0;
  };
  auto tile_view = [&](int s) {
    const StageDesc d = s_desc[s];
    const uint32_t stage = lay.stage0 + (uint32_t)s * lay.stage_stride;
    const unsigned char* st = smem + stage;
    TileView T;
    T.nl2s = reinterpret_cast<const uint32_t*>(st);
    T.nl1s = reinterpret_cast<const uint32_t*>(st + lay.nl1);
    T.metas = reinterpret_cast<const RecMeta*>(st + lay.meta);
    T.bins = reinterpret_cast<const uint32_t*>(st + lay.bins);
    T.o2 = stage + lay.bytes - d.base2;
    T.o1 = stage + lay.bytes + d.off1 - d.base1;
    T.n = d.n;
    T.r0 = d.r0;
    return T;
  };
  auto geom = [&](const TileView& T, uint32_t i) {
    RecGeom G;
    const uint4 a = *reinterpret_cast<const uint4*>(T.nl2s + 4u * i + 4u);
    G.h2 = T.nl2s[4u * i + 3u] + 1u + T.o2;
    G.s2 = a.x + 1u + T.o2;
    G.p2 = a.y + 1u + T.o2;
    G.q2 = a.z + 1u + T.o2;
    G.e2 = a.w + T.o2;
    G.h1 = G.s1 = G.p1 = G.q1 = G.e1 = 0;
    if (paired) {
      const uint4 b = *reinterpret_cast<const uint4*>(T.nl1s + 4u * i + 4u);
      G.h1 = T.nl1s[4u * i + 3u] + 1u + T.o1;
      G.s1 = b.x + 1u + T.o1;
      G.p1 = b.y + 1u + T.o1;
      G.q1 = b.z + 1u + T.o1;
      G.e1 = b.w + T.o1;
    }
    return G;
  };
  auto geom_one = [&](const TileView& T, uint32_t i, int which) {   // only the rows of one read (the other's fields 0)
    RecGeom G;
    const uint32_t* nls = which == 2 ? T.nl2s : T.nl1s;
    const uint32_t o = which == 2 ? T.o2 : T.o1;
    const uint4 a = *reinterpret_cast<const uint4*>(nls + 4u * i + 4u);
    const uint32_t h = nls[4u * i + 3u] + 1u + o, s1 = a.x + 1u + o, p1 = a.y + 1u + o, q1 = a.z + 1u + o, e1 = a.w + o;
    const bool br = which == 2;
    G.h2 = br ? h : 0u; G.s2 = br ? s1 : 0u; G.p2 = br ? p1 : 0u; G.q2 = br ? q1 : 0u; G.e2 = br ? e1 : 0u;
    G.h1 = br ? 0u : h; G.s1 = br ? 0u : s1; G.p1 = br ? 0u : p1; G.q1 = br ? 0u : q1; G.e1 = br ? 0u : e1;
    return G;
  };
  auto dest = [&](const TileView& T, const RecMeta& mt, int which) -> uint8_t* {   // where the record's output starts
    const int bin = mt.sample < P.S ? s_writing[mt.sample] : (int)mt.sample;
    return which == 2 ? c.out2 + T.bins[bin] + mt.off2 : c.out1 + T.bins[total + bin] + mt.off1;
  };

  // ---- descriptors of a staged tile, thread t of 2 * EMIT_NMAX per (record, read): the record's byte ranges (where
  // from, where to), so that the movers start after a single 16-byte load.  Two tables (tile parity): the table of
  // tile k + 1 is made while tile k is being moved.
  constexpr uint32_t DESC_TABLE = EMIT_PIECES * EMIT_NMAX * (uint32_t)sizeof(PieceDesc);
  auto make_descs = [&](const TileView& T, uint32_t par, uint32_t t) {
    PieceDesc* descs = reinterpret_cast<PieceDesc*>(smem + lay.desc + par * DESC_TABLE);
    const uint32_t i = t % EMIT_NMAX;
    const int which = t < EMIT_NMAX ? 2 : 1;
    if (i >= T.n) return;
    const RecMeta mt = T.metas[i];
    auto put = [&](int kind, uint32_t dst, uint32_t src, uint32_t n, uint32_t aux) {
      PieceDesc x;
      x.dst = dst;
      x.src = src;
      x.n = n;
      x.aux = aux;
      descs[kind * EMIT_NMAX + i] = x;
    };
    const int kb = which == 2 ? 0 : 1, kh = kb + 2;   // this read's body and staged header
    if ((mt.flags & 1) || (which == 1 && !paired)) {   // layout errors are reported through status, nothing is emitted
      put(kb, 0u, 0u, 0u, 0u);
      put(kh, 0u, 0u, 0u, 0u);
      return;
    }
    const RecGeom G = geom_one(T, i, which);
    const int bar_t = (P.reformat || (mt.tpl & 15) == 15) ? -1 : (mt.tpl & 15);
    const uint32_t rf = rf_pack(mt.tpl, (mt.flags >> 1) & 1u);   // read in reformat mode only
    const int bin = mt.sample < P.S ? s_writing[mt.sample] : (int)mt.sample;
    const uint32_t buf = which == 2 ? 0u : 1u << 31;
    const uint32_t dst = which == 2 ? T.bins[bin] + mt.off2 : T.bins[total + bin] + mt.off1;
    const ReadPlan r = plan_read(P, G, which, mt.sample, bar_t, which == 2 ? mt.len2 : mt.len1, rf);
    const Piece a = read_piece(r, 0), b = read_piece(r, 1);   // b follows a in the output
    put(kb, dst + a.dst_off, a.src | buf, b.src | ((r.trim && r.len) ? 1u << 31 : 0u), a.n | (b.n << 16));
    const bool staged = r.len && r.hdr != HDR_VERBATIM && r.hl <= HDR_SLOT_BYTES;
    const uint32_t patch = r.patch_at != 0xffffffffu ? (r.patch_at + 1u) | ((uint32_t)smem[r.patch_src] << 16) : 0u;
    put(kh, dst, lay.hdr + hdr_slot(staged ? (uint32_t)r.hdr - 1u : 0u, i), staged ? r.hl : 0u, buf | patch);
  };

  // A QUARTER-warp (8 lanes) moves one descriptor's bytes in 16-byte units aligned on the destination, adjacent
  // lanes adjacent units (the 8 lanes write 128 contiguous bytes per store), the bytes around the units one by one.
  auto move_body = [&](const uint4 pd, uint32_t l8) {
    const uint32_t n0 = pd.w & 0xffffu, N = n0 + (pd.w >> 16);
    const uint32_t sa = smem_s + (pd.y & 0x7fffffffu), sb = smem_s + (pd.z & 0x7fffffffu) - n0;
    uint8_t* g = ((int32_t)pd.y < 0 ? c.out1 : c.out2) + pd.x;
    const uint32_t hb = min((0u - pd.x) & 15u, N);   // the output buffers are 16-byte aligned
    const uint32_t mid = N - hb, nfull = mid >> 4, z = N - (mid & 15u) + l8;
    const uint32_t d0 = n0 - hb;
    const bool in0 = n0 > hb;
    const uint32_t ubf = in0 ? d0 >> 4 : 0u;
    const uint32_t ubs = (in0 && (d0 & 15u) && ubf < nfull) ? ubf : 0xffffffffu;
    move_pair_lane(g, l8, hb, nfull, N, n0, sa, sb, ubf, ubs, z, pd.z >> 31);
    if (nfull > 24u) {   // ranges longer than 24 units: the rest, plainly
      for (uint32_t u = l8 + 24u; u < nfull; u += 8u) {
        if (u == ubs) continue;
        const uint32_t x = hb + (u << 4);
        *reinterpret_cast<uint4*>(g + x) = ld16u(smem, (u < ubf ? sa : sb) + x - smem_s);
      }
    }
  };
  auto move_header = [&](const uint4 pd, uint32_t l8) {
    const uint32_t n = pd.z, aux = pd.w;
    uint8_t* g = ((int32_t)aux < 0 ? c.out1 : c.out2) + pd.x;
    const uint32_t hb = min((0u - pd.x) & 15u, n);
    const uint32_t mid = n - hb, nfull = mid >> 4, tail0 = n - (mid & 15u), z = tail0 + l8;
    move_small_lane(g, l8, hb, nfull, n, smem_s + pd.y, z);
    // the paired read's copy: its own read-number digit, stored by the lane that has just stored the bytes
    // around it (same thread, program order)
    const uint32_t pa = (aux & 0xffffu) - 1u, pv = (aux >> 16) & 0xffu;   // pa == 0xffffffff: nothing to replace
    const uint32_t owner = pa < hb ? pa : pa >= tail0 ? pa - tail0 : (pa - hb) >> 4;
    if (pa < n && (owner & 7u) == l8) g[pa] = (uint8_t)pv;
  };
  const bool desc_warp = warp >= EMIT_SWARPS && warp < EMIT_SWARPS + 2 * (EMIT_NMAX / 32);
  const uint32_t desc_t = (uint32_t)(tid - EMIT_SWARPS * 32);

  // prologue: descriptors of this CTA's first tile
  mbar_wait(&s_bar[0], 0u);
  if (desc_warp && s_desc[0].n) make_descs(tile_view(0), 0u, desc_t);
  __syncthreads();

  unsigned long long t_wait = 0, t_s1 = 0, t_role = 0, t_mv = 0, t_b2 = 0, t_s3 = 0;   // MGK_PROFILE=1
  for (uint32_t k = 0;; ++k) {
    const int s = (int)(k & 1u);
    if (warp == EMIT_PRODUCER && lane == 0) produce(s ^ 1);
    const long long c1 = tick(c.prof);
    mbar_wait(&s_bar[s], (k >> 1) & 1u);   // every thread: the stage's bytes and its s_desc entry are visible after it
    if (s_desc[s].n == 0) break;
    const long long c2 = tick(c.prof);
    t_wait += (unsigned long long)(c2 - c1);
    const TileView T = tile_view(s);
    const uint32_t table_s = desc_s + (k & 1u) * DESC_TABLE;
    const long long c4 = c2;

    // ======== scalar roles; then the bodies straight out of the staged input, the next tile's descriptors, the headers
    if (warp < EMIT_SWARPS) {
      // one thread per (record, role)
      const int role = warp / (EMIT_NMAX / 32);
      const uint32_t i = (uint32_t)tid % EMIT_NMAX;
      if (i < T.n) {
        const RecMeta mt = T.metas[i];
        const int bar_t = (P.reformat || (mt.tpl & 15) == 15) ? -1 : (mt.tpl & 15), umi_t = (P.reformat || (mt.tpl >> 4) == 15) ? -1 : (mt.tpl >> 4);
        const uint32_t rf = rf_pack(mt.tpl, (mt.flags >> 1) & 1u);
        if (role < 2) {
          if (!(mt.flags & 1) && mt.sample < P.S && P.out_mode != MGK_OUT_MGI) {   // rewritten headers
            const RecGeom G = geom(T, i);
            const ReadPlan p2 = plan_read(P, G, 2, mt.sample, bar_t, mt.len2, rf);
            ReadPlan p1 = p2;
            p1.len = 0;
            if (paired) p1 = plan_read(P, G, 1, mt.sample, bar_t, mt.len1, rf);
#pragma unroll 1
            for (int which = 2; which >= 1; --which) {
              const ReadPlan& r = which == 2 ? p2 : p1;
              const int slot = which == 2 ? HDR_SLOT0 : HDR_SLOT1;
              const bool own = r.hdr == slot && r.len != 0, lent = which == 2 && p1.hdr == HDR_SLOT0 && p1.len != 0;
              if (!own && !lent) continue;
              const uint32_t hl = own ? r.hl : p1.hl;   // illumina: both reads carry the same header
              if (hl <= HDR_SLOT_BYTES) {
                build_header(P, smem, lay.consts, G, which, bar_t, umi_t, hl, smem + lay.hdr + hdr_slot((uint32_t)slot - 1u, i), role, rf);
              } else {   // too long for a slot: byte by byte at its destination(s)
#pragma unroll 1
                for (int to = 2; to >= 1; --to) {
                  const ReadPlan& t = to == 2 ? p2 : p1;
                  if (t.hdr != slot || t.len == 0) continue;
                  uint8_t* g = dest(T, mt, to);
                  build_header(P, smem, lay.consts, G, which, bar_t, umi_t, hl, g, role, rf);
                  if (role == 1 && t.patch_at != 0xffffffffu) g[t.patch_at] = smem[t.patch_src];
                }
              }
            }
          }
        } else if (!(mt.flags & 1)) {
          const int val_role = paired ? 2 : 3;   // --validate goes to the QC thread every record has
          const bool qc = P.report_level > 0 && (role == 3 || paired);
          if (qc || (P.validate && role == val_role)) {
            const RecGeom G = geom(T, i);
            if (qc) {   // sum_qc x3 + update_stats x6, src/lib.rs:1188-1210
              const QcSix q = qc_pair(P, smem, G, role == 2 ? 1 : 2);
#pragma unroll
              for (int j = 0; j < QC_SLOTS; ++j)
                if (q.v[j]) atomicAdd(&s_stats[mt.sample * QC_SLOTS + j], q.v[j]);
            }
            if (P.validate && role == val_role) {
              const int v = validate_pair(P, smem, G);
              if (v) atomicMin(&c.status->validate_key, ((unsigned long long)(T.r0 + i) << 4) | (unsigned)v);
            }
          }
        }
      }
    }
    if (warp < 2 * (EMIT_NMAX / 32)) {   // a header-builder warp is through with the tile: its slots may be read
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(&s_headers_built, 1u);
      }
    }
    const long long c5 = tick(c.prof);
    t_role += (unsigned long long)(c5 - c4);
    // Every warp then draws rounds of four descriptors from the tile's cursors until both lists are exhausted (the
    // next draw is in flight while a round is moved), so a warp joins whenever its other work is done: first the
    // bodies; then, once all eight header-builder warps have signalled (they have, long before, except in a tile's
    // last instants), the staged headers.  Eight of the warps without a role make the next tile's descriptors after
    // their first three rounds (its stage has landed by then).  No CTA-wide barrier between bodies and headers.
    if (warp < EMIT_PRODUCER) {
      uint32_t* cursors = s_claim[k & 1u];
      if (tid == 0) s_claim[(k + 1u) & 1u][0] = s_claim[(k + 1u) & 1u][1] = 0;   // untouched during this tile
      auto claim = [&](uint32_t* cursor, uint32_t first, uint32_t max_rounds, bool headers) {
        const uint32_t l8 = (uint32_t)lane & 7u, q = (uint32_t)lane >> 3;
        uint32_t raw = 0;
        if (lane == 0) raw = atomicAdd(cursor, 4u);
        uint32_t cur = __shfl_sync(0xffffffffu, raw, 0);
        MGK_LOOP_ROLLED
        for (uint32_t r = 1; cur < 2u * EMIT_NMAX; ++r) {
          const bool more = r < max_rounds;
          if (more && lane == 0) raw = atomicAdd(cursor, 4u);
          const uint32_t it = cur + q;
          if (it % EMIT_NMAX < T.n) {
            const uint4 pd = lds_v4(table_s + (first + it) * 16u);
            if (headers) {
              if (pd.z) move_header(pd, l8);
            } else if (pd.w) {
              move_body(pd, l8);
            }
          }
          if (!more) break;
          cur = __shfl_sync(0xffffffffu, raw, 0);
        }
      };
      if (desc_warp) {
        claim(&cursors[0], 0u, 3u, false);
        const long long c8 = tick(c.prof);
        mbar_wait(&s_bar[s ^ 1], ((k + 1u) >> 1) & 1u);
        if (s_desc[s ^ 1].n) make_descs(tile_view(s ^ 1), (k + 1u) & 1u, desc_t);
        __syncwarp();
        t_s1 += (unsigned long long)(tick(c.prof) - c8);
      }
      claim(&cursors[0], 0u, 0xffffffffu, false);
      const long long c6 = tick(c.prof);
      t_mv += (unsigned long long)(c6 - c5);
      if (lane == 0) {
        const uint32_t want = (k + 1u) * 2u * (EMIT_NMAX / 32);
        while (*reinterpret_cast<volatile uint32_t*>(&s_headers_built) < want) __nanosleep(40);   // do not take the LSU from the working warps
        __threadfence_block();
      }
      __syncwarp();
      const long long c7 = tick(c.prof);
      t_b2 += (unsigned long long)(c7 - c6);
      claim(&cursors[1], 2u * EMIT_NMAX, 0xffffffffu, true);
      t_s3 += (unsigned long long)(tick(c.prof) - c7);
    }
    __syncthreads();   // stage s and the header slots are free again; the next tile's descriptors are published
    if (P.report_level > 0 && (k % EMIT_FLUSH_TILES) == EMIT_FLUSH_TILES - 1u) {
      flush_stats();
      __syncthreads();
    }
  }
  __syncthreads();
  if (P.report_level > 0) flush_stats();
  if (c.prof && lane == 0) {
    unsigned long long* o = c.prof + ((size_t)blockIdx.x * (EMIT_THREADS / 32) + warp) * 6;
    o[0] = t_wait;
    o[1] = t_s1;
    o[2] = t_role;
    o[3] = t_mv;
    o[4] = t_b2;
    o[5] = t_s3;
  }
}

// ------------------------------------------------------------------ stage kernels (parity tests)
__global__ void __launch_bounds__(TILE_RECORDS) k_match_only(const Params* __restrict__ Pg, const uint8_t* __restrict__ bars, uint32_t n,
                                                             int32_t* sample, int32_t* mism, uint32_t* error) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Params P;
  __shared__ MatchShared ms;
  copy_params_to_smem(&P, Pg);
  __syncthreads();
  uint64_t* s_codes = reinterpret_cast<uint64_t*>(smem_raw);
  for (int i = threadIdx.x; i < P.n_codes; i += blockDim.x) s_codes[i] = P.codes[i];
  __syncthreads();
  const uint32_t rounds = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (uint32_t r = 0; r < rounds; ++r) {   // uniform trip count: cta_match is CTA-wide
    const uint32_t i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const MatchOut mo = cta_match(P, s_codes, i < n ? bars + (size_t)i * P.BL : nullptr, ms);
    if (i < n) {
      if (mo.panicked) atomicOr(error, ERR_PANIC);
      sample[i] = mo.sample;
      mism[i] = mo.mism;
    }
    __syncthreads();
  }
}

}  // namespace mgk

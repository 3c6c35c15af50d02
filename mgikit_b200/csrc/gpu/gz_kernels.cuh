// gzip on the device: every per-sample output segment of a chunk leaves the GPU as a run of gzip
// members instead of text, so the host only appends (replaces compress_buffer / libdeflate, /root/reference/
// src/sample_data.rs:180-207, :472-490; byte-exact .gz is not part of the contract -- the reference's own tests
// compare decompressed text, tests/integration_test.rs:44-53 -- the decompressed bytes are).
//
// One member per GZ_BLOCK = 32 KiB of text, BGZF-framed (extra field "BC", so bgzip-aware tools can index the
// files), holding ONE dynamic-Huffman deflate block.  Per 256-thread CTA: the text is staged in shared memory, every
// thread owns a run of 128 bytes; a first-occurrence table over the 12-byte strings of the block (position | tag per
// slot, filled with atomicMin: the result does not depend on timing, so the member's bytes are the same on every
// run) gives each run its repeats of 12 bytes or more (the rewritten headers of consecutive records of a sample, runs
// of equal qualities; random bases repeat shorter strings all the time and are cheaper as literals); histograms ->
// two Huffman codes (two-queue construction, lengths capped at 15 like zlib's gen_bitlen) -> bit offsets by a
// block-wide scan -> every thread ORs the codes of its run into the shared-memory image of the member.  CRC-32 per
// thread, combined with carry-less multiplications by x^(8*128*2^j).  A block that would not shrink is stored.
//
//   k_gz_plan    blocks per (read, bin) run, their prefix
//   k_gz_blocks  persistent, one 32 KiB block per CTA at a time -> a member in its worst-case sized slot
//   k_gz_scan    prefix of the member sizes -> packed offsets, per-bin (offset, length) of the members
//   k_gz_pack    members from their slots to their packed places (over the text, which is dead by then)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mgk {

constexpr uint32_t GZ_BLOCK = 32768;                     // text bytes per member
constexpr uint32_t GZ_THREADS = 256;
constexpr uint32_t GZ_RUN = GZ_BLOCK / GZ_THREADS;       // text bytes per thread
constexpr uint32_t GZ_ROW_WORDS = GZ_RUN / 4 + 1;        // rows of the staged text are skewed by a word: thread t, word w -> bank (t + w) % 32
constexpr uint32_t GZ_HDR = 18, GZ_TRL = 8;
constexpr uint32_t GZ_SLOT = GZ_BLOCK + 64;              // worst case (stored): 18 + 5 + 32768 + 8 bytes
constexpr uint32_t GZ_OUT_WORDS = GZ_SLOT / 4;
constexpr uint32_t GZ_POLY = 0xedb88320u;

struct GzConsts {
  uint32_t xshift[6];   // x^(8 * GZ_RUN * 2^j) mod P (reflected), j = 0..5: CRC of a left half moved over its right half
};

struct GzDev {
  const unsigned long long* seg;    // [4][total] text: off2, len2, off1, len1
  uint8_t* text2;
  uint8_t* text1;
  uint8_t* slots;                   // [blocks_cap][GZ_SLOT]
  uint32_t* blk_base;               // [2 * total + 1] first block of every (read, bin) run
  uint32_t* blk_size;               // [blocks_cap] bytes of the member
  unsigned long long* blk_off;      // [blocks_cap + 1] where the member goes in its read's packed buffer
  unsigned long long* gzseg;        // [4][total] packed members: off2, len2, off1, len1
  uint32_t total, blocks_cap;
  unsigned long long cap2, cap1;    // bytes the text buffers hold
  uint32_t* error;                  // Status::error
  uint32_t err_capacity;            // the bit to raise when something does not fit
};

// a * b mod P, reflected representation (zlib's multmodp)
__host__ __device__ inline uint32_t gz_multmodp(uint32_t a, uint32_t b) {
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ GZ_POLY : b >> 1;
  }
  return p;
}
inline GzConsts gz_make_consts() {
  GzConsts c;
  // x^(2^k): x2n[0] = x^1
  uint32_t x2n[32];
  uint32_t p = 1u << 30;
  x2n[0] = p;
  for (int n = 1; n < 32; ++n) x2n[n] = p = gz_multmodp(p, p);
  for (int j = 0; j < 6; ++j) {   // x^(8 * GZ_RUN * 2^j) = x^(GZ_RUN * 2^j * 2^3)
    uint64_t n = (uint64_t)GZ_RUN << j;
    uint32_t k = 3, r = 1u << 31;
    while (n) {
      if (n & 1) r = gz_multmodp(x2n[k & 31], r);
      n >>= 1;
      ++k;
    }
    c.xshift[j] = r;
  }
  return c;
}

// ---------------------------------------------------------------- plan
__global__ void __launch_bounds__(1024) k_gz_plan(GzDev g) {
  __shared__ uint32_t s_part[1024];
  const uint32_t n = 2u * g.total, per = (n + 1023u) / 1024u;
  const uint32_t i0 = threadIdx.x * per, i1 = min(n, i0 + per);
  auto blocks_of = [&](uint32_t q) {
    const uint32_t r = q / g.total, b = q % g.total;
    const unsigned long long len = g.seg[(r ? 3u : 1u) * g.total + b];
    return (uint32_t)((len + GZ_BLOCK - 1u) / GZ_BLOCK);
  };
  uint32_t sum = 0;
  for (uint32_t q = i0; q < i1; ++q) sum += blocks_of(q);
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {   // inclusive scan of the 1024 partial sums
    const uint32_t v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = s_part[threadIdx.x] - sum;
  for (uint32_t q = i0; q < i1; ++q) {
    g.blk_base[q] = run;
    run += blocks_of(q);
  }
  if (threadIdx.x == 1023) {
    g.blk_base[n] = s_part[1023];
    if (s_part[1023] > g.blocks_cap) atomicOr(g.error, g.err_capacity);
  }
}

// ---------------------------------------------------------------- blocks
constexpr uint32_t GZ_MAX_TOK = 16;        // matches per thread run (at least GZ_MIN_MATCH bytes each)
constexpr uint32_t GZ_MIN_MATCH = 12;      // shorter repeats do not pay for their distance bits (and random bases repeat 8-mers all the time)
constexpr uint32_t GZ_HASH_BITS = 12;
constexpr uint32_t GZ_NLIT = 286, GZ_NDIST = 30;

struct GzShared {
  uint32_t text[GZ_THREADS * GZ_ROW_WORDS + 4];   // the block, right-aligned: text byte i at position (GZ_BLOCK - n) + i
  // the member being built; before that: first-occurrence table [4096] | literal/length histograms [8][286] | distance
  // histograms [8][30], and (over the dead table) symbol counts, sort order and the Huffman scratch
  uint32_t out[GZ_OUT_WORDS];
  uint32_t tok[GZ_THREADS][GZ_MAX_TOK];           // a run's matches: position in the run | length << 7 | distance << 16
  uint32_t ntok[GZ_THREADS];
  uint32_t code[GZ_NLIT];                         // reversed code | length << 16
  uint32_t dcode[GZ_NDIST];
  uint32_t crc_tab[256];
  uint32_t warp_bits[8], warp_crc[8];
  uint32_t q, stored, total_bits, hlit, hdist;
};
constexpr uint32_t GZ_O_TABLE = 0, GZ_O_HLIT = 1u << GZ_HASH_BITS, GZ_O_HDIST = GZ_O_HLIT + 8u * GZ_NLIT;
static_assert(GZ_O_HDIST + 8u * GZ_NDIST <= GZ_OUT_WORDS, "scratch fits the image area");

// word g of the block's text (rows of 32 words, skewed by one)
__device__ __forceinline__ uint32_t gz_word(const GzShared& S, uint32_t g) { return S.text[g + (g >> 5)]; }
__device__ __forceinline__ uint32_t gz_text_byte(const GzShared& S, uint32_t pos) {   // pos in right-aligned coordinates
  return (gz_word(S, pos >> 2) >> (8u * (pos & 3u))) & 0xffu;
}
// the 4 bytes at pos (any alignment; words behind the block are zero padding)
__device__ __forceinline__ uint32_t gz_load4(const GzShared& S, uint32_t pos) {
  const uint32_t g = pos >> 2, sh = (pos & 3u) * 8u;
  const uint32_t a = gz_word(S, min(g, GZ_BLOCK / 4u));
  return sh ? __funnelshift_r(a, gz_word(S, min(g + 1u, GZ_BLOCK / 4u)), sh) : a;
}
// slot and tag of a 12-byte string: the table keeps, per slot, position << 16 | tag of the FIRST string that hashes there
__device__ __forceinline__ uint32_t gz_hash(uint32_t w0, uint32_t w1, uint32_t w2) {
  return (w0 * 0x9E3779B1u) ^ (w1 * 0x85EBCA77u) ^ (w2 * 0xC2B2AE3Du);
}
__device__ __forceinline__ uint32_t gz_slot(uint32_t h) { return h >> (32u - GZ_HASH_BITS); }
__device__ __forceinline__ uint32_t gz_tag(uint32_t h) { return (h >> 3) & 0xffffu; }
// deflate's length / distance alphabets (RFC 1951 3.2.5)
__device__ __forceinline__ void gz_len_symbol(uint32_t len, uint32_t& sym, uint32_t& eb, uint32_t& ev) {
  if (len == 258u) {
    sym = 285u;
    eb = ev = 0u;
    return;
  }
  const uint32_t l = len - 3u;
  if (l < 8u) {
    sym = 257u + l;
    eb = ev = 0u;
    return;
  }
  const uint32_t k = 31u - (uint32_t)__clz((int)l);
  eb = k - 2u;
  sym = 261u + 4u * eb + ((l >> eb) & 3u);
  ev = l & ((1u << eb) - 1u);
}
__device__ __forceinline__ void gz_dist_symbol(uint32_t dist, uint32_t& sym, uint32_t& eb, uint32_t& ev) {
  const uint32_t d = dist - 1u;
  if (d < 4u) {
    sym = d;
    eb = ev = 0u;
    return;
  }
  const uint32_t k = 31u - (uint32_t)__clz((int)d);
  eb = k - 1u;
  sym = 2u * k + ((d >> eb) & 1u);
  ev = d & ((1u << eb) - 1u);
}
__device__ __forceinline__ uint32_t gz_len_extra(uint32_t sym) { return (sym < 265u || sym == 285u) ? 0u : (sym - 261u) >> 2; }
__device__ __forceinline__ uint32_t gz_dist_extra(uint32_t sym) { return sym < 4u ? 0u : (sym >> 1) - 1u; }

// One thread: Huffman code lengths (<= 15) and canonical, bit-reversed codes for the symbols with cnt > 0.
// sorted[0 .. m) = the used symbols by (count, symbol) ascending.  Returns sum of count * length.
__device__ uint32_t gz_build_code(const uint32_t* cnt, uint32_t n_sym, const uint32_t* sorted, uint32_t m, uint32_t* scratch, uint32_t* code) {
  uint32_t bl_count[17];
  for (int i = 0; i < 17; ++i) bl_count[i] = 0;
  int overflow = 0;
  if (m == 1) {
    bl_count[1] = 1;
  } else {
    // two-queue Huffman over the leaves (ascending): nodes 0..m-1 leaves, m..2m-2 internal
    uint32_t* iw = scratch;                 // [m - 1] weights of the internal nodes, in creation order (ascending)
    uint32_t* parent = scratch + 290u;      // [2m - 1]
    uint32_t* depth = parent + 580u;        // [2m - 1]
    uint32_t li = 0, ii = 0, ni = 0;
    for (uint32_t step = 0; step + 1 < m; ++step) {
      uint32_t pick[2], wsum = 0;
      for (int t = 0; t < 2; ++t) {
        const bool leaf = li < m && (ii >= ni || cnt[sorted[li]] <= iw[ii]);
        if (leaf) {
          pick[t] = li;
          wsum += cnt[sorted[li]];
          ++li;
        } else {
          pick[t] = m + ii;
          wsum += iw[ii];
          ++ii;
        }
      }
      iw[ni] = wsum;
      parent[pick[0]] = parent[pick[1]] = m + ni;
      ++ni;
    }
    depth[2u * m - 2u] = 0;
    for (int node = (int)(2u * m - 3u); node >= 0; --node) {
      uint32_t d = depth[parent[node]] + 1u;
      depth[node] = d;
      if ((uint32_t)node < m) {
        if (d > 15u) {
          d = 15u;
          ++overflow;
        }
        ++bl_count[d];
      }
    }
  }
  while (overflow > 0) {   // zlib's gen_bitlen: move a leaf down one level, two overflowing leaves up
    uint32_t bits = 14;
    while (bl_count[bits] == 0) --bits;
    --bl_count[bits];
    bl_count[bits + 1] += 2;
    --bl_count[15];
    overflow -= 2;
  }
  uint32_t total = 0;
  {   // the rarest symbols take the longest codes
    uint32_t i = 0;
    for (uint32_t bits = 15; bits >= 1; --bits)
      for (uint32_t c = 0; c < bl_count[bits]; ++c) {
        const uint32_t s = sorted[i++];
        code[s] = bits << 16;
        total += bits * cnt[s];
      }
  }
  // canonical codes, bit-reversed (deflate packs Huffman codes starting with their most significant bit)
  uint32_t next_code[16], c0 = 0;
  bl_count[0] = 0;
  for (uint32_t bits = 1; bits <= 15; ++bits) {
    c0 = (c0 + bl_count[bits - 1]) << 1;
    next_code[bits] = c0;
  }
  for (uint32_t s = 0; s < n_sym; ++s) {
    const uint32_t len = code[s] >> 16;
    if (!len) continue;
    const uint32_t c = next_code[len]++;
    code[s] = (__brev(c) >> (32u - len)) | (len << 16);
  }
  return total;
}

__global__ void __launch_bounds__(GZ_THREADS, 2) k_gz_blocks(GzDev g, GzConsts K) {
  extern __shared__ __align__(16) unsigned char gz_smem[];
  GzShared& S = *reinterpret_cast<GzShared*>(gz_smem);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  {   // CRC table
    uint32_t c = tid;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (GZ_POLY & (0u - (c & 1u)));
    S.crc_tab[tid] = c;
  }
  if (*g.error & g.err_capacity) return;   // the plan does not fit (uniform: raised by an earlier launch)
  const uint32_t n_runs = 2u * g.total, n_blocks = g.blk_base[n_runs];
  for (uint32_t id = blockIdx.x; id < n_blocks; id += gridDim.x) {
    __syncthreads();
    if (tid == 0) {   // the run this block belongs to: last q with blk_base[q] <= id
      uint32_t lo = 0, hi = n_runs;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (g.blk_base[mid] <= id) lo = mid;
        else hi = mid;
      }
      S.q = lo;
    }
    __syncthreads();
    const uint32_t q = S.q, r = q / g.total, b = q % g.total, k = id - g.blk_base[q];
    const unsigned long long seg_off = g.seg[(r ? 2u : 0u) * g.total + b], seg_len = g.seg[(r ? 3u : 1u) * g.total + b];
    const uint32_t n = (uint32_t)min((unsigned long long)GZ_BLOCK, seg_len - (unsigned long long)k * GZ_BLOCK);
    const uint8_t* src = (r ? g.text1 : g.text2) + seg_off + (unsigned long long)k * GZ_BLOCK;
    const uint32_t pad = GZ_BLOCK - n;
    // ---- stage the text: smem word j holds positions 4j .. 4j+3; aligned global words, funnel-shifted
    for (uint32_t j = tid; j < GZ_BLOCK / 4; j += GZ_THREADS) {
      const int i0 = (int)(4u * j) - (int)pad;   // text index of the word's first byte
      uint32_t w = 0;
      if (i0 >= 0 && (uint32_t)i0 + 4u <= n) {
        const uintptr_t a = (uintptr_t)(src + i0);
        const uint32_t* al = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 3u) * 8u;
        w = sh ? __funnelshift_r(al[0], al[1], sh) : al[0];
      } else if (i0 + 3 >= 0 && i0 < (int)n) {   // the word the text starts in (or a text shorter than a word)
        for (int bi = 0; bi < 4; ++bi)
          if (i0 + bi >= 0 && i0 + bi < (int)n) w |= (uint32_t)src[i0 + bi] << (8 * bi);
      }
      S.text[(j >> 5) * GZ_ROW_WORDS + (j & 31u)] = w;
    }
    for (uint32_t i = tid; i < GZ_O_HLIT; i += GZ_THREADS) S.out[i] = 0xffffffffu;                               // first-occurrence table
    for (uint32_t i = GZ_O_HLIT + tid; i < GZ_O_HDIST + 8u * GZ_NDIST; i += GZ_THREADS) S.out[i] = 0;            // histograms, one per warp
    if (tid < 4) S.text[GZ_THREADS * GZ_ROW_WORDS + tid] = 0;
    __syncthreads();
    const uint32_t p0 = tid * GZ_RUN, run_lo = max(p0, pad), run_hi = p0 + GZ_RUN;   // run_lo >= run_hi: nothing of the text here
    // ---- CRC of this thread's run; first occurrence of every 8-byte string that starts in it.  The run's words pass
    // through three registers (the 8 bytes at a position span at most three words)
    uint32_t crc = 0;
    if (run_lo < run_hi) {
      uint32_t c = 0xffffffffu;
      const uint32_t g0 = tid * (GZ_RUN / 4u);
      uint32_t a = gz_word(S, g0), b = gz_word(S, g0 + 1u), d = gz_word(S, g0 + 2u), e = gz_word(S, min(g0 + 3u, GZ_BLOCK / 4u));
      for (uint32_t i = 0; i < GZ_RUN / 4u; ++i) {
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
          const uint32_t p = p0 + 4u * i + k;
          if (p >= run_lo) {
            c = S.crc_tab[(c ^ (a >> (8u * k))) & 0xffu] ^ (c >> 8);
            if (p + GZ_MIN_MATCH <= GZ_BLOCK) {
              const uint32_t h = k ? gz_hash(__funnelshift_r(a, b, 8u * k), __funnelshift_r(b, d, 8u * k), __funnelshift_r(d, e, 8u * k)) : gz_hash(a, b, d);
              atomicMin(&S.out[GZ_O_TABLE + gz_slot(h)], (p << 16) | gz_tag(h));
            }
          }
        }
        a = b;
        b = d;
        d = e;
        e = gz_word(S, min(g0 + i + 4u, GZ_BLOCK / 4u));   // (the words behind the block are zero padding)
      }
      crc = ~c;
    }
    // CRC of the block: tree over the threads (a left part is moved over its right part by a multiplication)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const uint32_t left = __shfl_up_sync(0xffffffffu, crc, 1u << j);
      if ((lane & ((2u << j) - 1u)) == (2u << j) - 1u) crc = gz_multmodp(K.xshift[j], left) ^ crc;
    }
    if (lane == 31) S.warp_crc[warp] = crc;
    __syncthreads();
    // ---- greedy parse of the run: a repeat of at least GZ_MIN_MATCH bytes of anything earlier in the block (its first
    // occurrence: the table does not depend on timing, so the member's bytes are the same on every run), ending
    // inside the run; histograms of what it emits
    {
      uint32_t* hl = S.out + GZ_O_HLIT + warp * GZ_NLIT;
      uint32_t* hd = S.out + GZ_O_HDIST + warp * GZ_NDIST;
      uint32_t nt = 0, p = run_lo;
      uint32_t cg = 0xfffffff0u, a = 0, b = 0, d = 0, e = 0;   // words cg .. cg + 3 of the text
      while (p < run_hi) {
        const uint32_t g = p >> 2, k8 = (p & 3u) * 8u;
        if (g != cg) {
          if (g == cg + 1u) {
            a = b;
            b = d;
            d = e;
          } else {
            a = gz_word(S, g);
            b = gz_word(S, g + 1u);
            d = gz_word(S, min(g + 2u, GZ_BLOCK / 4u));
          }
          e = gz_word(S, min(g + 3u, GZ_BLOCK / 4u));
          cg = g;
        }
        uint32_t len = 0, c = 0;
        if (p + GZ_MIN_MATCH <= run_hi && nt < GZ_MAX_TOK) {
          const uint32_t h = k8 ? gz_hash(__funnelshift_r(a, b, k8), __funnelshift_r(b, d, k8), __funnelshift_r(d, e, k8)) : gz_hash(a, b, d);
          const uint32_t entry = S.out[GZ_O_TABLE + gz_slot(h)];
          c = entry >> 16;
          if (c < p && (entry & 0xffffu) == gz_tag(h)) {   // very likely the same 12 bytes: compare, then extend
            const uint32_t max_len = min(258u, run_hi - p);
            while (len < max_len) {   // four bytes at a time
              const uint32_t x = gz_load4(S, c + len) ^ gz_load4(S, p + len);
              if (x) {
                len += (uint32_t)(__ffs((int)x) - 1) >> 3;
                break;
              }
              len += 4u;
            }
            len = min(len, max_len);
            if (len < GZ_MIN_MATCH) len = 0;
          }
        }
        if (len) {
          S.tok[tid][nt++] = (p - p0) | (len << 7) | ((p - c) << 16);
          uint32_t sym, eb, ev;
          gz_len_symbol(len, sym, eb, ev);
          atomicAdd(&hl[sym], 1u);
          gz_dist_symbol(p - c, sym, eb, ev);
          atomicAdd(&hd[sym], 1u);
          p += len;
        } else {
          atomicAdd(&hl[(a >> k8) & 0xffu], 1u);
          ++p;
        }
      }
      S.ntok[tid] = nt;
    }
    __syncthreads();
    // ---- symbol counts (over the dead table), then the two Huffman codes
    uint32_t* cnt = S.out;                        // [286] literal/length, [30] distance behind them
    uint32_t* cntd = cnt + GZ_NLIT;
    uint32_t* sorted = cntd + GZ_NDIST + 2u;      // [286] | [30]
    uint32_t* sortedd = sorted + GZ_NLIT;
    uint32_t* scratch = sortedd + GZ_NDIST + 2u;  // thread 0's arrays (1450 words)
    for (uint32_t s = tid; s < GZ_NLIT + GZ_NDIST; s += GZ_THREADS) {
      uint32_t c = 0;
      if (s < GZ_NLIT) {
        for (uint32_t w = 0; w < 8; ++w) c += S.out[GZ_O_HLIT + w * GZ_NLIT + s];
        if (s == 256u) c = 1;   // end of block
      } else {
        for (uint32_t w = 0; w < 8; ++w) c += S.out[GZ_O_HDIST + w * GZ_NDIST + (s - GZ_NLIT)];
      }
      // (the counts overwrite the table area only: the histograms sit behind it)
      cnt[s < GZ_NLIT ? s : GZ_NLIT + (s - GZ_NLIT)] = c;
    }
    __syncthreads();
    // the used symbols of either alphabet, listed (any order), then ranked among themselves by (count, symbol)
    uint32_t* used = scratch + 1500u;             // [286] | [30]
    uint32_t* usedd = used + GZ_NLIT;
    if (tid == 0) S.hlit = S.hdist = 0;           // (list lengths for now)
    __syncthreads();
    for (uint32_t s = tid; s < GZ_NLIT + GZ_NDIST; s += GZ_THREADS) {
      const bool lit = s < GZ_NLIT;
      const uint32_t me = lit ? s : s - GZ_NLIT;
      if ((lit ? cnt : cntd)[me]) (lit ? used : usedd)[atomicAdd(lit ? &S.hlit : &S.hdist, 1u)] = me;
      (lit ? S.code : S.dcode)[me] = 0;
    }
    __syncthreads();
    {
      const uint32_t m = S.hlit, md = S.hdist;
      for (uint32_t i = tid; i < m + md; i += GZ_THREADS) {
        const bool lit = i < m;
        const uint32_t* cc = lit ? cnt : cntd;
        const uint32_t* uu = lit ? used : usedd;
        const uint32_t nn = lit ? m : md, me = uu[lit ? i : i - m], c = cc[me];
        uint32_t rank = 0;
        for (uint32_t u = 0; u < nn; ++u) {
          const uint32_t o = uu[u], cu = cc[o];
          rank += (cu < c) | ((cu == c) & (o < me));
        }
        (lit ? sorted : sortedd)[rank] = me;
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t m = 0, md = 0, hlit = 257, hdist = 1;
      for (uint32_t s = 0; s < GZ_NLIT; ++s)
        if (cnt[s]) {
          ++m;
          if (s >= 257u) hlit = s + 1u;
        }
      for (uint32_t s = 0; s < GZ_NDIST; ++s)
        if (cntd[s]) {
          ++md;
          hdist = s + 1u;
        }
      uint32_t bits = gz_build_code(cnt, GZ_NLIT, sorted, m, scratch, S.code);
      if (md >= 2u) {
        bits += gz_build_code(cntd, GZ_NDIST, sortedd, md, scratch, S.dcode);
      } else {   // none or one distance in use: two one-bit codes make the set complete for every inflater
        const uint32_t used = md ? sortedd[0] : 0u, other = used ? 0u : 1u;
        S.dcode[used] = 0u | (1u << 16);
        S.dcode[other] = 1u | (1u << 16);
        if (used > other) {   // canonical order: the lower symbol takes code 0
          S.dcode[other] = 0u | (1u << 16);
          S.dcode[used] = 1u | (1u << 16);
        }
        hdist = max(hdist, 2u);
        bits += md ? cntd[used] : 0u;
      }
      for (uint32_t s = 265u; s < 285u; ++s) bits += cnt[s] * gz_len_extra(s);
      for (uint32_t s = 4u; s < GZ_NDIST; ++s) bits += cntd[s] * gz_dist_extra(s);
      S.hlit = hlit;
      S.hdist = hdist;
      S.total_bits = 3u + 14u + 19u * 3u + 4u * (hlit + hdist) + bits;
      S.stored = (S.total_bits + 7u) / 8u >= n + 5u;
      // the whole block's CRC: the warps' parts, left to right
      uint32_t c = S.warp_crc[0];
      for (int w = 1; w < 8; ++w) c = gz_multmodp(K.xshift[5], c) ^ S.warp_crc[w];
      S.warp_crc[0] = c;
    }
    __syncthreads();
    const uint32_t total_bits = S.total_bits, stored = S.stored, hlit = S.hlit, hdist = S.hdist;
    const uint32_t table_bits = 3u + 14u + 19u * 3u + 4u * (hlit + hdist);
    const uint32_t deflate_bytes = stored ? n + 5u : (total_bits + 7u) / 8u;
    const uint32_t member = GZ_HDR + deflate_bytes + GZ_TRL;
    for (uint32_t i = tid; i < GZ_OUT_WORDS; i += GZ_THREADS) S.out[i] = 0;
    __syncthreads();
    uint8_t* ob = reinterpret_cast<uint8_t*>(S.out);
    if (!stored) {
      // ---- bits of every thread's run, block-wide exclusive scan
      const uint32_t nt = S.ntok[tid];
      uint32_t my_bits = 0;
      uint32_t cgw = 0xffffffffu, cw = 0;   // the text word the walk is in
      {
        uint32_t kt = 0, p = run_lo;
        uint32_t t = nt ? S.tok[tid][0] : 0xffffffffu;   // the next match (its position in the low 7 bits), kept in a register
        while (p < run_hi) {
          if (kt < nt && (t & 127u) == p - p0) {
            const uint32_t len = (t >> 7) & 511u, dist = t >> 16;
            t = ++kt < nt ? S.tok[tid][kt] : 0xffffffffu;
            uint32_t sym, eb, ev;
            gz_len_symbol(len, sym, eb, ev);
            my_bits += (S.code[sym] >> 16) + eb;
            gz_dist_symbol(dist, sym, eb, ev);
            my_bits += (S.dcode[sym] >> 16) + eb;
            p += len;
          } else {
            if ((p >> 2) != cgw) {
              cgw = p >> 2;
              cw = gz_word(S, cgw);
            }
            my_bits += S.code[(cw >> ((p & 3u) * 8u)) & 0xffu] >> 16;
            ++p;
          }
        }
      }
      uint32_t incl = my_bits;
#pragma unroll
      for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      if (lane == 31) S.warp_bits[warp] = incl;
      __syncthreads();
      uint32_t before = 0;
      for (uint32_t w = 0; w < warp; ++w) before += S.warp_bits[w];
      const uint32_t at = 8u * GZ_HDR + table_bits + before + incl - my_bits;   // bit position of this thread's first code
      // ---- OR the codes into the image (zeroed): a 64-bit accumulator, whole words out
      uint32_t wi = at >> 5, nb = at & 31u;
      unsigned long long acc = 0;
      auto put = [&](uint32_t bits, uint32_t len) {
        acc |= (unsigned long long)bits << nb;
        nb += len;
        if (nb >= 32u) {
          atomicOr(&S.out[wi++], (uint32_t)acc);
          acc >>= 32;
          nb -= 32u;
        }
      };
      if (tid == 0) {   // block header in front of the first code: BFINAL=1, BTYPE=10, HLIT, HDIST, HCLEN=15 (19)
        uint32_t hw = 8u * GZ_HDR >> 5, hb = (8u * GZ_HDR) & 31u;
        unsigned long long hacc = 0;
        auto hput = [&](uint32_t bits, uint32_t len) {
          hacc |= (unsigned long long)bits << hb;
          hb += len;
          if (hb >= 32u) {
            atomicOr(&S.out[hw++], (uint32_t)hacc);
            hacc >>= 32;
            hb -= 32u;
          }
        };
        hput(1u | (2u << 1), 3);
        hput(hlit - 257u, 5);
        hput(hdist - 1u, 5);
        hput(15u, 4);
        // code-length code: the lengths 0..15 as 4-bit codes, the repeat symbols 16-18 unused (order 16,17,18,0,8,...)
        for (int i = 0; i < 19; ++i) hput(i < 3 ? 0u : 4u, 3);
        for (uint32_t s = 0; s < hlit; ++s) hput(__brev(S.code[s] >> 16) >> 28, 4);   // canonical 4-bit code of value v is v, sent MSB first
        for (uint32_t s = 0; s < hdist; ++s) hput(__brev(S.dcode[s] >> 16) >> 28, 4);
        if (hb) atomicOr(&S.out[hw], (uint32_t)hacc);
      }
      {
        uint32_t kt = 0, p = run_lo;
        uint32_t t = nt ? S.tok[tid][0] : 0xffffffffu;
        while (p < run_hi) {
          if (kt < nt && (t & 127u) == p - p0) {
            const uint32_t len = (t >> 7) & 511u, dist = t >> 16;
            t = ++kt < nt ? S.tok[tid][kt] : 0xffffffffu;
            uint32_t sym, eb, ev;
            gz_len_symbol(len, sym, eb, ev);
            uint32_t e = S.code[sym];
            put((e & 0xffffu) | (ev << (e >> 16)), (e >> 16) + eb);
            gz_dist_symbol(dist, sym, eb, ev);
            e = S.dcode[sym];
            put((e & 0xffffu) | (ev << (e >> 16)), (e >> 16) + eb);
            p += len;
          } else {
            if ((p >> 2) != cgw) {
              cgw = p >> 2;
              cw = gz_word(S, cgw);
            }
            const uint32_t e = S.code[(cw >> ((p & 3u) * 8u)) & 0xffu];
            put(e & 0xffffu, e >> 16);
            ++p;
          }
        }
      }
      if (tid == GZ_THREADS - 1) put(S.code[256] & 0xffffu, S.code[256] >> 16);   // end of block after the last symbol
      if (nb) atomicOr(&S.out[wi], (uint32_t)acc);
      __syncthreads();
    } else {
      if (tid == 0) {
        ob[GZ_HDR] = 1;   // BFINAL=1, BTYPE=00
        ob[GZ_HDR + 1] = (uint8_t)(n & 0xffu);
        ob[GZ_HDR + 2] = (uint8_t)(n >> 8);
        ob[GZ_HDR + 3] = (uint8_t)(~n & 0xffu);
        ob[GZ_HDR + 4] = (uint8_t)((~n >> 8) & 0xffu);
      }
      for (uint32_t i = tid; i < n; i += GZ_THREADS) ob[GZ_HDR + 5u + i] = (uint8_t)gz_text_byte(S, pad + i);
      __syncthreads();
    }
    if (tid == 0) {   // gzip header with the BGZF extra field, trailer
      const uint8_t hdr[GZ_HDR] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, (uint8_t)((member - 1u) & 0xffu), (uint8_t)((member - 1u) >> 8)};
      for (uint32_t i = 0; i < GZ_HDR; ++i) ob[i] = hdr[i];
      const uint32_t c = S.warp_crc[0];
      uint8_t* tr = ob + GZ_HDR + deflate_bytes;
      for (int i = 0; i < 4; ++i) {
        tr[i] = (uint8_t)(c >> (8 * i));
        tr[4 + i] = (uint8_t)(n >> (8 * i));
      }
      g.blk_size[id] = member;
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(g.slots + (size_t)id * GZ_SLOT);
    const uint4* so = reinterpret_cast<const uint4*>(S.out);
    for (uint32_t i = tid; i < (member + 15u) / 16u; i += GZ_THREADS) dst[i] = so[i];
  }
}

// ---------------------------------------------------------------- packed offsets
__global__ void __launch_bounds__(1024) k_gz_scan(GzDev g) {
  __shared__ unsigned long long s_part[1024];
  if (*g.error & g.err_capacity) return;
  const uint32_t n_runs = 2u * g.total, n = g.blk_base[n_runs];
  const uint32_t per = (n + 1023u) / 1024u, i0 = min(n, threadIdx.x * per), i1 = min(n, i0 + per);
  unsigned long long sum = 0;
  for (uint32_t i = i0; i < i1; ++i) sum += g.blk_size[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t d = 1; d < 1024; d <<= 1) {
    const unsigned long long v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0ull;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned long long run = s_part[threadIdx.x] - sum;
  for (uint32_t i = i0; i < i1; ++i) {   // global offsets first (both reads in one sequence)
    g.blk_off[i] = run;
    run += g.blk_size[i];
  }
  if (threadIdx.x == 1023) g.blk_off[n] = s_part[1023];
  __syncthreads();   // the writes above are to global memory: make them visible to the block
  __threadfence_block();
  const uint32_t first1 = g.blk_base[g.total];          // first block of the paired read
  const unsigned long long t2 = g.blk_off[first1];      // bytes of the barcode read's members
  for (uint32_t q = threadIdx.x; q < n_runs; q += 1024u) {
    const uint32_t r = q / g.total, b = q % g.total;
    const unsigned long long o0 = g.blk_off[g.blk_base[q]], o1 = g.blk_off[g.blk_base[q + 1]];
    g.gzseg[(r ? 2u : 0u) * g.total + b] = o0 - (r ? t2 : 0ull);
    g.gzseg[(r ? 3u : 1u) * g.total + b] = o1 - o0;
  }
  if (threadIdx.x == 0 && (t2 > g.cap2 || g.blk_off[n] - t2 > g.cap1)) atomicOr(g.error, g.err_capacity);
}

// members from their slots to their packed places; the text buffers are reused (the text is dead by now)
__global__ void __launch_bounds__(256) k_gz_pack(GzDev g) {
  if (*g.error & g.err_capacity) return;
  const uint32_t n = g.blk_base[2u * g.total], first1 = g.blk_base[g.total];
  const unsigned long long t2 = g.blk_off[first1];
  for (uint32_t id = blockIdx.x; id < n; id += gridDim.x) {
    const uint32_t size = g.blk_size[id];
    const uint8_t* src = g.slots + (size_t)id * GZ_SLOT;
    uint8_t* dst = id < first1 ? g.text2 + g.blk_off[id] : g.text1 + (g.blk_off[id] - t2);
    const uint32_t head = min(size, (uint32_t)((0u - (uintptr_t)dst) & 3u));
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    // destination-aligned words: four source bytes each, the source is 16-byte aligned + head
    const uint32_t n_words = (size - head) >> 2;
    const uint32_t sh = (head & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
    uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t i = threadIdx.x; i < n_words; i += 256u) dw[i] = sh ? __funnelshift_r(sw[i], sw[i + 1], sh) : sw[i];
    const uint32_t done = head + (n_words << 2);
    if (threadIdx.x < size - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
  }
}

}  // namespace mgk

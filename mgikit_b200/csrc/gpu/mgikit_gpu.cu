// libmgikit_gpu.so — the C ABI of include/mgikit_gpu.h over the sm_100a kernels.
// One context per GPU; each in-flight chunk owns a slot (its own stream, device
// scratch and pinned result buffers), so copies of chunk k+1 overlap the kernels
// of chunk k.  No CPU fallback: without a usable B200 mgk_create() fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "gz_kernels.cuh"
#include "kernels.cuh"
#include "params_build.hpp"

using namespace mgk;

namespace {

thread_local char g_create_err[512] = "";

struct Slot {
  int state = 0;  // 0 free, 1 submitted, 2 collected (waiting for release)
  uint64_t seq = 0;
  uint32_t flags = 0;
  uint64_t len2 = 0, len1 = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};  // start, after scan, after match, after bin scan, after scatter, end
  cudaEvent_t ev_status = nullptr;
  // device
  uint8_t *d_r2 = nullptr, *d_r1 = nullptr, *out2 = nullptr, *out1 = nullptr;
  uint32_t *nl2 = nullptr, *nl1 = nullptr, *ctrl = nullptr, *tile_bins = nullptr, *group_sums = nullptr, *unassigned = nullptr;
  uint32_t *stg2 = nullptr, *stg1 = nullptr;
  unsigned long long* seg = nullptr;
  RecMeta* meta = nullptr;
  Status* status = nullptr;
  // pinned host
  uint8_t *h_out2 = nullptr, *h_out1 = nullptr;
  Status* h_status = nullptr;
  unsigned long long* h_seg = nullptr;
  uint32_t* h_unassigned = nullptr;
  std::vector<uint64_t> off2, len2v, off1, len1v;
  ChunkDev cd{};
  float ms[5] = {0, 0, 0, 0, 0};
  float gz_ms = 0;
  bool in_span = false;   // ev[5] was recorded since the span was opened
  // MGK_OUT_GZIP (allocated at the slot's first such chunk)
  GzDev gz{};
  uint8_t* gz_slots = nullptr;
  uint32_t *gz_base = nullptr, *gz_size = nullptr;
  unsigned long long *gz_off = nullptr, *gz_seg = nullptr;
  cudaEvent_t ev_gz = nullptr;
};

}  // namespace

struct mgk_ctx {
  int device = 0;
  Params hp{};              // host copy (pointers are device pointers)
  Params* d_params = nullptr;
  uint64_t* d_codes = nullptr;
  int32_t *d_pairs = nullptr, *d_xlat = nullptr, *d_writing = nullptr, *d_hids = nullptr;
  uint64_t* d_hkeys = nullptr;
  uint8_t* d_consts = nullptr;
  unsigned long long *d_stats = nullptr, *d_mism = nullptr, *d_prof = nullptr;
  unsigned long long *d_red_stats = nullptr, *d_red_mism = nullptr;   // result of the last counter all-reduce
  bool reduced = false;
  Status* h_status_init = nullptr;
  int total = 0, mism_w = 0;
  uint64_t cap_bytes = 0, cap_records = 0, out_cap = 0;
  uint32_t scan_grid = 0, scan_cap = 0, tiles_max = 0, groups_max = 0;
  size_t match_smem = 0;
  int mism_smem = 0;
  int grid_match = 0, grid_emit = 0, n_sm = 0;
  uint32_t bins_stride = 0;
  EmitLayout lay{};
  std::vector<Slot> slots;
  std::deque<int> fifo;     // submitted slots, oldest first
  uint64_t launches = 0;
  float last_ms[5] = {0, 0, 0, 0, 0};
  float last_gz_ms = 0;
  GzConsts gzk{};
  ncclComm_t comm = nullptr;
  cudaEvent_t span_begin = nullptr;
  bool span_open = false;
  char err[512] = "";
};

namespace {

int fail(mgk_ctx* c, int code, const char* fmt, ...) {
  char* dst = c ? c->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(ctx, MGK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
  return cudaMalloc(reinterpret_cast<void**>(p), n ? n * sizeof(T) : sizeof(T));
}

void free_slot(Slot& s) {
  cudaFree(s.d_r2); cudaFree(s.d_r1); cudaFree(s.out2); cudaFree(s.out1);
  cudaFree(s.nl2); cudaFree(s.nl1); cudaFree(s.ctrl); cudaFree(s.tile_bins); cudaFree(s.group_sums);
  cudaFree(s.unassigned); cudaFree(s.stg2); cudaFree(s.stg1); cudaFree(s.seg); cudaFree(s.meta); cudaFree(s.status);
  cudaFreeHost(s.h_out2); cudaFreeHost(s.h_out1); cudaFreeHost(s.h_status); cudaFreeHost(s.h_seg); cudaFreeHost(s.h_unassigned);
  cudaFree(s.gz_slots); cudaFree(s.gz_base); cudaFree(s.gz_size); cudaFree(s.gz_off); cudaFree(s.gz_seg);
  if (s.ev_gz) cudaEventDestroy(s.ev_gz);
  for (auto& e : s.ev) if (e) cudaEventDestroy(e);
  if (s.ev_status) cudaEventDestroy(s.ev_status);
  if (s.stream) cudaStreamDestroy(s.stream);
}

// ---------------------------------------------------------------- NCCL (resolved at run time)
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
#define SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
      SYM(GetUniqueId, "ncclGetUniqueId");
      SYM(CommInitRank, "ncclCommInitRank");
      SYM(CommInitAll, "ncclCommInitAll");
      SYM(AllReduce, "ncclAllReduce");
      SYM(GroupStart, "ncclGroupStart");
      SYM(GroupEnd, "ncclGroupEnd");
      SYM(CommDestroy, "ncclCommDestroy");
      SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.AllReduce || !api.GroupStart || !api.GroupEnd) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  });
  return api.lib ? &api : nullptr;
}

// MGK_IN_DEVICE: is [p, p + len + slack) inside the allocation p belongs to?  true when the driver cannot tell.
bool device_range_ok(const uint8_t* p, uint64_t len) {
  if (!p || !len) return true;
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<range_fn>(f);
    cudaGetLastError();
  });
  if (!fn) return true;
  unsigned long long base = 0;
  size_t size = 0;
  if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return true;
  return (unsigned long long)(uintptr_t)p + len + MGK_DEVICE_INPUT_SLACK <= base + size;
}

}  // namespace

// ====================================================================== create / destroy
extern "C" const char* mgk_last_error(const mgk_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" void mgk_destroy(mgk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->d_prof) {   // last k_emit launch: mean cycles per warp index over the CTAs
    const int nw = EMIT_THREADS / 32;
    std::vector<unsigned long long> h((size_t)ctx->n_sm * nw * 6);
    cudaMemcpy(h.data(), ctx->d_prof, h.size() * 8, cudaMemcpyDeviceToHost);
    for (int w = 0; w < nw; ++w) {
      double a[6] = {0, 0, 0, 0, 0, 0};
      for (int b = 0; b < ctx->n_sm; ++b)
        for (int k = 0; k < 6; ++k) a[k] += (double)h[((size_t)b * nw + w) * 6 + k] / ctx->n_sm;
      fprintf(stderr, "[mgk profile] k_emit warp %2d: stage wait %.0f, descriptors+barrier %.0f, role %.0f, bodies %.0f, barrier %.0f, headers %.0f cycles\n", w, a[0], a[1], a[2],
              a[3], a[4], a[5]);
    }
    cudaFree(ctx->d_prof);
  }
  for (auto& s : ctx->slots) free_slot(s);
  cudaFree(ctx->d_params); cudaFree(ctx->d_codes); cudaFree(ctx->d_pairs); cudaFree(ctx->d_xlat);
  cudaFree(ctx->d_hkeys); cudaFree(ctx->d_hids);
  cudaFree(ctx->d_writing); cudaFree(ctx->d_consts); cudaFree(ctx->d_stats); cudaFree(ctx->d_mism);
  cudaFree(ctx->d_red_stats); cudaFree(ctx->d_red_mism);
  cudaFreeHost(ctx->h_status_init);
  if (ctx->span_begin) cudaEventDestroy(ctx->span_begin);
  if (ctx->comm) {
    NcclApi* n = nccl_api();
    if (n && n->CommDestroy) n->CommDestroy(ctx->comm);
  }
  delete ctx;
}

extern "C" int mgk_create(const mgk_config* cfg, mgk_ctx** out) {
  mgk_ctx* ctx = nullptr;  // errors before the context exists go to the thread-local buffer
  if (!cfg || !out) return fail(ctx, MGK_ERR_ARG, "mgk_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != MGK_ABI_VERSION) return fail(ctx, MGK_ERR_ARG, "mgk_create: ABI version %d, library speaks %d", cfg->abi_version, MGK_ABI_VERSION);
  if (cfg->n_templates < (cfg->reformat ? 0 : 1) || cfg->n_templates > MGK_MAX_TEMPLATES || cfg->n_samples < 1 ||
      (!cfg->templates && cfg->n_templates > 0) || !cfg->writing_sample)
    return fail(ctx, MGK_ERR_ARG, "mgk_create: bad sample/template tables");
  if (cfg->n_samples + 2 > 60000) return fail(ctx, MGK_ERR_ARG, "mgk_create: more than 60000 samples are not supported");
  if (cfg->max_chunk_bytes == 0 || cfg->max_chunk_bytes >= (1ull << 31) || cfg->max_chunk_records == 0)
    return fail(ctx, MGK_ERR_ARG, "mgk_create: max_chunk_bytes must be in (0, 2^31) and max_chunk_records > 0");
  if (cfg->allowed_mismatches < 0 || cfg->allowed_mismatches > 100) return fail(ctx, MGK_ERR_ARG, "mgk_create: bad mismatch count");

  const bool trace = getenv("MGK_TRACE") != nullptr;   // developer aid: where the set-up time goes
  const auto t_create = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (trace) fprintf(stderr, "[mgk trace] create: %s at %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_create).count());
  };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(ctx, MGK_ERR_NO_DEVICE, "no CUDA device is visible: the demultiplex hot path has no CPU fallback");
  lap("device count");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(ctx, MGK_ERR_NO_DEVICE, "CUDA device %d does not exist (%d visible)", cfg->device, ndev);
  int cc_major = 0, cc_minor = 0, n_sm = 0;   // attributes one by one: cudaGetDeviceProperties queries everything the driver knows
  if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, cfg->device) != cudaSuccess ||
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device) != cudaSuccess)
    return fail(ctx, MGK_ERR_NO_DEVICE, "can not query device %d", cfg->device);
  if (cc_major != 10)
    return fail(ctx, MGK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only (B200)", cfg->device, cc_major, cc_minor);
  lap("device attributes");

  ctx = new mgk_ctx();
  ctx->device = cfg->device;
  ctx->n_sm = n_sm;
#define CUC(call)                                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      fail(nullptr, MGK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                               \
      mgk_destroy(ctx);                                                                                   \
      return MGK_ERR_CUDA;                                                                                \
    }                                                                                                     \
  } while (0)
#define BAD(...)                                 \
  do {                                           \
    fail(nullptr, MGK_ERR_ARG, __VA_ARGS__);     \
    mgk_destroy(ctx);                            \
    return MGK_ERR_ARG;                          \
  } while (0)
  CUC(cudaSetDevice(cfg->device));
  CUC(cudaFree(nullptr));
  lap("device context");
  if (const char* g = getenv("MGK_L2_FETCH")) {   // developer aid: L2 fetch granularity hint (32 / 64 / 128 bytes)
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    fprintf(stderr, "[mgk] L2 fetch granularity: asked %s, device says %zu\n", g, got);
  }

  Params& P = ctx->hp;
  HostTables tb;
  {
    std::string why;
    if (!build_params_host(cfg, P, tb, why)) BAD("mgk_create: %s", why.c_str());
  }
  ctx->total = P.total;
  ctx->mism_w = P.mism_w;
  const std::vector<uint64_t>& codes = tb.codes;
  const std::vector<int32_t>&pairs = tb.pairs, &xlat = tb.xlat;
  const std::string& consts = tb.consts;

  CUC(dalloc(&ctx->d_codes, codes.size() + 1));
  CUC(dalloc(&ctx->d_pairs, pairs.size() + 1));
  CUC(dalloc(&ctx->d_xlat, xlat.size() + 1));
  CUC(dalloc(&ctx->d_writing, (size_t)P.total));
  CUC(dalloc(&ctx->d_hkeys, tb.hkeys.size() + 2));
  CUC(dalloc(&ctx->d_hids, tb.hids.size() + 1));
  if (!tb.hids.empty()) {
    CUC(cudaMemcpy(ctx->d_hkeys, tb.hkeys.data(), tb.hkeys.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CUC(cudaMemcpy(ctx->d_hids, tb.hids.data(), tb.hids.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  CUC(dalloc(&ctx->d_consts, consts.size()));
  CUC(dalloc(&ctx->d_stats, (size_t)P.total * MGK_N_STATS));
  CUC(dalloc(&ctx->d_mism, (size_t)P.total * P.mism_w));
  CUC(dalloc(&ctx->d_red_stats, (size_t)P.total * MGK_N_STATS));
  CUC(dalloc(&ctx->d_red_mism, (size_t)P.total * P.mism_w));
  CUC(dalloc(&ctx->d_params, 1));
  CUC(cudaMemcpy(ctx->d_codes, codes.data(), codes.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(ctx->d_pairs, pairs.data(), pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(ctx->d_xlat, xlat.data(), xlat.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(ctx->d_writing, cfg->writing_sample, (size_t)P.total * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUC(cudaMemcpy(ctx->d_consts, consts.data(), consts.size(), cudaMemcpyHostToDevice));
  CUC(cudaMemset(ctx->d_stats, 0, (size_t)P.total * MGK_N_STATS * sizeof(uint64_t)));
  CUC(cudaMemset(ctx->d_mism, 0, (size_t)P.total * P.mism_w * sizeof(uint64_t)));
  P.codes = ctx->d_codes;
  P.pairs = ctx->d_pairs;
  P.xlat = ctx->d_xlat;
  P.writing = ctx->d_writing;
  P.consts = ctx->d_consts;
  P.hkeys = ctx->d_hkeys;
  P.hids = ctx->d_hids;
  CUC(cudaMemcpy(ctx->d_params, &P, sizeof P, cudaMemcpyHostToDevice));

  // capacities
  ctx->cap_bytes = cfg->max_chunk_bytes;
  ctx->cap_records = cfg->max_chunk_records;
  const uint64_t growth = (uint64_t)P.prefix_len + 2ull * P.BL + 32ull + (uint64_t)P.rf_bc_len;
  ctx->out_cap = ctx->cap_bytes + ctx->cap_records * growth;
  if (ctx->out_cap >= (1ull << 32) - 4096) BAD("mgk_create: chunk capacity too large for 32-bit output offsets; lower max_chunk_bytes");
  // newline scan: one CTA per SM and stream, each with its own run of the staging array.  A run holds four times
  // the mean share of a range plus a whole stage of newlines; lanes whose newline density is more skewed than
  // that are refused (MGK_ERR_CAPACITY), real FASTQ is nowhere near.
  ctx->scan_grid = (uint32_t)ctx->n_sm * 3u / (P.paired ? 2u : 1u);   // 3 CTAs per SM over both streams, one wave
  ctx->scan_cap = (uint32_t)((4 * (ctx->cap_records * 4 + 4)) / ((uint64_t)ctx->scan_grid * SCAN_WARPS) + SCAN_WTILE + 64) & ~3u;   // per range (one per warp)
  CUC(cudaFuncSetAttribute(k_scan_newlines, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_STAGES * SCAN_TILE));
  ctx->tiles_max = (uint32_t)((ctx->cap_records + TILE_RECORDS / 2 - 1) / (TILE_RECORDS / 2)) + 1;   // plan tiles hold 128..256 records (device_tile_records)
  ctx->groups_max = (ctx->tiles_max + GROUP_TILES - 1) / GROUP_TILES;
  ctx->match_smem = (size_t)P.n_codes * 8 + (size_t)2 * P.total * 4 + 16;
  ctx->mism_smem = (size_t)P.total * P.mism_w <= 8192 ? 1 : 0;   // per-CTA mismatch counters in shared memory
  if (ctx->mism_smem) ctx->match_smem += (size_t)P.total * P.mism_w * 4;
  if (ctx->match_smem > 200 * 1024) BAD("mgk_create: index tables (%zu B) do not fit the shared-memory staging area", ctx->match_smem);
  CUC(cudaFuncSetAttribute(k_match_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->match_smem));
  CUC(cudaFuncSetAttribute(k_match_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->match_smem));
  int occ_m = 1;
  CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, k_match_plan, TILE_RECORDS, ctx->match_smem));
  ctx->grid_match = ctx->n_sm * (occ_m > 0 ? occ_m : 1);
  ctx->bins_stride = (2u * (uint32_t)P.total + 3u) & ~3u;
  {  // shared-memory plan of the emit kernel: everything that is not a stage is fixed, the two stages share the rest
    cudaFuncAttributes fa;
    CUC(cudaFuncGetAttributes(&fa, k_emit));
    int max_optin = 0;
    CUC(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    auto up = [](uint32_t v, uint32_t a) { return (v + a - 1u) / a * a; };
    EmitLayout& L = ctx->lay;
    uint32_t at = 0;
    L.consts = at;
    at += up((uint32_t)P.prefix_len + LIT_BYTES + (uint32_t)P.rf_bc_len + 48u, 16u);
    L.stats = at;
    at += up((uint32_t)P.total * QC_SLOTS * 4u, 16u);
    L.writing = at;
    at += up((uint32_t)P.total * 4u, 16u);
    at += 32u;                                   // slack in front of the header staging areas (gather16 reads around them)
    L.hdr = at;
    at += (P.out_mode == MGK_OUT_MGI_FULL_HEADER && P.paired ? 2u : 1u) * HDR_KIND_BYTES + 32u;   // the paired read has its own header only with --mgi-full-header
    L.desc = at;
    at += 2u * EMIT_NMAX * EMIT_PIECES * (uint32_t)sizeof(PieceDesc);   // two tables (tile parity)
    at = up(at, 128u);
    const uint32_t nl_bytes = 16u * (EMIT_NMAX + 1u);
    L.nl1 = up(nl_bytes, 16u);
    L.meta = L.nl1 + up(nl_bytes, 16u);
    L.bins = L.meta + 16u * EMIT_NMAX;
    L.bytes = up(L.bins + 4u * ctx->bins_stride + 16u, 128u);
    const int64_t avail = (int64_t)max_optin - (int64_t)fa.sharedSizeBytes - (int64_t)at - 256 - 512;   // 512: the movers' unpredicated loads run past a range
    const int64_t cap = (avail / 2 - (int64_t)L.bytes) / 128 * 128;
    if (cap < 4096) BAD("mgk_create: %d samples leave no shared memory for the record stages", P.S);
    L.cap = (uint32_t)cap;
    L.stage0 = at;
    L.stage_stride = L.bytes + L.cap;
    L.total_bytes = L.stage0 + 2u * L.stage_stride + 512u;
    CUC(cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total_bytes));
    int occ_e = 0;
    CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_emit, EMIT_THREADS, L.total_bytes));
    if (occ_e < 1) BAD("mgk_create: the emit kernel does not fit an SM (%u bytes of shared memory)", L.total_bytes);
    ctx->grid_emit = ctx->n_sm;
  }

  ctx->gzk = gz_make_consts();
  CUC(cudaFuncSetAttribute(k_gz_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GzShared)));
  if (getenv("MGK_PROFILE")) {   // developer aid: per-warp cycle counters of k_emit, dumped at destroy
    CUC(dalloc(&ctx->d_prof, (size_t)ctx->n_sm * (EMIT_THREADS / 32) * 6));
    CUC(cudaMemset(ctx->d_prof, 0, (size_t)ctx->n_sm * (EMIT_THREADS / 32) * 6 * 8));
  }
  CUC(cudaEventCreate(&ctx->span_begin));
  CUC(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_status_init), sizeof(Status), cudaHostAllocDefault));
  memset(ctx->h_status_init, 0, sizeof(Status));
  ctx->h_status_init->fmt_record = ~0ull;
  ctx->h_status_init->validate_key = ~0ull;

  lap("tables and kernel attributes");
  const int n_slots = cfg->n_slots < 1 ? 1 : (cfg->n_slots > 8 ? 8 : cfg->n_slots);
  ctx->slots.resize((size_t)n_slots);
  const size_t B2 = (size_t)2 * P.total;
  const uint32_t nl_sentinel[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  for (auto& s : ctx->slots) {
    CUC(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (auto& e : s.ev) CUC(cudaEventCreate(&e));
    CUC(cudaEventCreateWithFlags(&s.ev_status, cudaEventDisableTiming));
    CUC(dalloc(&s.d_r2, ctx->cap_bytes + 256));
    if (P.paired) CUC(dalloc(&s.d_r1, ctx->cap_bytes + 256));
    CUC(dalloc(&s.out2, ctx->out_cap + 256));
    if (P.paired) CUC(dalloc(&s.out1, ctx->out_cap + 256));
    // newline tables: 4 sentinel entries (-1: "a newline just before byte 0") in front of entry 0, so that
    // record r's previous newline is always nl[4r-1] and the rows stay 16-byte aligned for the bulk copies
    CUC(dalloc(&s.nl2, ctx->cap_records * 4 + 16));
    CUC(cudaMemcpy(s.nl2, nl_sentinel, sizeof nl_sentinel, cudaMemcpyHostToDevice));
    if (P.paired) {
      CUC(dalloc(&s.nl1, ctx->cap_records * 4 + 16));
      CUC(cudaMemcpy(s.nl1, nl_sentinel, sizeof nl_sentinel, cudaMemcpyHostToDevice));
    }
    CUC(dalloc(&s.stg2, (size_t)ctx->scan_grid * SCAN_WARPS * ctx->scan_cap));
    if (P.paired) CUC(dalloc(&s.stg1, (size_t)ctx->scan_grid * SCAN_WARPS * ctx->scan_cap));
    CUC(dalloc(&s.ctrl, 8 + 2 * (size_t)ctx->scan_grid * SCAN_WARPS));
    CUC(dalloc(&s.meta, ctx->cap_records + 1));
    CUC(dalloc(&s.tile_bins, (size_t)ctx->tiles_max * ctx->bins_stride));
    CUC(dalloc(&s.group_sums, (size_t)ctx->groups_max * B2));
    CUC(dalloc(&s.seg, (size_t)4 * P.total));
    CUC(dalloc(&s.unassigned, ctx->cap_records + 1));
    CUC(dalloc(&s.status, 1));
    CUC(cudaHostAlloc(reinterpret_cast<void**>(&s.h_out2), ctx->out_cap + 256, cudaHostAllocDefault));
    if (P.paired) CUC(cudaHostAlloc(reinterpret_cast<void**>(&s.h_out1), ctx->out_cap + 256, cudaHostAllocDefault));
    CUC(cudaHostAlloc(reinterpret_cast<void**>(&s.h_status), sizeof(Status), cudaHostAllocDefault));
    CUC(cudaHostAlloc(reinterpret_cast<void**>(&s.h_seg), (size_t)4 * P.total * 8, cudaHostAllocDefault));
    CUC(cudaHostAlloc(reinterpret_cast<void**>(&s.h_unassigned), (ctx->cap_records + 1) * 4, cudaHostAllocDefault));
    s.off2.resize((size_t)P.total); s.len2v.resize((size_t)P.total); s.off1.resize((size_t)P.total); s.len1v.resize((size_t)P.total);
    ChunkDev& c = s.cd;
    c.nl2 = s.nl2 + 4; c.nl1 = s.nl1 ? s.nl1 + 4 : nullptr;
    c.bins_stride = ctx->bins_stride;
    c.emit_cap = ctx->lay.cap;
    c.nl_cap = (uint32_t)(ctx->cap_records * 4 + 4);
    c.max_records = (uint32_t)ctx->cap_records;
    c.stg2 = s.stg2; c.stg1 = s.stg1;
    c.scan_grid = ctx->scan_grid; c.scan_cap = ctx->scan_cap;
    c.counts = s.ctrl; c.range_counts = s.ctrl + 8;
    c.meta = s.meta; c.tile_bins = s.tile_bins; c.group_sums = s.group_sums; c.seg = s.seg;
    c.out2 = s.out2; c.out1 = s.out1;
    c.out_cap2 = ctx->out_cap; c.out_cap1 = ctx->out_cap;
    c.unassigned = s.unassigned; c.status = s.status;
    c.stats = ctx->d_stats; c.mism = ctx->d_mism;
    c.prof = ctx->d_prof;
  }
#undef CUC
#undef BAD
  lap("slots (device scratch, pinned result buffers)");
  *out = ctx;
  return MGK_OK;
}

// ====================================================================== submit / collect / release
extern "C" int mgk_submit(mgk_ctx* ctx, uint64_t seq, const uint8_t* r2, uint64_t r2_len, const uint8_t* r1, uint64_t r1_len,
                          uint32_t flags) {
  if (!ctx) return MGK_ERR_ARG;
  const Params& P = ctx->hp;
  if (!r2 && r2_len) return fail(ctx, MGK_ERR_ARG, "mgk_submit: null barcode-read buffer");
  if (P.paired && !r1 && r1_len) return fail(ctx, MGK_ERR_ARG, "mgk_submit: paired run needs both reads");
  if (!P.paired) r1_len = 0;
  if (r2_len > ctx->cap_bytes || r1_len > ctx->cap_bytes)
    return fail(ctx, MGK_ERR_CAPACITY, "mgk_submit: chunk of %llu/%llu bytes exceeds max_chunk_bytes %llu", (unsigned long long)r2_len,
                (unsigned long long)r1_len, (unsigned long long)ctx->cap_bytes);
  if (flags & MGK_IN_DEVICE) {
    if ((((uintptr_t)r2) | ((uintptr_t)r1)) & 15u) return fail(ctx, MGK_ERR_ARG, "mgk_submit: device input pointers must be 16-byte aligned");
    // the kernels fetch whole 16-byte units (bulk copies, word loads): the caller's allocation must stay readable
    // for MGK_DEVICE_INPUT_SLACK bytes past the text (include/mgikit_gpu.h); checked where the driver can tell
    if (!device_range_ok(r2, r2_len) || !device_range_ok(r1, r1_len))
      return fail(ctx, MGK_ERR_ARG, "mgk_submit: a device input buffer ends less than %d readable bytes after its text", MGK_DEVICE_INPUT_SLACK);
  }
  int si = -1;
  for (size_t i = 0; i < ctx->slots.size(); ++i)
    if (ctx->slots[i].state == 0) {
      si = (int)i;
      break;
    }
  if (si < 0) return fail(ctx, MGK_ERR_STATE, "mgk_submit: all %zu slots are in flight; collect and release first", ctx->slots.size());
  CU(cudaSetDevice(ctx->device));
  Slot& s = ctx->slots[(size_t)si];
  s.seq = seq;
  s.flags = flags;
  s.len2 = r2_len;
  s.len1 = r1_len;
  ChunkDev& c = s.cd;
  c.len2 = (uint32_t)r2_len;
  c.len1 = (uint32_t)r1_len;
  cudaStream_t st = s.stream;
  static const bool trace = getenv("MGK_TRACE") != nullptr;
  const auto t_sub = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (trace) fprintf(stderr, "[mgk trace] submit %llu: %s at %.3f ms\n", (unsigned long long)seq, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_sub).count());
  };
  CU(cudaEventRecord(s.ev[0], st));
  if (!ctx->span_open) {
    CU(cudaEventRecord(ctx->span_begin, st));
    ctx->span_open = true;
  }
  if (flags & MGK_IN_DEVICE) {
    c.r2 = r2;
    c.r1 = r1;
  } else {
    if (r2_len) CU(cudaMemcpyAsync(s.d_r2, r2, r2_len, cudaMemcpyHostToDevice, st));
    if (r1_len) CU(cudaMemcpyAsync(s.d_r1, r1, r1_len, cudaMemcpyHostToDevice, st));
    c.r2 = s.d_r2;
    c.r1 = s.d_r1;
  }
  CU(cudaMemcpyAsync(s.status, ctx->h_status_init, sizeof(Status), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(s.ev[1], st));  // [0..1): input copy
  lap("copies queued");
  {
    const dim3 grid(ctx->scan_grid, P.paired ? 2 : 1);
    k_scan_newlines<<<grid, SCAN_THREADS, SCAN_STAGES * SCAN_TILE, st>>>(c, P.paired ? 2 : 1);
    k_gather_newlines<<<dim3(grid.x * SCAN_WARPS, grid.y), 256, 0, st>>>(c, P.paired ? 2 : 1);
    ctx->launches += 2;
  }
  CU(cudaEventRecord(s.ev[2], st));
  // the grids below are sized from the byte length (a record is >= 8 bytes); the kernels
  // themselves read the record count the scan left on the device
  const uint64_t rec_bound = std::min<uint64_t>(ctx->cap_records, r2_len / 8 + 1);
  const uint32_t tiles_bound = (uint32_t)((rec_bound + TILE_RECORDS / 2 - 1) / (TILE_RECORDS / 2));
  const uint32_t groups_bound = (tiles_bound + GROUP_TILES - 1) / GROUP_TILES;
  const uint32_t B2 = 2u * (uint32_t)P.total;
  k_match_plan<<<std::min<uint32_t>(tiles_bound, (uint32_t)ctx->grid_match), TILE_RECORDS, ctx->match_smem, st>>>(ctx->d_params, c, ctx->mism_smem);
  CU(cudaEventRecord(s.ev[3], st));
  const uint32_t scan_threads = groups_bound * B2;
  k_scan_groups<<<(scan_threads + 255) / 256, 256, 0, st>>>(ctx->d_params, c);
  k_scan_bins<<<1, 1024, 0, st>>>(ctx->d_params, c);
  k_scan_tiles<<<(scan_threads + 255) / 256, 256, 0, st>>>(ctx->d_params, c);
  CU(cudaEventRecord(s.ev[4], st));
  k_emit<<<std::min<uint32_t>(tiles_bound, (uint32_t)ctx->grid_emit), EMIT_THREADS, ctx->lay.total_bytes, st>>>(ctx->d_params, c, ctx->lay);
  ctx->launches += 5;
  CU(cudaEventRecord(s.ev[5], st));
  s.in_span = true;
  lap("kernels queued");
  if (flags & MGK_OUT_GZIP) {   // the segments as runs of gzip members, packed over the text
    if (!s.gz_slots) {
      const uint64_t blocks_cap = (ctx->out_cap * (P.paired ? 2u : 1u)) / GZ_BLOCK + 2ull * P.total + 2ull;
      CU(dalloc(&s.gz_slots, (size_t)blocks_cap * GZ_SLOT));
      CU(dalloc(&s.gz_base, (size_t)2 * P.total + 1));
      CU(dalloc(&s.gz_size, (size_t)blocks_cap));
      CU(dalloc(&s.gz_off, (size_t)blocks_cap + 1));
      CU(dalloc(&s.gz_seg, (size_t)4 * P.total));
      CU(cudaEventCreate(&s.ev_gz));
      GzDev& g = s.gz;
      g.seg = s.seg; g.text2 = s.out2; g.text1 = s.out1; g.slots = s.gz_slots; g.blk_base = s.gz_base; g.blk_size = s.gz_size;
      g.blk_off = s.gz_off; g.gzseg = s.gz_seg; g.total = (uint32_t)P.total; g.blocks_cap = (uint32_t)blocks_cap;
      g.cap2 = ctx->out_cap; g.cap1 = ctx->out_cap; g.error = &s.status->error; g.err_capacity = ERR_OUT_CAPACITY;
    }
    const uint64_t out_bound = std::min<uint64_t>(ctx->out_cap, r2_len + r1_len + rec_bound * ((uint64_t)P.prefix_len + 2ull * P.BL + 32ull + (uint64_t)P.rf_bc_len));
    const uint32_t blocks_bound = (uint32_t)std::min<uint64_t>(s.gz.blocks_cap, out_bound / GZ_BLOCK + 2ull * P.total + 2ull);
    k_gz_plan<<<1, 1024, 0, st>>>(s.gz);
    k_gz_blocks<<<std::min<uint32_t>(blocks_bound, (uint32_t)ctx->n_sm * 2u), GZ_THREADS, sizeof(GzShared), st>>>(s.gz, ctx->gzk);
    k_gz_scan<<<1, 1024, 0, st>>>(s.gz);
    k_gz_pack<<<std::min<uint32_t>(blocks_bound, (uint32_t)ctx->n_sm * 8u), 256, 0, st>>>(s.gz);
    ctx->launches += 4;
    CU(cudaEventRecord(s.ev_gz, st));
  }
  CU(cudaMemcpyAsync(s.h_status, s.status, sizeof(Status), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(s.h_seg, (flags & MGK_OUT_GZIP) ? s.gz_seg : s.seg, (size_t)4 * P.total * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(s.ev_status, st));
  CU(cudaGetLastError());
  lap("all queued");
  s.state = 1;
  ctx->fifo.push_back(si);
  return MGK_OK;
}

extern "C" int mgk_collect(mgk_ctx* ctx, mgk_result* res) {
  if (!ctx || !res) return MGK_ERR_ARG;
  if (ctx->fifo.empty()) return fail(ctx, MGK_ERR_STATE, "mgk_collect: nothing in flight");
  CU(cudaSetDevice(ctx->device));
  const int si = ctx->fifo.front();
  Slot& s = ctx->slots[(size_t)si];
  const Params& P = ctx->hp;
  CU(cudaEventSynchronize(s.ev_status));
  ctx->fifo.pop_front();
  s.state = 2;
  const Status& hs = *s.h_status;
  for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&s.ms[i], s.ev[i + 1], s.ev[i + 2]);
  cudaEventElapsedTime(&s.ms[4], s.ev[0], s.ev[5]);
  memcpy(ctx->last_ms, s.ms, sizeof s.ms);
  s.gz_ms = 0.f;
  if (s.flags & MGK_OUT_GZIP) cudaEventElapsedTime(&s.gz_ms, s.ev[5], s.ev_gz);
  ctx->last_gz_ms = s.gz_ms;
  if (hs.error & ERR_PANIC)
    return fail(ctx, MGK_ERR_FORMAT, "chunk %llu: a read hits a sample-sheet combination on which the reference panics (stale template dictionary)",
                (unsigned long long)s.seq);
  if (hs.error & ERR_NL_CAPACITY)
    return fail(ctx, MGK_ERR_CAPACITY,
                "chunk %llu: more line ends than the record tables hold (over max_chunk_records = %llu, or one 1/%u-th of the chunk carries "
                "more than four times its share of them)",
                (unsigned long long)s.seq, (unsigned long long)ctx->cap_records, ctx->scan_grid * (uint32_t)SCAN_WARPS);
  if (hs.error & ERR_REC_CAPACITY)
    return fail(ctx, MGK_ERR_CAPACITY, "chunk %llu holds more records than max_chunk_records = %llu", (unsigned long long)s.seq,
                (unsigned long long)ctx->cap_records);
  if (hs.error & ERR_OUT_CAPACITY) return fail(ctx, MGK_ERR_CAPACITY, "chunk %llu: output does not fit the output buffer", (unsigned long long)s.seq);
  if (hs.error & ERR_REC_TOO_LONG)
    return fail(ctx, MGK_ERR_CAPACITY, "chunk %llu: a read pair is longer than the %u bytes one shared-memory stage holds", (unsigned long long)s.seq, ctx->lay.cap);
  if (hs.fmt_record != ~0ull)
    return fail(ctx, MGK_ERR_FORMAT, "chunk %llu, read %llu: lines shorter than the barcode/header layout", (unsigned long long)s.seq, hs.fmt_record);
  const int total = P.total;
  for (int b = 0; b < total; ++b) {
    s.off2[(size_t)b] = s.h_seg[0 * total + b];
    s.len2v[(size_t)b] = s.h_seg[1 * total + b];
    s.off1[(size_t)b] = s.h_seg[2 * total + b];
    s.len1v[(size_t)b] = s.h_seg[3 * total + b];
  }
  const uint64_t tot2 = s.off2[(size_t)total - 1] + s.len2v[(size_t)total - 1];
  const uint64_t tot1 = s.off1[(size_t)total - 1] + s.len1v[(size_t)total - 1];
  memset(res, 0, sizeof *res);
  if (!(s.flags & MGK_OUT_DEVICE)) {
    if (tot2) CU(cudaMemcpyAsync(s.h_out2, s.out2, tot2, cudaMemcpyDeviceToHost, s.stream));
    if (tot1) CU(cudaMemcpyAsync(s.h_out1, s.out1, tot1, cudaMemcpyDeviceToHost, s.stream));
    res->out_r2 = s.h_out2;
    res->out_r1 = P.paired ? s.h_out1 : nullptr;
  } else {
    res->out_r2 = s.out2;
    res->out_r1 = P.paired ? s.out1 : nullptr;
  }
  if (hs.n_unassigned) CU(cudaMemcpyAsync(s.h_unassigned, s.unassigned, (size_t)hs.n_unassigned * 4, cudaMemcpyDeviceToHost, s.stream));
  CU(cudaStreamSynchronize(s.stream));
  res->seq = s.seq;
  res->n_records = hs.n_records;
  res->consumed_r2 = hs.consumed2;
  res->consumed_r1 = hs.consumed1;
  res->out_r2_bytes = tot2;
  res->out_r1_bytes = tot1;
  res->seg_off_r2 = s.off2.data();
  res->seg_len_r2 = s.len2v.data();
  res->seg_off_r1 = s.off1.data();
  res->seg_len_r1 = s.len1v.data();
  res->unassigned = s.h_unassigned;
  res->n_unassigned = hs.n_unassigned;
  if (hs.validate_key != ~0ull) {
    res->validate_error = validate_kind((int)(hs.validate_key & 15u));
    res->validate_record = hs.validate_key >> 4;
  }
  return MGK_OK;
}

extern "C" int mgk_release(mgk_ctx* ctx, uint64_t seq) {
  if (!ctx) return MGK_ERR_ARG;
  for (auto& s : ctx->slots)
    if (s.state == 2 && s.seq == seq) {
      s.state = 0;
      return MGK_OK;
    }
  return fail(ctx, MGK_ERR_STATE, "mgk_release: chunk %llu is not waiting for release", (unsigned long long)seq);
}

// ====================================================================== counters
extern "C" int mgk_counters(mgk_ctx* ctx, uint64_t* stats, uint64_t* mism) {
  if (!ctx) return MGK_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  if (stats) CU(cudaMemcpy(stats, ctx->d_stats, (size_t)ctx->total * MGK_N_STATS * 8, cudaMemcpyDeviceToHost));
  if (mism) CU(cudaMemcpy(mism, ctx->d_mism, (size_t)ctx->total * ctx->mism_w * 8, cudaMemcpyDeviceToHost));
  return MGK_OK;
}

extern "C" int mgk_reset_counters(mgk_ctx* ctx) {
  if (!ctx) return MGK_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  ctx->reduced = false;
  CU(cudaMemset(ctx->d_stats, 0, (size_t)ctx->total * MGK_N_STATS * 8));
  CU(cudaMemset(ctx->d_mism, 0, (size_t)ctx->total * ctx->mism_w * 8));
  return MGK_OK;
}

#define NC(call)                                                                                              \
  do {                                                                                                        \
    ncclResult_t r_ = (call);                                                                                 \
    if (r_ != ncclSuccess) return fail(ctx, MGK_ERR_NCCL, "%s: %s", #call, n->GetErrorString ? n->GetErrorString(r_) : "nccl error"); \
  } while (0)

// The merge is out of place: the running tables of every context stay what they were (so more chunks can follow and
// the merge can be repeated), the sum over the ranks lands in the context's `reduced` tables (mgk_reduced_counters).
extern "C" int mgk_allreduce_counters(mgk_ctx* const* ctxs, int n_ctx) {
  if (!ctxs || n_ctx < 1 || !ctxs[0]) return MGK_ERR_ARG;
  mgk_ctx* ctx = ctxs[0];
  const size_t ns = (size_t)ctx->total * MGK_N_STATS, nm = (size_t)ctx->total * ctx->mism_w;
  for (int i = 0; i < n_ctx; ++i)
    if (!ctxs[i] || ctxs[i]->total != ctx->total || ctxs[i]->mism_w != ctx->mism_w) return fail(ctx, MGK_ERR_ARG, "contexts describe different runs");
  if (n_ctx == 1) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(ctx->d_red_stats, ctx->d_stats, ns * 8, cudaMemcpyDeviceToDevice));
    CU(cudaMemcpy(ctx->d_red_mism, ctx->d_mism, nm * 8, cudaMemcpyDeviceToDevice));
    ctx->reduced = true;
    return MGK_OK;
  }
  NcclApi* n = nccl_api();
  if (!n) return fail(ctx, MGK_ERR_NCCL, "libnccl.so.2 can not be loaded");
  std::vector<int> devs((size_t)n_ctx);
  for (int i = 0; i < n_ctx; ++i) devs[(size_t)i] = ctxs[i]->device;
  std::vector<ncclComm_t> comms((size_t)n_ctx, nullptr);
  NC(n->CommInitAll(comms.data(), n_ctx, devs.data()));
  auto drop_comms = [&] {
    for (int i = 0; i < n_ctx; ++i)
      if (comms[(size_t)i] && n->CommDestroy) {
        cudaSetDevice(ctxs[i]->device);
        n->CommDestroy(comms[(size_t)i]);
      }
  };
  for (int i = 0; i < n_ctx; ++i) {
    cudaSetDevice(ctxs[i]->device);
    cudaDeviceSynchronize();
  }
  ncclResult_t bad = ncclSuccess;
  ncclResult_t r = n->GroupStart();
  if (r != ncclSuccess) bad = r;
  for (int i = 0; i < n_ctx && bad == ncclSuccess; ++i) {
    r = n->AllReduce(ctxs[i]->d_stats, ctxs[i]->d_red_stats, ns, ncclUint64, ncclSum, comms[(size_t)i], 0);
    if (r != ncclSuccess) bad = r;
    r = n->AllReduce(ctxs[i]->d_mism, ctxs[i]->d_red_mism, nm, ncclUint64, ncclSum, comms[(size_t)i], 0);
    if (r != ncclSuccess) bad = r;
  }
  r = n->GroupEnd();
  if (r != ncclSuccess && bad == ncclSuccess) bad = r;
  cudaError_t ce = cudaSuccess;
  for (int i = 0; i < n_ctx; ++i) {
    cudaSetDevice(ctxs[i]->device);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) ce = e;
  }
  drop_comms();
  if (bad != ncclSuccess) return fail(ctx, MGK_ERR_NCCL, "counter all-reduce: %s", n->GetErrorString ? n->GetErrorString(bad) : "nccl error");
  if (ce != cudaSuccess) return fail(ctx, MGK_ERR_CUDA, "counter all-reduce: %s", cudaGetErrorString(ce));
  for (int i = 0; i < n_ctx; ++i) ctxs[i]->reduced = true;
  return MGK_OK;
}

extern "C" int mgk_reduced_counters(mgk_ctx* ctx, uint64_t* stats, uint64_t* mism) {
  if (!ctx) return MGK_ERR_ARG;
  if (!ctx->reduced) return fail(ctx, MGK_ERR_STATE, "mgk_reduced_counters: no counter all-reduce has run on this context");
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  if (stats) CU(cudaMemcpy(stats, ctx->d_red_stats, (size_t)ctx->total * MGK_N_STATS * 8, cudaMemcpyDeviceToHost));
  if (mism) CU(cudaMemcpy(mism, ctx->d_red_mism, (size_t)ctx->total * ctx->mism_w * 8, cudaMemcpyDeviceToHost));
  return MGK_OK;
}

extern "C" int mgk_nccl_unique_id(uint8_t id[MGK_NCCL_ID_BYTES]) {
  mgk_ctx* ctx = nullptr;
  NcclApi* n = nccl_api();
  if (!n) return fail(ctx, MGK_ERR_NCCL, "libnccl.so.2 can not be loaded");
  static_assert(sizeof(ncclUniqueId) <= MGK_NCCL_ID_BYTES, "nccl id size");
  ncclUniqueId uid;
  NC(n->GetUniqueId(&uid));
  memset(id, 0, MGK_NCCL_ID_BYTES);
  memcpy(id, &uid, sizeof uid);
  return MGK_OK;
}

extern "C" int mgk_nccl_join(mgk_ctx* ctx, const uint8_t id[MGK_NCCL_ID_BYTES], int rank, int nranks) {
  if (!ctx) return MGK_ERR_ARG;
  NcclApi* n = nccl_api();
  if (!n) return fail(ctx, MGK_ERR_NCCL, "libnccl.so.2 can not be loaded");
  CU(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  NC(n->CommInitRank(&ctx->comm, nranks, uid, rank));
  return MGK_OK;
}

extern "C" int mgk_nccl_allreduce_counters(mgk_ctx* ctx) {
  if (!ctx) return MGK_ERR_ARG;
  if (!ctx->comm) return fail(ctx, MGK_ERR_STATE, "mgk_nccl_allreduce_counters: call mgk_nccl_join first");
  NcclApi* n = nccl_api();
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  NC(n->GroupStart());
  const ncclResult_t r1 = n->AllReduce(ctx->d_stats, ctx->d_red_stats, (size_t)ctx->total * MGK_N_STATS, ncclUint64, ncclSum, ctx->comm, 0);
  const ncclResult_t r2 = n->AllReduce(ctx->d_mism, ctx->d_red_mism, (size_t)ctx->total * ctx->mism_w, ncclUint64, ncclSum, ctx->comm, 0);
  NC(n->GroupEnd());
  NC(r1);
  NC(r2);
  CU(cudaDeviceSynchronize());
  ctx->reduced = true;
  return MGK_OK;
}

// ====================================================================== misc
extern "C" void* mgk_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void mgk_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" int mgk_last_timings(const mgk_ctx* ctx, float ms[5]) {
  if (!ctx || !ms) return MGK_ERR_ARG;
  memcpy(ms, ctx->last_ms, sizeof ctx->last_ms);
  return MGK_OK;
}
extern "C" int mgk_last_gzip_ms(const mgk_ctx* ctx, float* ms) {
  if (!ctx || !ms) return MGK_ERR_ARG;
  *ms = ctx->last_gz_ms;
  return MGK_OK;
}
extern "C" int mgk_span_ms(mgk_ctx* ctx, float* ms) {
  if (!ctx || !ms) return MGK_ERR_ARG;
  *ms = 0.f;
  if (!ctx->span_open) return MGK_OK;
  CU(cudaSetDevice(ctx->device));
  // chunks in flight run on independent streams and may finish out of order: the span ends with the LAST of the
  // end-of-kernels events recorded since it was opened
  for (auto& s : ctx->slots) {
    if (!s.in_span) continue;
    float t = 0.f;
    CU(cudaEventSynchronize(s.ev[5]));
    CU(cudaEventElapsedTime(&t, ctx->span_begin, s.ev[5]));
    if (t > *ms) *ms = t;
    s.in_span = false;
  }
  ctx->span_open = false;
  return MGK_OK;
}
extern "C" uint64_t mgk_launch_count(const mgk_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ====================================================================== single-stage entry points
extern "C" int mgk_scan_newlines(mgk_ctx* ctx, const uint8_t* buf, uint64_t len, uint32_t* positions, uint64_t capacity, uint64_t* count) {
  if (!ctx || (!buf && len) || !count || len >= (1ull << 31)) return MGK_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  uint8_t* d_buf = nullptr;
  uint32_t *d_nl = nullptr, *d_ctrl = nullptr, *d_stg = nullptr;
  Status* d_status = nullptr;
  const uint32_t grid = ctx->scan_grid;
  const uint64_t ranges = (uint64_t)grid * SCAN_WARPS, tiles = (len + SCAN_WTILE - 1) / SCAN_WTILE, per = (tiles + ranges - 1) / ranges;
  const uint32_t cap = (uint32_t)std::min<uint64_t>(per * SCAN_WTILE + 4, 0xfffffff0u);   // a range cannot hold more newlines than bytes
  int rc = MGK_OK;
  std::vector<uint32_t> h_ctrl(8, 0);
  do {
#define TRY(call)                                                                  \
  if ((call) != cudaSuccess) {                                                     \
    rc = fail(ctx, MGK_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(cudaGetLastError())); \
    break;                                                                         \
  }
    TRY(dalloc(&d_buf, len + 256));
    TRY(dalloc(&d_nl, capacity + 4));
    TRY(dalloc(&d_stg, (size_t)ranges * cap));
    TRY(dalloc(&d_ctrl, 8 + (size_t)ranges));
    TRY(dalloc(&d_status, 1));
    TRY(cudaMemcpy(d_buf, buf, len, cudaMemcpyHostToDevice));
    TRY(cudaMemset(d_ctrl, 0, (8 + (size_t)ranges) * 4));
    TRY(cudaMemcpy(d_status, ctx->h_status_init, sizeof(Status), cudaMemcpyHostToDevice));
    ChunkDev c{};
    c.r2 = d_buf;
    c.len2 = (uint32_t)len;
    c.nl2 = d_nl;
    c.nl_cap = (uint32_t)std::min<uint64_t>(capacity, 0xffffffffu);
    c.stg2 = d_stg;
    c.scan_grid = grid;
    c.scan_cap = cap;
    c.counts = d_ctrl;
    c.range_counts = d_ctrl + 8;
    c.status = d_status;
    k_scan_newlines<<<dim3(grid, 1), SCAN_THREADS, SCAN_STAGES * SCAN_TILE>>>(c, 1);
    k_gather_newlines<<<dim3((uint32_t)ranges, 1), 256>>>(c, 1);
    ctx->launches += 2;
    TRY(cudaDeviceSynchronize());
    TRY(cudaMemcpy(h_ctrl.data(), d_ctrl, 32, cudaMemcpyDeviceToHost));
    *count = h_ctrl[0];
    const uint64_t ncopy = std::min<uint64_t>(*count, capacity);
    if (ncopy) TRY(cudaMemcpy(positions, d_nl, ncopy * 4, cudaMemcpyDeviceToHost));
    if (*count > capacity) rc = fail(ctx, MGK_ERR_CAPACITY, "mgk_scan_newlines: %llu newlines, capacity %llu", (unsigned long long)*count, (unsigned long long)capacity);
  } while (0);
  cudaFree(d_buf); cudaFree(d_nl); cudaFree(d_ctrl); cudaFree(d_stg); cudaFree(d_status);
  return rc;
}

extern "C" int mgk_match_barcodes(mgk_ctx* ctx, const uint8_t* barcodes, uint64_t n, int32_t* sample_id, int32_t* mismatches) {
  if (!ctx || (!barcodes && n) || n >= (1ull << 31)) return MGK_ERR_ARG;
  if (n == 0) return MGK_OK;
  CU(cudaSetDevice(ctx->device));
  uint8_t* d_bars = nullptr;
  int32_t *d_s = nullptr, *d_m = nullptr;
  uint32_t* d_err = nullptr;
  int rc = MGK_OK;
  uint32_t herr = 0;
  const size_t bytes = (size_t)n * ctx->hp.BL;
  do {
    TRY(dalloc(&d_bars, bytes + 64));
    TRY(dalloc(&d_s, (size_t)n));
    TRY(dalloc(&d_m, (size_t)n));
    TRY(dalloc(&d_err, 1));
    TRY(cudaMemcpy(d_bars, barcodes, bytes, cudaMemcpyHostToDevice));
    TRY(cudaMemset(d_err, 0, 4));
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->n_sm * 8);
    k_match_only<<<blocks, 256, ctx->match_smem>>>(ctx->d_params, d_bars, (uint32_t)n, d_s, d_m, d_err);
    ++ctx->launches;
    TRY(cudaDeviceSynchronize());
    TRY(cudaMemcpy(sample_id, d_s, (size_t)n * 4, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(mismatches, d_m, (size_t)n * 4, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(&herr, d_err, 4, cudaMemcpyDeviceToHost));
    if (herr & ERR_PANIC) rc = fail(ctx, MGK_ERR_FORMAT, "a barcode hits a sample-sheet combination on which the reference panics");
#undef TRY
  } while (0);
  cudaFree(d_bars); cudaFree(d_s); cudaFree(d_m); cudaFree(d_err);
  return rc;
}

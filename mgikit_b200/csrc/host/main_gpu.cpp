// `mgikit` — the product binary: host pipeline + the CUDA hot path.  There is
// no CPU fallback: without libmgikit_gpu.so and a B200 the command fails.
#include <unistd.h>

#include <cstdio>

#include "cli.hpp"

namespace {
int b_create(const mgk_config* c, void** o) { return mgk_create(c, (mgk_ctx**)o); }
void b_destroy(void* c) { mgk_destroy((mgk_ctx*)c); }
const char* b_err(const void* c) { return mgk_last_error((const mgk_ctx*)c); }
int b_submit(void* c, uint64_t s, const uint8_t* a, uint64_t al, const uint8_t* b, uint64_t bl, uint32_t f) {
  return mgk_submit((mgk_ctx*)c, s, a, al, b, bl, f);
}
int b_collect(void* c, mgk_result* r) { return mgk_collect((mgk_ctx*)c, r); }
int b_release(void* c, uint64_t s) { return mgk_release((mgk_ctx*)c, s); }
int b_counters(void* c, uint64_t* a, uint64_t* b) { return mgk_counters((mgk_ctx*)c, a, b); }
int b_allreduce(void* const* c, int n) { return mgk_allreduce_counters((mgk_ctx* const*)c, n); }
int b_reduced(void* c, uint64_t* a, uint64_t* b) { return mgk_reduced_counters((mgk_ctx*)c, a, b); }
}  // namespace

int main(int argc, char** argv) {
  const mgikit::Backend be = {"cuda-sm_100a", b_create, b_destroy, b_err, b_submit, b_collect,
                              b_release, b_counters, b_allreduce, b_reduced, mgk_host_alloc, mgk_host_free, MGK_OUT_GZIP, true};
  const int rc = mgikit::cli_main(argc, argv, be);
  // everything is on disk and every file is closed: leave without the driver's exit handlers (unmapping gigabytes of
  // pinned and device memory one allocation at a time takes half a second that a short lane would notice)
  std::fflush(stdout);
  std::fflush(stderr);
  _exit(rc);
}

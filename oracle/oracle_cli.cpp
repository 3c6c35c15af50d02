// TEST INFRASTRUCTURE ONLY: the product's host pipeline bound to the CPU oracle
// instead of the CUDA library, so the oracle (and the host code) can be pinned
// against the reference's fixtures and binary on a machine without a GPU.
// Never shipped, never used by the product path.
#include "../mgikit_b200/csrc/host/cli.hpp"
#include "mgikit_oracle.h"

namespace {
int b_create(const mgk_config* c, void** o) { return orc_create(c, (orc_ctx**)o); }
void b_destroy(void* c) { orc_destroy((orc_ctx*)c); }
const char* b_err(const void* c) { return orc_last_error((const orc_ctx*)c); }
int b_submit(void* c, uint64_t s, const uint8_t* a, uint64_t al, const uint8_t* b, uint64_t bl, uint32_t f) {
  return orc_submit((orc_ctx*)c, s, a, al, b, bl, f);
}
int b_collect(void* c, mgk_result* r) { return orc_collect((orc_ctx*)c, r); }
int b_release(void* c, uint64_t s) { return orc_release((orc_ctx*)c, s); }
int b_counters(void* c, uint64_t* a, uint64_t* b) { return orc_counters((orc_ctx*)c, a, b); }
int b_allreduce(void* const* c, int n) { return orc_allreduce_counters((orc_ctx* const*)c, n); }
int b_reduced(void* c, uint64_t* a, uint64_t* b) { return orc_reduced_counters((orc_ctx*)c, a, b); }
}  // namespace

int main(int argc, char** argv) {
  const mgikit::Backend be = {"cpu-oracle", b_create, b_destroy, b_err, b_submit, b_collect,
                              b_release, b_counters, b_allreduce, b_reduced, orc_host_alloc, orc_host_free, 0u, false};
  return mgikit::cli_main(argc, argv, be);
}

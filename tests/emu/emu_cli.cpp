// TEST INFRASTRUCTURE ONLY: host pipeline bound to the CPU emulation of the
// device logic (tests/emu/emu.cpp), to run the reference's fixture matrix
// through the exact per-record code the GPU kernels execute.
#include <cstdlib>

#include "../../mgikit_b200/csrc/host/cli.hpp"

struct emu_ctx;
extern "C" {
int emu_create(const mgk_config*, emu_ctx**);
void emu_destroy(emu_ctx*);
const char* emu_last_error(const emu_ctx*);
int emu_submit(emu_ctx*, uint64_t, const uint8_t*, uint64_t, const uint8_t*, uint64_t, uint32_t);
int emu_collect(emu_ctx*, mgk_result*);
int emu_release(emu_ctx*, uint64_t);
int emu_counters(emu_ctx*, uint64_t*, uint64_t*);
}
namespace {
int b_create(const mgk_config* c, void** o) { return emu_create(c, (emu_ctx**)o); }
void b_destroy(void* c) { emu_destroy((emu_ctx*)c); }
const char* b_err(const void* c) { return emu_last_error((const emu_ctx*)c); }
int b_submit(void* c, uint64_t s, const uint8_t* a, uint64_t al, const uint8_t* b, uint64_t bl, uint32_t f) {
  return emu_submit((emu_ctx*)c, s, a, al, b, bl, f);
}
int b_collect(void* c, mgk_result* r) { return emu_collect((emu_ctx*)c, r); }
int b_release(void* c, uint64_t s) { return emu_release((emu_ctx*)c, s); }
int b_counters(void* c, uint64_t* a, uint64_t* b) { return emu_counters((emu_ctx*)c, a, b); }
int b_allreduce(void* const*, int) { return MGK_ERR_ARG; }
int b_reduced(void*, uint64_t*, uint64_t*) { return MGK_ERR_ARG; }
void* b_alloc(size_t n) { return malloc(n); }
void b_free(void* p) { free(p); }
}  // namespace

int main(int argc, char** argv) {
  const mgikit::Backend be = {"cpu-emulation-of-device-logic", b_create, b_destroy, b_err, b_submit, b_collect,
                              b_release, b_counters, b_allreduce, b_reduced, b_alloc, b_free, 0u, false};
  return mgikit::cli_main(argc, argv, be);
}

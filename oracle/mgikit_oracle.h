/*
 * mgikit_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of mgikit's demultiplex hot path, used as the
 * checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product
 * (mgikit_b200/, libmgikit_gpu.so, the `mgikit` CLI) never links or calls it.
 *
 * The entry points mirror include/mgikit_gpu.h one to one (orc_ instead of
 * mgk_) and take the very same mgk_config / mgk_result structs, so a parity
 * test is "run both, compare the structs field by field".
 *
 * Parity pinning: oracle/README.md — checked against every fixture under the
 * reference's testing_data/expected that belongs to `demultiplex`, and against
 * the reference's own prebuilt binary (oracle/_ref/mgikit) on seeded lanes.
 */
#ifndef MGIKIT_ORACLE_H
#define MGIKIT_ORACLE_H

#include "../include/mgikit_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

int orc_create(const mgk_config* cfg, orc_ctx** out);
void orc_destroy(orc_ctx* ctx);
const char* orc_last_error(const orc_ctx* ctx);
int orc_submit(orc_ctx* ctx, uint64_t seq, const uint8_t* r2, uint64_t r2_len,
               const uint8_t* r1, uint64_t r1_len, uint32_t flags);
int orc_collect(orc_ctx* ctx, mgk_result* res);
int orc_release(orc_ctx* ctx, uint64_t seq);
int orc_counters(orc_ctx* ctx, uint64_t* stats, uint64_t* mism);
int orc_reset_counters(orc_ctx* ctx);
int orc_allreduce_counters(orc_ctx* const* ctxs, int n);
int orc_reduced_counters(orc_ctx* ctx, uint64_t* stats, uint64_t* mism);
void* orc_host_alloc(size_t bytes);
void orc_host_free(void* p);
int orc_scan_newlines(orc_ctx* ctx, const uint8_t* buf, uint64_t len,
                      uint32_t* positions, uint64_t capacity, uint64_t* count);
int orc_match_barcodes(orc_ctx* ctx, const uint8_t* barcodes, uint64_t n,
                       int32_t* sample_id, int32_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif

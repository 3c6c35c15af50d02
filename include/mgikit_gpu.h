/*
 * mgikit_gpu.h — C ABI of the B200 demultiplex hot path.
 *
 * This is the drop-in boundary for ONE call of the reference:
 *     process_buffer()                     /root/reference/src/lib.rs:840-868 (signature)
 *     its only call site                   src/lib.rs:1496-1536 (reached from demultiplex, :511-838, and,
 *                                          with demultiplex == false, from reformat, :2266-2464)
 *     per-worker state it mutates          Vec<SampleData> (src/sample_data.rs:16-23)
 *                                          ReportManager   (src/report_manager.rs:254-261)
 * The reference has no FFI of its own (pure Rust crate); a Rust host would bind
 * these symbols from a `-sys` crate (see INTEGRATION.md for the stub).
 *
 * Vocabulary follows the reference: "barcode read" = the read that carries the
 * barcode (R2 when paired, R1 when single-end), "paired read" = R1 of a pair,
 * S = number of sample-sheet rows, sample ids S and S+1 are Undetermined and
 * Ambiguous (src/sample_manager.rs:83-104), BL = barcode length (template[9]).
 *
 * Contract
 *  - plain C, no exceptions or aborts cross the boundary; every call returns
 *    MGK_OK (0) or a negative mgk_status, text via mgk_last_error().
 *  - a context belongs to one CUDA device and is NOT thread-safe; different
 *    contexts are independent.
 *  - there is no CPU fallback: mgk_create() fails with MGK_ERR_NO_DEVICE when
 *    no sm_100 device is usable.
 *  - the library never frees caller memory; result buffers belong to the
 *    context until mgk_release().
 */
#ifndef MGIKIT_GPU_H
#define MGIKIT_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGK_ABI_VERSION 2
#define MGK_MAX_TEMPLATES 16
#define MGK_MAX_INDEX_LEN 32      /* bases per i7 / i5 (2-bit packed into 64 bit) */
#define MGK_MAX_BARCODE_LEN 255
#define MGK_N_STATS 10            /* ReportManager::sample_statistics row, src/report_manager.rs:269 */

typedef enum {
  MGK_OK = 0,
  MGK_ERR_ARG = -1,          /* bad argument / unsupported configuration */
  MGK_ERR_NO_DEVICE = -2,    /* no usable CUDA device (no CPU fallback exists) */
  MGK_ERR_CUDA = -3,         /* CUDA runtime error, see mgk_last_error */
  MGK_ERR_CAPACITY = -4,     /* chunk larger than the capacity given at create */
  MGK_ERR_STATE = -5,        /* call order violated (collect without submit, ...) */
  MGK_ERR_FORMAT = -6,       /* chunk is not parseable as 4-line records */
  MGK_ERR_NCCL = -7
} mgk_status;

/* output header style: src/lib.rs:1240-1319 */
enum { MGK_OUT_ILLUMINA = 0, MGK_OUT_MGI = 1, MGK_OUT_MGI_FULL_HEADER = 2 };

/* One barcode template = one entry of SampleManager::all_template_data
 * (src/sample_manager.rs:24-34), already reverse-complemented
 * (src/sample_manager.rs:573-594) and de-duplicated.                      */
typedef struct {
  int32_t geometry[10];        /* parse_template(), src/sample_manager.rs:249-288:
                                  [has_i7,len_i7,off_i7, has_i5,len_i5,off_i5,
                                   has_umi,len_umi,off_umi, BL]; offsets count
                                  back from the END of the barcode            */
  int32_t has_i5;              /* tuple field .5 */
  int32_t n_i7;                /* unique i7 strings of the template (field .1) */
  const char* i7;              /* n_i7 * geometry[1] bytes, [ACGT] */
  int32_t n_i5;                /* unique i5 strings (field .2); 0 if !has_i5 */
  const char* i5;              /* n_i5 * geometry[4] bytes */
  const int32_t* pair_sample;  /* field .4: has_i5 ? [n_i7*n_i5] : [n_i7];
                                  sample index or -1 when the pair is no sample */
} mgk_template;

/* Immutable description of a run; deep-copied by mgk_create(). */
typedef struct {
  int32_t abi_version;             /* MGK_ABI_VERSION */
  int32_t device;                  /* CUDA ordinal */
  int32_t paired;                  /* RunManager::paired_read_input(), src/run_manager.rs:192 */
  int32_t read2_has_sequence;      /* src/lib.rs:1724-1727: 0 => samples get no barcode-read file */
  int32_t out_mode;                /* MGK_OUT_* */
  int32_t keep_barcode;            /* src/lib.rs:883-886 */
  int32_t comprehensive_scan;      /* src/lib.rs:433-435 */
  int32_t validate;                /* --validate, check_read_content src/lib.rs:476-509 */
  int32_t report_level;            /* 0: no QC sums (src/lib.rs:1188); >1: barcode histograms (:1082) */
  int32_t mgi_data;                /* !--not-mgi; header equality check src/lib.rs:1163 */
  int32_t allowed_mismatches;      /* m */
  int32_t total_allowed_mismatches;/* m, or 2m with --per-index-error (src/lib.rs:1500-1504) */
  int32_t l_position;              /* src/run_manager.rs:400-418 */
  const char* illumina_prefix;     /* "@instr:run:flowcell:", src/run_manager.rs:166-186 (may be NULL unless illumina) */
  int32_t n_samples;               /* S, without the two pseudo samples */
  const int32_t* writing_sample;   /* [S+2], src/sample_manager.rs:692-730 */
  int32_t n_templates;
  const mgk_template* templates;   /* in matching order (src/sample_manager.rs:673-687) */
  uint64_t max_chunk_bytes;        /* capacity per read file per chunk (< 2^31) */
  uint64_t max_chunk_records;      /* capacity in records per chunk */
  int32_t n_slots;                 /* chunks that may be in flight (>=1; 2 = double buffering) */
  /* `mgikit reformat` (src/lib.rs:2266-2464): the same call with demultiplex == false (src/lib.rs:1090-1147,
   * :1240-1293).  Every read belongs to sample 0 (n_samples = 1, n_templates = 0, allowed_mismatches = 5 as
   * at :2431); the barcode is the one already in the header (" n:N:0:BARCODE", or none: MGI raw "/n").     */
  int32_t reformat;                /* 0: demultiplex */
  int32_t reformat_umi_length;     /* bases at the end of the barcode read's sequence that move into the header
                                      (:1117-1124); it is the barcode_length of the trim and of the r3 QC
                                      sums (:878-882) */
  const char* reformat_barcode;    /* --barcode ("" / NULL: none): mismatches are counted against the header's
                                      barcode (:1103-1113, more than 4 count as 4); headers without one get
                                      " n:N:0:" + this (:1256-1264) */
} mgk_config;

/* mgk_submit flags */
#define MGK_IN_DEVICE   1u   /* r2/r1 are device pointers (already resident in HBM): 16-byte aligned, and
                                the caller's allocation must stay READABLE for MGK_DEVICE_INPUT_SLACK
                                bytes past r*_len -- the kernels fetch whole 16-byte units (bulk copies
                                round a tail up, index probes load aligned words); the content of the
                                slack is never used.  Checked against the allocation's range where the
                                driver reports it (MGK_ERR_ARG).  Host buffers need no slack: the
                                library's own device copies are padded.                              */
#define MGK_DEVICE_INPUT_SLACK 32
#define MGK_OUT_DEVICE  2u   /* leave out_r2/out_r1 in HBM (result pointers are device pointers) */
#define MGK_OUT_GZIP    4u   /* out_r2/out_r1 hold, per bin, the segment as a run of gzip members made on
                                the device (BGZF-framed, 32 KiB of text each) instead of the text itself:
                                what compress_buffer() would hand to write_data() (src/sample_data.rs:
                                472-503).  seg_off/seg_len and out_*_bytes then describe the compressed
                                runs; gunzip of bin b's run == the text segment.  The .gz bytes are not
                                the reference's (libdeflate level 1); its tests compare decompressed text. */

typedef struct mgk_ctx mgk_ctx;

/* Result of one chunk.  Segment b (b = writing sample, 0..S+1) of the barcode
 * read is out_r2[seg_off_r2[b] .. +seg_len_r2[b]]: every record routed to that
 * sample, rewritten exactly as SampleData would have buffered it
 * (src/lib.rs:1224-1337), in input order.                                    */
typedef struct {
  uint64_t seq;                    /* as given to mgk_submit */
  uint64_t n_records;              /* whole records found in BOTH reads */
  uint64_t consumed_r2;            /* bytes of whole records consumed == process_buffer's */
  uint64_t consumed_r1;            /*   return value (header_start, header_start_pr)      */
  const uint8_t* out_r2;           /* barcode-read output text */
  const uint8_t* out_r1;           /* paired-read output text (NULL if single-end) */
  uint64_t out_r2_bytes, out_r1_bytes;
  const uint64_t* seg_off_r2;      /* [S+2] host memory */
  const uint64_t* seg_len_r2;
  const uint64_t* seg_off_r1;
  const uint64_t* seg_len_r1;
  /* reads that went to Undetermined / Ambiguous (src/lib.rs:1024-1088): byte
   * offset, inside the submitted barcode-read chunk, of the first barcode base
   * (plus_start-1-BL); bit 31 set => Ambiguous.  Host memory, unordered.      */
  const uint32_t* unassigned;
  uint64_t n_unassigned;
  /* --validate (src/lib.rs:1154-1186, :462-466): 0 = clean, else MGK_VAL_* of
   * the earliest offending record and its index in the chunk.                 */
  int32_t validate_error;
  uint64_t validate_record;
} mgk_result;

enum {
  MGK_VAL_OK = 0,
  MGK_VAL_HEADER_AT = 1,       /* header does not start with '@'  (src/lib.rs:477) */
  MGK_VAL_BASE = 2,            /* base outside ACGTNacgtn          (src/lib.rs:483-499) */
  MGK_VAL_QUALITY = 3,         /* quality outside [33,73]          (src/lib.rs:501-507, :462-466) */
  MGK_VAL_HEADER_MISMATCH = 4  /* R1/R2 headers differ (MGI)       (src/lib.rs:1163-1178) */
};

int mgk_create(const mgk_config* cfg, mgk_ctx** out);
void mgk_destroy(mgk_ctx* ctx);
const char* mgk_last_error(const mgk_ctx* ctx);   /* ctx may be NULL: last create error */

/* Queue one chunk pair.  r2 = barcode read bytes, r1 = paired read bytes (NULL,0
 * when single-end).  Chunks should hold whole records; a trailing partial record
 * is left unconsumed exactly like the reference (src/lib.rs:991-1002).
 * Host pointers should be pinned for the copy to overlap compute. Asynchronous. */
int mgk_submit(mgk_ctx* ctx, uint64_t seq,
               const uint8_t* r2, uint64_t r2_len,
               const uint8_t* r1, uint64_t r1_len, uint32_t flags);
/* Block for the OLDEST chunk in flight (FIFO => input order is preserved). */
int mgk_collect(mgk_ctx* ctx, mgk_result* res);
/* Give the slot of chunk `seq` back. */
int mgk_release(mgk_ctx* ctx, uint64_t seq);

/* Running totals over every chunk collected so far:
 * stats = ReportManager::sample_statistics  [S+2][10]   (src/report_manager.rs:297-305)
 * mism  = ReportManager::sample_mismatches  [S+2][2m+2] (src/report_manager.rs:319-329),
 *         slot 0 left 0 (set_sample_total_reads is the host's job, :367-375). */
int mgk_counters(mgk_ctx* ctx, uint64_t* stats, uint64_t* mism);
int mgk_reset_counters(mgk_ctx* ctx);

/* Multi-GPU merge of the counter tables == ReportManager::update
 * (src/report_manager.rs:377-416) across ranks, as one NCCL sum over NVLink.
 * (a) one process, n contexts on n devices: */
int mgk_allreduce_counters(mgk_ctx* const* ctxs, int n);
/* Both variants are out of place: the running tables (mgk_counters) keep this
 * context's own share, so more chunks may follow and the merge may be repeated;
 * the sum over all ranks is read here (same layout; MGK_ERR_STATE before the
 * first merge).                                                               */
int mgk_reduced_counters(mgk_ctx* ctx, uint64_t* stats, uint64_t* mism);
/* (b) one process per GPU (torchrun): rank 0 makes the id, every rank joins. */
#define MGK_NCCL_ID_BYTES 128
int mgk_nccl_unique_id(uint8_t id[MGK_NCCL_ID_BYTES]);
int mgk_nccl_join(mgk_ctx* ctx, const uint8_t id[MGK_NCCL_ID_BYTES], int rank, int nranks);
int mgk_nccl_allreduce_counters(mgk_ctx* ctx);

/* Page-locked host memory for chunk buffers (cudaHostAlloc / cudaFreeHost). */
void* mgk_host_alloc(size_t bytes);
void mgk_host_free(void* p);

/* Device time (ms, CUDA events on the context's stream) of the kernels of the
 * most recently collected chunk: [0] record scan, [1] match+lengths,
 * [2] bin scan, [3] scatter(+QC), [4] whole chunk incl. copies.               */
int mgk_last_timings(const mgk_ctx* ctx, float ms[5]);
/* Device time (ms) of the gzip kernels of the most recently collected chunk (0 without MGK_OUT_GZIP). */
int mgk_last_gzip_ms(const mgk_ctx* ctx, float* ms);
/* Device time (ms) from the start of the first chunk submitted after the last
 * call (or after create) to the end of the last chunk submitted: one pair of
 * CUDA events on the launching streams, so overlapped chunks are not counted
 * twice.  Waits for that last chunk.                                          */
int mgk_span_ms(mgk_ctx* ctx, float* ms);
/* Number of kernel launches issued by this context so far. */
uint64_t mgk_launch_count(const mgk_ctx* ctx);

/* Standalone pieces of the path, for parity tests of a single stage.
 * All pointers are HOST pointers; the call copies in, runs the kernel on the
 * context's device and copies out.                                            */
/* a1: positions of every '\n' (src/file_utils.rs:95, src/lib.rs:1526). */
int mgk_scan_newlines(mgk_ctx* ctx, const uint8_t* buf, uint64_t len,
                      uint32_t* positions, uint64_t capacity, uint64_t* count);
/* a6: find_matching_sample() (src/lib.rs:223-443) on n barcodes of BL bytes
 * each (stride BL); sample_id[i] in [0,S+1], mismatches[i] as returned.       */
int mgk_match_barcodes(mgk_ctx* ctx, const uint8_t* barcodes, uint64_t n,
                       int32_t* sample_id, int32_t* mismatches);

#ifdef __cplusplus
}
#endif
#endif /* MGIKIT_GPU_H */

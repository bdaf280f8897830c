/*
 * dbt_b200.h -- the thin extern "C" layer of the B200-native tuple operators.
 *
 * This is the drop-in boundary beneath the four dbtproj.h entry points (include/dbtproj.h):
 * plain pointers and sizes, int status, no C++ or torch types, no exceptions, and no hidden
 * device allocation on the device-scope path (the caller owns every buffer, including the
 * workspace).  Each group cites the part of the reference it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative dbt_status on failure;
 *     dbt_last_error() gives a human-readable message for the calling thread;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - "image" = a block file in memory: a flat array of 14016-byte block_t (dbtproj.h);
 *   - `field` is the ASCII selector '0'..'3' of the reference (DatabaseProject.cpp:23-40);
 *   - device-scope operators (dbt_dev_*) take device pointers, enqueue on `stream` and
 *     synchronise that stream before returning (they hand row counts back to the host);
 *   - host-scope operators (dbt_host_*) take host pointers and do the host<->device copies
 *     themselves (pinned staging when the caller's memory is pageable);
 *   - device images, column buffers and workspaces must be 16-byte aligned (anything from cudaMalloc is; a block
 *     offset inside an image keeps it, 14016 = 876 x 16); misaligned pointers fail with DBT_ERR_ARG.  Only
 *     dbt_sort_pairs_u32 accepts 4-byte aligned buffers (it then runs its non-TMA kernel);
 *   - there is NO CPU fallback anywhere: without a CUDA device every entry fails with
 *     DBT_ERR_CUDA.
 */
#ifndef DBT_B200_H
#define DBT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBT_BLOCK_BYTES 14016u
#define DBT_RECORD_BYTES 140u
#define DBT_RECORDS_PER_BLOCK 100u

typedef enum dbt_status {
    DBT_OK = 0,
    DBT_ERR_ARG = -1,       /* bad argument (field, sizes, NULL pointers) */
    DBT_ERR_CUDA = -2,      /* CUDA runtime error, incl. "no device" */
    DBT_ERR_WORKSPACE = -3, /* workspace / output capacity too small */
    DBT_ERR_IO = -4,        /* file open/read/write failure */
    DBT_ERR_UNSUPPORTED = -5,
    DBT_ERR_NEED_WIDE_KEYS = -6, /* a string has no NUL in its first 32 bytes and the workspace was sized for 8-word
                                    keys: call again with a workspace from dbt_dev_ws_bytes_kw(..., 30) */
    DBT_ERR_TIMEOUT = -7    /* a multi-GPU rendezvous or peer signal did not arrive in time */
} dbt_status;

const char *dbt_last_error(void);
int dbt_abi_version(void);
/* number of visible CUDA devices (0 => every other call fails loudly) */
int dbt_device_count(void);

/* ----------------------------------------------------------------------------------------------
 * Counters of the external merge sort the reference would have run.
 * Replaces the bookkeeping in reference DatabaseProject.cpp:216,227,298,334,346,364-376 and
 * :522,565,579,639 (HashJoin) / :395,405,441,465,475,491 (MergeJoin).  Pure host arithmetic.
 * ---------------------------------------------------------------------------------------------- */
int dbt_sort_counters(uint64_t nblocks, uint32_t nmem_blocks, uint64_t *nsorted_segs, uint64_t *npasses,
                      uint64_t *nios);
uint64_t dbt_dedup_nios(uint64_t nblocks, uint32_t nmem_blocks, uint64_t nunique);
uint64_t dbt_hashjoin_nios(uint64_t nblocks_r, uint64_t nblocks_s, uint32_t nmem_blocks, uint64_t nres);
/* res = the 4 values dbt_dev_mergejoin returns (nres, nunique_R, nunique_S, later block reads) */
uint64_t dbt_mergejoin_nios(uint64_t nblocks_r, uint64_t nblocks_s, uint32_t nmem_blocks, const uint64_t *res);

/* ----------------------------------------------------------------------------------------------
 * Building-block kernels (device scope).
 * ---------------------------------------------------------------------------------------------- */

/* LSD onesweep radix sort of (u32 key, u32 value) pairs over key bits [begin_bit, end_bit).
 * Replaces qsort run generation + priority_queue k-way merge (DatabaseProject.cpp:207-214,
 * 255-343) at the pair level.  keys/vals are double buffers of n elements each; on return
 * *result_in_alt is 1 when the sorted data sits in the *_alt buffers.  Stable.
 * Digit positions where all keys agree are skipped.  n <= 2^32 - 2^16 (32-bit row ids); from 2^30 elements on the
 * look-back chain runs on 64-bit tile states. */
size_t dbt_sort_pairs_ws_bytes(uint64_t n);
int dbt_sort_pairs_u32(uint32_t *d_keys, uint32_t *d_keys_alt, uint32_t *d_vals, uint32_t *d_vals_alt, uint64_t n,
                       int begin_bit, int end_bit, void *d_ws, size_t ws_bytes, void *stream, int *result_in_alt);

/* Record gather + block packer: out image block k receives rows d_rows[100k .. 100k+99] of the
 * input image (row = index into the live rows of the input in file order), with CANON headers
 * (blockid=k, nreserved, valid=1, misc=0, dummy=nreserved; unused slots zero).
 * Replaces the 140-byte memcpy per record per pass (DatabaseProject.cpp:202,223,303,338).
 * d_row_slot may be NULL when every input block except the last is full (slot == row). */
int dbt_gather_records(const void *d_in_image, const uint32_t *d_rows, const uint32_t *d_row_slot, uint64_t nrows_out,
                       void *d_out_image, void *stream);
/* ----------------------------------------------------------------------------------------------
 * Device-scope operators: image in HBM -> image in HBM.  `d_out` must hold
 * ceil(rows/100) blocks where rows is the worst case for the operator (input rows for sort /
 * dedup, min side for mergejoin, dbt_dev_hashjoin's own capacity argument).
 * ---------------------------------------------------------------------------------------------- */
typedef enum dbt_op { DBT_OP_SORT = 0, DBT_OP_DEDUP = 1, DBT_OP_MERGEJOIN = 2, DBT_OP_HASHJOIN = 3 } dbt_op;

/* upper bound of the workspace an operator needs for inputs of these sizes (blocks) */
size_t dbt_dev_ws_bytes(int op, uint64_t nblocks_r, uint64_t nblocks_s, int field);
/* same, for string keys of `kw` 32-bit words: 8 (strings shorter than 32 bytes, the default) or 30
 * (full 120-byte keys; an operator that meets a longer string with the small workspace fails with
 * DBT_ERR_NEED_WIDE_KEYS) */
size_t dbt_dev_ws_bytes_kw(int op, uint64_t nblocks_r, uint64_t nblocks_s, int field, uint32_t kw);
/* HashJoin whose output may be larger than S (field '3': an S row is emitted once per matching R row,
 * DatabaseProject.cpp:616-629): the workspace for an output capacity of out_capacity_blocks blocks.  The plain
 * bound above covers capacities up to nblocks_s. */
size_t dbt_dev_hashjoin_ws_bytes(uint64_t nblocks_r, uint64_t nblocks_s, int field, uint32_t kw,
                                 uint64_t out_capacity_blocks);

/* MergeSort (DatabaseProject.cpp:172-381): every live row once, ordered by (key(field), recid). */
int dbt_dev_mergesort(const void *d_in, uint64_t nblocks, int field, void *d_out, void *d_ws, size_t ws_bytes,
                      void *stream, uint64_t *nrows);

/* EliminateDuplicates (DatabaseProject.cpp:94-170): the min-recid row of every distinct key,
 * ordered by key. */
int dbt_dev_dedup(const void *d_in, uint64_t nblocks, int field, void *d_out, void *d_ws, size_t ws_bytes,
                  void *stream, uint64_t *nrows, uint64_t *nunique);

/* MergeJoin (DatabaseProject.cpp:384-502): dedup(R), dedup(S) into d_out_ur / d_out_us (the
 * "1outfile.bin"/"2outfile.bin" side files), then R's row for every key present in both.
 * res[0]=nres res[1]=nunique_R res[2]=nunique_S res[3]=later block reads of the reference's
 * two-pointer walk (for nios). */
int dbt_dev_mergejoin(const void *d_in_r, uint64_t nblocks_r, const void *d_in_s, uint64_t nblocks_s, int field,
                      void *d_out_ur, void *d_out_us, void *d_out, void *d_ws, size_t ws_bytes, void *stream,
                      uint64_t *res);

/* HashJoin (DatabaseProject.cpp:504-647): S's rows, in S file order, whose key is in keys(R);
 * field '3': once per matching R row.  out_capacity_blocks bounds d_out; when the result needs
 * more, the call fails with DBT_ERR_WORKSPACE and *nres holds the required row count. */
int dbt_dev_hashjoin(const void *d_in_r, uint64_t nblocks_r, const void *d_in_s, uint64_t nblocks_s, int field,
                     void *d_out, uint64_t out_capacity_blocks, void *d_ws, size_t ws_bytes, void *stream,
                     uint64_t *nres);

/* HashJoin with the build side given as a key column instead of an image (fields '0' and '1'): the S rows, in S
 * file order, whose key is in d_rkeys[0..nr).  This is what a multi-GPU semi-join wants: the ranks all-gather R's
 * KEYS (4 bytes each) and probe their own shard of S in place -- no S record crosses the fabric, key skew cannot
 * unbalance anything, and the concatenation of the ranks' outputs is in S file order like the reference's. */
int dbt_dev_semijoin_keys(const uint32_t *d_rkeys, uint64_t nr, const void *d_in_s, uint64_t nblocks_s, int field,
                          void *d_out, uint64_t out_capacity_blocks, void *d_ws, size_t ws_bytes, void *stream,
                          uint64_t *nres);

/* Pair-producing inner join -- an EXTENSION (north_star compares joins as (recid_R, recid_S) multisets; the
 * reference itself only emits records).  d_pairs[2k] = recid of the R row, d_pairs[2k+1] = recid of
 * the S row, for every pair of rows with equal key(field); S file order, then R rows by (key, recid).
 * pairs_capacity bounds d_pairs (in pairs); *npairs is always the true count (DBT_ERR_WORKSPACE if it did not fit). */
int dbt_dev_innerjoin_pairs(const void *d_in_r, uint64_t nblocks_r, const void *d_in_s, uint64_t nblocks_s, int field,
                            uint32_t *d_pairs, uint64_t pairs_capacity, void *d_ws, size_t ws_bytes, void *stream,
                            uint64_t *npairs);

/* ----------------------------------------------------------------------------------------------
 * Multi-GPU building blocks, used by the C++ layer below and usable on their own: the routing word of
 * every row (the key's most significant word: recid, num, or the first four str bytes -- equal keys share it, so
 * every field shards on it) and the rows grouped by destination.
 * ---------------------------------------------------------------------------------------------- */
/* routing word of every live row in file order (recid | num | first 4 bytes of str, NUL-normalised, big-endian);
 * *nrows receives the count */
int dbt_dev_extract_keys_u32(const void *d_in, uint64_t nblocks, int field, uint32_t *d_keys, void *d_ws,
                             size_t ws_bytes, void *stream, uint64_t *nrows);
/* Destination of every row and the rows grouped by destination (stable: file order inside a group).
 * mode 0: dest = number of splitters <= key   (h_splitters: nparts-1 ascending keys, host memory)
 * mode 1: dest = mixhash(key) mod nparts      (h_splitters ignored)
 * d_rows_grouped[n]: row ids, group 0 first; h_counts[nparts]: group sizes (host memory). */
int dbt_dev_partition_rows(const uint32_t *d_keys, uint64_t n, int mode, const uint32_t *h_splitters, uint32_t nparts,
                           uint32_t *d_rows_grouped, uint64_t *h_counts, void *d_ws, size_t ws_bytes, void *stream);
size_t dbt_dev_partition_ws_bytes(uint64_t nblocks);

/* ----------------------------------------------------------------------------------------------
 * Multi-GPU operators (C++ host layer, csrc/dist.cu).  One RANK per GPU of one box.  Ranks are
 * processes (dbt_dist_init: they rendezvous through a POSIX shared-memory control block named after `session` and
 * map each other's staging buffers with CUDA IPC) or threads of one process (dbt_dist_init_local: peer access, no
 * IPC; this is how the file entry points use every visible GPU).  Every operator is COLLECTIVE: all ranks of the
 * group call it, each with its own shard (device pointers on its own GPU, any cudaMalloc'ed memory).  Records and
 * key columns cross NVLink as stores of the library's own kernels into the owner's staging buffer (gather + exchange
 * in one kernel); completion travels as stream-ordered flags in peer memory; no NCCL, no torch on the data path.
 *   sort / dedup   shard by key RANGE: splitters from a global sample cut the key space into P x Q sub-ranges (split
 *                  on the key only, so equal keys meet); rows are pushed sub-range by sub-range and the owner runs
 *                  the single-GPU operator on sub-range q while q+1.. are still on the wire.  The concatenation of
 *                  the ranks' outputs, in rank order, is the globally sorted (duplicate-free) file
 *                  (DatabaseProject.cpp:172-381, :94-170 across P GPUs).
 *   hashjoin       fields '0'/'1': the build side's KEYS are replicated (4 bytes per R row) and every rank probes its
 *                  own S shard in place (fused streaming semi-join): no S record moves, key skew cannot unbalance the
 *                  ranks, the outputs concatenate in S file order like the reference's.  Fields '2'/'3': both
 *                  relations are hash-partitioned on the key's first word, then the ordinary operator runs per rank.
 *   mergejoin      both relations are range-partitioned with the SAME splitters, then dbt_dev_mergejoin per rank; the
 *                  ranks' outputs concatenate in ascending key order (res[] are this rank's counts).
 * Replaces nothing in the reference (it is single-threaded); boundary: dbtproj.h:55-96.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dbt_dist dbt_dist;
int dbt_dist_init(const char *session, int rank, int world, int device, dbt_dist **out);
int dbt_dist_init_local(int world, const int *devices, dbt_dist **out /*[world]*/);
int dbt_dist_destroy(dbt_dist *d);
int dbt_dist_rank(const dbt_dist *d);
int dbt_dist_world(const dbt_dist *d);
int dbt_dist_barrier(dbt_dist *d); /* host barrier over the control block */
int dbt_dist_trim(dbt_dist *d);    /* collective: give the staging / send / list / workspace buffers back */
/* key sub-ranges per owner for sort/dedup (the pipeline depth); 0 = automatic: 4, or 1 for small shards */
int dbt_dist_set_sub_ranges(dbt_dist *d, uint32_t q);
/* every rank contributes `bytes` (<= 32 KB) of host memory; out receives world * bytes in rank order */
int dbt_dist_allgather_host(dbt_dist *d, const void *mine, size_t bytes, void *out);
/* d_out must hold out_capacity_blocks blocks (a rank receives about 1/P of the rows; splitters come from a sample, so
 * leave headroom); *out_rows = rows in this rank's output, *rows_received = rows this rank owned before dedup */
int dbt_dist_sort(dbt_dist *d, const void *d_in, uint64_t nblocks, int field, int dedup, void *d_out,
                  uint64_t out_capacity_blocks, void *stream, uint64_t *out_rows, uint64_t *rows_received);
int dbt_dist_hashjoin(dbt_dist *d, const void *d_r, uint64_t nblocks_r, const void *d_s, uint64_t nblocks_s, int field,
                      void *d_out, uint64_t out_capacity_blocks, void *stream, uint64_t *nres);
int dbt_dist_mergejoin(dbt_dist *d, const void *d_r, uint64_t nblocks_r, const void *d_s, uint64_t nblocks_s, int field,
                       void *d_out, uint64_t out_capacity_blocks, void *stream, uint64_t *res /*[4]*/);
/* last operator on this rank: [0] ms of the NVLink phase (first push .. last flag), [1] bytes stored into other GPUs,
 * [2] bytes stored in total (incl. own staging), [3] sub-ranges used; sort/dedup host timeline in ms since the call:
 * [4] routing keys extracted, [5] splitters agreed, [6] rows grouped and pushes enqueued, [7] first sub-range done,
 * [8] everything done */
int dbt_dist_stats(const dbt_dist *d, double out[16]);
/* host-only self test of the control block (rendezvous, barriers, all-gathers, splitter choice, layout arithmetic);
 * no CUDA call: runs on machines without a GPU.  *checksum is identical on all ranks. */
int dbt_dist_selftest_host(const char *session, int rank, int world, uint64_t *checksum);

/* ----------------------------------------------------------------------------------------------
 * Host-scope operators: image in host memory -> image in host memory, copies included.
 * These are what the file-based dbtproj entry points call after reading the block files into
 * pinned staging.  `device` is the CUDA device index.  h_out sized like the device-scope case.
 * ---------------------------------------------------------------------------------------------- */
int dbt_host_mergesort(const void *h_in, uint64_t nblocks, int field, void *h_out, int device, uint64_t *nrows);
int dbt_host_dedup(const void *h_in, uint64_t nblocks, int field, void *h_out, int device, uint64_t *nrows,
                   uint64_t *nunique);
int dbt_host_mergejoin(const void *h_in_r, uint64_t nblocks_r, const void *h_in_s, uint64_t nblocks_s, int field,
                       void *h_out_ur, void *h_out_us, void *h_out, int device, uint64_t *res);
int dbt_host_hashjoin(const void *h_in_r, uint64_t nblocks_r, const void *h_in_s, uint64_t nblocks_s, int field,
                      void *h_out, uint64_t out_capacity_blocks, int device, uint64_t *nres);

/* Pipelined host-scope jobs.  A query engine runs operators back to back; one job cannot use both
 * directions of the PCIe link at once (its first output record is known only after its last input
 * record has arrived -- the reference's MergeSort has the same barrier between run formation and
 * merging, DatabaseProject.cpp:192-236 vs 245-369), but two jobs can: job i's download overlaps
 * job i+1's upload.  Each of the DBT_HOST_SLOTS slots owns a stream, device buffers and a workspace.
 *   dbt_host_*_begin(slot, ...)  same arguments as the synchronous call minus the result pointers;
 *                                returns once upload, kernels and download are enqueued (the device
 *                                operators read small counters back, so the kernels have run).
 *   dbt_host_job_wait(slot, r)   blocks until the slot's result is in h_out; r[4] receives
 *                                sort {nrows}, dedup {nrows, nunique}, mergejoin {nres, nunique_r,
 *                                nunique_s, later_reads}, hashjoin {nres}.
 * The host buffers should be pinned (dbt_host_alloc); pageable memory works but is copied through
 * staging synchronously, so nothing overlaps.  The synchronous calls above are begin + wait on
 * slot 0.  One host thread drives the slots.  dbt_host_trim() frees every cached buffer. */
#define DBT_HOST_SLOTS 4
int dbt_host_mergesort_begin(int slot, const void *h_in, uint64_t nblocks, int field, void *h_out, int device);
int dbt_host_dedup_begin(int slot, const void *h_in, uint64_t nblocks, int field, void *h_out, int device);
int dbt_host_mergejoin_begin(int slot, const void *h_in_r, uint64_t nblocks_r, const void *h_in_s, uint64_t nblocks_s,
                             int field, void *h_out_ur, void *h_out_us, void *h_out, int device);
int dbt_host_hashjoin_begin(int slot, const void *h_in_r, uint64_t nblocks_r, const void *h_in_s, uint64_t nblocks_s,
                            int field, void *h_out, uint64_t out_capacity_blocks, int device);
int dbt_host_job_wait(int slot, uint64_t result[4]);
int dbt_host_job_slots(void);
int dbt_host_trim(void);

/* Out-of-core mode (the analogue of the reference's nmem_blocks-bounded external sort,
 * DatabaseProject.cpp:182-369, and chunked join reads, :521,:564).  Host-scope MergeSort / EliminateDuplicates /
 * HashJoin (and the file entry points above them) switch to it when their images do not fit on the device:
 * sort/dedup = in-core sort of chunk-sized runs + one global sort of the resident key columns + chunked gather from
 * the runs; hash join = R's key columns resident, S streamed in chunks.  `blocks` (> 0) forces the chunk size (the
 * tests use this to drive the path with small images; env DBT_OOC_CHUNK_BLOCKS does the same), 0 restores the
 * automatic choice from the device's memory size.  MergeJoin out of core = both dedups as above + a streamed
 * semi-join of dedup(R) against the keys of dedup(S).  Limit: < 2^30 rows per out-of-core sort. */
int dbt_host_set_chunk_blocks(uint64_t blocks);
/* what the last out-of-core call did: {runs (R chunks for a join), output chunks, chunk shrinks, key-width
 * restarts (120-byte strings found late), blocks staged for the gathers, S chunks} */
int dbt_host_ooc_stats(uint64_t out[6]);

/* pinned host memory helpers for callers that want zero-staging copies */
int dbt_host_alloc(void **p, size_t bytes);
int dbt_host_free(void *p);

/* ----------------------------------------------------------------------------------------------
 * Synthetic inputs generated directly in HBM (SURVEY.md 8d "G_syn"; the distribution follows
 * the reference generator main.cpp:41-77: 100 live rows per block, recid = row index, 5-letter
 * strings, "Hola" at row 1 of every block).  kind: 0 = exactly U distinct num keys over n rows,
 * 1 = uniform over [0,U), 2 = heavy-head power law over [0,U), 3 = half of the rows copy (num, str) of a random
 * row of the partner relation (kind 1, seed ^ 0x5EED) so that composite-key joins match, 4 = exact Zipf(1.1):
 * num = rho(rank), rank ~ Zipf(s = 1.1) over [1, U] by rejection-inversion, rho a bijection of [0, U).  Same arithmetic as
 * oracle/dbt_oracle.c orc_gen_syn so any sub-range can be reproduced on the CPU.
 * ---------------------------------------------------------------------------------------------- */
int dbt_gen_syn(uint64_t seed, uint64_t n_total, uint64_t U, int kind, uint64_t row0, uint64_t nrows, uint32_t recid0,
                void *d_image, void *stream);

/* ----------------------------------------------------------------------------------------------
 * Stage timing (CUDA events on the operator's stream, resolved at the operator's final sync).
 * ---------------------------------------------------------------------------------------------- */
void dbt_stage_timing_enable(int on);
void dbt_stage_timing_reset(void);
int dbt_stage_count(void);
const char *dbt_stage_name(int i);
double dbt_stage_ms(int i);    /* accumulated milliseconds since the last reset */
uint64_t dbt_stage_launches(int i);
uint64_t dbt_kernel_launches(void); /* kernels launched by this library since process start */

#ifdef __cplusplus
}
#endif
#endif /* DBT_B200_H */

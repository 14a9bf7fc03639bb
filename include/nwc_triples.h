/*
 * libnwc_triples -- B200-native (sm_100a) CCSD(T) perturbative-triples kernels behind NWChem TCE's
 * Fortran->C call surface for src/tce/ccsd_t.  Plain C ABI: pointers and sizes only.
 *
 * Tier 1 ("compat"): the exact symbols the reference's GPU driver binds
 *   (ccsd_t_gpu.F, ccsd_t_singles_gpu.F, ccsd_t_doubles_gpu.F -> sd_t_total.cu, memory.cu, hybrid.c).
 *   Fortran calling convention: lower case, one trailing underscore, every argument by reference,
 *   integers are 64-bit (`typedef long Integer`, header.h:51).
 *   Semantics differ from the reference in one documented way: execution is DEFERRED.  Each sd_t_*_cuda_
 *   call copies its host operands to the device (they may be freed by the caller on return, as in
 *   ccsd_t_doubles_gpu.F:723-726) and records the contraction; compute_en_ then runs ONE fused kernel
 *   over everything recorded for the tuple and returns the two energies.  The t3 tile is never
 *   materialised, so the host `triplesx` arguments are ignored exactly as in the reference.
 *   Errors: print + exit(1) (header.h:27-37).
 *
 * Tier 2 ("native"): resident block stores, task partitioning and multi-GPU reduction, for hosts
 *   that keep T1/T2/V2 in HBM (replaces the Global Arrays gets of get_block.F:79-81, the nxtask
 *   counter of util_gnxtval.c:31 and the ga_dgop of ccsd_t.F:297).  Functions return 0 on success,
 *   nonzero on error with a message available from nwc_triples_last_error(); after an error the context has
 *   dropped whatever batch it was building and can be used again (CUDA errors, a missing block key, an arena
 *   over its cap never terminate the process in this tier).
 */
#ifndef NWC_TRIPLES_H
#define NWC_TRIPLES_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef long Integer; /* src/tce/ccsd_t/header.h:51 */

/* ------------------------------------------------------------------------------------------
 * Tier 1: drop-in symbols
 * ---------------------------------------------------------------------------------------- */
/* hybrid.c:24 -- 1 if this rank drives a GPU (rank-on-node < *icuda) */
int check_device_(Integer *icuda);
/* hybrid.c:31 -- bind the rank to a device; writes 30 to *cuda_device_number if the node has fewer
 * than *icuda devices (ccsd_t_gpu.F:58-60 turns that into errquit) */
int device_init_(Integer *icuda, Integer *cuda_device_number);
/* memory.cu:70 / :165 -- pool life cycle (per tuple, ccsd_t_gpu.F:135,220) */
void initmemmodule_(void);
void finalizememmodule_(void);
/* sd_t_total.cu:5392 / :12 -- open the singles / doubles part of a tuple; dims are the TASK tuple's ranges */
void dev_mem_s_(Integer *h1d, Integer *h2d, Integer *h3d, Integer *p4d, Integer *p5d, Integer *p6d);
void dev_mem_d_(Integer *h1d, Integer *h2d, Integer *h3d, Integer *p4d, Integer *p5d, Integer *p6d);
/* sd_t_total.cu:24 */
void dev_release_(void);

/* sd_t_total.cu:5522.. (s1), :312.. (d1), :2885.. (d2).  dims are the PERMUTED tuple's ranges.
 * t1sub(p4,h1), t2sub(h7,p4,p5,h1) | t2sub(p7,p4,h1,h2), v2sub(h3,h2,p6,p5) | (h3,h2,p6,h7) | (p7,h3,p6,p5),
 * all Fortran column-major host arrays. */
#define NWC_DECL_S1(K)                                                                                   \
  void sd_t_s1_##K##_cuda_(Integer *h1d, Integer *h2d, Integer *h3d, Integer *p4d, Integer *p5d, Integer *p6d, \
                           double *triplesx_unused, double *t1sub, double *v2sub);
#define NWC_DECL_D1(K)                                                                                   \
  void sd_t_d1_##K##_cuda_(Integer *h1d, Integer *h2d, Integer *h3d, Integer *h7d, Integer *p4d, Integer *p5d, \
                           Integer *p6d, double *triplesx_unused, double *t2sub, double *v2sub);
#define NWC_DECL_D2(K)                                                                                   \
  void sd_t_d2_##K##_cuda_(Integer *h1d, Integer *h2d, Integer *h3d, Integer *p4d, Integer *p5d, Integer *p6d, \
                           Integer *p7d, double *triplesx_unused, double *t2sub, double *v2sub);
NWC_DECL_S1(1) NWC_DECL_S1(2) NWC_DECL_S1(3) NWC_DECL_S1(4) NWC_DECL_S1(5) NWC_DECL_S1(6) NWC_DECL_S1(7) NWC_DECL_S1(8) NWC_DECL_S1(9)
NWC_DECL_D1(1) NWC_DECL_D1(2) NWC_DECL_D1(3) NWC_DECL_D1(4) NWC_DECL_D1(5) NWC_DECL_D1(6) NWC_DECL_D1(7) NWC_DECL_D1(8) NWC_DECL_D1(9)
NWC_DECL_D2(1) NWC_DECL_D2(2) NWC_DECL_D2(3) NWC_DECL_D2(4) NWC_DECL_D2(5) NWC_DECL_D2(6) NWC_DECL_D2(7) NWC_DECL_D2(8) NWC_DECL_D2(9)

/* sd_t_total.cu:5373 -- energy[0] = E[T] part, energy[1] = E(T) part of this tuple (ccsd_t_gpu.F:205-217) */
void compute_en_(double *factor, double *energy, double *eval_h1, double *eval_h2, double *eval_h3, double *eval_p4,
                 double *eval_p5, double *eval_p6, Integer *h1d, Integer *h2d, Integer *h3d, Integer *p4d,
                 Integer *p5d, Integer *p6d, double *host_doubles_unused, double *host_singles_unused);

/* not in the reference: tell the library this process's rank on its node when util_my_smp_index() is not linked */
void nwc_triples_set_local_rank(Integer local_rank);
/* Upload policy of the sd_t_*_cuda_ calls.  Default 0: a call returns only when its host operands may be freed or
 * overwritten (the reference's contract).  1: the caller promises that PINNED operands stay untouched until
 * compute_en_ returns, so their copies run as asynchronous DMA overlapped with the caller's next sort. */
void nwc_compat_set_async_uploads(int on);
/* validation aid: like compute_en_, but also writes the two t3 tiles T3(h3,h2,h1,p6,p5,p4) to host arrays */
void nwc_compute_en_dump_(double *factor, double *energy, double *eval_h1, double *eval_h2, double *eval_h3,
                          double *eval_p4, double *eval_p5, double *eval_p6, Integer *h1d, Integer *h2d, Integer *h3d,
                          Integer *p4d, Integer *p5d, Integer *p6d, double *host_doubles, double *host_singles);

/* ------------------------------------------------------------------------------------------
 * Host-side driver (C++ restatement of the Fortran above the kernel boundary; stands in for
 * ccsd_t_gpu.F + ccsd_t_singles_gpu.F + ccsd_t_doubles_gpu.F where no Fortran compiler exists).
 * It fetches and sorts blocks on the HOST exactly like the reference and calls the Tier-1 symbols.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  Integer noab, nvab;      /* tce.fh:19-20 */
  Integer restricted;      /* tce.fh:78   */
  Integer irrep_t, irrep_v;
  const Integer *spin;     /* k_spin   [noab+nvab], 1 alpha / 2 beta */
  const Integer *sym;      /* k_sym    irrep bit code */
  const Integer *range;    /* k_range */
  const Integer *offset;   /* k_offset into evl_sorted */
  const Integer *alpha;    /* k_alpha (1-based) */
  const double *evl_sorted;
  const Integer *t1_hash; const double *t1;  /* tce_t1_offset_new.F  : [n, keys.., offsets..] */
  const Integer *t2_hash; const double *t2;  /* tce_t2_offset_new.F */
  const Integer *v2_hash; const double *v2;  /* tce_mo2e_offset.F   */
} nwc_tce_state;

/* ccsd_t_gpu.F:2 -- whole (T) through the Tier-1 call surface on this rank; tasks dealt round-robin
 * as `my_rank`-th of `nranks` (nranks=1: all).  energy[0]=E[T], energy[1]=E(T), unreduced. */
int nwc_ccsd_t_gpu(const nwc_tce_state *st, Integer icuda, Integer my_rank, Integer nranks, double energy[2],
                   double *per_task /* 2*ntasks or NULL */);
/* the same for a given list of tasks (ntasks x 6 tile ids), e.g. a prefix of the heaviest-first list */
int nwc_ccsd_t_gpu_tasks(const nwc_tce_state *st, Integer icuda, const Integer *tasks6, Integer ntasks, double energy[2],
                         double *per_task /* 2*ntasks or NULL */);
/* 1: the host driver above behaves exactly like the unmodified Fortran call sites (one pageable scratch buffer refilled
 * per operand pair, nothing pinned, no promise to the library); 0 (default): pinned scratch + nwc_compat_set_async_uploads */
void nwc_driver_set_reference_contract(int on);
/* Route the host driver's Tier-1 calls (the 27 kernels, dev_mem_*, compute_en_, dev_release_, init/finalizememmodule_)
 * into another shared library that exports the reference's symbols -- e.g. the reference's own sd_t_total.cu + memory.cu
 * compiled unmodified -- instead of this library's.  NULL or "": back to this library.  0 on success. */
int nwc_driver_bind_backend(const char *so_path);
/* threads of the library's host-side loops (TCE_SORT_4 of the host driver, staging copies); <= 0: leave unchanged.
 * torch.distributed.run exports OMP_NUM_THREADS=1, so a launcher should pass cores / ranks-per-node here. */
void nwc_triples_set_host_threads(int n);
/* one tuple through the Tier-1 surface; optional t3 tiles out (host, T3(h3,h2,h1,p6,p5,p4)) */
int nwc_ccsd_t_gpu_tuple(const nwc_tce_state *st, const Integer tuple_p4p5p6h1h2h3[6], double energy[2],
                         double *host_doubles, double *host_singles);

/* host-only helpers (no device): task list of ccsd_t_neword.F and a dry run of one tuple's dispatch */
Integer nwc_host_task_list(const nwc_tce_state *st, Integer *klist7, Integer capacity_tasks);
/* sorted unique block keys of store `which` (1 T1, 2 T2, 3 spin-orbital V2) read by the given tasks (ntasks x 6 tile
 * ids): a caller can stage only those blocks on the host.  Returns their number (keys_out filled if cap suffices). */
Integer nwc_host_collect_blocks(const nwc_tce_state *st, const Integer *tasks6, Integer ntasks, int which,
                                Integer *keys_out, Integer cap);
/* host-only: the host driver's TCE_SORT_4 -- sorted(i,j,k,l order) = factor * unsorted(a,b,c,d order), last index
 * fastest (src/tce/sort/new_sort4.F semantics), cache-blocked and threaded */
void nwc_host_sort4(const double *unsorted, double *sorted, Integer a, Integer b, Integer c, Integer d, int i, int j,
                    int k, int l, double factor);
/* host-only view of nwc_triples_run_partition (below): ranges[2*i], ranges[2*i+1] = the sub-tile range of task
 * first_task+i that `rank` of `nranks` runs (ntasks <= 0: to the end of the list; ranges holds 2*ntasks entries) */
int nwc_host_block_partition(const nwc_tce_state *st, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                             long long *ranges);
int nwc_host_count_tuple(const nwc_tce_state *st, const Integer tuple_p4p5p6h1h2h3[6], Integer calls_s1_d1_d2[3],
                         double flops_s1_d1_d2[3]);

/* ------------------------------------------------------------------------------------------
 * Tier 2: native API
 * ---------------------------------------------------------------------------------------- */
typedef struct nwc_triples_ctx nwc_triples_ctx;

typedef struct {
  double fused_ms, repack_ms;          /* CUDA-event kernel time accumulated while timing is on */
  long long fused_launches, repack_launches, reduce_launches;
  long long work_items, descs, tuples;
  double flops;                        /* algorithmic FLOPs of the tuples run */
  double h2d_bytes, d2h_bytes;
  double resident_bytes;               /* T1+T2+V2 in HBM (this rank's shard when V2 is sharded) */
  double pull_ms;                      /* CUDA-event time of the peer-block pulls (NVLink) while timing is on */
  double peer_bytes;                   /* bytes pulled from other GPUs' V2 shards */
  long long pull_launches, antisym_launches;
} nwc_triples_stats;

const char *nwc_triples_last_error(void);
int nwc_triples_create(nwc_triples_ctx **out, int device);
int nwc_triples_destroy(nwc_triples_ctx *ctx);
/* Host-only TRACE context: no device, no arithmetic, nothing executes.  The host driver logic of the native tier (which
 * blocks a tuple reads, through which strides, into which kernel and with which sign: the restatement of
 * ccsd_t_singles_gpu.F / ccsd_t_doubles_gpu.F / lambda_ccsd_t_left.F / cr_ccsd_t_N.F / cr_ccsd_t_E.F in
 * csrc/host_driver.h + csrc/native_abi.cu) runs exactly as it does in front of the GPU, but the operand descriptors it
 * hands to the engine are recorded instead of being packed and launched.  set_state (replicated spin-orbital stores
 * only), set_lambda and set_cr keep the caller's host arrays by reference; trace_tuple records one tuple
 * (method 0: (T), 1: Lambda-CCSD(T), 2 / 3: CR-CCSD(T) numerator / denominator pass of the two-pass form, 4: the one-pass
 * dual tuple, 5 / 6 / 7: the three tuples of the composed form of CR-EOMCCSD(T), 8: its one-tuple form); trace_take hands the records
 * out.  Every compute entry point fails on a trace context.  It exists so the CPU test-suite can check the driver half
 * against the oracle's tiles without a GPU (tests/test_trace.py); the pointers in the records are the caller's. */
typedef struct {
  Integer kind;        /* 0 sd_t_s1_K, 1 sd_t_d1_K, 2 sd_t_d2_K (one contracted tile), 3 outer product, 9 end of tuple */
  Integer k0;          /* K - 1 (kinds 0-2) */
  Integer side;        /* kinds 1,2: 0 = the tuple's doubles tile, 1 = the second tile of a two-sided tuple;
                          kind 3: 0 = singles tile, 1 = side-1 tile, 2 = side-0 tile */
  Integer K;           /* contracted range (kinds 1,2); kind 9: 1 if the tuple is two-sided, 2 if it is a dual-energy
                          tuple (the side-0 outer products form a fourth tile of their own), 3 for the CR-EOMCCSD(T)
                          form of a dual tuple (side 1 = R, singles = L) */
  Integer neg;         /* kind 3: 1 = subtract */
  const double *a;     /* kinds 0-2: t1sub / t2sub source; kind 3: the two-index operand */
  const double *b;     /* kinds 0-2: v2sub source; kind 3: the four-index operand */
  long long sa[6];     /* element strides: kinds 0-2 per PERMUTED name (h1,h2,h3,p4,p5,p6); kind 3 per physical
                          position (h3,h2,h1,p6,p5,p4); kind 9: sa = the ranges by physical position, a / b / sb[2..5] =
                          the six orbital-energy vectors (h1,h2,h3,p4,p5,p6) as addresses, sb[0..1] the sub-tile range */
  long long sb[6];
  long long ka, kb;    /* strides of the contracted index */
  double scale;        /* kinds 1,2: factor of the T2 sort (tce_hashnsort.F); kind 9: the tuple factor */
} nwc_trace_rec;
int nwc_triples_create_trace(nwc_triples_ctx **out);
int nwc_triples_trace_tuple(nwc_triples_ctx *ctx, const Integer tuple_p4p5p6h1h2h3[6], int method);
/* copies up to cap records to out, clears the trace, *n = number of records there were */
int nwc_triples_trace_take(nwc_triples_ctx *ctx, nwc_trace_rec *out, size_t cap, size_t *n);
/* copies the tiling tables and uploads the three block stores into HBM (replicated per GPU).  In every set_state*
 * variant a NULL data pointer (st->t1, st->t2, st->v2, orb->v2orb) means "allocate the store, do not upload": the
 * caller fills it on the device (nwc_triples_synth_fill).  A new set_state* call releases whatever the previous one
 * held (other storage modes' buffers, peer mappings). */
int nwc_triples_set_state(nwc_triples_ctx *ctx, const nwc_tce_state *st);
/* Sharded V2 for shapes whose <pp||hp> class does not fit one HBM (replaces the ga_get of
 * get_block.F:79-81 by NVLink peer reads): block i of the V2 offset table is owned by rank i % nranks and
 * st->v2 holds THIS rank's blocks only (table order, compacted); T1/T2 stay replicated.  After set_state_sharded
 * every rank publishes nwc_triples_v2_ipc_handle (64 bytes), the host all-gathers them and calls
 * nwc_triples_v2_open_peers(handles[nranks*64]). */
int nwc_triples_set_state_sharded(nwc_triples_ctx *ctx, const nwc_tce_state *st, int rank, int nranks);
int nwc_triples_v2_ipc_handle(nwc_triples_ctx *ctx, char handle64[64]);
int nwc_triples_v2_open_peers(nwc_triples_ctx *ctx, const char *handles);
/* same-process alternative (several contexts in one process): exchange raw device pointers instead of IPC handles */
void *nwc_triples_v2_shard_ptr(nwc_triples_ctx *ctx);
int nwc_triples_v2_set_peer_ptr(nwc_triples_ctx *ctx, int rank, void *dev_ptr);
/* task list of ccsd_t_neword.F (7 Integers per task: p4b,p5b,p6b,h1b,h2b,h3b,weight), heaviest first */
Integer nwc_triples_num_tasks(nwc_triples_ctx *ctx);
int nwc_triples_task_list(nwc_triples_ctx *ctx, Integer *klist7);
/* static partition replacing nxtask: tasks first, first+stride, ... (< ntasks).  energy[2] accumulates
 * nothing across calls: it is set to this call's sums.  per_task: 2 doubles per task run, or NULL. */
int nwc_triples_run(nwc_triples_ctx *ctx, Integer first, Integer stride, Integer max_tasks, double energy[2],
                    double *per_task);
/* Static block partition of the task space over `nranks` GPUs, the stand-in for nxtask's dynamic counter
 * (ccsd_t.F:174-255, util_gnxtval.c:31-36): tasks [first_task, first_task+ntasks) of the heaviest-first list (ntasks <= 0:
 * to the end) are laid end to end, each 4^6 sub-tile weighted by the k4 steps its tuple contracts, and rank r runs the
 * r-th equal-cost contiguous piece; a tuple on a boundary is shared between two ranks at sub-tile granularity (the
 * energies are additive), so the balance does not depend on the number of tuples.  energy[2] = this rank's sums
 * (combine with nwc_triples_allreduce_energy); per_task (optional, 2*ntasks doubles indexed by task - first_task) = this
 * rank's possibly partial per-task energies. */
int nwc_triples_run_partition(nwc_triples_ctx *ctx, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                              double energy[2], double *per_task);
/* the same for an explicit list of n task indices (e.g. a strided sample of the list), partitioned in the order given;
 * per_task is indexed by position in the list */
int nwc_triples_run_partition_list(nwc_triples_ctx *ctx, Integer rank, Integer nranks, const Integer *task_ids, Integer n,
                                   double energy[2], double *per_task);
/* one tuple restricted to sub-tiles [item_lo, item_hi) of its linear sub-tile order (4-wide blocks, h3 block fastest,
 * p4 block slowest; nwc_triples_tuple_items = their number): e.g. the p4 slab [4a,4b) of the t3 tile is
 * [a*m, b*m), m = items / ceil(range(p4)/4).  Energies of disjoint ranges add up to the tuple's. */
int nwc_triples_run_items(nwc_triples_ctx *ctx, const Integer tuple_p4p5p6h1h2h3[6], long long item_lo, long long item_hi,
                          double energy[2]);
long long nwc_triples_tuple_items(nwc_triples_ctx *ctx, const Integer tuple_p4p5p6h1h2h3[6]);
/* Synthetic stores generated on the device (benchmarks, tests): fills T1, T2 and the V2 store the context holds (whole,
 * or this rank's shard) with scale * U(-1,1) values that are a pure function of (seed, store, block key, element index)
 * -- no rank ever holds a store on the host and every rank count sees the same tensors.  nwchem_b200/synth.py restates
 * the generator in numpy for the oracle. */
int nwc_triples_synth_fill(nwc_triples_ctx *ctx, unsigned long long seed, double scale_t1, double scale_t2,
                           double scale_v2);
/* validation aids: read back part of a resident store (which: 1 T1, 2 T2, 3 V2 shard, 4 orbital-form V2 shard), and
 * the spin-orbital block <g3 g4||g1 g2> as the (T) path sees it (from whichever storage the context holds) */
int nwc_triples_debug_read(nwc_triples_ctx *ctx, int which, size_t offset, size_t n, double *host_out);
int nwc_triples_export_v2_block(nwc_triples_ctx *ctx, const Integer g3g4g1g2[4], double *host_out);
/* `2eorb` V2 storage (tce.fh `intorb`): the two-electron integrals are kept spin-free over the ALPHA
   tiles and every spin-orbital block <g3 g4||g1 g2> = (g3 g1|g4 g2) - (g3 g2|g4 g1) is antisymmetrised on the
   device when a tuple needs it -- replaces get_hash_block_i (get_hash_block.F:47-118) -> get_block_ind_i
   (get_block_ind.F:818-1538) -> tce_hash_v2 (tce_hash.F:1-135).  The arrays are the reference's own:
   b2am = int_mb(k_b2am) (tce_tile.F:1156-1212), spin/sym/range_alpha = k_spin_alpha.. (tce_tile.F:1376-1383),
   v2orb_hash = int_mb(k_v2_alpha_offset), the checkpointed table of tce_mo2e_offset_intorb.F:52-150
   ([length1 | keys | offsets | g3b | g4b | g1b | g2b], 6*(length1+1)+1 integers), v2orb = the d_v2orb file,
   block (g3b<=g4b | g1b<=g2b) holding (k l|i j), k in g4b fastest, l in g3b, i in g2b, j in g1b
   (tce_mo2e_trans.F:707-723). */
typedef struct {
  Integer noa, nva;
  const Integer *b2am;          /* [noab+nvab] */
  const Integer *spin_alpha;    /* [noa+nva] */
  const Integer *sym_alpha;
  const Integer *range_alpha;
  const Integer *v2orb_hash;
  const double *v2orb;
} nwc_tce_orb_state;
/* host-only view of the device-side antisymmetrisation plan (tests): where the direct / exchange halves of
   <g3 g4||g1 g2> sit in the caller's d_v2orb (element offset, -1 = half absent for these spins) and the element
   strides of (g3,g4,g1,g2) in each */
int nwc_host_2eorb_plan(const nwc_tce_state *st, const nwc_tce_orb_state *orb, const Integer g3g4g1g2[4],
                        Integer off_host[2], Integer strides[8]);
/* like nwc_triples_set_state, but V2 comes from `orb`; st->v2_hash / st->v2 are not read (may be NULL) */
int nwc_triples_set_state_2eorb(nwc_triples_ctx *ctx, const nwc_tce_state *st, const nwc_tce_orb_state *orb);
/* `2eorb` AND sharded (the only form in which (H2O)10/aug-cc-pVTZ fits: 105 GB orbital-form vs 527 GB spin-orbital,
 * get_hash_block.F:47-118, get_block_ind.F:818-1538): the i-th orbital block (T) can touch, in storage order, is owned
 * by rank i % nranks.  orb->v2orb, if not NULL, is the caller's full d_v2orb file (only this rank's blocks are read).
 * Peers are mapped with the same nwc_triples_v2_ipc_handle / nwc_triples_v2_open_peers exchange as above.  The remote
 * orbital blocks a batch of tuples needs are pulled whole over NVLink into the batch arena (contiguous 16-byte loads)
 * and antisymmetrised locally. */
int nwc_triples_set_state_2eorb_sharded(nwc_triples_ctx *ctx, const nwc_tce_state *st, const nwc_tce_orb_state *orb,
                                        int rank, int nranks);
/* Restartable (T): replaces ccsd_t_restart.F:57-290.  *restart_begin and table[nvab] are the RTDB entries
   'tce:ccsd_t_restart_begin' (1-based outer virtual tile index, :57-66) and 'tce:restart_triples_table' (:84-95).
   For outer = *restart_begin .. nvab (at most max_outer of them when max_outer > 0) the CCSD(T) partial of every
   tuple with t_p4b = noab+outer is computed -- this rank takes the tuples first, first+stride, ... of that outer
   tile's loop order (:120-150), the static stand-in for the per-tile nxtask deal -- then, if a communicator was set up
   with nwc_triples_nccl_init, summed over ranks (the ga_dgop at :255; collective: every rank must make the same
   call), stored in table[outer-1] (:274) and *restart_begin advanced to outer+1 (:247).  The caller persists both
   between calls (max_outer = 1 gives one checkpoint per outer tile like the reference).  table_bracket (optional,
   nvab doubles) receives the CCSD[T] partials the same way.  *t_energy = sum(table) (:288-290). */
int nwc_triples_run_restart(nwc_triples_ctx *ctx, Integer first, Integer stride, Integer *restart_begin, double *table,
                            double *table_bracket, Integer max_outer, double *t_energy);
/* Lambda-CCSD(T) (src/tce/ccsd_t/lambda_ccsd_t.F, tce_energy.F:3404-3437): the sibling correction that shares the 27
 * contractions.  set_lambda uploads lambda_1 (h,p), lambda_2 (hh,pp) and the (h,p) Fock blocks in the reference's block
 * layout with their offset tables (keys: lambda_ccsd_t_left.F:154-155, :378-380, :390-391); run_lambda returns
 * energy[0] = sum f Td Yd/Delta (Lambda-CCSD[T]) and energy[1] = sum f Td (Ys+Yd)/Delta (Lambda-CCSD(T)) over tasks
 * first, first+stride, ... of the heaviest-first list, per_task (optional) 2 doubles per task run.  The left-hand tiles
 * are paired with the right-hand tile in T3 index order, i.e. with the L3->T3 sort lambda_ccsd_t.F:35-36 announces (the
 * file as written multiplies them with one running index; both readings coincide at tilesize 1 -- see DESIGN.md 8). */
int nwc_triples_set_lambda(nwc_triples_ctx *ctx, const Integer *y1_hash, const double *y1, const Integer *y2_hash,
                           const double *y2, const Integer *f1_hash, const double *f1);
int nwc_triples_run_lambda(nwc_triples_ctx *ctx, Integer first, Integer stride, Integer max_tasks, double energy[2],
                           double *per_task);
/* the same over the static block partition of nwc_triples_run_partition (combine with nwc_triples_allreduce_energy) */
int nwc_triples_run_lambda_partition(nwc_triples_ctx *ctx, Integer rank, Integer nranks, Integer first_task,
                                     Integer ntasks, double energy[2], double *per_task);
/* CR-CCSD(T) (src/tce/ccsd_t/cr_ccsd_t.F, tce_energy.F `cr-ccsd(t)`): the tuple loop cr_ccsd_t.F:93-233.  Per tuple it
 * needs the (T) tiles S (ccsd_t_singles_l) and D (ccsd_t_doubles_l), the moment M (cr_ccsd_t_N_1/_N_2, cr_ccsd_t_N.F:296,
 * :3540 -- the doubles contractions with V2 replaced by dressed intermediates) and the denominator tile E
 * (cr_ccsd_t_E_1/_E_2, cr_ccsd_t_E.F:74,:408 -- outer products), and forms num1 = sum f M D/Delta, num2 = sum f M (S+D)/Delta,
 * den1 = sum f E D/Delta, den2 = sum f E (S+D)/Delta (:176-207).  set_cr uploads the three intermediates the loop reads
 * -- what cr_ccsd_t_N(...,toggle 1) / cr_ccsd_t_E(...,toggle 1) leave in d_i1_1, d_i1_2 and d_i1_3, or the files
 * gr1_1 / gr1_2 / ei1_2 of read_in3 (cr_ccsd_t_N.F:98-104) -- in the reference's block layout with their offset tables:
 *   n1: i1(h11 p4 h1 h2), blocks (p4b, h11b, h1b<=h2b), key h2b-1+noab*(h1b-1+noab*(h11b-1+noab*(p4b-noab-1)))      (cr_ccsd_t_N.F:773-841)
 *   n2: i1(p4 p5 h1 p12), blocks (p4b<=p5b, h1b, p12b), key p12b-noab-1+nvab*(h1b-1+noab*(p5b-noab-1+nvab*(p4b-noab-1))) (:4011-4079)
 *   e2: i1(p4 p5 h1 h2)_tt, the T2 block structure and key                                                       (cr_ccsd_t_E.F:907-960)
 * run_cr returns sums[4] = (num1, num2, den1, den2) over tasks first, first+stride, ... of the heaviest-first list,
 * per_task (optional) 4 doubles per task run.  The caller adds den0 (the scalar of cr_ccsd_t_D, cr_ccsd_t.F:66-69) after
 * the sum over ranks and forms CR-CCSD[T] = num1/(1+den1+den0), CR-CCSD(T) = num2/(1+den2+den0) (:260-263). */
int nwc_triples_set_cr(nwc_triples_ctx *ctx, const Integer *n1_hash, const double *n1, const Integer *n2_hash,
                       const double *n2, const Integer *e2_hash, const double *e2);
/* The pphp intermediate is as large as V2's <pp||hp> class (2.5*o*v^3 doubles), so like V2 it can be dealt over the GPUs:
 * block i of its offset table lives on rank i % nranks and n2_shard holds THIS rank's blocks only (table order,
 * compacted); the hphh and pphh intermediates stay replicated.  Then exchange the shards exactly as for a sharded V2:
 * cr_ipc_handle (64 bytes) all-gathered + cr_open_peers, or inside one process cr_shard_ptr / cr_set_peer_ptr. */
int nwc_triples_set_cr_sharded(nwc_triples_ctx *ctx, const Integer *n1_hash, const double *n1, const Integer *n2_hash,
                               const double *n2_shard, const Integer *e2_hash, const double *e2, int rank, int nranks);
int nwc_triples_cr_ipc_handle(nwc_triples_ctx *ctx, char handle64[64]);
int nwc_triples_cr_open_peers(nwc_triples_ctx *ctx, const char *handles);
void *nwc_triples_cr_shard_ptr(nwc_triples_ctx *ctx);
int nwc_triples_cr_set_peer_ptr(nwc_triples_ctx *ctx, int rank, void *dev_ptr);
int nwc_triples_run_cr(nwc_triples_ctx *ctx, Integer first, Integer stride, Integer max_tasks, double sums[4],
                       double *per_task);
/* the same over the static block partition of nwc_triples_run_partition (combine with nwc_triples_allreduce_sum, n = 4) */
int nwc_triples_run_cr_partition(nwc_triples_ctx *ctx, Integer rank, Integer nranks, Integer first_task, Integer ntasks,
                                 double sums[4], double *per_task);
/* CR-EOMCCSD(T) (src/tce/cr-eomccsd_t/cr_eomccsd_t.F, tce_energy.F:8775-8786): the tuple loop :325-493.  Its six per-tuple
 * routines are the CR-CCSD(T) ones with other operands and constant factors (creomsd_t_n2_mem_1..4 == cr_ccsd_t_N_1/_N_2 on
 * (t2 | x2) x (d_i2_1..4), creomccsd_t_n2_mem.F:674,:5665,:9657,:12905; q3rexpt2_1/_2 == cr_ccsd_t_E_1/_E_2 on (t2, x1) and
 * (t1, d_i3_1), q3rexpt2.F:80,:414).  Per tuple: right = r0*cr_ccsd_t_N + creomsd_t_n2_mem, left = r0*cr_ccsd_t_E + q3rexpt2,
 * denex = Delta + excit, and  num1 += f*right^2/denex + f*left*right,  den1 += f*left*right/denex + f*left^2 (:455-464).
 * set_creom takes what the loop reads beyond T1/T2 and the CR-CCSD(T) intermediates of nwc_triples_set_cr (needed when
 * |r0| >= 1e-7, :146-147): x1 / x2 (offset tables equal to T1's / T2's: irrep_x = 0), d_i2_1..4 with their offset tables
 * (the layouts of the CR ones: hphh stored (p,h,h<=h), pphp stored (p<=p,h,p)), d_i3_1 (the T2 block structure), r0 and
 * the excitation energy -- all produced upstream with toggle 1, or read from files with read_in3.  run_creom returns
 * sums[4] = (sum f R R/denex, sum f L R, sum f L R/denex, sum f L L); num1 = sums[0]+sums[1], den1 = sums[2]+sums[3],
 * energy1 = num1/(r0^2 + d12 + den1) (:564) with the caller's d12. */
int nwc_triples_set_creom(nwc_triples_ctx *ctx, const Integer *x1_hash, const double *x1, const Integer *x2_hash,
                          const double *x2, const Integer *i2_1_hash, const double *i2_1, const Integer *i2_2_hash,
                          const double *i2_2, const Integer *i2_3_hash, const double *i2_3, const Integer *i2_4_hash,
                          const double *i2_4, const Integer *i3_1_hash, const double *i3_1, double r0, double excit);
int nwc_triples_run_creom(nwc_triples_ctx *ctx, Integer first, Integer stride, Integer max_tasks, double sums[4],
                          double *per_task);
int nwc_triples_run_creom_partition(nwc_triples_ctx *ctx, Integer rank, Integer nranks, Integer first_task,
                                    Integer ntasks, double sums[4], double *per_task);
/* one tuple, optionally materialising the t3 tiles (validation only) */
int nwc_triples_run_tuple(nwc_triples_ctx *ctx, const Integer tuple_p4p5p6h1h2h3[6], double energy[2],
                          double *host_doubles, double *host_singles);
int nwc_triples_set_timing(nwc_triples_ctx *ctx, int on);
int nwc_triples_get_stats(nwc_triples_ctx *ctx, nwc_triples_stats *out, int reset);
/* device-side stopwatch: CUDA events recorded on the stream the kernels are launched on */
int nwc_triples_timer_start(nwc_triples_ctx *ctx);
int nwc_triples_timer_stop_ms(nwc_triples_ctx *ctx, double *ms);
/* pin / unpin a host range (cudaHostRegister) so Tier-1 operand uploads are true async DMA */
int nwc_host_register(void *ptr, size_t bytes);
int nwc_host_unregister(void *ptr);
/* counters of the Tier-1 engine (the one behind sd_t_*_cuda_ / compute_en_) */
int nwc_compat_get_stats(nwc_triples_stats *out, int reset);
int nwc_compat_set_timing(int on);
int nwc_compat_timer_start(void);
int nwc_compat_timer_stop_ms(double *ms);
/* panel arena budget per batch in bytes (default 8 GiB; two batches are in flight) and the hard cap of one batch
 * arena (default 150 GiB): beyond it a call fails with an error instead of exhausting the device */
int nwc_triples_set_batch_bytes(nwc_triples_ctx *ctx, size_t bytes);
int nwc_triples_set_arena_cap(nwc_triples_ctx *ctx, size_t bytes);
/* give the batch arenas back to the device (tens of GB after large tuples; re-allocated on demand).  The resident
 * stores stay.  nwc_compat_trim does the same for the Tier-1 engine. */
int nwc_triples_trim(nwc_triples_ctx *ctx);
int nwc_compat_trim(void);
/* index order inside the operand panels chosen for this tiling: 0 holes first, 1 particles first (the less ragged tile
 * type goes first so that more padding rows can be skipped; env NWC_ORDER overrides) */
int nwc_triples_get_order(nwc_triples_ctx *ctx);

/* roofline denominator measured in-process: rate (TFLOP/s) of a register-resident DMMA.8x8x4 loop on `device`
 * (the FP64 tensor instruction of the K loop; MEASURED_PEAKS.json has no FP64 figure).  ~0.1 s. */
int nwc_fp64_peak_probe(int device, double *dmma_tflops);

/* multi-GPU: one process per GPU.  The host distributes the 128-byte id (MPI/GA broadcast in NWChem,
 * torch.distributed in bench.py), then every rank calls init; allreduce replaces ga_dgop (ccsd_t.F:297). */
int nwc_triples_nccl_unique_id(char id128[128]);
int nwc_triples_nccl_init(nwc_triples_ctx *ctx, const char id128[128], int rank, int nranks);
int nwc_triples_allreduce_energy(nwc_triples_ctx *ctx, double energy[2]);
/* the same collective on n doubles (per-task energies of a partitioned run) */
int nwc_triples_allreduce_sum(nwc_triples_ctx *ctx, double *buf, size_t n);

#ifdef __cplusplus
}
#endif
#endif

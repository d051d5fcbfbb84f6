/* include/qtorch_b200.h -- C ABI of libqtorch_b200.so, the B200 (sm_100a) contraction engine that
 * replaces the arithmetic and tensor storage behind qTorch's Network/Node API.
 *
 * The reference has no FFI/plugin interface for this path; the boundary is the pair of C++ seams
 *   Network::ContractNodes  -> Network::ContractIndices   /root/reference/src/Network.h:715, :876
 *   Node storage (mVals, Index/Access/GetTensorVals/ClearNodeData)  /root/reference/src/Node.h:108-194
 * Every entry point below cites the reference behaviour it replaces.  INTEGRATION.md shows the shim a
 * reference maintainer adds inside Network.h/Node.h; qtorch_b200/host/ is that shim written out as a
 * full drop-in header set.
 *
 * Conventions (identical to the reference):
 *   - every leg (Wire) has dimension 4; a rank-r tensor has 4^r elements of interleaved (re, im)
 *     IEEE binary64, element index = sum_j digit_j * 4^j (leg 0 fastest)          Node.h:178-186
 *   - a step contracts ALL shared legs of A and B; shared pair j = (pos_a[j], pos_b[j]) with pos_a
 *     strictly increasing (A-wire order)                                           Network.h:739-758
 *   - C's legs = A's free legs in A order, then B's free legs in B order            Network.h:809-812
 *   - C is overwritten (the reference zero-fills then accumulates)                  Node.h:112-113
 * Plain pointers and sizes only; no C++ exceptions cross this ABI; every call returns a status.
 * There is no CPU fallback: without a usable CUDA device every compute entry returns
 * QTB_ERR_NO_DEVICE and the host shim throws.
 */
#ifndef QTORCH_B200_H
#define QTORCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTB_ABI_VERSION 1
#define QTB_MAX_RANK 16            /* 16 * 4^16 B = 68.7 GB: the largest single tensor one B200 can hold */

typedef struct qtb_ctx_s qtb_ctx;            /* one engine context: device, stream, pool, pending micro-steps */
typedef struct qtb_tensor_s *qtb_tensor;     /* opaque device tensor handle (NULL = none) */
typedef struct qtb_plan_s qtb_plan;          /* a compiled contraction plan (CUDA-graph backed) */

enum qtb_status {
    QTB_OK = 0,
    QTB_ERR_NO_DEVICE = 1,      /* no CUDA device / driver: the product refuses to run                        */
    QTB_ERR_INVALID = 2,        /* bad argument (rank, leg map, NULL)  -> InvalidFunctionInput (Exceptions.h:44) */
    QTB_ERR_EMPTY_INPUT = 3,    /* operand data was cleared            -> InvalidFunctionInput (Network.h:938-940) */
    QTB_ERR_OOM = 4,            /* device memory exhausted             -> ContractionFailure (Exceptions.h:36)  */
    QTB_ERR_CUDA = 5,           /* a CUDA call failed (see qtb_last_error) -> ContractionFailure               */
    QTB_ERR_NCCL = 6,           /* NCCL unavailable or failed                                                   */
    QTB_ERR_UNSUPPORTED = 7
};

/* ---- library ---------------------------------------------------------------------------------- */
int         qtb_abi_version(void);
const char *qtb_status_string(int status);
const char *qtb_last_error(void);                 /* thread-local detail of the last failure */
int         qtb_device_count(int *count);         /* QTB_ERR_NO_DEVICE when CUDA is absent    */

/* ---- context: replaces nothing in the reference (it has no device); one per Network user thread
 * group.  All calls on one ctx are serialised internally, so the two planner threads of
 * ContractionTools::ParallelContract (ContractionTools.h:360-368) may share it.                     */
int qtb_ctx_create(int device, qtb_ctx **out);
int qtb_ctx_destroy(qtb_ctx *ctx);
int qtb_ctx_sync(qtb_ctx *ctx);                   /* flush deferred micro-steps and wait for the stream */
int qtb_ctx_flush(qtb_ctx *ctx);                  /* flush deferred micro-steps, do not wait            */
void *qtb_ctx_stream(qtb_ctx *ctx);               /* the cudaStream_t all work of this ctx runs on     */

/* ---- tensor storage: replaces Node::mVals (Node.h:160) ------------------------------------------ */
/* Node(int rank) (Node.h:112-113).  Contents are UNINITIALISED: the reference's zero-fill is only ever
 * observed through host access, which the shim serves from its host mirror.                          */
int qtb_tensor_alloc(qtb_ctx *ctx, int rank, qtb_tensor *out);
/* Node::ClearNodeData (Node.h:137) / ~Node.  Stream-ordered: safe right after enqueuing a step.     */
int qtb_tensor_free(qtb_ctx *ctx, qtb_tensor t);
int qtb_tensor_rank(qtb_tensor t);
void *qtb_tensor_device_ptr(qtb_tensor t);        /* raw device address (for tests / interop)          */
/* Gate constructors writing through Node::Index (Node.h:197-898): host -> device, 2*4^rank doubles.
 * The host buffer may be reused as soon as the call returns.                                          */
int qtb_tensor_upload(qtb_ctx *ctx, qtb_tensor t, const double *host_re_im);
/* Node::GetTensorVals / Access (Node.h:151,192): device -> host, blocking.                           */
int qtb_tensor_download(qtb_ctx *ctx, qtb_tensor t, double *host_re_im);
/* LGContract reading GetTensorVals()[0] (LineGraph.h:388) / mFinalVal = C[0] (Network.h:964).        */
int qtb_read_scalar(qtb_ctx *ctx, qtb_tensor t, double out_re_im[2]);
/* The same read in two halves, for callers that keep several networks in flight: `begin` launches everything still
 * deferred and enqueues the 16-byte device->host copy right behind the work that produces t (no synchronisation);
 * `end` waits for THAT copy only -- not for steps of other networks enqueued after it -- and releases the handle.
 * The tensor may be freed (qtb_tensor_free) between the two calls.                                                */
typedef struct qtb_scalar_read_s qtb_scalar_read;
int qtb_read_scalar_begin(qtb_ctx *ctx, qtb_tensor t, qtb_scalar_read **out);
int qtb_read_scalar_end(qtb_ctx *ctx, qtb_scalar_read *read, double out_re_im[2]);

/* ---- the hot path: replaces Network::ContractIndices (Network.h:876-971) ------------------------- */
/* C = contract(A, B) over k shared legs.  Asynchronous on the ctx stream; tiny steps may be deferred
 * and grouped into one launch until the next flush/sync/download.  rank(C) must equal
 * rank(A)+rank(B)-2k.  k = 0 is accepted (outer product; the reference only reaches it for two rank-0
 * nodes, Network.h:772).                                                                             */
/* Tensors are SINGLE-USE as operands, like the reference's nodes (mContracted, Network.h:719-723): a big tile-kernel step
 * whose result is immediately contracted with another tensor over all of its legs runs fused with that inner product and
 * its intermediate is never written -- the intermediate's handle then holds no data and any later use of it returns
 * QTB_ERR_EMPTY_INPUT.  a == b and c == a / c == b are rejected (QTB_ERR_INVALID).  An output or upload target that
 * deferred steps still read or write is ordered after them.                                                          */
int qtb_contract(qtb_ctx *ctx, qtb_tensor a, qtb_tensor b, int k,
                 const int *pos_a, const int *pos_b, qtb_tensor c);

/* ---- compiled plans: the plan executor (many ContractNodes calls -> few launches, CUDA graph) ----- */
typedef struct qtb_plan_step {
    int32_t a, b;                    /* operand tensor ids: 0..n_inputs-1 are inputs, n_inputs+i is the  */
                                     /* result of step i (the reference's mCreatedFrom numbering relative */
                                     /* to the plan, Network.h:853-857)                                   */
    int32_t k;                       /* number of shared legs                                             */
    int8_t pos_a[QTB_MAX_RANK];      /* shared leg positions in A (increasing)                            */
    int8_t pos_b[QTB_MAX_RANK];      /* matching positions in B                                           */
} qtb_plan_step;

/* Compile a plan: validates leg maps, assigns pooled device buffers with liveness-based reuse and
 * groups micro-steps.  The last step's result is the plan output (any rank).                         */
int qtb_plan_create(qtb_ctx *ctx, int n_inputs, const int *input_ranks,
                    int n_steps, const qtb_plan_step *steps, qtb_plan **out);
int qtb_plan_destroy(qtb_ctx *ctx, qtb_plan *plan);
/* Host buffers in, host result out (H2D of every input, execute, D2H of the output): the end-to-end call. */
int qtb_plan_run_host(qtb_ctx *ctx, qtb_plan *plan, const double *const *host_inputs, double *host_out);
/* Inputs already resident in the plan's device buffers (after one qtb_plan_upload_inputs); async.       */
int qtb_plan_upload_inputs(qtb_ctx *ctx, qtb_plan *plan, const double *const *host_inputs);
int qtb_plan_run_device(qtb_ctx *ctx, qtb_plan *plan);
int qtb_plan_read_output(qtb_ctx *ctx, qtb_plan *plan, double *host_out);
int qtb_plan_output_rank(qtb_plan *plan);
/* Several input sets resident in HBM at once (e.g. the per-edge measurement caps of all <ZiZj> terms):
 * stage set `slot` once, then run the plan on it with a device-to-device copy of the small-input blob.
 * Only for plans whose inputs all have rank <= 5 (gate / state / measurement tensors).                  */
int qtb_plan_stage_inputs(qtb_ctx *ctx, qtb_plan *plan, int slot, const double *const *host_inputs);
int qtb_plan_run_device_slot(qtb_ctx *ctx, qtb_plan *plan, int slot);
/* Index-sliced networks (SURVEY 8e "Slices"): the 4^s slices of one network share every step that does not touch a cut
 * wire.  The caller orders the plan so that its first n_invariant_steps steps depend only on inputs that are identical
 * in all slots (qtorch_b200/slicing.py:hoist_invariant does that); their results are kept alive for the whole call.
 * qtb_plan_run_slots then runs that prefix ONCE, the remaining steps once per listed slot, and returns the sum of the
 * scalar outputs (host_each, if not NULL, receives the n individual (re, im) pairs) with a single synchronisation.
 * n_invariant_steps = 0 gives an ordinary plan; qtb_plan_run_slots on it just runs the whole plan per slot.        */
int qtb_plan_create_sliced(qtb_ctx *ctx, int n_inputs, const int *input_ranks, int n_steps, const qtb_plan_step *steps,
                           int n_invariant_steps, qtb_plan **out);
int qtb_plan_run_slots(qtb_ctx *ctx, qtb_plan *plan, const int *slots, int n, double *host_sum, double *host_each);
/* sum of 4^(rC+k) over the invariant prefix (done once per qtb_plan_run_slots call, not once per slot) */
long long qtb_plan_prefix_units(qtb_plan *plan);

/* The sliced-amplitude executor: the same sliced plan, but an amplitude never synchronises with the host while it runs.
 * `n_lanes` (1..8) replicas of the plan (own buffers, graphs and stream) share a rank's slices, so that two big launches
 * are always queued (one slice's tile kernel takes over each SM the previous one has left) and the invariant prefix of
 * the next amplitude overlaps the slices of the current one.  qtb_sliced_begin enqueues everything and returns at once:
 * prefix per used lane, the suffix per listed slot, slot scalars accumulated on the device in a fixed order, then -- if
 * `allreduce` is non-zero (needs qtb_comm_init) -- ONE in-stream ncclAllReduce of the complex scalar (replaces
 * f_pVal += ..., maxcut.cpp:196, across ranks) and the 16-byte device->host copy.  qtb_read_scalar_end waits for that
 * copy.  n = 0 is legal (a rank that owns no slice still joins the reduction).  A slot may be re-staged only after the
 * amplitudes that used it have been read; staging goes through its own upload stream.                               */
typedef struct qtb_sliced_s qtb_sliced;
int qtb_sliced_create(qtb_ctx *ctx, int n_inputs, const int *input_ranks, int n_steps, const qtb_plan_step *steps,
                      int n_invariant_steps, int n_lanes, qtb_sliced **out);
int qtb_sliced_destroy(qtb_ctx *ctx, qtb_sliced *sliced);
int qtb_sliced_stage(qtb_ctx *ctx, qtb_sliced *sliced, int slot, const double *const *host_inputs);
int qtb_sliced_begin(qtb_ctx *ctx, qtb_sliced *sliced, const int *slots, int n, int allreduce, qtb_scalar_read **out);
int qtb_sliced_lanes(qtb_sliced *sliced);
long long qtb_sliced_units(qtb_sliced *sliced);            /* per slice, prefix included */
long long qtb_sliced_prefix_units(qtb_sliced *sliced);
int qtb_sliced_launches(qtb_sliced *sliced, int *prefix_launches);   /* kernel launches of one slice (prefix included) */
/* Grouped evaluation of n independent plans with scalar outputs (e.g. the 45 per-edge <ZiZj> networks of one QAOA
 * objective evaluation, maxcut.cpp:171-198): all inputs are uploaded, plans that consist of micro-steps only run in
 * ONE launch (one CTA -- or, for heavy plans when SMs are free, one thread-block cluster -- per plan), the n scalars come
 * back with one synchronisation.
 * host_inputs[i] is plan i's array of input pointers; host_out receives n (re, im) pairs.                       */
int qtb_plans_run_batched(qtb_ctx *ctx, qtb_plan *const *plans, int n, const double *const *const *host_inputs, double *host_out);
/* Term batches: n independent scalar plans whose inputs stay resident in HBM and of which only a few small gate tensors
 * change between evaluations -- the QAOA objective (maxcut.cpp:162-204: same per-edge networks, 2p angles per call).
 * qtb_batch_set_inputs uploads a plan's inputs once; qtb_batch_bind declares that input `input` of plan `plan` is table
 * `table` of the per-evaluation table set (n_tables tensors of rank table_rank, e.g. Rz(-gamma_l) and Rx(2 beta_l)).
 * One evaluation = one CUDA-graph launch: H2D of the tables (n_tables * 16 * 4^table_rank bytes), scatter into the bound
 * inputs, all plans (one CTA or one thread-block cluster per plan when they are grouped micro-steps only, forked streams
 * otherwise), gather of the n
 * scalars and their sum in a fixed order, optionally ONE in-stream ncclAllReduce of the sum (replaces f_pVal += ...,
 * maxcut.cpp:196), one D2H.  begin returns at once; end waits and returns the sum (over all ranks after an allreduce) and,
 * if `terms` is not NULL, this rank's n individual (re, im) pairs.                                                      */
typedef struct qtb_batch_s qtb_batch;
int qtb_batch_create(qtb_ctx *ctx, qtb_plan *const *plans, int n, int n_tables, int table_rank, qtb_batch **out);
int qtb_batch_destroy(qtb_ctx *ctx, qtb_batch *batch);
int qtb_batch_set_inputs(qtb_ctx *ctx, qtb_batch *batch, int plan, const double *const *host_inputs);
int qtb_batch_bind(qtb_ctx *ctx, qtb_batch *batch, int plan, int input, int table);
int qtb_batch_begin(qtb_ctx *ctx, qtb_batch *batch, const double *tables_re_im, int allreduce);
int qtb_batch_end(qtb_ctx *ctx, qtb_batch *batch, double sum_re_im[2], double *terms_re_im);
int qtb_batch_run(qtb_ctx *ctx, qtb_batch *batch, const double *tables_re_im, int allreduce, double sum_re_im[2], double *terms_re_im);
int qtb_batch_launches(qtb_batch *batch);          /* kernel launches inside one evaluation */
long long qtb_batch_units(qtb_batch *batch);
/* sum_steps 4^(rC+k): the reference's getNumFloatOps() contribution of this plan (Network.h:884-885). */
long long qtb_plan_units(qtb_plan *plan);
/* Number of kernel launches one qtb_plan_run_device enqueues. */
int qtb_plan_launches(qtb_plan *plan);

/* ---- multi-GPU: the scalar reduction after sharded terms / slices (replaces f_pVal += ..., maxcut.cpp:196) */
#define QTB_UNIQUE_ID_BYTES 128
int qtb_comm_unique_id(char id[QTB_UNIQUE_ID_BYTES]);                 /* rank 0, then broadcast by the host */
int qtb_comm_init(qtb_ctx *ctx, int n_ranks, int rank, const char id[QTB_UNIQUE_ID_BYTES]);
int qtb_comm_destroy(qtb_ctx *ctx);
/* In-place sum over ranks of n complex scalars held in HOST memory (staged through the ctx stream,
 * one ncclAllReduce over NVLink).                                                                     */
int qtb_allreduce_sum(qtb_ctx *ctx, double *host_re_im, int n_complex);
/* The same sum on n complex scalars that already live in DEVICE memory: enqueued on the ctx stream, returns at once. */
int qtb_allreduce_sum_device(qtb_ctx *ctx, void *device_re_im, int n_complex);

/* ---- introspection (bench.py: gpu_launches, roofline bookkeeping) --------------------------------- */
typedef struct qtb_stats {
    long long launches;          /* kernels of this library launched on the ctx stream                  */
    long long steps;             /* contraction steps executed                                           */
    long long micro_steps;       /* ... of which ran inside grouped micro-plan launches                  */
    long long units;             /* sum 4^(rC+k) over executed steps                                     */
    long long bytes_h2d, bytes_d2h;
    long long pool_bytes_reserved, pool_bytes_peak_live;
    long long tma_launches;      /* ... of the launches: tile-kernel steps whose operand tiles were fed by TMA         */
} qtb_stats;
int qtb_ctx_stats(qtb_ctx *ctx, qtb_stats *out);
int qtb_ctx_reset_stats(qtb_ctx *ctx);
/* Per-step timing trace: when enabled every non-deferred step is bracketed by CUDA events.            */
typedef struct qtb_step_trace {
    int32_t rank_a, rank_b, k, kernel;   /* kernel: 0 grouped micro-steps (k = number of steps in the group), 1 thread-per-output,
                                            2 DMMA tile kernel (k_gett), 3 warp-per-output, 5 split-K / inner product,
                                            6 DMMA tile kernel fused with the inner product that follows it,
                                            7 streaming kernel (big tensor x tensor of <= 64 elements)                    */
    float ms;
} qtb_step_trace;
/* Largest step (4^n complex multiply-adds, n = 0..10) that may ride in a grouped micro-step launch.  A micro-step runs
 * on one SM, so the limit trades launch count against single-SM load bandwidth: default 6; callers that evaluate many
 * independent plans side by side (qtb_plans_run_batched, qtb_batch_*) raise it (up to 10) before creating their plans.                 */
int qtb_ctx_set_micro_limit(qtb_ctx *ctx, int log4_units);
int qtb_ctx_get_micro_limit(qtb_ctx *ctx);
/* CUDA-event stopwatch on the ctx stream (bench.py times the hot path with it): start flushes deferred work
 * and records an event; stop flushes, records, waits and returns the elapsed milliseconds.             */
int qtb_ctx_timer_start(qtb_ctx *ctx);
int qtb_ctx_timer_stop(qtb_ctx *ctx, float *ms);
int qtb_ctx_trace_enable(qtb_ctx *ctx, int on);
int qtb_ctx_trace_read(qtb_ctx *ctx, qtb_step_trace *out, int max_entries, int *n_entries);

#ifdef __cplusplus
}
#endif
#endif /* QTORCH_B200_H */

/*
 * desilike_b200 -- C ABI of the B200-native batched full-shape likelihood path.
 *
 * The reference (cosmodesi/desilike) is pure Python and has no FFI; this header is what a `backend='cuda'` branch of
 *   desilike.base.vmap                      (reference desilike/base.py:232-383)
 * would bind through ctypes.  One call evaluates N parameter points through
 *   emulator predict / Kaiser / damped-BAO+FFTLog   (reference emulators/__init__.py:408-409,471-507; theories/galaxy_clustering/
 *                                                    full_shape.py:488-500,545-550,628-634; bao.py:117-151; base.py:127-136)
 *   -> window matrix apply                          (reference observables/galaxy_clustering/window.py:459-473, :717-728)
 *   -> Gaussian chi2                                (reference likelihoods/base.py:13-17, :658-664)
 *   -> analytic marginalisation of linear params    (reference likelihoods/base.py:129-200, :314-413)
 *   -> prior                                        (reference parameter.py:1889-1897, :1994-2017)
 *
 * Plain pointers and sizes only; no torch / Python types.  All matrices are row-major FP64, index tables int32.
 * Buffers passed to the *_batch calls may live in host memory (pageable or pinned) or in device memory
 * (detected with cudaPointerGetAttributes); device buffers are used in place, host buffers are staged.
 *
 * Ownership: the caller owns every buffer it passes; dlb_plan_create copies what it needs to the device, the plan owns
 * those copies and its workspace.  A plan is immutable after creation; calls on the SAME plan are serialised
 * internally (one workspace per plan); different plans are independent.  One plan lives on one device (the current
 * CUDA device at creation).
 *
 * Errors: every entry point returns 0 on success, a negative dlb_status otherwise; dlb_last_error() gives the message
 * (thread-local).  Per-point problems never fail the call: they are reported in the `status` bitmask
 * and mapped by the host layer onto the reference's `errors='raise'|'return'|'nan'` conventions (reference base.py:193-258).
 * There is NO CPU fallback: without a usable CUDA device every compute entry point fails with DLB_ERR_CUDA.
 */
#ifndef DESILIKE_B200_H
#define DESILIKE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLB_ABI_VERSION 10

typedef enum {
  DLB_OK = 0,
  DLB_ERR_ARG = -1,      /* bad shape / NULL pointer / inconsistent descriptor */
  DLB_ERR_CUDA = -2,     /* CUDA runtime error (no device, launch failure, out of memory) */
  DLB_ERR_UNSUPPORTED = -3
} dlb_status;

/* per-point status bits (int32 per point) */
#define DLB_PT_NONFINITE 1   /* theory / chi2 not finite                      -> logposterior = NaN          */
#define DLB_PT_OUT_OF_PRIOR 2/* a sampled parameter is outside its limits     -> logposterior = -inf         */
#define DLB_PT_NOT_SPD 4     /* -Hessian of the solved parameters not SPD     -> logposterior = NaN          */

/* ---- feature generators: theta -> per-point feature segment -------------------------------------------------------
 * Every generator's segment starts with the constant pair (1.0, 0.0) -- feature 0 lets theta-dependent scalars multiply
 * constant vectors (shot-noise / broadband terms) -- followed by the kind-specific features.                            */
typedef enum {
  DLB_GEN_TAYLOR = 1,  /* monomials prod_j (theta_j - c_j)^p_tj   (reference emulators/__init__.py:471-507)               */
  DLB_GEN_MLP = 2,     /* last hidden layer of an MLP emulator (+1) (reference emulators/conversion.py:18-33)              */
  DLB_GEN_KAISER = 3,  /* Kaiser pt tables dd/dt/tt/pk11 on (ell, k) (reference full_shape.py:488-500, power_template.py:747-761) */
  DLB_GEN_BAO = 4      /* damped BAO P_ell(k) [-> FFTLog -> xi_ell(s)] (reference bao.py:117-151, base.py:127-136)          */
} dlb_gen_kind;

typedef struct {
  int32_t kind;            /* dlb_gen_kind */
  int32_t nparams;         /* number of input columns this generator reads */
  const int32_t* param_idx;/* [nparams] column index into the batch `params` (>= 0), or -1: use param_default */
  const double* param_default; /* [nparams] */
  int32_t nfeat;           /* features written per point (before padding) */
  /* kind-specific integer / double payloads (layouts documented next to each kernel in csrc/features.cu) */
  int32_t ni; const int32_t* idata;
  int64_t nd; const double* ddata;
} dlb_generator_desc;

/* ---- per-point scalar coefficients: c = scale * prod_f theta[idx_f]^pow_f  (bias monomials, reference full_shape.py:545-550) */
#define DLB_COEF_MAXF 4
typedef struct {
  double scale;
  int32_t nf;
  int32_t idx[DLB_COEF_MAXF];
  int32_t pow[DLB_COEF_MAXF];
} dlb_coef_desc;

/* ---- contraction terms: slot[s] += coef * features[foff : foff + K] . M[K, ncols] --------------------------------- */
typedef struct {
  int32_t slot;       /* 0: (whitened) residual / theory; s >= 1: theta-dependent Jacobian row of a solved parameter */
  int32_t coef;       /* index into the coefficient table, -1 for 1.0 */
  int32_t gen;        /* generator whose feature segment is contracted */
  int32_t foff;       /* offset inside that segment */
  int32_t K;          /* contraction length */
  int64_t moff;       /* offset (doubles) of M[K, ncols_white] (row-major) in `matrices_white` */
  int64_t moff_plain; /* offset (doubles) of M[K, ncols] (row-major) in `matrices_plain` */
} dlb_term_desc;

/* ---- output blocks: independent groups of data-vector columns (one per tracer when the precision is block-diagonal) */
typedef struct {
  int32_t ncols;            /* n_b */
  int32_t nslots;           /* 1 + number of theta-dependent Jacobian rows */
  int32_t nterms;
  const dlb_term_desc* terms;
  const double* matrices_white;  /* whitened (L . W . ...) term matrices, for dlb_eval_batch   */
  const double* matrices_plain;  /* un-whitened (W . ...) term matrices, for dlb_theory_batch  (may be NULL) */
  const double* r0_white;        /* [ncols] constant subtracted from slot 0:  L (data - offset - const part)  */
  const double* t0_plain;        /* [ncols] constant ADDED to slot 0 in theory mode (offset + const part)      */
  const int32_t* slot_solved;    /* [nslots - 1] solved-parameter index of slots 1.. */
  int32_t nconst;                /* constant (theta-independent) whitened Jacobian rows */
  const int32_t* const_solved;   /* [nconst] solved-parameter index */
  const double* gconst_white;    /* [nconst, ncols] */
  /* Optional observable transform (reference observables/galaxy_clustering/power_spectrum.py:400-404, 'cubic':
   * u = (3 (t / d)^(1/3) - 2) d).  It is non-linear in the theory vector, so a plan with a transformed block is evaluated in two
   * stages: t = slot 0 + t0_plain through the un-whitened matrices (required), u on the columns with transform_data != 0, then the
   * whitened residual r = u . whitening_t - r0_data_white.  Either no block or every block of a plan provides these three; solved
   * parameters are not supported together with a transform (ns must be 0). */
  const double* transform_data;  /* [ncols] flat data d of the transformed columns, 0 on untransformed columns (NULL: one-stage plan) */
  const double* whitening_t;     /* [ncols, ncols] row-major L^T:  r_c = sum_j u_j whitening_t[j, c] */
  const double* r0_data_white;   /* [ncols] L . data */
  /* Columns of the WHITENED matrices / r0_white (0: ncols).  The host may hand over a rank-reduced whitened problem: chi2 only needs
   * |r|^2, and r = g . M - r0 lives in the row space of the stacked term matrices, so a QR factorisation at plan time replaces the n
   * data-space columns by sum_t K_t + 1 columns (desilike_b200/lowering.py); requires nslots == 1 and nconst == 0. */
  int32_t ncols_white;
  /* Optional explicit meaning of the block's Gram partials (default NULL: slot 0 = residual, slots >= 1 / constant rows = Jacobian rows, every
   * pair contributes).  Entry e of the nslots (nslots + 1) / 2 + nslots * nconst partials (pairs (a, b), a <= b, row-major; then, per constant row i,
   * its dot product with every slot a) means: kind 0 chi2 += v; 1 gradient[i] -= v; 2 M[i][j] += v (and M[j][i] when i != j); 3 ignored.
   * Used by plans that split a dense joint precision over several passes of <= 8 slots (desilike_b200/lowering.py: un-whitened slots S_i next to
   * precision-weighted ones y_j = P S_j, chi2 / gradient / Hessian entries = S_i . y_j). */
  const int32_t* entry_kind; const int32_t* entry_i; const int32_t* entry_j;
  int32_t skip_const_gram;   /* != 0: the Gram of the constant rows of this block is NOT g_i . g_j (it is supplied through dlb_plan_desc::m0_extra) */
  /* Optional quadratic-form tables (q_ncols > 0; every block of the plan or none; desilike_b200/lowering.py: qform_tables).  All terms of the block
   * contract the same K features f; instead of the nterms matrices [K, ncols] the kernel multiplies f by ONE matrix q_matrix [K, q_ncols] whose columns
   * come in n-tiles of 8, one group per tile: q_tiles[3 t ..] = (group, first feature column of the tile relative to the generator segment, number of
   * valid columns); q_groups[4 g ..] = (kind, Gram entry, coefficient a, coefficient b; -1 = 1).  kind 0 (quadratic): column k' is multiplied by
   * feature k', the group's sum weighted by coef_a coef_b is added to its Gram entry.  kind 1 (direct): column j is added to Gram entry q_dent[24 g + 3 j] (-1: unused), weighted by the
   * coefficients q_dent[24 g + 3 j + 1] and [.. + 2] (-1 = 1).  q_const [npart] is added to the partials (theta-independent parts).  Same partials as the slot form. */
  int32_t q_ncols; int32_t q_ngroups; const double* q_matrix; const int32_t* q_tiles; const int32_t* q_groups; const int32_t* q_dent; const double* q_const;
  /* q_matrix is [q_K, q_ncols] over the feature columns q_foff .. q_foff + q_K - 1 of the generator segment (the NON-constant features: constant ones --
   * the (1, 0) head of an MLP segment -- are folded into direct columns and into the q_nw weighted constants: entry q_wconst_idx[3 w] += q_wconst_val[w] x
   * coefficient q_wconst_idx[3 w + 1] x coefficient q_wconst_idx[3 w + 2] (-1 = 1), sorted by entry, added once per point). */
  int32_t q_foff; int32_t q_K; int32_t q_nw; const double* q_wconst_val; const int32_t* q_wconst_idx;
} dlb_block_desc;

typedef struct {
  int32_t abi_version;      /* DLB_ABI_VERSION */
  int32_t d;                /* number of columns of the batch `params` (the sampled inputs) */
  /* prior over the d input columns (reference parameter.py:1994-2017): kind 0 none, 1 uniform, 2 normal */
  const int32_t* prior_kind; const double* prior_lo; const double* prior_hi; const double* prior_loc; const double* prior_scale;
  int32_t ngen; const dlb_generator_desc* gens;
  int32_t ncoef; const dlb_coef_desc* coefs;
  int32_t nblocks; const dlb_block_desc* blocks;
  /* solved (analytically marginalised / profiled) linear parameters, reference likelihoods/base.py:314-413.
   * Ordered with the `nmarg` '.marg' parameters FIRST, then the '.best' ones. */
  int32_t ns; int32_t nmarg;
  const double* x0;          /* [ns] value at which d0 and J are evaluated (the parameter defaults) */
  const int32_t* xprior_kind;/* [ns] 0 none (improper), 1 uniform (limits only), 2 normal */
  const double* xprior_loc; const double* xprior_scale; const double* xprior_lo; const double* xprior_hi; /* [ns] */
  /* Derived-expression parameters (reference parameter.py:795-808 Parameter.eval, :1872-1887 ParameterCollection.eval, applied per point at
   * base.py:537): parameter j = d + i of the EXTENDED parameter row is a function of the d input columns (and of earlier derived ones), e.g.
   * b1 = b**2.  Every index table above (generator param_idx, coefficient idx, the prior tables, which then have d + nderived entries) addresses
   * the extended row.  Expression i is the reverse-Polish program derived_code[2 * derived_off[i] : 2 * derived_off[i + 1]] of (opcode, argument)
   * pairs: 0 push column `argument` of the extended row (< d + i), 1 push derived_const[argument], 2 add, 3 sub, 4 mul, 5 div, 6 pow, 7 neg,
   * 8 exp, 9 log, 10 sqrt, 11 log10, 12 abs, 13 sin, 14 cos, 15 tanh (stack depth <= 16); evaluated by a small kernel in front of every pass. */
  int32_t nderived; const int32_t* derived_off; const int32_t* derived_code; const double* derived_const;
  const double* m0_extra;   /* optional [ns, ns] (plan order) added to the constant part of the normal matrix (constant-row Gram of split dense plans) */
  /* > 0 (opt-in, 2 .. 8): hidden layers of MLP emulators with fan-in >= 64 run as FP64-grade GEMMs on the tcgen05 int8 tensor cores with this many
   * exact 7-bit slices per operand (csrc/ozaki.cu; 6 slices ~ 1e-11 relative, |delta chi2| << 1e-4 on the 5-tracer configuration); 0: FP64 DMMA. */
  int32_t mlp_slices;
} dlb_plan_desc;

typedef struct dlb_plan dlb_plan;

/* Library / device ----------------------------------------------------------------------------------------------- */
int dlb_abi_version(void);
const char* dlb_last_error(void);
int dlb_device_count(void);

/* Plan -------------------------------------------------------------------------------------------------------------- */
int dlb_plan_create(const dlb_plan_desc* desc, dlb_plan** out);
void dlb_plan_destroy(dlb_plan* plan);
int dlb_plan_ntheory(const dlb_plan* plan);      /* total data-vector length n */
int dlb_plan_nsolved(const dlb_plan* plan);
/* which kernels the plan runs: bit 0 Taylor monomials inside the contraction kernel, 1 two-stage (observable transform), 2 fused MLP kernel,
 * 3 MLP hidden layers on the tcgen05 int8 path, 4 likelihood blocks as tcgen05 int8 GEMMs + Gram reduction (both opt-in through mlp_slices),
 * 5 likelihood blocks evaluated as quadratic forms of their features (dlb_block_desc::q_*) */
int dlb_plan_flags(const dlb_plan* plan);
/* FP64 tensor-pipe FLOPs per evaluation EXECUTED by the contraction launch(es) of dlb_eval_batch: 2 * k-chunk length * tile width over
 * the (column block, k-chunk) tiles actually multiplied (zero padding inside tiles included, skipped leading zero tiles excluded).
 * Diagnostic for bench.py: the gap to the algorithmic count is padding. */
double dlb_plan_executed_flops(const dlb_plan* plan);

/* Batched evaluation (replaces the per-point loop of reference base.py:239-273 + likelihoods/base.py:242-245,314-413).
 * params [N, d].  Outputs (each may be NULL): logposterior/loglikelihood/logprior [N], x_solved [N, ns] (in the plan's solved
 * order), status [N].  `stream` is a cudaStream_t (NULL = default stream); the call returns after the work is
 * ENQUEUED when every buffer is device memory, and after completion when any buffer is host memory. */
int dlb_eval_batch(dlb_plan* plan, const double* params, int64_t N,
                   double* logposterior, double* loglikelihood, double* logprior,
                   double* x_solved, int32_t* status, void* stream);

/* Same, additionally returning (hessian may be NULL) the NEGATIVE Hessian of the log-posterior in the solved parameters at every point,
 * M = diag(1 / scale^2) + J~ J~^T (reference likelihoods/base.py:165-187: -posterior_hessian; its entries are what the reference
 * exports as derivs-tagged derived values, :342-352, :388-390, and what Chain.sample_solved draws from): packed upper triangle,
 * row-major over the plan's solved order, [N, ns (ns + 1) / 2]. */
int dlb_eval_batch_hessian(dlb_plan* plan, const double* params, int64_t N,
                           double* logposterior, double* loglikelihood, double* logprior,
                           double* x_solved, double* hessian, int32_t* status, void* stream);

/* Theory vectors (parity / debug entry; reference `likelihood.flattheory`, likelihoods/base.py:662-664): theory [N, n],
 * evaluated at the solved parameters' defaults x0. */
int dlb_theory_batch(dlb_plan* plan, const double* params, int64_t N, double* theory, void* stream);

/* Per-point features of one generator [N, nfeat] (parity / debug entry for the theory kernels, e.g. Kaiser tables). */
int dlb_features_batch(dlb_plan* plan, int32_t gen, const double* params, int64_t N, double* features, void* stream);

/* One iteration of the affine-invariant ensemble sampler (stretch move, red-blue split) with walkers resident on the device:
 * replaces one `emcee.EnsembleSampler.sample` iteration driving the vectorised log-posterior (reference desilike/samplers/emcee.py:68,
 * :100-110; desilike/samplers/base.py:144-200).  walkers [nwalkers, d] and logposterior [nwalkers] are DEVICE buffers updated in
 * place (logposterior must hold the current values, e.g. from dlb_eval_batch); the first nwalkers / 2 rows are one half, the rest
 * the other.  Randomness: Philox4x32-10 keyed by `seed`, counter (rank_offset + walker index in its half, iteration, half):
 * reproducible for any launch geometry; ranks sharding an ensemble pass distinct rank_offset.  `a` is the stretch scale (emcee
 * default 2).  naccept (device uint64, may be NULL) is incremented by the number of accepted proposals.  Work is enqueued on
 * `stream`; no host synchronisation. */
int dlb_ensemble_step(dlb_plan* plan, double* walkers, double* logposterior, int64_t nwalkers, uint64_t seed, uint64_t iteration, double a,
                      int64_t rank_offset, uint64_t* naccept, void* stream);

/* Multi-GPU result exchange over NVLink peer memory ---------------------------------------------------------------------------------
 * The path shards by parameter point (reference desilike/base.py:310-316, partition rule mpi.py:145-150); its only exchange is returning
 * the per-point results to every rank (reference base.py:335 pickled reduce + bcast; samplers/base.py:196-200).  One process per GPU.
 * Every rank creates a *window* (device memory: `nbuf` buffers of `capacity` doubles + flag words), exports its CUDA IPC handle
 * (DLB_IPC_HANDLE_BYTES bytes; the host exchanges the handles with whatever it has -- torch.distributed, MPI, a file) and connects.
 * dlb_eval_batch_gather is dlb_eval_batch for this rank's slice [slice_start, slice_start + N) of a batch of `total` points whose finish
 * kernel also stores logposterior (and, with_solved != 0, x_solved) straight into EVERY rank's window (buffer `buf`: logposterior [total],
 * then x_solved [total, ns] in the plan's solved order) and publishes a per-buffer sequence number; no collective kernel runs.  Every rank
 * must issue the same sequence of publishing calls per buffer (collective semantics).  dlb_gather_wait enqueues the acquire of all ranks'
 * sequence numbers on `stream`: kernels enqueued after it see the complete buffer (dlb_gather_buffer).  A rank that does not publish within
 * the timeout (default 20 s) makes later calls fail with DLB_ERR_CUDA instead of hanging the GPU.  Use one stream per gather. */
typedef struct dlb_gather dlb_gather;
#define DLB_IPC_HANDLE_BYTES 64
int dlb_gather_create(int32_t rank, int32_t world, int64_t capacity, int32_t nbuf, dlb_gather** out);
int dlb_gather_ipc_handle(const dlb_gather* gather, void* handle /* [DLB_IPC_HANDLE_BYTES] */);
int dlb_gather_connect(dlb_gather* gather, const void* handles /* [world, DLB_IPC_HANDLE_BYTES], rank-major */);
double* dlb_gather_buffer(dlb_gather* gather, int32_t buf);   /* this rank's copy of buffer `buf` (device pointer, [capacity]) */
int dlb_eval_batch_gather(dlb_plan* plan, const double* params, int64_t N, dlb_gather* gather, int32_t buf, int64_t slice_start, int64_t total,
                          int32_t with_solved, double* loglikelihood, double* logprior, double* x_solved, int32_t* status, int32_t wait_buf, void* stream);
/* loglikelihood / logprior / x_solved [N, ns] / status: optional LOCAL outputs of this rank's own points (device or host; host buffers are filled by
 * the pipelined copy-out of the sub-batches, overlapping the evaluation). */
/* wait_buf >= 0: additionally enqueue dlb_gather_wait(gather, wait_buf) behind the evaluation (one call per step of a pipelined loop: the acquire
 * of the previous step's buffer overlaps this step's NVLink stores). */
/* Generic form: store src[0 : count] (device) at [offset, offset + count) of buffer `buf` of every rank's window and publish. */
int dlb_gather_publish(dlb_gather* gather, int32_t buf, const double* src, int64_t count, int64_t offset, void* stream);
int dlb_gather_wait(dlb_gather* gather, int32_t buf, void* stream);
int dlb_gather_status(const dlb_gather* gather);   /* 0, or 1 + the first rank a wait timed out on */
int dlb_gather_set_timeout(dlb_gather* gather, double seconds);
void dlb_gather_destroy(dlb_gather* gather);

/* FP64-grade GEMM on the tcgen05 int8 tensor cores (csrc/ozaki.cu; exact integer slicing, `slices` in [2, 8] signed 7-bit slices per operand):
 * out [rows, N] = act(X [rows, K] . W [K, N] + bias [N]), act 0 none / 1 silu / 2 relu / 3 tanh; HOST buffers (staged by the call).  The parity /
 * debug entry of the kernel pair the plan uses for wide MLP hidden layers when dlb_plan_desc::mlp_slices > 0 (reference emulators/conversion.py:
 * 18-33: v <- act(v @ kernel + bias)). */
int dlb_ozaki_matmul(const double* X, int64_t rows, int32_t K, const double* W, const double* bias, int32_t N, int32_t slices, int32_t act, double* out);

/* Instrumentation ----------------------------------------------------------------------------------------------- */
int64_t dlb_launch_count(void);                  /* kernels launched by this library since load (all plans) */
/* Achieved FP64 rate (TFLOP/s) of a register-resident DMMA.8x8x4 loop (kind 0) or DFMA loop (kind 1) on the current
 * device: the denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 entry).  kind = 100 + w: DMMA loop with w
 * resident warps per SM; kind = 1000 + 100 * tile + 10 * flags + w / 4: DMMA loop fed from shared memory like the
 * contraction kernel's main loop (tile 0: 2x7, 1: 4x7, 2: 2x4, 3: 1x7 m8n8 tiles per warp; w warps per SM); kind = 2000 + 100 * ilp + w:
 * register-resident DMMA loop with `ilp` independent accumulators per warp (dependent chains; ilp in {1, 2, 4, 6, 10, 14}), w warps per SM. */
int dlb_measure_fp64_peak(int32_t kind, double* tflops);
/* Time (ms, CUDA events on `stream`) of the last dlb_eval_batch's contraction kernel launches; for bench.py's roofline. */
int dlb_last_kernel_ms(const dlb_plan* plan, double* features_ms, double* contract_ms, double* solve_ms);
int dlb_set_timing(dlb_plan* plan, int32_t enabled);
/* Time (ms) of the most recent FFTLog kernel launch (part of the `features` phase) while timing is enabled. */
int dlb_last_fftlog_ms(double* ms);

#ifdef __cplusplus
}
#endif
#endif /* DESILIKE_B200_H */

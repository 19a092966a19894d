// Internal declarations shared by the kernels and the C-ABI layer of libdesilike_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <algorithm>

#include "../../include/desilike_b200.h"

namespace dlb {

// ---------------------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;
extern std::atomic<bool> g_timing;     // per-kernel CUDA-event timing requested (dlb_set_timing)
int last_fftlog_ms(double* ms);

#define DLB_CUDA(call)                                                                                      \
  do {                                                                                                      \
    cudaError_t _e = (call);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      dlb::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return DLB_ERR_CUDA;                                                                                  \
    }                                                                                                       \
  } while (0)

#define DLB_TRY(call)                 \
  do {                                \
    int _rc = (call);                 \
    if (_rc) return _rc;              \
  } while (0)

#define DLB_REQUIRE(cond, msg)                                              \
  do {                                                                      \
    if (!(cond)) { dlb::set_error(std::string("invalid argument: ") + msg); return DLB_ERR_ARG; } \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// tiling constants of the contraction kernel (csrc/contract.cu)
// ---------------------------------------------------------------------------------------------------------------------
// k-chunk length == shared-memory row stride (doubles) of the B tiles is a per-variant constant (Variant::KC): 28 or 36, both = 4 (mod 8)
// -> conflict-free 64-bit fragment loads
constexpr int ROW_ALIGN = 384;  // batch rows are padded to a multiple of every variant's row unit (64, 48) and tile

struct Variant { int R, MT, NT, WM, CPS, KC; }; // accumulator slots, m-tiles (of 8 rows) and n-tiles (of 8 columns) per warp; warp grid is WM (m) x 2 (n);
                                            // CPS = co-resident CTAs per SM; KC = k-chunk length
inline int variant_unit(const Variant& v) { return 8 * v.WM; }          // rows per unit (one m-tile of every warp row)
inline int variant_bm(const Variant& v) { return 8 * v.WM * v.MT; }
inline int variant_bn(const Variant& v) { return 16 * v.NT; }

struct Chunk {        // one k-chunk of one term (device table, one per output block); read as one int4
  int slot;           // accumulator slot
  int coef;           // coefficient index, -1 -> 1
  int fcol;           // first feature column: absolute in the per-point feature row (streaming A) or relative to the
                      // block's generator segment (resident A)
  int kc;             // chunk length (multiple of 4, <= Variant::KC)
};

constexpr int MAXP = 16;  // max input parameters of one generator

struct GenDev {       // Taylor generator constants for the in-kernel monomial evaluation (resident A)
  int nparams, maxdeg;
  double cval[MAXP + 2];   // subtracted from the factor's parameter column: center_j; [MAXP] = 0 (absent factor: column "one" - 0);
                           // fixed parameters use the constant-one column with cval = 1 - (default_j - center_j)
};

struct BlockDev {     // one output block as seen by the contraction kernel (device table)
  const Chunk* chunks;
  const double* tiles;                 // [ncb][nchunks][BN][KC] pre-tiled term matrices (n-major, zero padded)
  const double* epi;                   // [ncb][BN] epilogue constants per column block: r0 (subtracted from slot 0 in eval mode / added in
                                       // output mode); the nconst auxiliary columns ncols .. ncols + nconst - 1 carry r0 . g_i
  const unsigned long long* paths;     // [ast][2] monomial factor lists of the block's Taylor generator (resident A), else NULL:
                                       // word 0: byte k = parameter-tile column of factor k (top byte 0xFF: structural zero column),
                                       // word 1: byte k = index into GenDev::cval
  long long poff;                      // first Gram-partial row of the block: partial[(poff + e) * npad + point]
  int nchunks, ncb, nslots, nconst, npart, ncols, ocol0, gen;
  int fill, fbase, acols;              // resident A: 0 = evaluate Taylor monomials, 1 = copy acols columns of F from column fbase
  const unsigned long long* mono2;     // two-pass monomial table of the block's Taylor generator (maxdeg <= 4 and every half product of a
                                       // degree 3-4 monomial is itself a column), else NULL:  [n1] pass-1 words (columns of degree <= 2:
                                       // bytes colA, colB, cvalA, cvalB, then the column as 16 bits, top byte != 0: structural zero) followed
                                       // by [n2] pass-2 words (16-bit column of the left half, of the right half, of the product)
  int n1, n2;
  const int* cb_first;                 // [ncb] first k-chunk with a non-zero tile in each column block (leading all-zero tiles of a
                                       // rank-reduced, lower-triangular problem are neither loaded nor multiplied), or NULL
  const unsigned long long* tmask;     // [ncb][nchunks][2] per (column block, k-chunk, warp of the pair): up to 8 runs of k4 steps, 8 bits each (5 bits
                                       // length, 3 bits = number of leading n-tiles of that warp with non-zero 4 x 8 sub-tiles); ~0: everything is multiplied; or NULL
  // quadratic-form blocks (dlb_block_desc::q_*; contract.cu QF epilogue): per 8-column n-tile a 16-byte record (group, valid columns, kind, first feature
  // column, Gram entry, coefficients), padded to ncb * BN / 8 tiles with group -1; direct groups: Gram entry of each of their 8 columns
  const int* qtile; const int* qdent; const double* qconst;   // qtile: one int4 record per n-tile (contract.cu QF epilogue)
  const int* qw_off; const int* qw_coef; const double* qw_val;  // weighted constants of Gram entry e: items qw_off[e] .. qw_off[e + 1] - 1, value x coef[2 t] x coef[2 t + 1]
  int oslot;                           // output mode of a transformed multi-slot block: column stride between the slots' outputs (0: slot 0 only)
  const double* tdata;                 // output mode: [ncols] data d of the 'cubic'-transformed columns (0: untransformed column), or NULL
};

struct DevCoef { double scale; int nf; int idx[DLB_COEF_MAXF]; int pow[DLB_COEF_MAXF]; };

struct SchedEntry {   // one TMA transaction of the per-tile chunk sequence (k-chunks and epilogue constants of every block / column block)
  const double* src;
  unsigned bytes;
  int chunk;          // streaming A: index of the chunk in its block (-1 for epilogue constants), block in the high 16 bits
};

struct ContractParams {
  const BlockDev* blocks; int nblocks;
  const SchedEntry* sched; int nsched;
  const GenDev* gens; int ngen;
  const DevCoef* coefs; int ncoef;
  const double* params; int d; long long N;
  const double* F; int fstride;          // features [npad, fstride] (streaming A)
  double* partial; long long npad;       // Gram partials, summed over column blocks
  double* out; long long ostride; long long out_rows;  // output mode: out[row * ostride + ocol0 + col]
  int act;                               // output mode: activation applied to (slot 0 + r0): 0 none, 1 silu, 2 relu, 3 tanh
  int mode;                              // 0 eval (Gram partials), 1 output (theory vectors / MLP layer)
  int nstages, resident, ast, dpad, npart_max;
  int masked;                            // some block of the launch has sub-tile masks (BlockDev::tmask): use the kernel instances with the masked path
  int qform;                             // quadratic-form blocks (BlockDev::qtile): the QF kernel instance
  int qf_offset_ns;                      // 2 CTAs / SM: start delay of the second-resident CTA (phase offset between the two)
  long long U;                           // number of row units (variant_unit rows each) to process
  long long* prof;                       // DLB_PROFILE builds: 8 cycle counters (summed over warps)
};

#ifdef __CUDACC__
// exp(x) for |x| <= 708 with ~1e-15 relative error in ~18 FP64 instructions (the activation epilogue of the MLP layers is bound by
// the FP64 pipe: CUDA's exp() + IEEE division cost ~60 FP64 instructions per element): k = rint(x log2 e) by the magic-number
// trick, r = x - k ln 2 in two steps, degree-11 Taylor polynomial of e^r on |r| <= ln 2 / 2 (remainder 6e-15), 2^k by exponent add.
__device__ __forceinline__ double fast_exp(double x) {
  const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(-kf, 6.93147180369123816490e-01, x);
  r = fma(-kf, 1.90821492927058770002e-10, r);
  double p = 2.50521083854417187751e-08;            // 1 / 11!
  p = fma(p, r, 2.75573192239858906526e-07);        // 1 / 10!
  p = fma(p, r, 2.75573192239858906526e-06);        // 1 / 9!
  p = fma(p, r, 2.48015873015873015873e-05);        // 1 / 8!
  p = fma(p, r, 1.98412698412698412698e-04);        // 1 / 7!
  p = fma(p, r, 1.38888888888888888889e-03);        // 1 / 6!
  p = fma(p, r, 8.33333333333333333333e-03);        // 1 / 5!
  p = fma(p, r, 4.16666666666666666667e-02);        // 1 / 4!
  p = fma(p, r, 1.66666666666666666667e-01);        // 1 / 3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.);
  p = fma(p, r, 1.);
  return __longlong_as_double(__double_as_longlong(p) + ((long long)k << 52));
}
// 1 / y for finite y >= 1: hardware approximation (MUFU.RCP64H, ~20 bits) + two Newton steps
__device__ __forceinline__ double fast_rcp(double y) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  r = r * fma(-y, r, 2.);
  r = r * fma(-y, r, 2.);
  return r;
}
#endif

// Function attributes (opt-in dynamic shared memory) are per device: the "already configured" caches of the launchers are indexed by device
// so that one process may drive several GPUs (one thread + one plan per GPU).
inline int current_device_slot() { int dev = 0; cudaGetDevice(&dev); return dev & 63; }

size_t contract_smem_limit(int ctas_per_sm);
size_t contract_smem_bytes(const Variant& v, int nstages, bool resident, int ast, int dpad, int ncoef, int npart_max, int ngen);
int launch_contract(const Variant& v, const ContractParams& p, int nsm, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------------
// feature generators (csrc/features.cu)
// ---------------------------------------------------------------------------------------------------------------------
struct DevGen {
  int kind, nparams, nfeat, fpad, foff;
  int pidx[MAXP];
  double pdef[MAXP];
  const int* idata; int ni;
  const double* ddata; long long nd;
  int aux[8];             // kind-specific layout offsets computed at plan time
};

int launch_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);
int launch_theory_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);  // csrc/theory_kernels.cu
int launch_derived(const double* params, int d_in, int nder, const int* off, const int* code, const double* consts, long long N, double* out,
                   cudaStream_t stream);   // csrc/features.cu: extended parameter rows
int launch_mlp_input(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);
// fused MLP forward (csrc/mlp.cu)
size_t mlp_fused_smem(const int* idata_host, int groups);
int launch_mlp_fused(const DevGen* d_gens, const int* d_mlp_index, int nmlp, size_t smem, int groups, const double* params, int d, long long N, long long npad,
                     double* F, int fstride, int nsm, cudaStream_t stream);

// FP64-grade GEMM on tcgen05 int8 tensor cores (csrc/ozaki.cu): wide MLP hidden layers, opt-in
size_t ozaki_slice_bytes(long long rows_pad, int K, int ns);
long long ozaki_rows_pad(long long rows);
void ozaki_slice_weights(const double* W, int K, int N, int ns, std::vector<int8_t>& slices, std::vector<int>& exps);
int launch_ozaki_layer(const double* X, long long rows, int K, long long ld, const int8_t* Bslices, const double* Sb, const double* bias, int ncols, int ns, int act,
                       double* out, long long ldo, int8_t* work_slices, int* work_exps, int nsm, cudaStream_t stream);
int launch_ozaki_slice(const double* X, long long rows, int K, long long ld, int ns, int8_t* work_slices, int* work_exps, cudaStream_t stream);
int launch_ozaki_gemm(long long rows, int K, const int8_t* Aslices, const int* Ea, const int8_t* Bslices, const double* Sb, const double* bias, int ncols, int ns, int act,
                      double* out, long long ldo, int nsm, cudaStream_t stream);
constexpr int OZ_GRAM_MAXT = 16;
struct OzakiGramParams {   // chi2 / normal-equation reduction of a GEMM output D[point, term * ncolx + col] (csrc/ozaki.cu: ozaki_gram_kernel)
  const double* D; long long ldD; long long N, npad;
  const double* params; int d;
  const DevCoef* coefs;
  int nterms, ncols, nconst, ncolx, R;
  int tslot[OZ_GRAM_MAXT], tcoef[OZ_GRAM_MAXT];
  const double* r0;        // [ncolx] whitened data over the data columns, r0 . g_i over the auxiliary columns
  double* partial; long long poff;
};
int launch_ozaki_gram(const OzakiGramParams& p, cudaStream_t stream);
int launch_mlp_small_layer(const double* X, long long rows, int K, long long ld, const double* W, const double* bias, int ncols, int act, double* out, long long ldo,
                           cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------------
// prior + marginalisation solve (csrc/solve.cu)
// ---------------------------------------------------------------------------------------------------------------------
struct EntryMap { int kind; int i; int j; };   // kind 0: chi2 += v; 1: g[i] -= v; 2: M[i][j] += v (and M[j][i])
struct SolveBlock { long long poff; int ncb; int npart; const EntryMap* map; };
// thread-per-point solve (small independent groups): Gram-partial row -> entry of its group's local system
struct GroupEntry { long long prow; int kind; int li; int lj; };   // kind 1: g[li] -= v; 2: M[li][lj] += v (li >= lj)

// Result exchange over NVLink peer memory (csrc/gather.cu): when world > 0 the finish kernels also store logposterior (and x_solved) of
// their points straight into every rank's gather window, and the last thread block to finish publishes `seq` to every rank's flag word.
constexpr int DLB_MAX_RANKS = 16;
struct GatherDev {
  int world;                               // 0: no gather
  int signal;                              // publish the sequence number at the end of this launch (last launch of an evaluation)
  int with_solved;
  int xs_only;                             // >= 0: x_solved is stored into this rank's window only (the peers get it by DMA afterwards, gather_push_xs); -1: every rank
  double* lp[DLB_MAX_RANKS];               // every rank's window, at this launch's first point: logposterior [n]
  double* xs[DLB_MAX_RANKS];               // ... x_solved [n, ns] (plan order), or NULL
  unsigned long long* flag[DLB_MAX_RANKS]; // this rank's flag word in every rank's window
  unsigned long long seq;
  unsigned* counter;                       // thread blocks that have finished (local device memory)
};

struct SolveParams {
  const double* params; int d; long long N; long long npad;
  const int* prior_kind; const double* prior_lo; const double* prior_hi; const double* prior_loc; const double* prior_scale;
  const double* partial;
  const SolveBlock* blocks; int nblocks;
  int ns, nmarg;
  const double* M0;       // [ns, ns] prior precision + constant-row Gram
  const double* g0;       // [ns] prior gradient at x0
  const double* gL0;      // [ns] likelihood gradient part that is constant (none at present: zeros)
  const double* x0; const double* xprec; const double* xloc; const double* xlo; const double* xhi; const int* xkind;
  // independent groups of solved parameters (connected components over the likelihood blocks); ncomp = 0 -> one dense problem
  int ncomp; const int* comp_off; const int* comp_idx; const int* comp_nmarg;
  int npos; const int* comp_pos; const double* comp_m0;   // intra-group matrix positions (i * (ns + 1) + j) and their constant part
  double* logpost; double* loglike; double* logprior; double* xsolved; int* status;
  double* hess; int nhess; const int* hess_ij; const int* comp_of;   // optional packed upper triangle of M = -H_posterior per point
  // thread-per-point form (every group <= 8 parameters): entries sorted by group, chi2 rows, offsets of the groups' constant blocks in comp_m0
  const GroupEntry* gent; const int* gent_off; const long long* chi_rows; int nchi; const int* m0_off; int comp_max;
  GatherDev gather;
};
int launch_solve(const SolveParams& p, cudaStream_t stream);
// csrc/gather.cu
int gather_fill(dlb_gather* g, int buf, long long start, long long n, long long total, int ns, bool with_solved, bool signal, GatherDev* out);
void gather_next_sequence(dlb_gather* g, int buf);
int gather_publish_empty(dlb_gather* g, int buf, long long total, cudaStream_t stream);
// DMA form of the x_solved exchange (large payloads): the finish kernel stores x_solved into the OWN window only and does not publish; after the
// evaluation gather_push_xs copies the rank's slice to every peer's window with the copy engines on a side stream (no SM, no NVLink stores stalling the
// next evaluation's kernels) and publishes the sequence number behind the copies.  gather_before_write makes `stream` wait for the previous push out of
// the same buffer.
bool gather_use_dma(const dlb_gather* g, long long n, int ns, bool with_solved);
int gather_rank(const dlb_gather* g);
int gather_before_write(dlb_gather* g, int buf, cudaStream_t stream);
int gather_push_xs(dlb_gather* g, int buf, long long start, long long n, long long total, int ns, cudaStream_t stream);

int measure_fp64_peak(int kind, double* tflops);

// ensemble sampler kernels (csrc/sampler.cu)
int launch_stretch_propose(const double* X, int d, long long n0, long long n1, int half, uint64_t seed, uint64_t iteration, double a,
                           long long rank_offset, double* Y, cudaStream_t stream);
int launch_stretch_accept(double* X, double* logp, const double* Y, const double* logpy, int d, long long n0, long long n1, int half, uint64_t seed,
                          uint64_t iteration, double a, long long rank_offset, unsigned long long* naccept, cudaStream_t stream);

}  // namespace dlb

// Internal declarations shared by the kernels and the C-ABI layer of libdesilike_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>

#include "../../include/desilike_b200.h"

namespace dlb {

// ---------------------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

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
constexpr int KC = 36;        // k-chunk length == shared-memory row stride in doubles; 36 = 4 (mod 8) -> conflict-free 64-bit fragment loads
constexpr int NCONSUMER_WARPS = 8;
constexpr int NTHREADS = (NCONSUMER_WARPS + 1) * 32;  // + 1 TMA producer warp
constexpr int ROW_ALIGN = 128;  // batch rows are padded to a multiple of the largest point tile

struct Variant { int R, MT, NT; };          // accumulator slots, m-tiles and n-tiles (of 8) per warp; warp grid is 4 (m) x 2 (n)
inline int variant_bm(const Variant& v) { return 4 * v.MT * 8; }
inline int variant_bn(const Variant& v) { return 2 * v.NT * 8; }

struct Chunk {        // one k-chunk of one term (device table, one per output block)
  int slot;           // accumulator slot
  int coef;           // coefficient column, -1 -> 1
  int fcol;           // first feature column (absolute, in the per-point feature row)
  int kc;             // chunk length (multiple of 4, <= KC)
};

struct ContractParams {
  const double* F; int fstride;          // features [Npad, fstride]
  const double* C; int cstride;          // coefficients [Npad, cstride]
  const Chunk* chunks; int nchunks;
  const double* tiles;                   // [ncb][nchunks][BN][KC] pre-tiled term matrices (n-major, zero padded)
  const double* r0;                      // [ncb * BN] subtracted from (eval) / added to (theory) slot 0
  const double* gconst; int nconst;      // [nconst][ncb * BN] constant whitened Jacobian rows
  int nslots;                            // slots actually used (<= R)
  int npart;                             // partial sums per (point, column block)
  double* partial; long long npad;       // [(poff + cb * npart + e) * npad + p]
  long long poff;
  double* out; long long ostride; int ocol0; int ncols;  // output mode: out[row * ostride + ocol0 + col] for col < ncols, row < out_rows
  long long out_rows;
  int act;                               // output mode: activation applied to (slot 0 + r0): 0 none, 1 silu, 2 relu, 3 tanh
  int mode;                              // 0 eval (Gram partials), 1 output (theory vectors / MLP layer)
  int nstages;
};

size_t contract_smem_bytes(const Variant& v, int nstages, int nchunks, int npart);
int launch_contract(const Variant& v, const ContractParams& p, int ntiles, int ncb, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------------
// feature generators (csrc/features.cu)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int MAXP = 16;  // max input parameters of one generator
struct DevGen {
  int kind, nparams, nfeat, fpad, foff;
  int pidx[MAXP];
  double pdef[MAXP];
  const int* idata; int ni;
  const double* ddata; long long nd;
  int aux[8];             // kind-specific layout offsets computed at plan time
};

struct DevCoef { double scale; int nf; int idx[DLB_COEF_MAXF]; int pow[DLB_COEF_MAXF]; };

int launch_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);
int launch_theory_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);  // csrc/theory_kernels.cu
int launch_mlp_input(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream);
int launch_coefs(const DevCoef* coefs, int ncoef, const double* params, int d, long long N, long long npad, double* C, int cstride, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------------
// prior + marginalisation solve (csrc/solve.cu)
// ---------------------------------------------------------------------------------------------------------------------
struct EntryMap { int kind; int i; int j; };   // kind 0: chi2 += v; 1: g[i] -= v; 2: M[i][j] += v (and M[j][i])
struct SolveBlock { long long poff; int ncb; int npart; const EntryMap* map; };

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
  double* logpost; double* loglike; double* logprior; double* xsolved; int* status;
};
int launch_solve(const SolveParams& p, cudaStream_t stream);

int measure_fp64_peak(int kind, double* tflops);

}  // namespace dlb

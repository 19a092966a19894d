// Device-resident ensemble MCMC step (affine-invariant stretch move, red-blue split) around the batched likelihood.
//
// The reference drives the batched likelihood from emcee's vectorised ensemble sampler
//   (reference desilike/samplers/emcee.py:68 `emcee.EnsembleSampler(nwalkers, ndim, self.logposterior, vectorize=True)`, :100-110;
//    desilike/samplers/base.py:144-200 `logposterior(values[N, nparams])`):
// per iteration emcee's default StretchMove splits the walkers in two halves and, for each half, proposes
//     y = c + z (x - c),   c a random walker of the OTHER half,   z ~ g(z) ∝ 1/sqrt(z) on [1/a, a]  (a = 2),
// evaluates log-posterior(y) for the whole half in one vectorised call and accepts with probability
//     min(1, z^(d-1) p(y) / p(x))                                                        (Goodman & Weare 2010).
// emcee is a third-party package absent from this image; the move is restated from its published definition
// (oracle/sampler.py is the NumPy restatement with the same counter-based random stream).
// Here walkers, log-posteriors and proposals never leave the GPU: propose kernel -> dlb_eval_batch on device buffers ->
// accept kernel, all on the caller's stream.  Randomness is Philox4x32-10 keyed by (seed) with counter
// (walker, iteration, half, purpose): results do not depend on the launch geometry.
#include "common.h"
#include <math_constants.h>

namespace dlb {

namespace {

struct Philox { uint32_t r[4]; };

__host__ __device__ inline Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox out;
  out.r[0] = c0; out.r[1] = c1; out.r[2] = c2; out.r[3] = c3;
  return out;
}

// 53-bit uniform in [0, 1)
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) { return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1. / 9007199254740992.); }

struct MoveRandoms { double z; double uacc; long long partner; };

// randoms of walker `i` (global index inside its half) at (iteration, half): stretch factor, acceptance uniform, partner index in [0, nc)
__device__ inline MoveRandoms move_randoms(uint64_t seed, uint64_t iteration, int half, long long i, long long nc, double a) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint64_t code = iteration * 2ull + (uint64_t)half;
  const Philox p = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)code, (uint32_t)(code >> 32) & 0x7FFFFFFFu, k0, k1);
  const Philox q = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)code, ((uint32_t)(code >> 32) & 0x7FFFFFFFu) | 0x80000000u, k0, k1);
  MoveRandoms m;
  const double t = (a - 1.) * u01(p.r[0], p.r[1]) + 1.;
  m.z = t * t / a;
  m.uacc = u01(p.r[2], p.r[3]);
  long long c = (long long)(u01(q.r[0], q.r[1]) * (double)nc);
  m.partner = c < nc ? c : nc - 1;
  return m;
}

// proposals of one half: Y[i, :] = X[c, :] + z (X[j, :] - X[c, :]), j = walker i of the active half, c = partner in the other half
__global__ void stretch_propose_kernel(const double* __restrict__ X, int d, long long n0, long long n1, int half, uint64_t seed, uint64_t iteration,
                                       double a, long long rank_offset, double* __restrict__ Y) {
  const long long na = half == 0 ? n0 : n1, nc = half == 0 ? n1 : n0;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = idx / d;
  const int j = (int)(idx - i * d);
  if (i >= na) return;
  const MoveRandoms m = move_randoms(seed, iteration, half, rank_offset + i, nc, a);
  const long long row = (half == 0 ? 0 : n0) + i, crow = (half == 0 ? n0 : 0) + m.partner;
  const double x = X[row * d + j], c = X[crow * d + j];
  Y[i * d + j] = c + m.z * (x - c);
}

// accept / reject: log u < (d - 1) log z + logp(y) - logp(x); non-finite logp(y) is a rejection (reference samplers/base.py:187-193
// maps NaN to -inf before the sampler sees it)
__global__ void stretch_accept_kernel(double* __restrict__ X, double* __restrict__ logp, const double* __restrict__ Y, const double* __restrict__ logpy,
                                      int d, long long n0, long long n1, int half, uint64_t seed, uint64_t iteration, double a, long long rank_offset,
                                      unsigned long long* __restrict__ naccept) {
  const long long na = half == 0 ? n0 : n1, nc = half == 0 ? n1 : n0;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool acc = false;
  if (i < na) {
    const MoveRandoms m = move_randoms(seed, iteration, half, rank_offset + i, nc, a);
    const long long row = (half == 0 ? 0 : n0) + i;
    const double lpy = logpy[i], lpx = logp[row];
    if (isfinite(lpy)) {
      const double lnr = (double)(d - 1) * log(m.z) + lpy - lpx;
      acc = log(m.uacc) < lnr;
    }
    if (acc) {
      for (int j = 0; j < d; ++j) X[row * d + j] = Y[i * d + j];
      logp[row] = lpy;
    }
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, acc);
  if ((threadIdx.x & 31) == 0 && ballot && naccept) atomicAdd(naccept, (unsigned long long)__popc(ballot));
}

}  // namespace

int launch_stretch_propose(const double* X, int d, long long n0, long long n1, int half, uint64_t seed, uint64_t iteration, double a,
                           long long rank_offset, double* Y, cudaStream_t stream) {
  const long long na = half == 0 ? n0 : n1;
  if (na == 0) return DLB_OK;
  const int threads = 256;
  stretch_propose_kernel<<<(unsigned)((na * d + threads - 1) / threads), threads, 0, stream>>>(X, d, n0, n1, half, seed, iteration, a, rank_offset, Y);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

int launch_stretch_accept(double* X, double* logp, const double* Y, const double* logpy, int d, long long n0, long long n1, int half, uint64_t seed,
                          uint64_t iteration, double a, long long rank_offset, unsigned long long* naccept, cudaStream_t stream) {
  const long long na = half == 0 ? n0 : n1;
  if (na == 0) return DLB_OK;
  const int threads = 256;
  stretch_accept_kernel<<<(unsigned)((na + threads - 1) / threads), threads, 0, stream>>>(X, logp, Y, logpy, d, n0, n1, half, seed, iteration, a,
                                                                                            rank_offset, naccept);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb

// Final per-point stage: sum the Gram partials of the contraction kernel over column blocks, add the prior, and -- when
// the likelihood has analytically solved linear parameters -- assemble and solve the normal equations.
//
// Restates, per point, reference desilike/likelihoods/base.py:129-200 (FastFisher.__call__: g, H, prior terms, solve) and
// :314-413 (BaseLikelihood._solve: logprior of the solved values, 1/2 dx^T H_L dx + g_L . dx, -1/2 log det(-H[marg, marg])),
// plus the prior of the sampled parameters (reference desilike/parameter.py:1889-1897, :1994-2017).
// With M = -H_posterior = J~ J~^T + diag(1/scale^2) (SPD), M dx = g, and the '.marg' parameters ordered first, the
// marginalisation log-determinant is 2 sum_{j < nmarg} log L_jj of the Cholesky factor M = L L^T.
// One warp per point: the matrix lives in shared memory, columns are eliminated cooperatively (warp-level batched Cholesky).
#include "common.h"
#include <math_constants.h>

namespace dlb {

namespace {

__device__ __forceinline__ double prior_logpdf(int kind, double x, double lo, double hi, double loc, double scale, bool* inside) {
  const bool isin = (lo <= x) && (x <= hi);
  if (!isin) { *inside = false; return -CUDART_INF; }
  if (kind == 2) { const double t = (x - loc) / scale; return -0.5 * t * t; }
  return 0.;
}

// ---- result stores: the caller's buffers and, when the evaluation takes part in a multi-GPU gather, every rank's window (peer memory
//      over NVLink; replaces the reduce + bcast of reference base.py:335, samplers/base.py:196-200) ----
__device__ __forceinline__ void store_logpost(const SolveParams& p, long long pt, double v) {
  if (p.logpost) p.logpost[pt] = v;
  for (int r = 0; r < p.gather.world; ++r) p.gather.lp[r][pt] = v;
}
__device__ __forceinline__ void store_xsolved(const SolveParams& p, long long pt, int i, double v) {
  if (p.xsolved) p.xsolved[pt * p.ns + i] = v;
  if (p.gather.with_solved)
    for (int r = 0; r < p.gather.world; ++r)
      if (p.gather.xs_only < 0 || r == p.gather.xs_only) p.gather.xs[r][pt * p.ns + i] = v;
}
__device__ __forceinline__ bool wants_xsolved(const SolveParams& p) { return p.xsolved != nullptr || p.gather.with_solved; }

// End of a finish kernel that takes part in a gather: make this thread's peer stores visible system-wide; the LAST thread block of the
// launch then publishes the evaluation's sequence number to this rank's flag word in every rank's window (release at system scope).
// Every thread of the block must reach this call.
__device__ __forceinline__ void gather_tail(const SolveParams& p) {
  if (p.gather.world <= 0) return;
  __threadfence_system();
  if (!p.gather.signal) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.gather.counter, 1u);
    if (prev == gridDim.x - 1) {
      *p.gather.counter = 0u;
      __threadfence_system();
      for (int r = 0; r < p.gather.world; ++r)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.gather.flag[r]), "l"(p.gather.seq) : "memory");
    }
  }
}

__device__ __forceinline__ void finish_nosolve_body(const SolveParams& p) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= p.N) return;
  double chi2 = 0.;
  for (int b = 0; b < p.nblocks; ++b) {
    const SolveBlock blk = p.blocks[b];
    for (int cb = 0; cb < blk.ncb; ++cb) chi2 += p.partial[(blk.poff + (long long)cb * blk.npart) * p.npad + pt];
  }
  double lp = 0.;
  bool inside = true;
  for (int j = 0; j < p.d; ++j)
    lp += prior_logpdf(p.prior_kind[j], p.params[pt * p.d + j], p.prior_lo[j], p.prior_hi[j], p.prior_loc[j], p.prior_scale[j], &inside);
  int st = 0;
  double ll = -0.5 * chi2;
  if (!inside) st |= DLB_PT_OUT_OF_PRIOR;
  if (!isfinite(chi2)) { st |= DLB_PT_NONFINITE; ll = CUDART_NAN; }
  if (p.loglike) p.loglike[pt] = ll;
  if (p.logprior) p.logprior[pt] = lp;
  store_logpost(p, pt, inside ? ll + lp : -CUDART_INF);
  if (p.status) p.status[pt] = st;
}

__global__ void finish_nosolve_kernel(const __grid_constant__ SolveParams p) {
  finish_nosolve_body(p);
  gather_tail(p);
}

constexpr int SOLVE_WARPS = 4;
constexpr int SOLVE_NREG = 8;   // groups of up to 8 solved parameters are factorised in registers

__device__ __forceinline__ void finish_solve_body(const SolveParams& p, double* sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = p.ns, ld = ns + 1;
  double* M = sm + (size_t)warp * (ns * ld + 4 * ns);
  double* g = M + ns * ld;
  double* gL = g + ns;
  double* dx = gL + ns;
  double* y = dx + ns;
  const long long pt = (long long)blockIdx.x * SOLVE_WARPS + warp;
  if (pt >= p.N) return;

  if (p.ncomp > 0) {   // only the intra-group entries are ever read: initialise those (prior precision + constant-row Gram)
    for (int t = lane; t < p.npos; t += 32) M[p.comp_pos[t]] = p.comp_m0[t];
  } else {
    for (int i = lane; i < ns * ns; i += 32) M[(i / ns) * ld + (i % ns)] = p.M0[i];
  }
  for (int i = lane; i < ns; i += 32) { gL[i] = 0.; }
  __syncwarp();
  double chi2 = 0.;
  for (int b = 0; b < p.nblocks; ++b) {
    const SolveBlock blk = p.blocks[b];
    for (int e = lane; e < blk.npart; e += 32) {
      double v = 0.;
      for (int cb = 0; cb < blk.ncb; ++cb) v += p.partial[(blk.poff + (long long)cb * blk.npart + e) * p.npad + pt];
      const EntryMap m = blk.map[e];
      if (m.kind == 0) chi2 += v;
      else if (m.kind == 1) gL[m.i] -= v;
      else if (m.kind == 2) { M[m.i * ld + m.j] += v; if (m.i != m.j) M[m.j * ld + m.i] += v; }
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) chi2 += __shfl_xor_sync(0xffffffffu, chi2, o);
  for (int i = lane; i < ns; i += 32) g[i] = p.g0[i] + gL[i];
  __syncwarp();

  if (p.hess) {   // -Hessian of the log-posterior in the solved parameters (prior precision + J~ J~^T), packed upper triangle, plan order
    for (int e = lane; e < p.nhess; e += 32) {
      const int ij = p.hess_ij[e], i = ij >> 16, j = ij & 0xFFFF;
      const bool coupled = p.ncomp == 0 || p.comp_of[i] == p.comp_of[j];   // entries between independent groups are exactly zero
      p.hess[pt * p.nhess + e] = coupled ? M[i * ld + j] : 0.;
    }
    __syncwarp();
  }
  bool spd = true;
  double logdet = 0.;
  double lp = 0.;
  bool inside = true;
  for (int j = lane; j < p.d; j += 32) {
    bool in_j = true;
    lp += prior_logpdf(p.prior_kind[j], p.params[pt * p.d + j], p.prior_lo[j], p.prior_hi[j], p.prior_loc[j], p.prior_scale[j], &in_j);
    inside = inside && in_j;
  }
  double quad = 0.;
  if (p.ncomp > 0) {
    // ---- block-diagonal normal equations: the solved parameters split into independent groups (connected components of
    //      the likelihood blocks sharing them -- the reference solves such groups jointly, likelihoods/base.py:60-97); one lane
    //      factorises one small group serially (Cholesky + two triangular solves), groups run in parallel across the warp ----
    for (int c = lane; c < p.ncomp; c += 32) {
      const int o = p.comp_off[c], n = p.comp_off[c + 1] - o, nm = p.comp_nmarg[c];
      const int* idx = p.comp_idx + o;
      if (n <= SOLVE_NREG) {
        // ---- register-resident factorisation (fully unrolled, guarded by the group size): no shared-memory round trips ----
        double a[SOLVE_NREG][SOLVE_NREG], yy[SOLVE_NREG];
        int id[SOLVE_NREG];
#pragma unroll
        for (int i = 0; i < SOLVE_NREG; ++i) id[i] = i < n ? idx[i] : 0;
#pragma unroll
        for (int i = 0; i < SOLVE_NREG; ++i) {
          yy[i] = i < n ? g[id[i]] : 0.;
#pragma unroll
          for (int j = 0; j <= i; ++j) a[i][j] = (i < n) ? M[id[i] * ld + id[j]] : (i == j ? 1. : 0.);
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < SOLVE_NREG; ++j) {
          const double djj = a[j][j];
          if (j < n && (!(djj > 0.) || !isfinite(djj))) ok = false;
          const double l = sqrt(djj), inv = 1. / l;
          if (j < nm) logdet += 2. * log(l);
          a[j][j] = l;
#pragma unroll
          for (int i = j + 1; i < SOLVE_NREG; ++i) a[i][j] *= inv;
#pragma unroll
          for (int i = j + 1; i < SOLVE_NREG; ++i)
#pragma unroll
            for (int k = j + 1; k <= i; ++k) a[i][k] -= a[i][j] * a[k][j];
        }
        if (!ok) { spd = false; continue; }
#pragma unroll
        for (int j = 0; j < SOLVE_NREG; ++j) {   // forward L y = g
          yy[j] /= a[j][j];
#pragma unroll
          for (int i = j + 1; i < SOLVE_NREG; ++i) yy[i] -= a[i][j] * yy[j];
        }
#pragma unroll
        for (int j = SOLVE_NREG - 1; j >= 0; --j) {   // backward L^T dx = y
          yy[j] /= a[j][j];
#pragma unroll
          for (int i = 0; i < j; ++i) yy[i] -= a[j][i] * yy[j];
        }
#pragma unroll
        for (int i = 0; i < SOLVE_NREG; ++i)
          if (i < n) dx[id[i]] = yy[i];
        continue;
      }
      bool ok = true;
      for (int j = 0; j < n && ok; ++j) {
        const int J = idx[j];
        const double djj = M[J * ld + J];
        if (!(djj > 0.) || !isfinite(djj)) { ok = false; break; }
        const double l = sqrt(djj), inv = 1. / l;
        if (j < nm) logdet += 2. * log(l);
        M[J * ld + J] = l;
        for (int i = j + 1; i < n; ++i) M[idx[i] * ld + J] *= inv;
        for (int i = j + 1; i < n; ++i) {
          const int I = idx[i];
          const double lij = M[I * ld + J];
          for (int k = j + 1; k <= i; ++k) M[I * ld + idx[k]] -= lij * M[idx[k] * ld + J];
        }
      }
      if (!ok) { spd = false; continue; }
      for (int j = 0; j < n; ++j) y[idx[j]] = g[idx[j]];
      for (int j = 0; j < n; ++j) {   // forward L y = g
        const int J = idx[j];
        const double yj = y[J] / M[J * ld + J];
        y[J] = yj;
        for (int i = j + 1; i < n; ++i) y[idx[i]] -= M[idx[i] * ld + J] * yj;
      }
      for (int j = n - 1; j >= 0; --j) {   // backward L^T dx = y
        const int J = idx[j];
        const double xj = y[J] / M[J * ld + J];
        dx[J] = xj;
        for (int i = 0; i < j; ++i) y[idx[i]] -= M[J * ld + idx[i]] * xj;
      }
    }
    spd = __all_sync(0xffffffffu, spd);
    __syncwarp();
  } else {
  // ---- Cholesky M = L L^T (lower, in place), cooperative over the warp (densely coupled solved parameters) ----
  for (int j = 0; j < ns; ++j) {
    const double djj = M[j * ld + j];
    if (!(djj > 0.) || !isfinite(djj)) { spd = false; break; }
    const double l = sqrt(djj);
    if (j < p.nmarg && lane == 0) logdet += 2. * log(l);
    const double inv = 1. / l;
    for (int i = j + 1 + lane; i < ns; i += 32) M[i * ld + j] *= inv;
    __syncwarp();
    if (lane == 0) M[j * ld + j] = l;
    for (int i = j + 1 + lane; i < ns; i += 32) {
      const double lij = M[i * ld + j];
      for (int k = j + 1; k <= i; ++k) M[i * ld + k] -= lij * M[k * ld + j];
    }
    __syncwarp();
  }
  if (spd) {
    // forward L y = g
    for (int i = lane; i < ns; i += 32) y[i] = g[i];
    __syncwarp();
    for (int j = 0; j < ns; ++j) {
      const double yj = y[j] / M[j * ld + j];
      __syncwarp();
      if (lane == 0) y[j] = yj;
      for (int i = j + 1 + lane; i < ns; i += 32) y[i] -= M[i * ld + j] * yj;
      __syncwarp();
    }
    // backward L^T dx = y
    for (int j = ns - 1; j >= 0; --j) {
      const double xj = y[j] / M[j * ld + j];
      __syncwarp();
      if (lane == 0) { dx[j] = xj; y[j] = xj; }
      for (int i = lane; i < j; i += 32) y[i] -= M[j * ld + i] * xj;
      __syncwarp();
    }
  }
  }
  if (spd) {
    for (int i = lane; i < ns; i += 32) {
      const double di = dx[i];
      const double xi = p.x0[i] + di;
      // 1/2 dx^T H_L dx + g_L . dx  with  H_L = -(M - diag(prec))  and  dx^T M dx = g . dx
      quad += 0.5 * (-g[i] * di + p.xprec[i] * di * di) + gL[i] * di;
      bool in_i = true;
      lp += prior_logpdf(p.xkind[i], xi, p.xlo[i], p.xhi[i], p.xloc[i], p.xkind[i] == 2 ? rsqrt(p.xprec[i]) : 1., &in_i);
      // NB: the reference adds -inf to logprior when a solved value leaves its limits, but does not veto the point
      store_xsolved(p, pt, i, xi);
    }
  } else if (wants_xsolved(p)) {
    for (int i = lane; i < ns; i += 32) store_xsolved(p, pt, i, CUDART_NAN);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    quad += __shfl_xor_sync(0xffffffffu, quad, o);
    logdet += __shfl_xor_sync(0xffffffffu, logdet, o);
    inside = __shfl_xor_sync(0xffffffffu, (int)inside, o) && inside;
  }
  if (lane == 0) {
    int st = 0;
    double ll = -0.5 * chi2 + quad - 0.5 * logdet;
    if (!inside) st |= DLB_PT_OUT_OF_PRIOR;
    if (!spd) { st |= DLB_PT_NOT_SPD; ll = CUDART_NAN; }
    if (!isfinite(chi2)) { st |= DLB_PT_NONFINITE; ll = CUDART_NAN; }
    if (p.loglike) p.loglike[pt] = ll;
    if (p.logprior) p.logprior[pt] = lp;
    store_logpost(p, pt, inside ? ll + lp : -CUDART_INF);
    if (p.status) p.status[pt] = st;
  }
}

__global__ void __launch_bounds__(SOLVE_WARPS * 32) finish_solve_kernel(const __grid_constant__ SolveParams p) {
  extern __shared__ double sm[];
  finish_solve_body(p, sm);
  gather_tail(p);
}

// ---- thread-per-point form: when every independent group of solved parameters has <= SOLVE_NREG members, ONE thread assembles and
//      factorises all the groups of its point.  Consecutive threads read consecutive points of every Gram-partial row (coalesced; the
//      warp-per-point kernel above reads one row per lane, a 32-way strided access), the per-thread local system lives in shared memory
//      interleaved over the block's threads (conflict free) while it is accumulated and in registers while it is factorised. ----
constexpr int ST_THREADS = 128;

// sm: [SOLVE_NREG * SOLVE_NREG + SOLVE_NREG][ST_THREADS]; xstage: [ST_THREADS][ns] solved values of the block's points (or NULL), written out
// coalesced by the kernel below (a thread's own ns values are 8-byte stores ns doubles apart: 4 x write amplification in HBM and over NVLink)
__device__ __forceinline__ void finish_solve_thread_body(const SolveParams& p, double* sm, double* xstage) {
  const int tid = threadIdx.x;
  const long long pt = (long long)blockIdx.x * ST_THREADS + tid;
  if (pt >= p.N) return;
  double* M = sm + tid;                                            // element (a, b), b <= a, at M[(a * SOLVE_NREG + b) * ST_THREADS]
  double* gl = sm + (size_t)SOLVE_NREG * SOLVE_NREG * ST_THREADS + tid;
  const int ns = p.ns;
  double chi2 = 0.;
  for (int t = 0; t < p.nchi; ++t) chi2 += p.partial[p.chi_rows[t] * p.npad + pt];
  double lp = 0.;
  bool inside = true;
  for (int j = 0; j < p.d; ++j)
    lp += prior_logpdf(p.prior_kind[j], p.params[pt * p.d + j], p.prior_lo[j], p.prior_hi[j], p.prior_loc[j], p.prior_scale[j], &inside);
  bool spd = true;
  double logdet = 0., quad = 0.;
  for (int c = 0; c < p.ncomp; ++c) {
    const int o = p.comp_off[c], n = p.comp_off[c + 1] - o, nm = p.comp_nmarg[c];
    const int* idx = p.comp_idx + o;
    const double* m0 = p.comp_m0 + p.m0_off[c];
    for (int a = 0; a < n; ++a) {
      gl[a * ST_THREADS] = 0.;
      for (int b = 0; b <= a; ++b) M[(a * SOLVE_NREG + b) * ST_THREADS] = m0[a * n + b];
    }
    const int t1 = p.gent_off[c + 1];
    int t = p.gent_off[c];
    for (; t + 8 <= t1; t += 8) {   // eight independent loads in flight (two warps per scheduler: memory-level parallelism has to come from the thread)
      GroupEntry e[8];
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { e[q] = p.gent[t + q]; v[q] = p.partial[e[q].prow * p.npad + pt]; }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (e[q].kind == 1) gl[e[q].li * ST_THREADS] -= v[q];
        else M[(e[q].li * SOLVE_NREG + e[q].lj) * ST_THREADS] += v[q];
      }
    }
    for (; t + 4 <= t1; t += 4) {   // four independent loads in flight
      GroupEntry e[4];
      double v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { e[q] = p.gent[t + q]; v[q] = p.partial[e[q].prow * p.npad + pt]; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e[q].kind == 1) gl[e[q].li * ST_THREADS] -= v[q];
        else M[(e[q].li * SOLVE_NREG + e[q].lj) * ST_THREADS] += v[q];
      }
    }
    for (; t < t1; ++t) {
      const GroupEntry e = p.gent[t];
      const double v = p.partial[e.prow * p.npad + pt];
      if (e.kind == 1) gl[e.li * ST_THREADS] -= v;
      else M[(e.li * SOLVE_NREG + e.lj) * ST_THREADS] += v;
    }
    double a[SOLVE_NREG][SOLVE_NREG], yy[SOLVE_NREG], gr[SOLVE_NREG], glr[SOLVE_NREG];
    int id[SOLVE_NREG];
#pragma unroll
    for (int i = 0; i < SOLVE_NREG; ++i) {
      id[i] = i < n ? idx[i] : 0;
      glr[i] = i < n ? gl[i * ST_THREADS] : 0.;
      gr[i] = i < n ? p.g0[id[i]] + glr[i] : 0.;
      yy[i] = gr[i];
#pragma unroll
      for (int j = 0; j <= i; ++j) a[i][j] = (i < n) ? M[(i * SOLVE_NREG + j) * ST_THREADS] : (i == j ? 1. : 0.);
    }
    // LDL^T instead of Cholesky: per pivot ONE reciprocal (no sqrt, no further divisions in the substitutions) and log det from the product of the
    // pivots (two partial products of <= 4 factors each: no overflow for pivots up to 1e+-75) -- the dependent sqrt / div / log chain of the
    // Cholesky form was most of this kernel's time (one thread per point, two warps per scheduler: nothing hides the latency)
    bool ok = true;
    double inv[SOLVE_NREG], det0 = 1., det1 = 1.;
#pragma unroll
    for (int j = 0; j < SOLVE_NREG; ++j) {
      const double djj = a[j][j];
      if (j < n && (!(djj > 0.) || !isfinite(djj))) ok = false;
      inv[j] = 1. / djj;
      if (j < nm) { if (j < SOLVE_NREG / 2) det0 *= djj; else det1 *= djj; }
      double col[SOLVE_NREG];
#pragma unroll
      for (int i = j + 1; i < SOLVE_NREG; ++i) { col[i] = a[i][j]; a[i][j] = col[i] * inv[j]; }   // a[i][j] <- l_ij
#pragma unroll
      for (int i = j + 1; i < SOLVE_NREG; ++i)
#pragma unroll
        for (int k = j + 1; k <= i; ++k) a[i][k] -= col[i] * a[k][j];                              // -= l_ij d_j l_kj
    }
    if (!ok) { spd = false; continue; }
    logdet += log(det0) + log(det1);
#pragma unroll
    for (int j = 0; j < SOLVE_NREG; ++j) {   // forward L z = g (unit lower triangular)
#pragma unroll
      for (int i = j + 1; i < SOLVE_NREG; ++i) yy[i] -= a[i][j] * yy[j];
    }
#pragma unroll
    for (int j = 0; j < SOLVE_NREG; ++j) yy[j] *= inv[j];   // D w = z
#pragma unroll
    for (int j = SOLVE_NREG - 1; j >= 0; --j) {   // backward L^T dx = w
#pragma unroll
      for (int i = 0; i < j; ++i) yy[i] -= a[j][i] * yy[j];
    }
#pragma unroll
    for (int i = 0; i < SOLVE_NREG; ++i)
      if (i < n) {
        const int I = id[i];
        const double di = yy[i], xi = p.x0[I] + di, prec = p.xprec[I];
        quad += 0.5 * (-gr[i] * di + prec * di * di) + glr[i] * di;
        bool in_i = true;
        lp += prior_logpdf(p.xkind[I], xi, p.xlo[I], p.xhi[I], p.xloc[I], p.xkind[I] == 2 ? rsqrt(prec) : 1., &in_i);
        if (xstage) xstage[tid * ns + I] = xi;
      }
  }
  if (!spd && xstage)
    for (int i = 0; i < ns; ++i) xstage[tid * ns + i] = CUDART_NAN;
  int st = 0;
  double ll = -0.5 * chi2 + quad - 0.5 * logdet;
  if (!inside) st |= DLB_PT_OUT_OF_PRIOR;
  if (!spd) { st |= DLB_PT_NOT_SPD; ll = CUDART_NAN; }
  if (!isfinite(chi2)) { st |= DLB_PT_NONFINITE; ll = CUDART_NAN; }
  if (p.loglike) p.loglike[pt] = ll;
  if (p.logprior) p.logprior[pt] = lp;
  store_logpost(p, pt, inside ? ll + lp : -CUDART_INF);
  if (p.status) p.status[pt] = st;
}

__global__ void __launch_bounds__(ST_THREADS) finish_solve_thread_kernel(const __grid_constant__ SolveParams p) {
  extern __shared__ double sm[];
  double* xstage = wants_xsolved(p) ? sm + (size_t)(SOLVE_NREG * SOLVE_NREG + SOLVE_NREG) * ST_THREADS : nullptr;
  finish_solve_thread_body(p, sm, xstage);
  if (xstage) {
    __syncthreads();
    const long long e0 = (long long)blockIdx.x * ST_THREADS * p.ns;
    const int cnt = (int)min((long long)ST_THREADS, p.N - (long long)blockIdx.x * ST_THREADS) * p.ns;
    for (int e = threadIdx.x; e < cnt; e += ST_THREADS) {
      const double v = xstage[e];
      if (p.xsolved) p.xsolved[e0 + e] = v;
      if (p.gather.with_solved)
        for (int r = 0; r < p.gather.world; ++r)
          if (p.gather.xs_only < 0 || r == p.gather.xs_only) p.gather.xs[r][e0 + e] = v;
    }
  }
  gather_tail(p);
}

}  // namespace

int launch_solve(const SolveParams& p, cudaStream_t stream) {
  if (p.ns == 0) {
    const int threads = 256;
    finish_nosolve_kernel<<<(unsigned)((p.N + threads - 1) / threads), threads, 0, stream>>>(p);
  } else if (p.gent && p.ncomp > 0 && p.comp_max <= SOLVE_NREG && !p.hess) {
    const bool stage = p.xsolved != nullptr || p.gather.with_solved;
    const size_t smem = (size_t)(SOLVE_NREG * SOLVE_NREG + SOLVE_NREG + (stage ? p.ns : 0)) * ST_THREADS * sizeof(double);
    static size_t configured_dev[64] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > configured) {
      DLB_CUDA(cudaFuncSetAttribute(finish_solve_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    finish_solve_thread_kernel<<<(unsigned)((p.N + ST_THREADS - 1) / ST_THREADS), ST_THREADS, smem, stream>>>(p);
  } else {
    const size_t smem = (size_t)SOLVE_WARPS * (p.ns * (p.ns + 1) + 4 * p.ns) * sizeof(double);
    static size_t configured_dev[64] = {};
    size_t& configured = configured_dev[current_device_slot()];
    if (smem > 48 * 1024 && smem > configured) {
      DLB_CUDA(cudaFuncSetAttribute(finish_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    finish_solve_kernel<<<(unsigned)((p.N + SOLVE_WARPS - 1) / SOLVE_WARPS), SOLVE_WARPS * 32, smem, stream>>>(p);
  }
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

}  // namespace dlb

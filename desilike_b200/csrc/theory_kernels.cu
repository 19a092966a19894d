// Direct (non-emulated) theory generators.
//
//  * kaiser_kernel : ShapeFit template  (reference desilike/theories/galaxy_clustering/power_template.py:747-761)
//                    -> not-a-knot cubic spline of pk_dd on log10 k (reference desilike/jax.py:263-265, the SciPy branch the
//                       reference runs here; bao.py:17-19)
//                    -> AP remap (reference theories/galaxy_clustering/base.py:211-223, :341-353)
//                    -> Kaiser tables dd / dt / tt by Gauss-Legendre mu quadrature (reference full_shape.py:488-500;
//                       base.py:201-208; utils.py:625-643)
//  * bao_kernel    : damped BAO wiggles P_ell(k), model 'standard' (reference bao.py:117-134)
//  * fftlog_kernel : P_ell(kin) -> log-grid resample + high-k extrapolation (reference base.py:127-133)
//                    -> FFTLog (cosmoprimo PowerToCorrelation, restated in oracle/fftlog.py) -> linear interpolation onto s (:136)
//
// One warp per parameter point for the spline/quadrature kernels (the per-point template table lives in shared memory);
// one CTA per (point, multipole) for the FFTLog kernel (4096-point complex FP64 transform staged in shared memory).
#include "common.h"
#include <math_constants.h>

namespace dlb {

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// payload header (idata[0..7]) shared by the theory generators
//   [0] nk   [1] nell   [2] nmu   [3] nt (template knots)   [4] apmode (0 qparqper, 1 qiso, 2 qap, 3 qisoqap)
//   [5] flags   [6] ntab (padded length of one output table)   [7] total segment length
// parameter slots (nparams = 8): kaiser: q0 q1 dm dn df sigmapar sigmaper -      bao: q0 q1 b1 dbeta sigmas sigmapar sigmaper df
// ---------------------------------------------------------------------------------------------------------------------
struct SplineConst {          // device views into ddata
  const double* x;            // knots [nt]
  const double* invh;         // 1 / (x[i+1] - x[i])            [nt - 1]
  const double* w;            // Thomas forward multipliers      [nt]   (indexed by knot)
  const double* invb;         // 1 / pivot                       [nt]
  const double* cb;           // super-diagonal / pivot          [nt]
  double r0, r1;              // not-a-knot boundary ratios
  double x0, inv_dx;          // for the initial interval guess
  double dx;                  // uniform grids (the reference's geomspace template grids are uniform in log10 k): knot spacing
  int nt;
  bool uniform;               // knots are x0 + i dx: interval index, width and offset in closed form (no table loads, no search)
};

__device__ __forceinline__ void ap_qpar_qper(int apmode, double q0, double q1, double eta, double& qpar, double& qper) {
  if (apmode == 0) { qpar = q0; qper = q1; }
  else if (apmode == 1) { qpar = q0; qper = q0; }
  else if (apmode == 2) { qpar = pow(q0, 1. - eta); qper = pow(q0, -eta); }
  else { qpar = q0 * pow(q1, 1. - eta); qper = q0 * pow(q1, -eta); }
}

// Second derivatives M of the not-a-knot cubic spline through (x_i, y_i): tridiagonal solve with constant (plan-time)
// pivots; both sweeps are affine recurrences evaluated as warp-level scans of (A, B) composites.  y, M in shared memory.
__device__ void spline_second_derivatives(const SplineConst& sc, const double* y, double* M, int lane) {
  const int n = sc.nt, m = n - 2;                 // unknowns M_1 .. M_{n-2}
  for (int i = 1 + lane; i <= m; i += 32)
    M[i] = 6. * ((y[i + 1] - y[i]) * sc.invh[i] - (y[i] - y[i - 1]) * sc.invh[i - 1]);
  __syncwarp();
  const int L = (m + 31) / 32;
  const int lo = 1 + lane * L, hi = min(lo + L, m + 1);
  // forward: d_i = rhs_i - w_i d_{i-1}
  double A = 1., B = 0.;
  for (int i = lo; i < hi; ++i) { const double wi = sc.w[i]; B = M[i] - wi * B; A = -wi * A; }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
    if (lane >= o) { B = A * Bp + B; A = A * Ap; }
  }
  double carry = __shfl_up_sync(0xffffffffu, B, 1);
  if (lane == 0) carry = 0.;
  for (int i = lo; i < hi; ++i) { carry = M[i] - sc.w[i] * carry; M[i] = carry; }
  __syncwarp();
  // backward: M_i = d_i / b_i - cb_i M_{i+1}   (reverse order: lane 0 takes the LAST chunk)
  const int rhi = m - lane * L, rlo = max(rhi - L, 0);   // indices rhi, rhi-1, .. rlo+1
  A = 1.; B = 0.;
  for (int i = rhi; i > rlo; --i) { const double c = sc.cb[i]; B = M[i] * sc.invb[i] - c * B; A = -c * A; }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
    if (lane >= o) { B = A * Bp + B; A = A * Ap; }
  }
  carry = __shfl_up_sync(0xffffffffu, B, 1);
  if (lane == 0) carry = 0.;
  for (int i = rhi; i > rlo; --i) { carry = M[i] * sc.invb[i] - sc.cb[i] * carry; M[i] = carry; }
  __syncwarp();
  if (lane == 0) {
    M[0] = (1. + sc.r0) * M[1] - sc.r0 * M[2];
    M[n - 1] = (1. + sc.r1) * M[n - 2] - sc.r1 * M[n - 3];
  }
  __syncwarp();
}

// Same sweeps with (a) the constants pre-transposed to lane-major order (coalesced loads) and (b) the second derivatives
// stored at padded positions i + (i >> lgL) (one pad per lane chunk of L = 2^lgL knots: chunk stride L + 1 is odd, so the
// per-lane sequential sweeps are shared-memory bank-conflict free).  Mp must hold nt + nt / L + 1 doubles.
__device__ __forceinline__ int pidx(int i, int lgL) { return i + (i >> lgL); }

__device__ void spline_second_derivatives_padded(const SplineConst& sc, const double* __restrict__ wT, const double* __restrict__ invbT,
                                                 const double* __restrict__ cbT, int lgL, const double* y, double* Mp, int lane) {
  const int n = sc.nt, m = n - 2, L = 1 << lgL;
  for (int i = 1 + lane; i <= m; i += 32)
    Mp[pidx(i, lgL)] = 6. * ((y[i + 1] - y[i]) * sc.invh[i] - (y[i] - y[i - 1]) * sc.invh[i - 1]);
  __syncwarp();
  const int lo = 1 + lane * L, hi = min(lo + L, m + 1);
  double A = 1., B = 0.;
  for (int i = lo, t = 0; i < hi; ++i, ++t) { const double wi = wT[t * 32 + lane]; B = Mp[pidx(i, lgL)] - wi * B; A = -wi * A; }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
    if (lane >= o) { B = A * Bp + B; A = A * Ap; }
  }
  double carry = __shfl_up_sync(0xffffffffu, B, 1);
  if (lane == 0) carry = 0.;
  for (int i = lo, t = 0; i < hi; ++i, ++t) { carry = Mp[pidx(i, lgL)] - wT[t * 32 + lane] * carry; Mp[pidx(i, lgL)] = carry; }
  __syncwarp();
  const int rhi = m - lane * L, rlo = max(rhi - L, 0);   // indices rhi, rhi-1, .. rlo+1
  A = 1.; B = 0.;
  for (int i = rhi, t = 0; i > rlo; --i, ++t) { const double c = cbT[t * 32 + lane]; B = Mp[pidx(i, lgL)] * invbT[t * 32 + lane] - c * B; A = -c * A; }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
    if (lane >= o) { B = A * Bp + B; A = A * Ap; }
  }
  carry = __shfl_up_sync(0xffffffffu, B, 1);
  if (lane == 0) carry = 0.;
  for (int i = rhi, t = 0; i > rlo; --i, ++t) { carry = Mp[pidx(i, lgL)] * invbT[t * 32 + lane] - cbT[t * 32 + lane] * carry; Mp[pidx(i, lgL)] = carry; }
  __syncwarp();
  if (lane == 0) {
    Mp[pidx(0, lgL)] = (1. + sc.r0) * Mp[pidx(1, lgL)] - sc.r0 * Mp[pidx(2, lgL)];
    Mp[pidx(n - 1, lgL)] = (1. + sc.r1) * Mp[pidx(n - 2, lgL)] - sc.r1 * Mp[pidx(n - 3, lgL)];
  }
  __syncwarp();
}

__device__ __forceinline__ double spline_eval_hp(const SplineConst& sc, const double* __restrict__ hh, const double* y, const double* Mp, int lgL, double xq) {
  const int n = sc.nt;
  int i = (int)((xq - sc.x0) * sc.inv_dx);
  i = max(0, min(i, n - 2));
  double h, ih, t;
  if (sc.uniform) {   // a rounding-induced off-by-one interval evaluates the neighbouring cubic an ulp outside its interval: the spline is C2, no search needed
    h = sc.dx; ih = sc.inv_dx; t = xq - fma((double)i, sc.dx, sc.x0);
  } else {
    while (i > 0 && xq < sc.x[i]) --i;
    while (i < n - 2 && xq >= sc.x[i + 1]) ++i;
    h = hh[i]; ih = sc.invh[i]; t = xq - sc.x[i];
  }
  const double Mi = Mp[pidx(i, lgL)], Mj = Mp[pidx(i + 1, lgL)], yi = y[i];
  const double b = (y[i + 1] - yi) * ih - h * (2. * Mi + Mj) * (1. / 6.);
  return yi + t * (b + t * (0.5 * Mi + t * (Mj - Mi) * ih * (1. / 6.)));
}

__device__ __forceinline__ double spline_eval(const SplineConst& sc, const double* y, const double* M, double xq) {
  const int n = sc.nt;
  int i = (int)((xq - sc.x0) * sc.inv_dx);
  i = max(0, min(i, n - 2));
  while (i > 0 && xq < sc.x[i]) --i;
  while (i < n - 2 && xq >= sc.x[i + 1]) ++i;
  const double h = 1. / sc.invh[i], t = xq - sc.x[i];
  const double Mi = M[i], Mj = M[i + 1];
  const double b = (y[i + 1] - y[i]) * sc.invh[i] - h * (2. * Mi + Mj) * (1. / 6.);
  return y[i] + t * (b + t * (0.5 * Mi + t * (Mj - Mi) * sc.invh[i] * (1. / 6.)));
}

// same cubic, interval width from a table (no division) -- used with the derived tables appended by theory_lowering.py
__device__ __forceinline__ double spline_eval_h(const SplineConst& sc, const double* __restrict__ hh, const double* y, const double* M, double xq) {
  const int n = sc.nt;
  int i = (int)((xq - sc.x0) * sc.inv_dx);
  i = max(0, min(i, n - 2));
  double h, ih, t;
  if (sc.uniform) {
    h = sc.dx; ih = sc.inv_dx; t = xq - fma((double)i, sc.dx, sc.x0);
  } else {
    while (i > 0 && xq < sc.x[i]) --i;
    while (i < n - 2 && xq >= sc.x[i + 1]) ++i;
    h = hh[i]; ih = sc.invh[i]; t = xq - sc.x[i];
  }
  const double Mi = M[i], Mj = M[i + 1], yi = y[i];
  const double b = (y[i + 1] - yi) * ih - h * (2. * Mi + Mj) * (1. / 6.);
  return yi + t * (b + t * (0.5 * Mi + t * (Mj - Mi) * ih * (1. / 6.)));
}

__device__ __forceinline__ SplineConst load_spline(const double* d, int nt, bool uniform = false) {
  // layout: x[nt] invh[nt] w[nt] invb[nt] cb[nt] then (r0, r1, x0, inv_dx)
  SplineConst sc;
  sc.nt = nt; sc.x = d; sc.invh = d + nt; sc.w = d + 2 * nt; sc.invb = d + 3 * nt; sc.cb = d + 4 * nt;
  sc.r0 = d[5 * nt]; sc.r1 = d[5 * nt + 1]; sc.x0 = d[5 * nt + 2]; sc.inv_dx = d[5 * nt + 3];
  sc.dx = 1. / sc.inv_dx; sc.uniform = uniform;
  return sc;
}
constexpr int SPLINE_DOUBLES(int nt) { return 5 * nt + 4; }

// ---------------------------------------------------------------------------------------------------------------------
// Kaiser.  ddata = [eta, a_inv_unused, f_fid, 0 | k[nk] | mu[nmu] | wmu[nell, nmu] | pk_fid[nt] | Tt[nt] | Lt[nt] | spline consts | log10 k[nk] | h[nt]
//                   | ratios]
//   Tt = tanh(a ln(k_t / kp)) / a,  Lt = ln(k_t / kp)   =>  pk_dd = pk_fid * exp(dm * Tt + dn * Lt)
// idata = [nk, nell, nmu, nt, apmode, flags | ell0 << 8, ntab, total, off_derived, -, nseg, off_seg, off_ratio, interp, pad, nw_max | segment records]
//
// Output = `nseg` SEGMENTS of ntab doubles after the (1, 0) pair.  Segment s is a per-point linear combination of the three Kaiser tables
//     seg_s[ell, k] = sum_mu w_ell(mu) (a_dd + a_dt f mu_ap^2 + a_tt f^2 mu_ap^4) jac D(k, mu) pk(k_ap(k, mu))
// with (a_dd, a_dt, a_tt) = sum over its records of ratio * prod_f theta[idx_f]^pow_f: the plain tables pk_dd / pk_dt / pk_tt are (1,0,0) / (0,1,0) /
// (0,0,1); pk11_l0 (the ell = 0 row of pk_dd tiled over ell) is (1,0,0) with flag bit 0; and the tracer combine b1^2 dd + 2 b1 dt + tt (reference
// full_shape.py:545-550) is ONE segment with weight (b1 + f mu_ap^2)^2 when the lowering merges the three terms (their window matrices coincide):
// a third of the quadrature sums, a quarter of the feature columns written and contracted.
// segment record (ints): flag, ncomb, then ncomb x { table (0 dd, 1 dt, 2 tt), nf, idx[4], pow[4] }; ratios (doubles) in record order.
//
// One warp per point.  (1) ShapeFit template ordinates y_i = pk_fid_i exp(dm Tt_i + dn Lt_i) only on the WINDOW of knots the AP-distorted nodes
// of this point can reach (+ `pad` knots each side); (2) interp 0: second derivatives of the not-a-knot spline (SciPy branch of reference
// jax.py:263-265) by the two sweeps restricted to the window -- a sweep restarted `pad` knots away from the first node differs from the full one by
// |w|^pad ~ 0.27^28 ~ 1e-16 relative; interp 1: no sweeps at all (interpax C1 local cubic, reference jax.py:258-261: knot slopes = mean of the two
// adjacent secant slopes); (3) the nk x nmu nodes are spread evenly over the lanes in blocks of KAISER_KB k nodes (node values to shared memory), then
// (4) the lanes form the mu sums of every (segment, ell, k) output of the block and store them coalesced.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int KAISER_WARPS = 4;
constexpr int KAISER_KB = 24;   // (32 -> 24 and the exact padding bound below: 8 instead of 6 co-resident blocks per SM)
constexpr int MAX_ELL = 3;
constexpr int KAISER_MAXSEG = 8;

// largest padded index pidx(j, lgL) = j + (j >> lgL) + 1 over windows of m <= nw unknowns, lgL = smallest with (32 << lgL) >= m (>= 1)
__host__ __device__ inline int kaiser_padded_len(int nw) {
  int best = 0;
  for (int lgL = 1; lgL < 16; ++lgL) {
    const int mmax = nw < (32 << lgL) ? nw : (32 << lgL);   // windows using this lgL have m <= 32 << lgL
    const int len = mmax + (mmax >> lgL) + 2;
    best = len > best ? len : best;
    if ((32 << lgL) >= nw) break;
  }
  return best;
}

__host__ __device__ inline int kaiser_warp_doubles(int nw, int nmu, int nell, int nseg, int interp) {
  // y | M (padded) | node values | segment weights | per-mu AP quantities | segment coefficients
  return (nw + 4) + (interp == 0 ? kaiser_padded_len(nw) + 6 : 0) + KAISER_KB * (nmu | 1) + nseg * nell * nmu + 3 * nmu + 3 * nseg + 2;
}

// NMU: number of mu nodes when known at compile time (8: the reference default, full_shape.py:484), 0: generic.  INTERP: 0 not-a-knot, 1 local cubic.
template <int NMU, int INTERP>
__global__ void __launch_bounds__(KAISER_WARPS * 32, 8) kaiser_kernel(const DevGen g, const double* __restrict__ params, int d, long long N,
                                                                   long long npad, double* __restrict__ F, int fstride, int warp_doubles) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int* __restrict__ id = g.idata;
  const int nk = id[0], nell = id[1], nmu = NMU ? NMU : id[2], nt = id[3], apmode = id[4], ntab = id[6];
  const int ell0 = id[5] >> 8;   // index of ell = 0 among the multipoles (or -1)
  const bool uniform = (id[5] & 1) != 0;
  const int nseg = id[10], pad = id[14], nw_max = id[15];
  const int* __restrict__ segi = id + id[11];
  const long long pt = (long long)blockIdx.x * KAISER_WARPS + warp;
  if (pt >= npad) return;
  double* row = F + pt * fstride + g.foff;
  if (pt >= N) {   // padding rows: zeros
    for (int j = lane; j < g.fpad; j += 32) row[j] = 0.;
    return;
  }
  const int pst = nmu | 1;
  double* y = sm + (size_t)warp * warp_doubles;
  double* M = y + nw_max + 4;
  double* P = M + (INTERP == 0 ? kaiser_padded_len(nw_max) + 6 : 0);
  double* wseg = P + KAISER_KB * pst;
  double* smu = wseg + nseg * nell * nmu;      // [nmu][3]: log10(F / qper), mu_ap^2, damping coefficient
  double* coef = smu + 3 * nmu;                // [nseg][3]
  const double* __restrict__ dd = g.ddata;
  const double eta = dd[0], f_fid = dd[2];
  const double* kk = dd + 4;
  const double* mu = kk + nk;
  const double* wmu = mu + nmu;
  const double* pk_fid = wmu + nell * nmu;
  const double* Tt = pk_fid + nt;
  const double* Lt = Tt + nt;
  const SplineConst sc = load_spline(Lt + nt, nt, uniform);
  const double* lkk = dd + id[8];
  const double* hh = lkk + nk;
  const double* ratio = dd + id[12];
  double par[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) par[j] = (j < g.nparams && g.pidx[j] >= 0) ? params[pt * d + g.pidx[j]] : g.pdef[j];
  double qpar, qper;
  ap_qpar_qper(apmode, par[0], par[1], eta, qpar, qper);
  const double dm = par[2], dn = par[3], f = f_fid * par[4], sigmapar = par[5], sigmaper = par[6];
  const double qap = qpar / qper, jac = 1. / (qpar * qper * qper), iq2 = 1. / (qap * qap) - 1.;
  const bool damped = (sigmapar != 0.) || (sigmaper != 0.);
  // ---- per-mu AP quantities, once per point (reference base.py:211-223, full_shape.py:491-495) ----
  double lkf = 0.;
  if (lane < nmu) {
    const double m_ = mu[lane];
    const double fap = sqrt(1. + m_ * m_ * iq2);
    const double kfac = fap / qper, muap = m_ / (qap * fap), m2 = muap * muap;
    lkf = log10(kfac);
    smu[lane * 3 + 0] = lkf;
    smu[lane * 3 + 1] = m2;
    smu[lane * 3 + 2] = 0.5 * kfac * kfac * (sigmapar * sigmapar * m2 + sigmaper * sigmaper * (1. - m2));
  }
  // ---- segment coefficients (a_dd, a_dt, a_tt) ----
  if (lane < nseg) {
    const int* rec = segi;
    const double* rt = ratio;
    for (int s_ = 0; s_ < lane; ++s_) { rt += rec[1]; rec += 2 + 10 * rec[1]; }
    double a0 = 0., a1 = 0., a2 = 0.;
    const int ncomb = rec[1];
    for (int c = 0; c < ncomb; ++c) {
      const int* cr = rec + 2 + 10 * c;
      double v = rt[c];
      for (int q = 0; q < cr[1]; ++q) {
        const double x = params[pt * d + cr[2 + q]];
        const int e = cr[6 + q];
        double pw = 1.;
        for (int i = 0; i < (e >= 0 ? e : -e); ++i) pw *= x;
        v *= e >= 0 ? pw : 1. / pw;
      }
      if (cr[0] == 0) a0 += v; else if (cr[0] == 1) a1 += v; else a2 += v;
    }
    coef[lane * 3 + 0] = a0; coef[lane * 3 + 1] = a1; coef[lane * 3 + 2] = a2;
  }
  // ---- window of template knots reached by the nodes: log10 k_ap in [lk_lo, lk_hi] ----
  double lmin = lane < nmu ? lkf : 1e300, lmax = lane < nmu ? lkf : -1e300;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
  int i_lo = 0, i_hi = nt - 1;
  if (uniform) {
    const double klo = dd[id[8] - 2], khi = dd[id[8] - 1];   // min / max of log10 k over the theory nodes (stored just before the log10 k table)
    i_lo = (int)floor((klo + lmin - sc.x0) * sc.inv_dx) - pad;
    i_hi = (int)floor((khi + lmax - sc.x0) * sc.inv_dx) + 2 + pad;
    i_lo = max(0, min(i_lo, nt - 4));
    i_hi = max(i_lo + 3, min(i_hi, nt - 1));
    if (i_hi - i_lo + 1 > nw_max) i_hi = i_lo + nw_max - 1;   // (only for points far outside the prior box the window was sized for)
  }
  const int nw = i_hi - i_lo + 1;
  __syncwarp();
  // fast_exp (common.h): ~1e-15 relative error in ~18 instructions instead of ~45.  Two knots per lane and iteration: their loads and FP64
  // chains are independent
  {
    const double* pf = pk_fid + i_lo; const double* tt_ = Tt + i_lo; const double* lt_ = Lt + i_lo;
    int j = lane;
    for (; j + 32 < nw; j += 64) {
      const double p0 = pf[j], p1 = pf[j + 32], a0 = dm * tt_[j] + dn * lt_[j], a1 = dm * tt_[j + 32] + dn * lt_[j + 32];
      y[j] = p0 * fast_exp(fmin(fmax(a0, -708.), 708.));
      y[j + 32] = p1 * fast_exp(fmin(fmax(a1, -708.), 708.));
    }
    if (j < nw) y[j] = pf[j] * fast_exp(fmin(fmax(dm * tt_[j] + dn * lt_[j], -708.), 708.));
  }
  __syncwarp();
  int lgL = 1;
  if constexpr (INTERP == 0) {
    // ---- second derivatives on the window: unknowns i = a .. b (global knot indices), M_j stored at pidx(j - i_lo); the window's end knots
    //      are either the ends of the template grid (not-a-knot conditions below) or lie `pad` knots away from every node (M := 0 there).
    //      Uniform grids: away from the two ends of the grid the pivots of the tridiagonal solve have converged to constants (their plan-time
    //      values sit in ddata): no table loads in the sweeps ----
    const int a = i_lo + 1, b = i_hi - 1, m = b - a + 1;
    while ((32 << lgL) < m) ++lgL;
    const int L = 1 << lgL;
    const double* cst = dd + id[12] - 6;   // w_inf, invb_inf, cb_inf, first / last knot with converged pivots (as doubles), -
    const bool interior = uniform && a >= (int)cst[3] && b <= (int)cst[4];
    const double w_c = cst[0], invb_c = cst[1], cb_c = cst[2];
    if (uniform) {
      const double c6 = 6. * sc.inv_dx;
      for (int j = 1 + lane; j <= m; j += 32) M[pidx(j, lgL)] = c6 * ((y[j + 1] - y[j]) - (y[j] - y[j - 1]));
    } else {
      for (int i = a + lane; i <= b; i += 32) {
        const int j = i - i_lo;
        M[pidx(j, lgL)] = 6. * ((y[j + 1] - y[j]) * sc.invh[i] - (y[j] - y[j - 1]) * sc.invh[i - 1]);
      }
    }
    if (lane == 0) { if (i_lo > 0) M[pidx(0, lgL)] = 0.; if (i_hi < nt - 1) M[pidx(nw - 1, lgL)] = 0.; }
    __syncwarp();
    const int lo = a + lane * L, hi = min(lo + L, b + 1);
    double A = 1., B = 0.;
    if (interior) { for (int i = lo; i < hi; ++i) { B = M[pidx(i - i_lo, lgL)] - w_c * B; A = -w_c * A; } }
    else { for (int i = lo; i < hi; ++i) { const double wi = sc.w[i]; B = M[pidx(i - i_lo, lgL)] - wi * B; A = -wi * A; } }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
      if (lane >= o) { B = A * Bp + B; A = A * Ap; }
    }
    double carry = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) carry = 0.;
    if (interior) { for (int i = lo; i < hi; ++i) { carry = M[pidx(i - i_lo, lgL)] - w_c * carry; M[pidx(i - i_lo, lgL)] = carry; } }
    else { for (int i = lo; i < hi; ++i) { carry = M[pidx(i - i_lo, lgL)] - sc.w[i] * carry; M[pidx(i - i_lo, lgL)] = carry; } }
    __syncwarp();
    const int rhi = b - lane * L, rlo = max(rhi - L, a - 1);   // indices rhi, rhi - 1, .. rlo + 1
    A = 1.; B = 0.;
    if (interior) { for (int i = rhi; i > rlo; --i) { B = M[pidx(i - i_lo, lgL)] * invb_c - cb_c * B; A = -cb_c * A; } }
    else { for (int i = rhi; i > rlo; --i) { const double c = sc.cb[i]; B = M[pidx(i - i_lo, lgL)] * sc.invb[i] - c * B; A = -c * A; } }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
      if (lane >= o) { B = A * Bp + B; A = A * Ap; }
    }
    carry = __shfl_up_sync(0xffffffffu, B, 1);
    if (lane == 0) carry = 0.;
    if (interior) { for (int i = rhi; i > rlo; --i) { carry = M[pidx(i - i_lo, lgL)] * invb_c - cb_c * carry; M[pidx(i - i_lo, lgL)] = carry; } }
    else { for (int i = rhi; i > rlo; --i) { carry = M[pidx(i - i_lo, lgL)] * sc.invb[i] - sc.cb[i] * carry; M[pidx(i - i_lo, lgL)] = carry; } }
    __syncwarp();
    if (lane == 0) {   // not-a-knot end conditions where the window reaches an end of the template grid
      if (i_lo == 0) M[pidx(0, lgL)] = (1. + sc.r0) * M[pidx(1, lgL)] - sc.r0 * M[pidx(2, lgL)];
      if (i_hi == nt - 1) M[pidx(nw - 1, lgL)] = (1. + sc.r1) * M[pidx(nw - 2, lgL)] - sc.r1 * M[pidx(nw - 3, lgL)];
    }
  }
  // ---- segment weights w_ell(mu) (a_dd + a_dt f mu_ap^2 + a_tt f^2 mu_ap^4) ----
  {
    const int* rec = segi;
    for (int s_ = 0; s_ < nseg; ++s_) {
      const bool l0 = (rec[0] & 1) != 0;
      rec += 2 + 10 * rec[1];
      const double c0 = coef[s_ * 3], c1 = coef[s_ * 3 + 1], c2 = coef[s_ * 3 + 2];
      for (int l = 0; l < nell; ++l) {
        const int lw = l0 ? ell0 : l;
        if (lane < nmu) {
          const double m2 = smu[lane * 3 + 1];
          wseg[(s_ * nell + l) * nmu + lane] = lw < 0 ? 0. : wmu[lw * nmu + lane] * (c0 + f * m2 * (c1 + f * m2 * c2));
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) { row[0] = 1.; row[1] = 0.; }
  // ---- nodes and outputs, KAISER_KB k nodes at a time.  Node e = ikl * nmu + im of the block is handled by lane e mod 32: (ikl, im) advance by
  //      (32 div nmu, 32 mod nmu) per iteration -- no integer divisions in the loops ----
  const int q32 = 32 / nmu, r32 = 32 - q32 * nmu;
  const int ikl0 = lane / nmu, im0 = lane - ikl0 * nmu;
  for (int k0 = 0; k0 < nk; k0 += KAISER_KB) {
    const int kc = min(KAISER_KB, nk - k0), nnode = kc * nmu;
    int ikl = ikl0, im = im0;
    for (int e = lane; e < nnode; e += 32) {
      const int ik = k0 + ikl;
      const double xq = lkk[ik] + smu[im * 3];
      double p;
      int i = (int)((xq - sc.x0) * sc.inv_dx);
      i = max(0, min(i, nt - 2));
      if (!uniform) {
        while (i > 0 && xq < sc.x[i]) --i;
        while (i < nt - 2 && xq >= sc.x[i + 1]) ++i;
      }
      const double xi = uniform ? fma((double)i, sc.dx, sc.x0) : sc.x[i];
      const double h = uniform ? sc.dx : hh[i], ih = uniform ? sc.inv_dx : sc.invh[i];
      const int j = max(0, min(i - i_lo, nw - 2));
      const double y0 = y[j], y1 = y[j + 1];
      if constexpr (INTERP == 0) {
        // cubic of the interval [x_i, x_i+1] from (y, M); a rounding-induced off-by-one interval evaluates the neighbouring cubic an ulp outside
        // its interval: the spline is C2, no search needed on uniform grids
        const double t = xq - xi;
        const double Mi = M[pidx(j, lgL)], Mj = M[pidx(j + 1, lgL)];
        const double bq = (y1 - y0) * ih - h * (2. * Mi + Mj) * (1. / 6.);
        p = y0 + t * (bq + t * (0.5 * Mi + t * (Mj - Mi) * ih * (1. / 6.)));
      } else {
        // C1 local cubic (interpax 'cubic'): Hermite interpolation with knot slopes = mean of the adjacent secant slopes (one-sided at the
        // ends); queries outside the template range are NaN (interpax extrap=False)
        const double s0 = (y1 - y0) * ih;
        const double sm1 = (i > 0 && j > 0) ? (y0 - y[j - 1]) * (uniform ? sc.inv_dx : sc.invh[i - 1]) : s0;
        const double sp1 = (i < nt - 2 && j + 2 < nw) ? (y[j + 2] - y1) * (uniform ? sc.inv_dx : sc.invh[i + 1]) : s0;
        const double d0 = i > 0 ? 0.5 * (sm1 + s0) : s0, d1 = i < nt - 2 ? 0.5 * (s0 + sp1) : s0;
        const double t = (xq - xi) * ih;
        const double c2 = 3. * (y1 - y0) - h * (2. * d0 + d1), c3 = -2. * (y1 - y0) + h * (d0 + d1);
        p = y0 + t * (h * d0 + t * (c2 + t * c3));
        if (xq < sc.x0 || xq > sc.x[nt - 1]) p = CUDART_NAN;
      }
      p *= jac;
      if (damped) { const double k = kk[ik]; p *= fast_exp(fmax(-k * k * smu[im * 3 + 2], -708.)); }
      P[ikl * pst + im] = p;
      ikl += q32; im += r32;
      if (im >= nmu) { im -= nmu; ++ikl; }
    }
    __syncwarp();
    if (lane < kc) {
      const double* pv = P + lane * pst;
      double pr[NMU ? NMU : 16];
#pragma unroll
      for (int q = 0; q < (NMU ? NMU : 16); ++q) pr[q] = q < nmu ? pv[q] : 0.;
      const double* wv = wseg;
      for (int s_ = 0; s_ < nseg; ++s_) {
        double* dst = row + 2 + s_ * ntab + k0 + lane;
        for (int l = 0; l < nell; ++l, wv += nmu) {
          double acc = 0.;
#pragma unroll
          for (int q = 0; q < (NMU ? NMU : 16); ++q)
            if (q < nmu) acc = fma(wv[q], pr[q], acc);
          dst[l * nk] = acc;   // k fastest across the lanes: coalesced stores
        }
      }
    }
    __syncwarp();
  }
  // zero the padding between / after the segments
  const int nin = nell * nk;
  for (int t = 0; t < nseg; ++t)
    for (int j = nin + lane; j < ntab; j += 32) row[2 + t * ntab + j] = 0.;
  for (int j = 2 + nseg * ntab + lane; j < g.fpad; j += 32) row[j] = 0.;
}


// ---------------------------------------------------------------------------------------------------------------------
// Damped BAO wiggles, model 'standard' (reference bao.py:117-134).  The BAO template tables are parameter independent
// (BAOPowerSpectrumTemplate, reference power_template.py:372-376 only rescales f), so their spline second derivatives are plan
// constants.  ddata = [eta, f_fid, smoothing_radius, 0 | k[nk] | mu[nmu] | wmu[nell, nmu] | pknow(k)[nk] | pk_t[nt] | M_pk[nt] |
//                      pknow_t[nt] | M_now[nt] | spline consts | (FFTLog block)]
// param slots: q0 q1 b1 dbeta sigmas sigmapar sigmaper df.   Output P_ell(k) [nell, nk] at segment offset idata[12].
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BAO_WARPS = 4;

__global__ void __launch_bounds__(BAO_WARPS * 32) bao_kernel(const DevGen g, const double* __restrict__ params, int d, long long N,
                                                             long long npad, double* __restrict__ F, int fstride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = g.idata[0], nell = g.idata[1], nmu = g.idata[2], nt = g.idata[3], apmode = g.idata[4], flags = g.idata[5];
  const int poff = g.idata[12];
  const long long pt = (long long)blockIdx.x * BAO_WARPS + warp;
  if (pt >= npad) return;
  double* row = F + pt * fstride + g.foff;
  if (pt >= N) {
    for (int j = lane; j < g.fpad; j += 32) row[j] = 0.;
    return;
  }
  const double* dd = g.ddata;
  const double eta = dd[0], f_fid = dd[1], rs = dd[2];
  const double* kk = dd + 4;
  const double* mu = kk + nk;
  const double* wmu = mu + nmu;
  const double* pknow_k = wmu + nell * nmu;
  const double* pk_t = pknow_k + nk;
  const double* M_pk = pk_t + nt;
  const double* pknow_t = M_pk + nt;
  const double* M_now = pknow_t + nt;
  const SplineConst sc = load_spline(M_now + nt, nt, (flags & 4) != 0);   // flags bit 2: template grid uniform in log10 k
  double par[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) par[j] = (j < g.nparams && g.pidx[j] >= 0) ? params[pt * d + g.pidx[j]] : g.pdef[j];
  double qpar, qper;
  ap_qpar_qper(apmode, par[0], par[1], eta, qpar, qper);
  const double b1 = par[2], f = par[3] * f_fid * par[7], sigmas = par[4], sigmapar = par[5], sigmaper = par[6];
  const double qap = qpar / qper, iq2 = 1. / (qap * qap) - 1.;
  // derived tables (theory_lowering.py): log10 k, S(k), k^2, wiggle spline ordinates / second derivatives, interval widths
  const double* lkk = dd + g.idata[6];
  const double* skt = lkk + nk;
  const double* k2t = skt + nk;
  const double* Dw = k2t + nk;
  const double* MDw = Dw + nt;
  const double* hh = MDw + nt;
  // per-mu quantities, once per point (reference base.py:211-223, bao.py:123-134)
  __shared__ double smu[BAO_WARPS][16][5];
  if (lane < nmu) {
    const double m_ = mu[lane];
    const double fap = sqrt(1. + m_ * m_ * iq2);
    const double kfac = fap / qper, muap = m_ / (qap * fap), m2 = muap * muap;
    smu[warp][lane][0] = log10(kfac);
    smu[warp][lane][1] = 0.5 * kfac * kfac * (sigmapar * sigmapar * m2 + sigmaper * sigmaper * (1. - m2));   // exp(-k^2 [1])
    smu[warp][lane][2] = f * m2;                    // (b1 + [2] (1 - S))^2 damped wiggles
    smu[warp][lane][3] = f * m_ * m_;               // (b1 + [3] (1 - S))^2 smooth part
    smu[warp][lane][4] = 0.5 * sigmas * sigmas * m_ * m_;   // FoG: 1 / (1 + [4] k^2)^2
  }
  __syncwarp();
  if (lane == 0) { row[0] = 1.; row[1] = 0.; }
  for (int ik = lane; ik < nk; ik += 32) {
    double acc[MAX_ELL] = {0., 0., 0.};
    const double lk0 = lkk[ik], k2 = k2t[ik], oms = 1. - skt[ik], pknow = pknow_k[ik];
#pragma unroll 2
    for (int im = 0; im < nmu; ++im) {   // two mu nodes in flight: their table loads and FP64 chains are independent
      const double pkwap = spline_eval_h(sc, hh, Dw, MDw, lk0 + smu[warp][im][0]);
      const double ca = b1 + smu[warp][im][2] * oms;
      const double Cap = ca * ca * fast_exp(fmax(-k2 * smu[warp][im][1], -708.));
      const double fg = 1. + smu[warp][im][4] * k2;
      const double cb = b1 + smu[warp][im][3] * oms;
      const double ifg = fast_rcp(fg);   // fg >= 1: hardware reciprocal + two Newton steps instead of an IEEE division
      const double B = cb * cb * ifg * ifg;
      const double pkmu = B * pknow + Cap * pkwap;
#pragma unroll
      for (int l = 0; l < MAX_ELL; ++l)
        if (l < nell) acc[l] += pkmu * wmu[l * nmu + im];
    }
#pragma unroll
    for (int l = 0; l < MAX_ELL; ++l)
      if (l < nell) row[poff + l * nk + ik] = acc[l];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// FFTLog P_ell(kin) -> xi_ell(s) (reference theories/galaxy_clustering/base.py:62-77,127-136; algorithm restated in
// oracle/fftlog.py).  One CTA of 128 threads per (point, ell):
//   1. resample P_ell(kin) onto the 2048-point log grid: interpolation and high-k extrapolation (reference base.py:127-133)
//      are both 2-point linear combinations  x[i] = c0[i] pin[j0[i]] + c1[i] pin[j0[i] + 1]  with the power-law prefactor
//      k^(3-q) / (2 pi)^(3/2) folded into c0, c1 at plan time; zero padding to Np = 4096 is implicit;
//   2. rfft(4096) as ONE 2048-point complex FFT of z[n] = x[2n] + i x[2n+1] (Stockham autosort, radix 16 x 16 x 8 butterflies in
//      registers, 3 shared-memory passes instead of 12 radix-2 ones) followed by the real-input untangling
//          X[m] = (Z[m] + conj Z[M-m]) / 2 - i/2 e^(-2 pi i m / Np) (Z[m] - conj Z[M-m]);
//   3. the Hankel kernel multiply Y[m] = u_ell(m) X[m] fused with both untanglings (numpy's hfft = the same operator applied to Y);
//   4. a second 2048-point complex FFT: xi_grid[2n] = Re g[n] / M, xi_grid[2n+1] = -Im g[n] / M, times s^-q (-1)^(ell/2);
//   5. un-pad and interpolate linearly onto the requested s (:136).
// FFT block of ddata at idata[10]: c0[Ng] c1[Ng] u[nell][Np/2+1][2] post[nell][Ng] tw[Np/2][2] outw[nell][ns]   (tw[j] = e^(-2 pi i j / Np))
// FFT block of idata at idata[11]: j0[Ng] outj[nell][ns]
// Shared memory: 2048 complex, index i stored at i + (i >> 4) (padding keeps the radix-16 scatter conflict free) + pin[nkin].
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FFT_THREADS = 128;
constexpr int FFT_M = 2048;                                  // complex transform length (Np = 4096)
__device__ __forceinline__ int fidx(int i) { return i + (i >> 4); }
constexpr int FFT_SMEM_COMPLEX = FFT_M + (FFT_M >> 4);

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// forward 4-point DFT in place: (x0, x1, x2, x3) -> (X0, X1, X2, X3)
__device__ __forceinline__ void dft4(double2& x0, double2& x1, double2& x2, double2& x3) {
  const double2 s02 = cadd(x0, x2), d02 = csub(x0, x2), s13 = cadd(x1, x3), d13 = csub(x1, x3);
  x0 = cadd(s02, s13);
  x2 = csub(s02, s13);
  x1 = make_double2(d02.x + d13.y, d02.y - d13.x);   // d02 - i d13
  x3 = make_double2(d02.x - d13.y, d02.y + d13.x);   // d02 + i d13
}

// forward 16-point DFT in registers; on return X[k] is in v[(k >> 2) + 4 * (k & 3)]
__device__ __forceinline__ void dft16(double2 (&v)[16]) {
  constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, H = 0.70710678118654752440;
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[b + 4], v[b + 8], v[b + 12]);   // v[b + 4 q] = I[b][q]
  // twiddles W16^(b q)
  v[1 + 4] = cmul(v[1 + 4], make_double2(C1, -S1));      // b=1,q=1 : W^1
  v[1 + 8] = cmul(v[1 + 8], make_double2(H, -H));        // b=1,q=2 : W^2
  v[1 + 12] = cmul(v[1 + 12], make_double2(S1, -C1));    // b=1,q=3 : W^3
  v[2 + 4] = cmul(v[2 + 4], make_double2(H, -H));        // b=2,q=1 : W^2
  v[2 + 8] = make_double2(v[2 + 8].y, -v[2 + 8].x);      // b=2,q=2 : W^4 = -i
  v[2 + 12] = cmul(v[2 + 12], make_double2(-H, -H));     // b=2,q=3 : W^6
  v[3 + 4] = cmul(v[3 + 4], make_double2(S1, -C1));      // b=3,q=1 : W^3
  v[3 + 8] = cmul(v[3 + 8], make_double2(-H, -H));       // b=3,q=2 : W^6
  v[3 + 12] = cmul(v[3 + 12], make_double2(-C1, S1));    // b=3,q=3 : W^9
#pragma unroll
  for (int q = 0; q < 4; ++q) dft4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);   // v[p + 4 q] = X[q + 4 p]
}

// forward 8-point DFT in registers; on return X[k] is in v[2 k] (k < 4) and v[2 (k - 4) + 1] (k >= 4)
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
  constexpr double H = 0.70710678118654752440;
  dft4(v[0], v[2], v[4], v[6]);   // v[2 q] = I[0][q]
  dft4(v[1], v[3], v[5], v[7]);   // v[1 + 2 q] = I[1][q]
  v[3] = cmul(v[3], make_double2(H, -H));
  v[5] = make_double2(v[5].y, -v[5].x);
  v[7] = cmul(v[7], make_double2(-H, -H));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double2 a = v[2 * q], b = v[2 * q + 1];
    v[2 * q] = cadd(a, b);
    v[2 * q + 1] = csub(a, b);
  }
}

// e^(-2 pi i t / 2048) from the table tw[j] = e^(-2 pi i j / 4096), j < 2048
__device__ __forceinline__ double2 w2048(const double2* __restrict__ tw, int t) {
  t &= (FFT_M - 1);
  if (t < FFT_M / 2) return __ldg(&tw[2 * t]);
  const double2 w = __ldg(&tw[2 * t - FFT_M]);
  return make_double2(-w.x, -w.y);
}

// powers w^1 .. w^(R-1) with multiplication depth <= 4 (rounding ~ 4 ulp)
template <int R>
__device__ __forceinline__ void twiddle_powers(double2 w1, double2 (&w)[R]) {
  w[0] = make_double2(1., 0.);
  w[1] = w1;
  w[2] = cmul(w1, w1);
  w[3] = cmul(w[2], w1);
  w[4] = cmul(w[2], w[2]);
  w[5] = cmul(w[4], w1);
  w[6] = cmul(w[4], w[2]);
  w[7] = cmul(w[4], w[3]);
  if constexpr (R > 8) {
    w[8] = cmul(w[4], w[4]);
#pragma unroll
    for (int r = 1; r < 8; ++r) w[8 + r] = cmul(w[8], w[r]);
  }
}

// passes 2 and 3 of the 2048-point Stockham FFT (pass 1 is done by the caller, which owns the loads); data in D (padded indexing)
__device__ __forceinline__ void fft2048_tail(double2* D, const double2* __restrict__ tw, int tid) {
  {  // pass 2: radix 16, Ns = 16
    const int j = tid, k = j & 15;
    double2 v[16], w[16];
    twiddle_powers<16>(w2048(tw, k * 8), w);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = D[fidx(j + r * 128)];
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], w[r]);
    dft16(v);
    __syncthreads();
    const int j0 = (j >> 4) * 256 + k;
#pragma unroll
    for (int r = 0; r < 16; ++r) D[fidx(j0 + r * 16)] = v[(r >> 2) + 4 * (r & 3)];
    __syncthreads();
  }
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {  // pass 3: radix 8, Ns = 256 (in place: every thread writes where it read)
    const int k = tid + h * 128;
    double2 v[8], w[8];
    twiddle_powers<8>(w2048(tw, k), w);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = D[fidx(k + r * 256)];
#pragma unroll
    for (int r = 1; r < 8; ++r) v[r] = cmul(v[r], w[r]);
    dft8(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) D[fidx(k + r * 256)] = r < 4 ? v[2 * r] : v[2 * (r - 4) + 1];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FFT_THREADS, 5) fftlog_kernel(const DevGen g, long long N, double* __restrict__ F, int fstride) {
  extern __shared__ double2 D[];
  const int nkin = g.idata[0], nell = g.idata[1], ns = g.idata[8], poff = g.idata[12];
  const int Ng = g.idata[13], Np = g.idata[14], padl = g.idata[15];
  const int tid = threadIdx.x;
  const long long pt = blockIdx.x;
  const int ell = blockIdx.y;
  double* row = F + pt * fstride + g.foff;
  if (pt >= N) {
    for (int i = tid; i < ns; i += FFT_THREADS) row[2 + ell * ns + i] = 0.;
    if (ell == 0 && tid < 2) row[tid] = 0.;
    return;
  }
  double* pin = reinterpret_cast<double*>(D + FFT_SMEM_COMPLEX);
  const double* fd = g.ddata + g.idata[10];
  const int* fi = g.idata + g.idata[11];
  const double *c0 = fd, *c1 = fd + Ng;
  const double2* u = reinterpret_cast<const double2*>(fd + 2 * Ng) + (size_t)ell * (Np / 2 + 1);
  const double* post = fd + 2 * Ng + (size_t)nell * (Np / 2 + 1) * 2 + (size_t)ell * Ng;
  const double2* tw = reinterpret_cast<const double2*>(fd + 2 * Ng + (size_t)nell * (Np / 2 + 1) * 2 + (size_t)nell * Ng);
  const double* outw = reinterpret_cast<const double*>(tw + Np / 2) + (size_t)ell * ns;
  const int* j0tab = fi;
  const int* outj = fi + Ng + ell * ns;
  for (int i = tid; i < nkin; i += FFT_THREADS) pin[i] = row[poff + ell * nkin + i];
  __syncthreads();

  // resampled, pre-factored, zero-padded input sample x[i], i in [0, Np)
  auto xval = [&](int i) -> double {
    i -= padl;
    if (i < 0 || i >= Ng) return 0.;
    const int j = __ldg(&j0tab[i]);
    return __ldg(&c0[i]) * pin[j] + __ldg(&c1[i]) * pin[j + 1];
  };

  {  // forward FFT, pass 1: radix 16, Ns = 1, inputs z[n] = x[2n] + i x[2n+1] built on the fly
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) { const int n = tid + r * 128; v[r] = make_double2(xval(2 * n), xval(2 * n + 1)); }
    dft16(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) D[fidx(tid * 16 + r)] = v[(r >> 2) + 4 * (r & 3)];
    __syncthreads();
  }
  fft2048_tail(D, tw, tid);

  // untangle (real-input FFT) -> multiply by u_ell -> tangle (Hermitian -> real inverse), pairs (m, M - m)
  constexpr int M = FFT_M;
  auto T = [](double2 a, double2 bconj, double2 e) {   // (a + b*) / 2 - i/2 e (a - b*), with bconj = conj(b)
    const double2 s = cadd(a, bconj), d = csub(a, bconj);
    const double2 ed = cmul(e, d);
    return make_double2(0.5 * (s.x + ed.y), 0.5 * (s.y - ed.x));
  };
  for (int m = tid; m <= M / 2; m += FFT_THREADS) {
    if (m == 0) {
      const double2 z = D[fidx(0)];
      const double2 y0 = cmul(__ldg(&u[0]), make_double2(z.x + z.y, 0.));
      const double2 yM = cmul(__ldg(&u[M]), make_double2(z.x - z.y, 0.));
      D[fidx(0)] = T(y0, cconj(yM), make_double2(1., 0.));
    } else if (m == M / 2) {
      const double2 z = D[fidx(m)];
      const double2 y = cmul(__ldg(&u[m]), cconj(z));
      D[fidx(m)] = cconj(y);
    } else {
      const double2 za = D[fidx(m)], zb = D[fidx(M - m)];
      const double2 e = __ldg(&tw[m]);                       // e^(-2 pi i m / Np)
      const double2 e2 = make_double2(-e.x, e.y);            // e^(-2 pi i (M - m) / Np) = -conj(e)
      const double2 xa = T(za, cconj(zb), e), xb = T(zb, cconj(za), e2);
      const double2 ya = cmul(__ldg(&u[m]), xa), yb = cmul(__ldg(&u[M - m]), xb);
      D[fidx(m)] = T(ya, cconj(yb), e);
      D[fidx(M - m)] = T(yb, cconj(ya), e2);
    }
  }
  __syncthreads();

  {  // second FFT, pass 1: radix 16, Ns = 1
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = D[fidx(tid + r * 128)];
    dft16(v);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) D[fidx(tid * 16 + r)] = v[(r >> 2) + 4 * (r & 3)];
    __syncthreads();
  }
  fft2048_tail(D, tw, tid);

  const double invm = 1. / (double)M;
  const int padout = Np - Ng - padl;   // pad_out_left = npad - npad // 2
  auto grid = [&](int j) -> double {   // xi on the FFTLog output grid, index j in [0, Ng)
    const int t = j + padout;
    const double2 gq = D[fidx(t >> 1)];
    return ((t & 1) ? -gq.y : gq.x) * invm * __ldg(&post[j]);
  };
  for (int i = tid; i < ns; i += FFT_THREADS) {
    const int j = outj[i];
    const double x0 = grid(j), x1 = grid(j + 1);
    row[2 + ell * ns + i] = x0 + outw[i] * (x1 - x0);
  }
  if (ell == 0 && tid == 0) { row[0] = 1.; row[1] = 0.; }
}

}  // namespace

std::atomic<bool> g_timing{false};
cudaEvent_t g_fft_ev[2] = {nullptr, nullptr};

int last_fftlog_ms(double* ms) {
  *ms = 0.;
  if (!g_fft_ev[0]) return DLB_OK;
  float t = 0.f;
  if (cudaEventElapsedTime(&t, g_fft_ev[0], g_fft_ev[1]) != cudaSuccess) { cudaGetLastError(); return DLB_OK; }
  *ms = t;
  return DLB_OK;
}

int launch_theory_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream) {
  if (g.kind == DLB_GEN_KAISER) {
    // sizes needed on the host for the shared-memory budget are mirrored in aux[] at plan creation (csrc/plan.cu)
    const int nmu = g.aux[5] & 0xFF, nseg = (g.aux[5] >> 8) & 0xFF, interp = (g.aux[5] >> 16) & 0xFF, nell = g.aux[3], nw_max = g.aux[7];
    const int wd = kaiser_warp_doubles(nw_max, nmu, nell, nseg, interp);
    const size_t smem = (size_t)KAISER_WARPS * wd * sizeof(double);
    auto launch = [&](auto kernel) -> int {
      DLB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
      kernel<<<(unsigned)((npad + KAISER_WARPS - 1) / KAISER_WARPS), KAISER_WARPS * 32, smem, stream>>>(g, params, d, N, npad, F, fstride, wd);
      return DLB_OK;
    };
    if (nmu == 8) DLB_TRY(interp == 0 ? launch(kaiser_kernel<8, 0>) : launch(kaiser_kernel<8, 1>));
    else DLB_TRY(interp == 0 ? launch(kaiser_kernel<0, 0>) : launch(kaiser_kernel<0, 1>));
    g_launches++;
    DLB_CUDA(cudaGetLastError());
    return DLB_OK;
  }
  if (g.kind == DLB_GEN_BAO) {
    bao_kernel<<<(unsigned)((npad + BAO_WARPS - 1) / BAO_WARPS), BAO_WARPS * 32, 0, stream>>>(g, params, d, N, npad, F, fstride);
    g_launches++;
    DLB_CUDA(cudaGetLastError());
    if (g.aux[4] & 2) {   // flags bit 1: P -> xi through FFTLog
      const int Np = g.aux[5], nkin = g.aux[2], nell = g.aux[3];
      if (Np != 2 * FFT_M) { set_error("FFTLog kernel is specialised for the reference's 2048-point grid padded to 4096"); return DLB_ERR_UNSUPPORTED; }
      const size_t smem = (size_t)FFT_SMEM_COMPLEX * sizeof(double2) + (size_t)((nkin + 1) & ~1) * sizeof(double);
      static size_t configured_dev[64] = {};
      size_t& configured = configured_dev[current_device_slot()];
      if (smem > 48 * 1024 && smem > configured) {
        DLB_CUDA(cudaFuncSetAttribute(fftlog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      dim3 grid((unsigned)npad, nell);
      if (g_timing.load()) {
        if (!g_fft_ev[0]) { DLB_CUDA(cudaEventCreate(&g_fft_ev[0])); DLB_CUDA(cudaEventCreate(&g_fft_ev[1])); }
        DLB_CUDA(cudaEventRecord(g_fft_ev[0], stream));
      }
      fftlog_kernel<<<grid, FFT_THREADS, smem, stream>>>(g, N, F, fstride);
      if (g_timing.load()) DLB_CUDA(cudaEventRecord(g_fft_ev[1], stream));
      g_launches++;
      DLB_CUDA(cudaGetLastError());
    }
    return DLB_OK;
  }
  set_error("feature generator kind " + std::to_string(g.kind) + " not implemented");
  return DLB_ERR_UNSUPPORTED;
}

}  // namespace dlb

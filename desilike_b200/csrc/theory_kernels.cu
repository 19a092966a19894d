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
  int nt;
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

__device__ __forceinline__ SplineConst load_spline(const double* d, int nt) {
  // layout: x[nt] invh[nt] w[nt] invb[nt] cb[nt] then (r0, r1, x0, inv_dx)
  SplineConst sc;
  sc.nt = nt; sc.x = d; sc.invh = d + nt; sc.w = d + 2 * nt; sc.invb = d + 3 * nt; sc.cb = d + 4 * nt;
  sc.r0 = d[5 * nt]; sc.r1 = d[5 * nt + 1]; sc.x0 = d[5 * nt + 2]; sc.inv_dx = d[5 * nt + 3];
  return sc;
}
constexpr int SPLINE_DOUBLES(int nt) { return 5 * nt + 4; }

// ---------------------------------------------------------------------------------------------------------------------
// Kaiser.  ddata = [eta, a_inv_unused, f_fid, 0 | k[nk] | mu[nmu] | wmu[nell, nmu] | pk_fid[nt] | Tt[nt] | Lt[nt] | spline consts]
//   Tt = tanh(a ln(k_t / kp)) / a,  Lt = ln(k_t / kp)   =>  pk_dd = pk_fid * exp(dm * Tt + dn * Lt)
// features = (1, 0 | pk_dd [nell, nk] | pk_dt | pk_tt | pk11_l0 = pk_dd[ell = 0] tiled over ell), tables padded to ntab.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int KAISER_WARPS = 4;
constexpr int MAX_ELL = 3;

__global__ void __launch_bounds__(KAISER_WARPS * 32) kaiser_kernel(const DevGen g, const double* __restrict__ params, int d, long long N,
                                                                   long long npad, double* __restrict__ F, int fstride) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = g.idata[0], nell = g.idata[1], nmu = g.idata[2], nt = g.idata[3], apmode = g.idata[4], ntab = g.idata[6];
  const int ell0 = g.idata[5] >> 8;   // index of ell = 0 among the multipoles (or -1)
  const long long pt = (long long)blockIdx.x * KAISER_WARPS + warp;
  if (pt >= npad) return;
  double* row = F + pt * fstride + g.foff;
  if (pt >= N) {   // padding rows: zeros
    for (int j = lane; j < g.fpad; j += 32) row[j] = 0.;
    return;
  }
  double* y = sm + (size_t)warp * 2 * nt;
  double* M = y + nt;
  const double* dd = g.ddata;
  const double eta = dd[0], f_fid = dd[2];
  const double* kk = dd + 4;
  const double* mu = kk + nk;
  const double* wmu = mu + nmu;
  const double* pk_fid = wmu + nell * nmu;
  const double* Tt = pk_fid + nt;
  const double* Lt = Tt + nt;
  const SplineConst sc = load_spline(Lt + nt, nt);
  double par[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) par[j] = (j < g.nparams && g.pidx[j] >= 0) ? params[pt * d + g.pidx[j]] : g.pdef[j];
  double qpar, qper;
  ap_qpar_qper(apmode, par[0], par[1], eta, qpar, qper);
  const double dm = par[2], dn = par[3], f = f_fid * par[4], sigmapar = par[5], sigmaper = par[6];
  for (int i = lane; i < nt; i += 32) y[i] = pk_fid[i] * exp(dm * Tt[i] + dn * Lt[i]);
  __syncwarp();
  spline_second_derivatives(sc, y, M, lane);
  const double qap = qpar / qper, jac = 1. / (qpar * qper * qper), iq2 = 1. / (qap * qap) - 1.;
  const bool damped = (sigmapar != 0.) || (sigmaper != 0.);
  if (lane == 0) { row[0] = 1.; row[1] = 0.; }
  for (int ik = lane; ik < nk; ik += 32) {
    double sdd[MAX_ELL] = {0., 0., 0.}, sdt[MAX_ELL] = {0., 0., 0.}, stt[MAX_ELL] = {0., 0., 0.};
    const double k = kk[ik];
    for (int im = 0; im < nmu; ++im) {
      const double m_ = mu[im];
      const double fap = sqrt(1. + m_ * m_ * iq2);
      const double kap = k / qper * fap, muap = m_ / qap / fap;
      double p = jac * spline_eval(sc, y, M, log10(kap));
      if (damped) p *= exp(-0.5 * kap * kap * (sigmapar * sigmapar * muap * muap + sigmaper * sigmaper * (1. - muap * muap)));
      const double m2 = muap * muap;
      const double pdt = f * m2 * p, ptt = f * f * m2 * m2 * p;
#pragma unroll
      for (int l = 0; l < MAX_ELL; ++l)
        if (l < nell) { const double w = wmu[l * nmu + im]; sdd[l] += p * w; sdt[l] += pdt * w; stt[l] += ptt * w; }
    }
#pragma unroll
    for (int l = 0; l < MAX_ELL; ++l)
      if (l < nell) {
        row[2 + l * nk + ik] = sdd[l];
        row[2 + ntab + l * nk + ik] = sdt[l];
        row[2 + 2 * ntab + l * nk + ik] = stt[l];
        row[2 + 3 * ntab + l * nk + ik] = ell0 >= 0 ? sdd[ell0 < MAX_ELL ? ell0 : 0] : 0.;
      }
  }
  // zero the padding between / after the tables
  const int nin = nell * nk;
  for (int t = 0; t < 4; ++t)
    for (int j = nin + lane; j < ntab; j += 32) row[2 + t * ntab + j] = 0.;
  for (int j = 2 + 4 * ntab + lane; j < g.fpad; j += 32) row[j] = 0.;
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
  const SplineConst sc = load_spline(M_now + nt, nt);
  double par[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) par[j] = (j < g.nparams && g.pidx[j] >= 0) ? params[pt * d + g.pidx[j]] : g.pdef[j];
  double qpar, qper;
  ap_qpar_qper(apmode, par[0], par[1], eta, qpar, qper);
  const double b1 = par[2], f = par[3] * f_fid * par[7], sigmas = par[4], sigmapar = par[5], sigmaper = par[6];
  const double qap = qpar / qper, iq2 = 1. / (qap * qap) - 1.;
  const bool reciso = flags & 1;
  if (lane == 0) { row[0] = 1.; row[1] = 0.; }
  for (int ik = lane; ik < nk; ik += 32) {
    double acc[MAX_ELL] = {0., 0., 0.};
    const double k = kk[ik];
    const double sk = reciso ? exp(-0.5 * (k * rs) * (k * rs)) : 0.;
    const double pknow = pknow_k[ik];
    for (int im = 0; im < nmu; ++im) {
      const double m_ = mu[im];
      const double fap = sqrt(1. + m_ * m_ * iq2);
      const double kap = k / qper * fap, muap = m_ / qap / fap;
      const double lk = log10(kap);
      const double pkwap = spline_eval(sc, pk_t, M_pk, lk) - spline_eval(sc, pknow_t, M_now, lk);
      const double snl2 = kap * kap * (sigmapar * sigmapar * muap * muap + sigmaper * sigmaper * (1. - muap * muap));
      const double ca = b1 + f * muap * muap * (1. - sk);
      const double Cap = ca * ca * exp(-0.5 * snl2);
      const double fg = 1. + 0.5 * (sigmas * k * m_) * (sigmas * k * m_);
      const double cb = b1 + f * m_ * m_ * (1. - sk);
      const double B = cb * cb / (fg * fg);
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
// oracle/fftlog.py).  One CTA per (point, ell): resample onto the 2048-point log grid (+ high-k extrapolation), multiply by
// k^(3-q)/(2 pi)^(3/2), zero-pad to 4096, radix-2 DIF forward FFT (natural -> bit-reversed), multiply by u_ell(m)
// (Hermitian extension handled analytically: for m > N/2 the factor is conj(u[N-m])), radix-2 DIT forward FFT
// (bit-reversed -> natural) == numpy's hfft, scale by s^-q (-1)^(ell/2) / N, un-pad, linear interpolation onto s.
// FFT block of ddata at idata[10]: wgt[Ng] lh[Ng] dh[Ng] pre[Ng] u[nell][Np/2+1][2] post[nell][Ng] (slope_inv, 0) tw[Np/2][2] outw[nell][ns]
// FFT block of idata at idata[11]: jmap[Ng] outj[nell][ns]
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FFT_THREADS = 256;

__device__ __forceinline__ unsigned bitrev(unsigned x, int bits) { return __brev(x) >> (32 - bits); }

__global__ void __launch_bounds__(FFT_THREADS) fftlog_kernel(const DevGen g, long long N, double* __restrict__ F, int fstride) {
  extern __shared__ double2 cdata[];
  const int nkin = g.idata[0], nell = g.idata[1], ns = g.idata[8], poff = g.idata[12];
  const int Ng = g.idata[13], Np = g.idata[14], padl = g.idata[15];
  const int bits = 31 - __clz(Np);
  const long long pt = blockIdx.x;
  const int ell = blockIdx.y;
  double* row = F + pt * fstride + g.foff;
  if (pt >= N) {
    for (int i = threadIdx.x; i < ns; i += FFT_THREADS) row[2 + ell * ns + i] = 0.;
    if (ell == 0 && threadIdx.x < 2) row[threadIdx.x] = 0.;
    return;
  }
  double* pin = reinterpret_cast<double*>(cdata + Np);
  const double* fd = g.ddata + g.idata[10];
  const int* fi = g.idata + g.idata[11];
  const double *wgt = fd, *lh = wgt + Ng, *dh = lh + Ng, *pre = dh + Ng;
  const double2* u = reinterpret_cast<const double2*>(pre + Ng) + (size_t)ell * (Np / 2 + 1);
  const double* post0 = pre + Ng + (size_t)nell * (Np / 2 + 1) * 2;
  const double* post = post0 + (size_t)ell * Ng;
  const double slope_inv = post0[(size_t)nell * Ng];            // followed by one pad double (keeps tw 16-byte aligned)
  const double2* tw = reinterpret_cast<const double2*>(post0 + (size_t)nell * Ng + 2);
  const double* outw = post0 + (size_t)nell * Ng + 2 + Np + (size_t)ell * ns;
  const int* jmap = fi;
  const int* outj = fi + Ng + ell * ns;
  for (int i = threadIdx.x; i < nkin; i += FFT_THREADS) pin[i] = row[poff + ell * nkin + i];
  __syncthreads();
  const double plast = pin[nkin - 1];
  const double slope = (plast - pin[nkin - 2]) * slope_inv;
  for (int n = threadIdx.x; n < Np; n += FFT_THREADS) {
    const int i = n - padl;
    double v = 0.;
    if (i >= 0 && i < Ng) {
      const int j = jmap[i];
      v = j >= 0 ? pin[j] + wgt[i] * (pin[j + 1] - pin[j]) : (plast + slope * lh[i]) * dh[i];
      v *= pre[i];
    }
    cdata[n] = make_double2(v, 0.);
  }
  __syncthreads();
  // forward DIF: natural order in, bit-reversed order out
  for (int half = Np >> 1; half >= 1; half >>= 1) {
    const int tstride = (Np >> 1) / half;
    for (int b = threadIdx.x; b < (Np >> 1); b += FFT_THREADS) {
      const int j = b & (half - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + half;
      const double2 a = cdata[i0], c = cdata[i1];
      const double2 w = __ldg(&tw[j * tstride]);
      const double dr = a.x - c.x, di = a.y - c.y;
      cdata[i0] = make_double2(a.x + c.x, a.y + c.y);
      cdata[i1] = make_double2(dr * w.x - di * w.y, dr * w.y + di * w.x);
    }
    __syncthreads();
  }
  // multiply by the kernel u_ell(m); position p holds X[bitrev(p)]
  for (int p = threadIdx.x; p < Np; p += FFT_THREADS) {
    const int m = (int)bitrev((unsigned)p, bits);
    double2 f = m <= (Np >> 1) ? __ldg(&u[m]) : __ldg(&u[Np - m]);
    if (m > (Np >> 1)) f.y = -f.y;
    const double2 x = cdata[p];
    cdata[p] = make_double2(x.x * f.x - x.y * f.y, x.x * f.y + x.y * f.x);
  }
  __syncthreads();
  // forward DIT: bit-reversed order in, natural order out
  for (int half = 1; half < Np; half <<= 1) {
    const int tstride = (Np >> 1) / half;
    for (int b = threadIdx.x; b < (Np >> 1); b += FFT_THREADS) {
      const int j = b & (half - 1);
      const int i0 = ((b - j) << 1) + j, i1 = i0 + half;
      const double2 a = cdata[i0], c = cdata[i1];
      const double2 w = __ldg(&tw[j * tstride]);
      const double tr = c.x * w.x - c.y * w.y, ti = c.x * w.y + c.y * w.x;
      cdata[i0] = make_double2(a.x + tr, a.y + ti);
      cdata[i1] = make_double2(a.x - tr, a.y - ti);
    }
    __syncthreads();
  }
  const double invn = 1. / (double)Np;
  const int padout = Np - Ng - padl;   // pad_out_left = npad - npad // 2
  for (int i = threadIdx.x; i < ns; i += FFT_THREADS) {
    const int j = outj[i];
    const double x0 = cdata[j + padout].x * invn * post[j], x1 = cdata[j + 1 + padout].x * invn * post[j + 1];
    row[2 + ell * ns + i] = x0 + outw[i] * (x1 - x0);
  }
  if (ell == 0 && threadIdx.x == 0) { row[0] = 1.; row[1] = 0.; }
}

}  // namespace

int launch_theory_features(const DevGen& g, const double* params, int d, long long N, long long npad, double* F, int fstride, cudaStream_t stream) {
  if (g.kind == DLB_GEN_KAISER) {
    // nt is needed on the host for the shared-memory size: it is mirrored in aux[1] at plan creation
    const int nt = g.aux[1];
    const size_t smem = (size_t)KAISER_WARPS * 2 * nt * sizeof(double);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      DLB_CUDA(cudaFuncSetAttribute(kaiser_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    kaiser_kernel<<<(unsigned)((npad + KAISER_WARPS - 1) / KAISER_WARPS), KAISER_WARPS * 32, smem, stream>>>(g, params, d, N, npad, F, fstride);
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
      const size_t smem = (size_t)Np * sizeof(double2) + (size_t)((nkin + 1) & ~1) * sizeof(double);
      static size_t configured = 0;
      if (smem > 48 * 1024 && smem > configured) {
        DLB_CUDA(cudaFuncSetAttribute(fftlog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      dim3 grid((unsigned)npad, nell);
      fftlog_kernel<<<grid, FFT_THREADS, smem, stream>>>(g, N, F, fstride);
      g_launches++;
      DLB_CUDA(cudaGetLastError());
    }
    return DLB_OK;
  }
  set_error("feature generator kind " + std::to_string(g.kind) + " not implemented");
  return DLB_ERR_UNSUPPORTED;
}

}  // namespace dlb

// C-ABI layer: plan construction (constant upload + pre-tiling of the term matrices into the contraction kernel's
// shared-memory image), workspace management and the batched entry points declared in include/desilike_b200.h.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <cstdio>
#include <cstdlib>

namespace dlb {

static thread_local std::string g_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

// A likelihood block whose terms all read the SAME wide feature segment (the last hidden layer of a wide MLP emulator) evaluated on the tcgen05 int8
// path (opt-in, dlb_plan_desc::mlp_slices): ONE FP64-grade GEMM D[point, term, column] = features . [M_1 | M_2 | ...] (csrc/ozaki.cu), then
// ozaki_gram_kernel forms the slot values from the per-point bias coefficients and reduces them to the Gram partials of contract_kernel's layout.
struct OzContract {
  bool on = false;
  int K = 0, fcol = 0, nterms = 0, ncols = 0, nconst = 0, ncolx = 0, ldD = 0, R = 1;
  int tslot[OZ_GRAM_MAXT], tcoef[OZ_GRAM_MAXT];
  int8_t* d_Bs = nullptr; double *d_Sb = nullptr, *d_r0 = nullptr;
};

struct Block {
  Variant var;
  OzContract oz;
  int ncols = 0, nslots = 1, nconst = 0, ncb = 0, nchunks = 0, npart = 0, coloff = 0, gen = -1, maxcol = 0;
  std::vector<Chunk> h_chunks;
  int* d_cb_first = nullptr; std::vector<int> cb_first;   // whitened tiles: first non-zero k-chunk per column block
  unsigned long long *d_tmask_white = nullptr, *d_tmask_plain = nullptr;   // [ncb][nchunks][2] sub-tile masks (contract.cu: mma_chunk_masked), NULL: dense
  std::vector<double> exec_white, exec_plain;   // [ncb * nchunks] multiplied (row, column) pairs of every (column block, k-chunk) tile
  int ncols_w = 0, ncb_plain = 0;   // columns of the whitened matrices (rank-reduced plans: < ncols) / column blocks of the un-whitened tiles
  int fill = 0, fbase = 0;   // resident A: 0 = Taylor monomials evaluated in the kernel, 1 = copied from the feature matrix at column fbase
  long long poff = 0;
  Chunk* d_chunks = nullptr;
  double *d_tiles_white = nullptr, *d_tiles_plain = nullptr, *d_epi_white = nullptr, *d_epi_plain = nullptr;
  EntryMap* d_map = nullptr;
  std::vector<EntryMap> h_map;
  int *d_qtile = nullptr, *d_qdent = nullptr; double* d_qconst = nullptr;   // quadratic-form tables (dlb_block_desc::q_*), tiles as packed 16-byte records
  int *d_qw_off = nullptr, *d_qw_coef = nullptr; double* d_qw_val = nullptr;
  double* d_tdata = nullptr;   // 'cubic' observable transform: data of the transformed columns (two-stage plans)
  int tcol = 0;                // first column of the block in the transformed-theory buffer of a two-stage plan
};

// launch configuration of one persistent contraction launch (a list of blocks sharing one kernel variant)
struct Launch {
  Variant var;
  int nblocks = 0, nstages = 2, resident = 0, ast = 0, npart_max = 1;
  BlockDev *d_white = nullptr, *d_plain = nullptr;
  BlockDev* d_plain_t = nullptr;   // two-stage plans: the un-whitened table writing into the transformed-theory buffer (ocol0 = Block::tcol)
  SchedEntry *d_sched_white = nullptr, *d_sched_plain = nullptr;   // per-tile TMA schedule
  int nsched = 0, nsched_plain = 0;
  bool masked_white = false, masked_plain = false;   // some block has sub-tile masks
  double flops_exec = 0., flops_exec_plain = 0.;   // 2 * kc * BN * slots-independent tile count per point (whitened / un-whitened tables)
};

// one dense layer of an MLP emulator on the tcgen05 int8 path (csrc/ozaki.cu) -- or, for narrow fan-in, the plain FP64 layer kernel
struct OzLayer { int K = 0, N = 0, in_off = 0, out_off = 0; bool ozaki = false; double *d_W = nullptr, *d_bias = nullptr, *d_Sb = nullptr; int8_t* d_Bs = nullptr; };
struct MlpLayers { int act = 0; std::vector<Block> layers; std::vector<Launch> launches; std::vector<OzLayer> oz; };

}  // namespace dlb

constexpr int MAX_SUB = 8;   // sub-batches of the pipelined host-buffer path

struct dlb_plan {
  int device = 0;
  int d = 0, ngen = 0, ncoef = 0, fstride = 4, ns = 0, nmarg = 0, ntheory = 0, nsm = 1, dpad = 1;   // d: columns of the EXTENDED parameter row (inputs + derived)
  int d_in = 0, nder = 0;              // input columns of the batch / derived-expression parameters appended to every row
  int *d_der_off = nullptr, *d_der_code = nullptr; double* d_der_const = nullptr;
  double* w_pext = nullptr;            // [cap, d] extended parameter rows (plans with derived parameters)
  int mlp_slices = 0, oz_maxk = 0;     // > 0: MLP hidden layers on the tcgen05 int8 path with this many slices (opt-in)
  int8_t* w_oz_slices = nullptr; int* w_oz_exps = nullptr;
  bool oz_contract = false; int oz_ldD = 0; double* w_oz_D = nullptr;   // every likelihood block runs as a tcgen05 GEMM + ozaki_gram_kernel
  bool resident = false;               // Taylor monomials evaluated inside the contraction kernel (no feature matrix in HBM)
  dlb::Launch launch;                  // the likelihood blocks
  // two-stage plans (an observable transform that is non-linear in the theory, reference power_spectrum.py:400-404): stage 1 = `launch` in output
  // mode with the un-whitened matrices and the transform in its epilogue -> w_T [npad, tstride]; stage 2 = `wlaunch`: r = u . L^T - L data
  // quadratic-form plans (dlb_block_desc::q_*): the likelihood evaluation runs `qlaunch` over `qblocks` (ONE term per block: features . q_matrix, single
  // accumulator per column, Gram entries formed in the epilogue); `launch` / `blocks` keep serving the theory-vector entry
  bool qform = false;
  dlb::Launch qlaunch;
  std::vector<dlb::Block> qblocks;
  bool two_stage = false;
  dlb::Launch wlaunch;
  std::vector<dlb::Block> wblocks;
  int tstride = 0;
  double* w_T = nullptr;
  dlb::GenDev* d_gens = nullptr;
  std::vector<unsigned long long*> d_paths;   // per generator (resident mode)
  std::vector<unsigned long long*> d_mono2;   // per generator: two-pass monomial table (NULL: factor-list evaluation)
  std::vector<int> mono_n1, mono_n2;
  std::vector<dlb::DevGen> gens;
  std::vector<dlb::Block> blocks;
  std::vector<dlb::MlpLayers> mlps;   // one per generator (empty unless kind == MLP)
  size_t mlp_fused_smem = 0;           // > 0: every MLP emulator runs in the fused kernel (csrc/mlp.cu)
  int mlp_groups = 1;                  // warp groups per CTA of the fused kernel
  dlb::DevGen* d_devgens = nullptr; int* d_mlp_index = nullptr; int nmlp = 0;
  std::vector<std::pair<dlb::Launch, int>> mlp_launches;   // hidden layers of ALL MLP generators, one launch per (depth, activation, A mode)
  dlb::DevCoef* d_coefs = nullptr;
  int* d_prior_kind = nullptr;
  double *d_prior_lo = nullptr, *d_prior_hi = nullptr, *d_prior_loc = nullptr, *d_prior_scale = nullptr;
  double *d_M0 = nullptr, *d_g0 = nullptr, *d_x0 = nullptr, *d_xprec = nullptr, *d_xloc = nullptr, *d_xlo = nullptr, *d_xhi = nullptr;
  int* d_xkind = nullptr;
  int ncomp = 0;                       // independent groups of solved parameters (0: solve one dense system)
  int *d_comp_off = nullptr, *d_comp_idx = nullptr, *d_comp_nmarg = nullptr, *d_comp_pos = nullptr;
  double* d_comp_m0 = nullptr; int npos = 0;
  int *d_hess_ij = nullptr, *d_comp_of = nullptr;
  dlb::GroupEntry* d_gent = nullptr; int *d_gent_off = nullptr, *d_m0_off = nullptr; long long* d_chi_rows = nullptr; int nchi = 0, comp_max = 0;
  double* w_hess = nullptr; long long hess_cap = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined host-buffer path
  cudaEvent_t ev_pipe[2 * MAX_SUB + 2] = {};      // [c] sub-batch c copied in, [MAX_SUB + c] its kernels done, then pipeline start / end
  dlb::SolveBlock* d_solve_blocks = nullptr;
  long long ptotal = 0;  // partial rows (in units of npad doubles)
  std::vector<void*> owned;
  // workspace
  long long cap = 0;
  double *w_F = nullptr, *w_partial = nullptr, *w_params = nullptr, *w_out = nullptr, *w_theory = nullptr;
  int* w_status = nullptr;
  long long theory_cap = 0;
  double *w_prop = nullptr, *w_logpy = nullptr;   // ensemble sampler: proposals of one half and their log-posteriors
  long long prop_cap = 0;
  std::mutex mutex, sampler_mutex;
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  double ms_features = 0., ms_contract = 0., ms_solve = 0.;
};

namespace dlb {

template <typename T>
static int upload(dlb_plan* plan, const T* host, size_t count, T** dev) {
  *dev = nullptr;
  if (count == 0) count = 1;
  DLB_CUDA(cudaMalloc((void**)dev, count * sizeof(T)));
  plan->owned.push_back(*dev);
  if (host) DLB_CUDA(cudaMemcpy(*dev, host, count * sizeof(T), cudaMemcpyHostToDevice));
  else DLB_CUDA(cudaMemset(*dev, 0, count * sizeof(T)));
  return DLB_OK;
}

static bool is_device_ptr(const void* ptr) {
  if (!ptr) return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// Kernel variant for a launch whose blocks need `nslots` accumulator slots: the first candidate whose shared memory fits with
// at least 2 TMA stages.  Long k-chunks (one per term when K <= 76) amortise the per-chunk cost (stage release, barrier wait,
// coefficient loads) best; `maxK` is the longest contraction of the launch.
static Variant pick_variant(const dlb_plan* plan, int nslots, bool resident, int ast, int ncoef, int npart_max, int maxK, bool narrow = false) {
  std::vector<Variant> cand;
  const bool longk = maxK > 36;
  if (nslots <= 1) {
    if (narrow && !getenv("DLB_NO_NARROW")) {   // 64 x 80 tiles: less zero padding when the blocks' column counts fit 80-column blocks much better than 112
      if (maxK > 28) cand.push_back(Variant{1, 2, 5, 4, 2, 36});
      cand.push_back(Variant{1, 2, 5, 4, 2, 28});
    }
    if (maxK > 28) cand.push_back(Variant{1, 2, 7, 4, 2, 36});   // 2 CTAs / SM x 8 warps, 64 x 112 tiles, longer k-chunks when they fit
    cand.push_back(Variant{1, 2, 7, 4, 2, 28});              // 2 CTAs / SM x 8 warps, 64 x 112 tiles (measured: C2 62 %, C5 70 %)
    if (longk) cand.push_back(Variant{1, 2, 7, 6, 1, 76});   // 12 warps,  96 x 112 tile                (C2 60 %, C5 68 %)
    cand.push_back(Variant{1, 2, 7, 8, 1, 36});              // 16 warps, 128 x 112 tile                (C2 62 %, C5 69 %)
    if (const char* e = getenv("DLB_R1_VARIANT")) { const Variant v = cand[std::min<size_t>(atoi(e), cand.size() - 1)]; cand.assign(1, v); }
  } else if (nslots <= 4) {
    // (measured and rejected: 96 x 64 tiles, 2 m-tiles x 4 n-tiles per warp -- fewer fragment loads per DMMA but twice the column
    //  blocks and Gram epilogues: C4 58.9 - 60.6 % vs 61.7 % of FP64 peak)
    if (longk && !getenv("DLB_R4_WIDE")) cand.push_back(Variant{4, 1, 4, 8, 1, 76});   // 16 warps (4 per scheduler: slack for the DMMA issue rate), 64 x 64 tile (C4 63.1 %, C5 + marg. 64.6 %)
    if (longk) cand.push_back(Variant{4, 1, 7, 6, 1, 76});   // 12 warps,  48 x 112 tile, one k-chunk per term (C5 + marg. 59 %; 52 % with KC = 36)
    cand.push_back(Variant{4, 1, 7, 6, 1, 36});
  } else {
    if (longk) cand.push_back(Variant{8, 1, 4, 6, 1, 76});   // 12 warps,  48 x  64 tile
    cand.push_back(Variant{8, 1, 4, 6, 1, 36});
  }
  for (const Variant& v : cand)
    if (contract_smem_bytes(v, 2, resident, ast, plan->dpad, ncoef, npart_max, plan->ngen) <= contract_smem_limit(v.CPS)) return v;
  return cand.back();
}

static int round_ast(int n) {   // row stride of the resident A tile: = 4 (mod 8) keeps the 64-bit fragment loads bank-conflict free
  int a = std::max(n, 4);
  while ((a & 7) != 4) ++a;
  return a;
}

// Extent (in feature columns) covered by the k-chunks of a term of length K (mirrors the chunking in build_block_core).
static int chunk_extent(int K) { return (K + 3) & ~3; }   // chunks are [k0, min(k0 + kc, Kpad)): they never reach beyond Kpad

// fcol: absolute first feature column of the term in the per-point feature row; frel: the same relative to its generator's segment
struct TermSpec { int slot, coef, fcol, frel, K; const double* white; const double* plain; };  // matrices [K, ncols] row-major

// Chunk the terms, pre-tile their matrices into the contraction kernel's shared-memory image and upload everything.
static int build_block_core(dlb_plan* plan, const Variant& var, bool resident, int ncols, int nslots, const std::vector<TermSpec>& terms,
                            const double* r0_white, const double* t0_plain, int nconst, const double* gconst_white, Block& blk, int ncols_white = 0) {
  if (ncols_white <= 0) ncols_white = ncols;
  blk.ncols = ncols; blk.ncols_w = ncols_white; blk.nslots = nslots; blk.nconst = nconst;
  blk.var = var;
  const int BN = variant_bn(blk.var), KC = var.KC;
  // Auxiliary columns: the dot products of every slot with the constant whitened Jacobian rows g_i are linear in the features,
  //   slot_a . g_i = sum_t coef_t f_t . (M_t g_i),
  // so they are obtained for free from the contraction itself by appending the nconst columns M_t g_i to the term matrices
  // (usually inside the zero padding of the last column block) instead of nconst extra dot products per column block.
  blk.ncb = (ncols_white + nconst + BN - 1) / BN;
  blk.ncb_plain = (ncols + nconst + BN - 1) / BN;
  std::vector<Chunk> chunks;
  struct Src { int term; int k0; };
  std::vector<Src> srcs;
  blk.maxcol = 0;
  for (size_t t = 0; t < terms.size(); ++t) {
    const TermSpec& td = terms[t];
    const int Kpad = (td.K + 3) & ~3;
    const int nch = (Kpad + KC - 1) / KC;
    int kc = (Kpad + nch - 1) / nch;
    kc = (kc + 3) & ~3;
    for (int k0 = 0; k0 < Kpad; k0 += kc) {
      Chunk ch;
      ch.slot = td.slot; ch.coef = td.coef; ch.fcol = (resident ? td.frel : td.fcol) + k0; ch.kc = std::min(kc, Kpad - k0);
      chunks.push_back(ch);
      srcs.push_back(Src{(int)t, k0});
      blk.maxcol = std::max(blk.maxcol, td.frel + k0 + ch.kc);
    }
  }
  blk.nchunks = (int)chunks.size();
  blk.npart = nslots * (nslots + 1) / 2 + nslots * nconst;
  DLB_TRY(upload(plan, chunks.data(), chunks.size(), &blk.d_chunks));
  blk.h_chunks = chunks;
  // pre-tile: [ncb][nchunks][BN][KC], n-major rows, zero padded
  const size_t tile = (size_t)BN * KC;
  for (int pass = 0; pass < 2; ++pass) {
    bool any = false;
    for (const TermSpec& td : terms) any = any || (pass == 0 ? td.white : td.plain) != nullptr;
    if (!any) continue;
    const int ncols = pass == 0 ? blk.ncols_w : blk.ncols, ncols_ext = ncols + nconst, ncb = pass == 0 ? blk.ncb : blk.ncb_plain;   // this pass's matrix width
    std::vector<double> tiles((size_t)ncb * blk.nchunks * tile, 0.);
    for (int cb = 0; cb < ncb; ++cb)
      for (int c = 0; c < blk.nchunks; ++c) {
        double* dst = tiles.data() + ((size_t)cb * blk.nchunks + c) * tile;
        const TermSpec& td = terms[srcs[c].term];
        const double* mat = pass == 0 ? td.white : td.plain;
        if (!mat) continue;
        for (int n = 0; n < BN; ++n) {
          const int col = cb * BN + n;
          if (col >= ncols_ext) break;
          for (int k = 0; k < chunks[c].kc; ++k) {
            const int kk = srcs[c].k0 + k;
            if (kk >= td.K) continue;
            if (col < ncols) dst[(size_t)n * KC + k] = mat[(long long)kk * ncols + col];
            else if (pass == 0) {   // auxiliary column i = col - ncols: (M_t g_i)[kk]
              const double* g = gconst_white + (size_t)(col - ncols) * ncols;
              double acc = 0.;
              for (int j = 0; j < ncols; ++j) acc += mat[(long long)kk * ncols + j] * g[j];
              dst[(size_t)n * KC + k] = acc;
            }
          }
        }
      }
    DLB_TRY(upload(plan, tiles.data(), tiles.size(), pass == 0 ? &blk.d_tiles_white : &blk.d_tiles_plain));
    {
      // Sub-tile masks.  Warp wn of a pair owns the 8-column n-tiles j = 2 nt + wn of a column block.  For every (column block, k-chunk,
      // warp): first k4 step (4 rows) at which n-tile nt holds a non-zero, made monotone in nt (min over nt' >= nt) so that the active
      // n-tiles at every step form a prefix; encoded as 3 bits per k4 step = number of active n-tiles.  ~0 = multiply everything.
      const int NT = var.NT, nj = 2 * NT;
      std::vector<unsigned long long> masks((size_t)ncb * blk.nchunks * 2, ~0ull);
      std::vector<double>& exec = pass == 0 ? blk.exec_white : blk.exec_plain;
      exec.assign((size_t)ncb * blk.nchunks, 0.);
      bool any_mask = false;
      for (int cb = 0; cb < ncb; ++cb)
        for (int c = 0; c < blk.nchunks; ++c) {
          const double* t = tiles.data() + ((size_t)cb * blk.nchunks + c) * tile;
          const int nk4 = chunks[c].kc >> 2;
          std::vector<int> kf(nj, nk4);   // first non-zero k4 step of n-tile j (nk4: none)
          for (int j = 0; j < nj; ++j)
            for (int k4 = 0; k4 < nk4 && kf[j] == nk4; ++k4)
              for (int n = 8 * j; n < 8 * j + 8 && kf[j] == nk4; ++n)
                for (int k = 4 * k4; k < 4 * k4 + 4; ++k)
                  if (t[(size_t)n * KC + k] != 0.) { kf[j] = k4; break; }
          for (int wn = 0; wn < 2; ++wn) {
            std::vector<int> start(NT);
            int run = nk4;
            for (int nt = NT - 1; nt >= 0; --nt) { run = std::min(run, kf[2 * nt + wn]); start[nt] = run; }
            unsigned long long w = 0ull;
            bool full = true;
            int nruns = 0, run_len = 0, run_nact = -1;
            auto flush = [&]() {
              if (run_len > 0) { w |= ((unsigned long long)run_len | ((unsigned long long)run_nact << 5)) << (8 * nruns); ++nruns; }
            };
            for (int k4 = 0; k4 < nk4; ++k4) {
              int nact = 0;
              for (int nt = 0; nt < NT; ++nt) nact += start[nt] <= k4 ? 1 : 0;
              full = full && nact == NT;
              if (nact != run_nact) { flush(); run_len = 0; run_nact = nact; }
              ++run_len;
              exec[(size_t)cb * blk.nchunks + c] += 4. * 8. * nact;
            }
            flush();
            if (nruns > 8) full = true;   // cannot happen: the count is non-decreasing in k4 and takes at most NT + 1 <= 8 values
            if (nk4 > 31) full = true;   // cannot happen (KC <= 76 -> nk4 <= 19 fits the 5-bit run length)
            if (!full && !getenv("DLB_NO_SUBTILE_SKIP")) { masks[((size_t)cb * blk.nchunks + c) * 2 + wn] = w; any_mask = true; }
            else exec[(size_t)cb * blk.nchunks + c] += 0.;
          }
          if (getenv("DLB_NO_SUBTILE_SKIP")) exec[(size_t)cb * blk.nchunks + c] = (double)chunks[c].kc * BN;
        }
      if (any_mask) DLB_TRY(upload(plan, masks.data(), masks.size(), pass == 0 ? &blk.d_tmask_white : &blk.d_tmask_plain));
    }
    if (pass == 0 && !getenv("DLB_NO_TILE_SKIP")) {   // leading all-zero tiles per column block (rank-reduced problems are lower triangular)
      blk.cb_first.assign(ncb, 0);
      bool any_skip = false;
      for (int cb = 0; cb < ncb; ++cb) {
        int c = 0;
        for (; c < blk.nchunks; ++c) {
          const double* t = tiles.data() + ((size_t)cb * blk.nchunks + c) * tile;
          bool zero = true;
          for (size_t i = 0; i < tile && zero; ++i) zero = t[i] == 0.;
          if (!zero) break;
        }
        blk.cb_first[cb] = c;
        any_skip = any_skip || c > 0;
      }
      if (any_skip) DLB_TRY(upload(plan, blk.cb_first.data(), blk.cb_first.size(), &blk.d_cb_first));
      else blk.cb_first.clear();
    }
  }
  // epilogue constants per column block: [ncb][BN] = r0 / t0, zero padded; auxiliary columns carry r0 . g_i
  for (int pass = 0; pass < 2; ++pass) {
    const double* vec = pass == 0 ? r0_white : t0_plain;
    if (!vec) continue;
    const int ncols = pass == 0 ? blk.ncols_w : blk.ncols, ncb = pass == 0 ? blk.ncb : blk.ncb_plain;
    std::vector<double> epi((size_t)ncb * BN, 0.);
    for (int col = 0; col < ncols; ++col) epi[col] = vec[col];
    if (pass == 0)
      for (int i = 0; i < nconst; ++i) {
        double acc = 0.;
        for (int j = 0; j < ncols; ++j) acc += vec[j] * gconst_white[(size_t)i * ncols + j];
        epi[ncols + i] = acc;
      }
    DLB_TRY(upload(plan, epi.data(), epi.size(), pass == 0 ? &blk.d_epi_white : &blk.d_epi_plain));
  }
  return DLB_OK;
}

static int build_block(dlb_plan* plan, const dlb_plan_desc* desc, const dlb_block_desc& bd, const Variant& var, Block& blk) {
  std::vector<TermSpec> terms;
  blk.gen = bd.terms[0].gen;
  for (int t = 0; t < bd.nterms; ++t) {
    const dlb_term_desc& td = bd.terms[t];
    terms.push_back(TermSpec{td.slot, td.coef, plan->gens[td.gen].foff + td.foff, td.foff, td.K, bd.matrices_white + td.moff,
                             bd.matrices_plain ? bd.matrices_plain + td.moff_plain : nullptr});
  }
  DLB_TRY(build_block_core(plan, var, plan->resident, bd.ncols, bd.nslots, terms, bd.r0_white, bd.t0_plain, bd.nconst, bd.gconst_white, blk, bd.ncols_white));
  // (features computed by their own kernels into F: everything except Taylor monomials evaluated inside the contraction kernel)
  if (plan->mlp_slices > 0 && !(plan->resident && plan->gens[blk.gen].aux[6]) && !bd.transform_data && !getenv("DLB_NO_OZAKI_CONTRACT")) {
    // eligible: every term contracts the same feature columns, long enough for the GEMM to beat the D round trip through HBM
    bool same = bd.nterms <= OZ_GRAM_MAXT && terms[0].K >= 128 && bd.nconst <= 8;
    for (const TermSpec& td : terms) same = same && td.fcol == terms[0].fcol && td.K == terms[0].K && td.white;
    if (same) {
      OzContract& oz = blk.oz;
      oz.on = true; oz.K = terms[0].K; oz.fcol = terms[0].fcol; oz.nterms = bd.nterms; oz.ncols = blk.ncols_w; oz.nconst = bd.nconst;
      oz.ncolx = (oz.ncols + oz.nconst + 1) & ~1; oz.R = bd.nslots;   // even: the Gram kernel reads column pairs
      const int ncolsD = oz.nterms * oz.ncolx;
      oz.ldD = (ncolsD + 63) / 64 * 64;
      std::vector<double> B((size_t)oz.K * ncolsD, 0.);
      for (int t = 0; t < oz.nterms; ++t) {
        oz.tslot[t] = terms[t].slot; oz.tcoef[t] = terms[t].coef;
        const double* mat = terms[t].white;
        for (int k = 0; k < oz.K; ++k) {
          double* dst = B.data() + (size_t)k * ncolsD + (size_t)t * oz.ncolx;
          for (int c = 0; c < oz.ncols; ++c) dst[c] = mat[(size_t)k * oz.ncols + c];
          for (int i = 0; i < oz.nconst; ++i) {   // auxiliary columns (M_t g_i), as in build_block_core
            const double* g = bd.gconst_white + (size_t)i * oz.ncols;
            double acc = 0.;
            for (int c = 0; c < oz.ncols; ++c) acc += mat[(size_t)k * oz.ncols + c] * g[c];
            dst[oz.ncols + i] = acc;
          }
        }
      }
      std::vector<int8_t> sl; std::vector<int> eb;
      ozaki_slice_weights(B.data(), oz.K, ncolsD, plan->mlp_slices, sl, eb);
      std::vector<double> sb(eb.size());
      for (size_t i = 0; i < eb.size(); ++i) sb[i] = std::ldexp(1., eb[i]);
      DLB_TRY(upload(plan, sl.data(), sl.size(), &oz.d_Bs));
      DLB_TRY(upload(plan, sb.data(), sb.size(), &oz.d_Sb));
      std::vector<double> r0x(oz.ncolx + 2, 0.);
      for (int c = 0; c < oz.ncols; ++c) r0x[c] = bd.r0_white[c];
      for (int i = 0; i < oz.nconst; ++i) {
        double acc = 0.;
        for (int c = 0; c < oz.ncols; ++c) acc += bd.r0_white[c] * bd.gconst_white[(size_t)i * oz.ncols + c];
        r0x[oz.ncols + i] = acc;
      }
      DLB_TRY(upload(plan, r0x.data(), r0x.size(), &oz.d_r0));
    }
  }
  // entry map of the Gram partials (order fixed by contract_kernel's epilogue)
  std::vector<EntryMap> map(blk.npart);
  const int R = bd.nslots;
  auto solved_of_slot = [&](int a) { return bd.slot_solved[a - 1]; };
  for (int a = 0; a < R; ++a)
    for (int b = a; b < R; ++b) {
      EntryMap m;
      if (a == 0 && b == 0) m = EntryMap{0, 0, 0};
      else if (a == 0) m = EntryMap{1, solved_of_slot(b), 0};
      else m = EntryMap{2, solved_of_slot(a), solved_of_slot(b)};
      map[a * R - (a * (a - 1)) / 2 + (b - a)] = m;
    }
  const int nS = R * (R + 1) / 2;
  for (int i = 0; i < bd.nconst; ++i)
    for (int a = 0; a < R; ++a) map[nS + i * R + a] = a == 0 ? EntryMap{1, bd.const_solved[i], 0} : EntryMap{2, solved_of_slot(a), bd.const_solved[i]};
  if (bd.entry_kind) {   // explicit meaning of every partial (split dense plans)
    DLB_REQUIRE(bd.entry_i && bd.entry_j, "block: entry_kind needs entry_i / entry_j");
    for (int e = 0; e < blk.npart; ++e) {
      DLB_REQUIRE(bd.entry_kind[e] >= 0 && bd.entry_kind[e] <= 3, "block: entry kind must be 0 .. 3");
      map[e] = EntryMap{bd.entry_kind[e], bd.entry_kind[e] == 0 || bd.entry_kind[e] == 3 ? 0 : bd.entry_i[e], bd.entry_kind[e] == 2 ? bd.entry_j[e] : 0};
    }
  }
  for (const EntryMap& m : map)
    DLB_REQUIRE(m.kind == 0 || m.kind == 3 || (m.i >= 0 && m.i < desc->ns && (m.kind == 1 || (m.j >= 0 && m.j < desc->ns))), "solved-parameter index out of range");
  DLB_TRY(upload(plan, map.data(), map.size(), &blk.d_map));
  blk.h_map = map;
  return DLB_OK;
}

// Device block tables + launch configuration (stage count from the shared-memory budget) of one persistent launch.
static int build_launch(dlb_plan* plan, const std::vector<Block*>& blocks, bool resident, int ast, int ncoef, Launch& L) {
  L.var = blocks[0]->var; L.nblocks = (int)blocks.size(); L.resident = resident ? 1 : 0; L.ast = resident ? ast : 0;
  L.npart_max = 1;
  for (const Block* b : blocks) L.npart_max = std::max(L.npart_max, b->npart);
  L.nstages = 0;
  for (int s = 8; s >= 2 && !L.nstages; --s)
    if (contract_smem_bytes(L.var, s, resident, L.ast, plan->dpad, ncoef, L.npart_max, plan->ngen) <= contract_smem_limit(L.var.CPS)) { L.nstages = s; break; }
  DLB_REQUIRE(L.nstages >= 2, "contraction launch needs too much shared memory (too many parameters, coefficients or solved templates)");
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<BlockDev> tab;
    bool ok = true;
    for (const Block* b : blocks) {
      BlockDev bd;
      std::memset(&bd, 0, sizeof(bd));
      bd.chunks = b->d_chunks; bd.tiles = pass == 0 ? b->d_tiles_white : b->d_tiles_plain; bd.epi = pass == 0 ? b->d_epi_white : b->d_epi_plain;
      if (!bd.tiles || !bd.epi) ok = false;
      bd.paths = (resident && b->gen >= 0) ? plan->d_paths[b->gen] : nullptr; bd.poff = b->poff;
      if (resident && b->gen >= 0 && !plan->d_mono2.empty() && plan->d_mono2[b->gen]) {
        bd.mono2 = plan->d_mono2[b->gen]; bd.n1 = plan->mono_n1[b->gen]; bd.n2 = plan->mono_n2[b->gen];
      }
      bd.nchunks = b->nchunks; bd.ncb = pass == 0 ? b->ncb : b->ncb_plain; bd.nslots = b->nslots; bd.nconst = b->nconst; bd.npart = b->npart;
      bd.ncols = pass == 0 ? b->ncols_w : b->ncols;
      bd.ocol0 = b->coloff; bd.gen = b->gen; bd.fill = b->fill; bd.fbase = b->fbase; bd.acols = std::min((b->maxcol + 1) & ~1, L.ast);
      bd.tdata = pass == 1 ? b->d_tdata : nullptr;
      bd.oslot = 0;
      bd.tmask = pass == 0 ? b->d_tmask_white : b->d_tmask_plain;
      if (bd.tmask) (pass == 0 ? L.masked_white : L.masked_plain) = true;
      bd.cb_first = pass == 0 ? b->d_cb_first : nullptr;
      bd.qtile = b->d_qtile; bd.qdent = b->d_qdent; bd.qconst = b->d_qconst; bd.qw_off = b->d_qw_off; bd.qw_coef = b->d_qw_coef; bd.qw_val = b->d_qw_val;
      tab.push_back(bd);
    }
    if (!ok) continue;
    DLB_TRY(upload(plan, tab.data(), tab.size(), pass == 0 ? &L.d_white : &L.d_plain));
    if (pass == 1 && blocks[0]->d_tdata) {
      std::vector<BlockDev> tab_t(tab);
      for (size_t ib = 0; ib < blocks.size(); ++ib) {   // stage 1 of a two-stage plan: [u | u' S_1 | ...] of the block at column tcol of w_T
        tab_t[ib].ocol0 = blocks[ib]->tcol;
        tab_t[ib].oslot = blocks[ib]->nslots > 1 ? ((blocks[ib]->ncols + 3) & ~3) : 0;
      }
      DLB_TRY(upload(plan, tab_t.data(), tab_t.size(), &L.d_plain_t));
    }
    // per-tile TMA schedule: for every block, column block: the k-chunk tiles in order, then the epilogue constants
    const int BN = variant_bn(L.var), KC = L.var.KC;
    std::vector<SchedEntry> sched;
    for (size_t ib = 0; ib < blocks.size(); ++ib) {
      const Block* b = blocks[ib];
      DLB_REQUIRE(b->nchunks < 65536 && blocks.size() < 32768, "too many chunks / blocks in one launch");
      for (int cb = 0; cb < tab[ib].ncb; ++cb) {
        for (int c = (pass == 0 && !b->cb_first.empty()) ? b->cb_first[cb] : 0; c < b->nchunks; ++c) {
          const std::vector<double>& ex = pass == 0 ? b->exec_white : b->exec_plain;
          (pass == 0 ? L.flops_exec : L.flops_exec_plain) += 2. * (ex.empty() ? (double)b->h_chunks[c].kc * BN : ex[(size_t)cb * b->nchunks + c]);
          sched.push_back(SchedEntry{tab[ib].tiles + ((size_t)cb * b->nchunks + c) * ((size_t)BN * KC), (unsigned)(BN * KC * 8), (int)((ib << 16) | (unsigned)c)});
        }
        if (!b->d_qtile) sched.push_back(SchedEntry{tab[ib].epi + (size_t)cb * BN, (unsigned)(BN * 8), -1});   // (quadratic-form blocks have no epilogue constants)
      }
    }
    (pass == 0 ? L.nsched : L.nsched_plain) = (int)sched.size();
    DLB_TRY(upload(plan, sched.data(), sched.size(), pass == 0 ? &L.d_sched_white : &L.d_sched_plain));
  }
  return DLB_OK;
}

// MLP generator (reference emulators/conversion.py:18-33): x-operation (affine min-max scale) -> hidden layers
// v <- act(v W_l + b_l), each one an output-mode launch of the DMMA contraction kernel writing the next feature segment.
// The last linear layer and the inverse y-operations are affine in the last hidden activations and are folded into
// the term matrices by the host (desilike_b200/mlp.py).
//   idata = [nhidden, act, d, H_1 .. H_nhidden];  ddata = [a (d), b (d), then per hidden layer W [Hin, Hout] row-major, bias [Hout]]
static int build_mlp(dlb_plan* plan, const dlb_generator_desc& gd, DevGen& g, MlpLayers& mlp) {
  DLB_REQUIRE(gd.ni >= 3 && gd.idata, "mlp generator: idata = [nhidden, act, d, H_1..]");
  const int nh = gd.idata[0], act = gd.idata[1], d = gd.idata[2];
  DLB_REQUIRE(nh >= 0 && gd.ni == 3 + nh && d == gd.nparams && act >= 0 && act <= 3, "mlp generator: inconsistent idata");
  std::vector<int> width(nh + 1);
  width[0] = d;
  long long need = 2LL * d;
  for (int l = 0; l < nh; ++l) { width[l + 1] = gd.idata[3 + l]; DLB_REQUIRE(width[l + 1] > 0, "mlp: layer width"); need += (long long)width[l] * width[l + 1] + width[l + 1]; }
  DLB_REQUIRE(gd.nd == need, "mlp generator: ddata size mismatch");
  DLB_REQUIRE(gd.nfeat == 2 + width[nh], "mlp generator: nfeat must be 2 + last hidden width");
  auto pad4 = [](int n) { return (n + 3) & ~3; };
  // segment layout: [1, 0, h_last | xs | h_1 .. h_{nh-1}]  (visible features first)
  std::vector<int> off(nh + 1);
  int cur = pad4(2 + width[nh]);
  if (nh == 0) off[0] = 2;
  else {
    off[nh] = 2;
    off[0] = cur; cur += pad4(width[0]);
    for (int l = 1; l < nh; ++l) { off[l] = cur; cur += pad4(width[l]); }
  }
  g.fpad = cur;
  g.aux[0] = off[0];
  mlp.act = act;
  const double* w = gd.ddata + 2 * d;
  if (plan->mlp_slices > 0) {   // opt-in tensor-core path: every hidden layer is its own (slice +) GEMM launch pair
    for (int l = 0; l < nh; ++l) {
      OzLayer oz;
      oz.K = width[l]; oz.N = width[l + 1]; oz.in_off = off[l]; oz.out_off = off[l + 1]; oz.ozaki = width[l] >= 64;
      const double* bias = w + (size_t)width[l] * width[l + 1];
      DLB_TRY(upload(plan, w, (size_t)oz.K * oz.N, &oz.d_W));
      DLB_TRY(upload(plan, bias, (size_t)oz.N, &oz.d_bias));
      if (oz.ozaki) {
        std::vector<int8_t> sl; std::vector<int> eb;
        ozaki_slice_weights(w, oz.K, oz.N, plan->mlp_slices, sl, eb);
        std::vector<double> sb(eb.size());
        for (size_t i = 0; i < eb.size(); ++i) sb[i] = std::ldexp(1., eb[i]);
        DLB_TRY(upload(plan, sl.data(), sl.size(), &oz.d_Bs));
        DLB_TRY(upload(plan, sb.data(), sb.size(), &oz.d_Sb));
        plan->oz_maxk = std::max(plan->oz_maxk, oz.K);
      }
      mlp.oz.push_back(oz);
      w = bias + width[l + 1];
    }
    return DLB_OK;
  }
  mlp.layers.resize(nh);
  for (int l = 0; l < nh; ++l) {
    Block& blk = mlp.layers[l];
    std::vector<TermSpec> terms{TermSpec{0, -1, g.foff + off[l], 0, width[l], nullptr, w}};
    const double* bias = w + (size_t)width[l] * width[l + 1];
    const int ast = round_ast(chunk_extent(width[l]));
    const bool resident = contract_smem_bytes(Variant{1, 2, 7, 8, 1, 36}, 2, true, ast, plan->dpad, 0, 1, plan->ngen) <= contract_smem_limit(1);
    const Variant var = resident ? Variant{1, 2, 7, 8, 1, 36} : pick_variant(plan, 1, resident, 0, 0, 1, width[l]);   // 128-row tiles amortise the per-(tile, layer) latency
    DLB_TRY(build_block_core(plan, var, resident, width[l + 1], 1, terms, nullptr, bias, 0, nullptr, blk));
    blk.coloff = g.foff + off[l + 1];
    blk.fill = 1; blk.fbase = g.foff + off[l]; blk.gen = -1;
    blk.maxcol = resident ? blk.maxcol : -1;   // -1: streaming launch
    w = bias + width[l + 1];
  }
  return DLB_OK;
}

static int ensure_workspace(dlb_plan* plan, long long npad) {
  if (npad <= plan->cap) return DLB_OK;
  for (void* ptr : {(void*)plan->w_F, (void*)plan->w_partial, (void*)plan->w_params, (void*)plan->w_out, (void*)plan->w_status, (void*)plan->w_T, (void*)plan->w_pext})
    if (ptr) cudaFree(ptr);
  plan->w_F = plan->w_partial = plan->w_params = plan->w_out = plan->w_T = plan->w_pext = nullptr; plan->w_status = nullptr;
  plan->cap = 0;
  if (plan->two_stage) {
    DLB_CUDA(cudaMalloc((void**)&plan->w_T, ((size_t)npad * plan->tstride + 64) * sizeof(double)));
    DLB_CUDA(cudaMemset(plan->w_T, 0, ((size_t)npad * plan->tstride + 64) * sizeof(double)));
  }
  DLB_CUDA(cudaMalloc((void**)&plan->w_F, ((size_t)npad * plan->fstride + 64) * sizeof(double)));
  DLB_CUDA(cudaMemset(plan->w_F, 0, ((size_t)npad * plan->fstride + 64) * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_partial, (size_t)std::max(plan->ptotal, 1LL) * npad * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_params, (size_t)npad * std::max(plan->d_in, 1) * sizeof(double)));
  if (plan->nder) DLB_CUDA(cudaMalloc((void**)&plan->w_pext, (size_t)npad * plan->d * sizeof(double)));
  if (plan->w_oz_slices) { cudaFree(plan->w_oz_slices); plan->w_oz_slices = nullptr; }
  if (plan->w_oz_exps) { cudaFree(plan->w_oz_exps); plan->w_oz_exps = nullptr; }
  if (plan->w_oz_D) { cudaFree(plan->w_oz_D); plan->w_oz_D = nullptr; }
  if (plan->mlp_slices > 0 && plan->oz_maxk > 0) {
    DLB_CUDA(cudaMalloc((void**)&plan->w_oz_slices, ozaki_slice_bytes(ozaki_rows_pad(npad), plan->oz_maxk, plan->mlp_slices)));
    DLB_CUDA(cudaMalloc((void**)&plan->w_oz_exps, (size_t)ozaki_rows_pad(npad) * sizeof(int)));
    if (plan->oz_contract) DLB_CUDA(cudaMalloc((void**)&plan->w_oz_D, (size_t)ozaki_rows_pad(npad) * plan->oz_ldD * sizeof(double)));
  }
  DLB_CUDA(cudaMalloc((void**)&plan->w_out, (size_t)npad * (3 + std::max(plan->ns, 0)) * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_status, (size_t)npad * sizeof(int)));
  plan->cap = npad;
  return DLB_OK;
}

static void fill_params(const dlb_plan* plan, const Launch& L, ContractParams& cp, const double* d_params, long long N, long long npad) {
  std::memset(&cp, 0, sizeof(cp));
  cp.nblocks = L.nblocks; cp.gens = plan->d_gens; cp.ngen = plan->ngen; cp.coefs = plan->d_coefs; cp.ncoef = plan->ncoef;
  cp.params = d_params; cp.d = plan->d; cp.N = N; cp.F = plan->w_F; cp.fstride = plan->fstride;
  cp.partial = plan->w_partial; cp.npad = npad;
  cp.nstages = L.nstages; cp.resident = L.resident; cp.ast = L.ast; cp.dpad = plan->dpad; cp.npart_max = L.npart_max;
}

// Feature generators that run as their own kernels (everything except Taylor monomials in resident mode).
static int run_front(dlb_plan* plan, const double* d_params, long long N, long long npad, cudaStream_t stream, bool force_taylor = false) {
  for (size_t ig = 0; ig < plan->gens.size(); ++ig) {
    const DevGen& g = plan->gens[ig];
    if (g.kind == DLB_GEN_MLP) {
      if (!plan->mlp_fused_smem) DLB_TRY(launch_mlp_input(g, d_params, plan->d, N, npad, plan->w_F, plan->fstride, stream));
      if (plan->mlp_slices > 0) {   // hidden layers: tcgen05 int8 GEMMs (fan-in >= 64) / plain FP64 (the narrow first layer)
        const MlpLayers& mlp = plan->mlps[ig];
        double* seg = plan->w_F + g.foff;
        for (const OzLayer& oz : mlp.oz) {
          if (oz.ozaki)
            DLB_TRY(launch_ozaki_layer(seg + oz.in_off, N, oz.K, plan->fstride, oz.d_Bs, oz.d_Sb, oz.d_bias, oz.N, plan->mlp_slices, mlp.act, seg + oz.out_off,
                                       plan->fstride, plan->w_oz_slices, plan->w_oz_exps, plan->nsm, stream));
          else
            DLB_TRY(launch_mlp_small_layer(seg + oz.in_off, N, oz.K, plan->fstride, oz.d_W, oz.d_bias, oz.N, mlp.act, seg + oz.out_off, plan->fstride, stream));
        }
      }
      continue;
    }
    if (g.kind == DLB_GEN_TAYLOR && g.aux[6] && plan->resident && !force_taylor) continue;   // evaluated inside the contraction kernel
    DLB_TRY(launch_features(g, d_params, plan->d, N, npad, plan->w_F, plan->fstride, stream));
  }
  if (plan->mlp_fused_smem)
    return launch_mlp_fused(plan->d_devgens, plan->d_mlp_index, plan->nmlp, plan->mlp_fused_smem, plan->mlp_groups, d_params, plan->d, N, npad, plan->w_F, plan->fstride,
                            plan->nsm, stream);
  for (const auto& la : plan->mlp_launches) {   // hidden layers, depth by depth, all MLP emulators of the likelihood in one launch
    const Launch& L = la.first;
    ContractParams cp;
    fill_params(plan, L, cp, d_params, N, npad);
    cp.blocks = L.d_plain; cp.sched = L.d_sched_plain; cp.nsched = L.nsched_plain; cp.ncoef = 0; cp.masked = L.masked_plain;
    cp.out = plan->w_F; cp.ostride = plan->fstride; cp.out_rows = npad; cp.act = la.second; cp.mode = 1;
    cp.U = npad / variant_unit(L.var);
    DLB_TRY(launch_contract(L.var, cp, plan->nsm, stream));
  }
  return DLB_OK;
}

// Two-stage plans: theory vectors (un-whitened matrices) with the observable transform applied in the epilogue -> w_T, then the
// whitening contraction r = u . L^T - L data with its chi2 reduction (one Gram partial per block).
static int run_two_stage(dlb_plan* plan, const double* d_params, long long N, long long npad, cudaStream_t stream) {
  {
    const Launch& L = plan->launch;
    ContractParams cp;
    fill_params(plan, L, cp, d_params, N, npad);
    cp.blocks = L.d_plain_t; cp.sched = L.d_sched_plain; cp.nsched = L.nsched_plain; cp.masked = L.masked_plain;
    cp.out = plan->w_T; cp.ostride = plan->tstride; cp.out_rows = npad; cp.act = 0; cp.mode = 1;
    cp.U = (N + variant_unit(L.var) - 1) / variant_unit(L.var);
    DLB_TRY(launch_contract(L.var, cp, plan->nsm, stream));
  }
  const Launch& L = plan->wlaunch;
  ContractParams cp;
  fill_params(plan, L, cp, d_params, N, npad);
  cp.blocks = L.d_white; cp.sched = L.d_sched_white; cp.nsched = L.nsched; cp.ncoef = 0; cp.masked = L.masked_white;
  cp.F = plan->w_T; cp.fstride = plan->tstride; cp.mode = 0;
  cp.U = (N + variant_unit(L.var) - 1) / variant_unit(L.var);
  return launch_contract(L.var, cp, plan->nsm, stream);
}

static int run_contract(dlb_plan* plan, const double* d_params, long long N, long long npad, int mode, double* d_theory, cudaStream_t stream) {
  if (mode == 0 && plan->two_stage) return run_two_stage(plan, d_params, N, npad, stream);
  if (mode == 0 && plan->oz_contract) {   // tcgen05 int8 path: per block one GEMM over all terms + the Gram reduction
    for (const Block& b : plan->blocks) {
      const OzContract& oz = b.oz;
      DLB_TRY(launch_ozaki_slice(plan->w_F + oz.fcol, N, oz.K, plan->fstride, plan->mlp_slices, plan->w_oz_slices, plan->w_oz_exps, stream));
      DLB_TRY(launch_ozaki_gemm(N, oz.K, plan->w_oz_slices, plan->w_oz_exps, oz.d_Bs, oz.d_Sb, nullptr, oz.ldD, plan->mlp_slices, 0, plan->w_oz_D, oz.ldD,
                                plan->nsm, stream));
      OzakiGramParams gp;
      std::memset(&gp, 0, sizeof(gp));
      gp.D = plan->w_oz_D; gp.ldD = oz.ldD; gp.N = N; gp.npad = npad; gp.params = d_params; gp.d = plan->d; gp.coefs = plan->d_coefs;
      gp.nterms = oz.nterms; gp.ncols = oz.ncols; gp.nconst = oz.nconst; gp.ncolx = oz.ncolx; gp.R = oz.R;
      for (int t = 0; t < oz.nterms; ++t) { gp.tslot[t] = oz.tslot[t]; gp.tcoef[t] = oz.tcoef[t]; }
      gp.r0 = oz.d_r0; gp.partial = plan->w_partial; gp.poff = b.poff;
      DLB_TRY(launch_ozaki_gram(gp, stream));
    }
    return DLB_OK;
  }
  const bool qf = mode == 0 && plan->qform;
  const Launch& L = qf ? plan->qlaunch : plan->launch;
  ContractParams cp;
  fill_params(plan, L, cp, d_params, N, npad);
  cp.qform = qf ? 1 : 0;
  static const int qf_offset = getenv("DLB_QF_OFFSET_NS") ? atoi(getenv("DLB_QF_OFFSET_NS")) : 3000;
  cp.qf_offset_ns = qf_offset;
  cp.blocks = mode == 0 ? L.d_white : L.d_plain;
  cp.sched = mode == 0 ? L.d_sched_white : L.d_sched_plain; cp.nsched = mode == 0 ? L.nsched : L.nsched_plain;
  cp.masked = mode == 0 ? L.masked_white : L.masked_plain;
  if (!cp.blocks) { set_error("plan was built without un-whitened matrices: theory_batch unavailable"); return DLB_ERR_ARG; }
  cp.out = d_theory; cp.ostride = plan->ntheory; cp.out_rows = N; cp.act = 0; cp.mode = mode;
  cp.U = (N + variant_unit(L.var) - 1) / variant_unit(L.var);
#ifdef DLB_PROFILE
  static long long* d_prof = nullptr;
  if (!d_prof) { cudaMalloc((void**)&d_prof, 8 * sizeof(long long)); }
  cudaMemsetAsync(d_prof, 0, 8 * sizeof(long long), stream);
  cp.prof = d_prof;
  int rc = launch_contract(L.var, cp, plan->nsm, stream);
  long long h[8];
  cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
  long long tot = 0; for (int i = 0; i < 8; ++i) tot += h[i];
  if (getenv("DLB_PROFILE_PRINT") && tot > 0)
    fprintf(stderr, "[prof] wait-epi-chunk %.1f%% wait-first-of-block %.1f%% agen %.1f%% mma %.1f%% wait-other %.1f%% release+barriers+setup+flush %.1f%% epilogue %.1f%% wait-first-of-cb %.1f%%\n",
            100. * h[0] / tot, 100. * h[1] / tot, 100. * h[2] / tot, 100. * h[3] / tot, 100. * h[4] / tot, 100. * h[5] / tot, 100. * h[6] / tot, 100. * h[7] / tot);
  return rc;
#else
  return launch_contract(L.var, cp, plan->nsm, stream);
#endif
}

}  // namespace dlb

using namespace dlb;

extern "C" {

int dlb_abi_version(void) { return DLB_ABI_VERSION; }
const char* dlb_last_error(void) { return g_error.c_str(); }
int64_t dlb_launch_count(void) { return g_launches.load(); }

int dlb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int dlb_plan_ntheory(const dlb_plan* plan) { return plan ? plan->ntheory : -1; }
int dlb_plan_nsolved(const dlb_plan* plan) { return plan ? plan->ns : -1; }
int dlb_plan_flags(const dlb_plan* plan) {
  if (!plan) return -1;
  bool ozl = false;
  for (const MlpLayers& m : plan->mlps) for (const OzLayer& l : m.oz) ozl = ozl || l.ozaki;
  return (plan->resident ? 1 : 0) | (plan->two_stage ? 2 : 0) | (plan->mlp_fused_smem ? 4 : 0) | (ozl ? 8 : 0) | (plan->oz_contract ? 16 : 0) | (plan->qform ? 32 : 0);
}
double dlb_plan_executed_flops(const dlb_plan* plan) {
  if (!plan) return -1.;
  if (plan->qform) return plan->qlaunch.flops_exec;
  return plan->two_stage ? plan->launch.flops_exec_plain + plan->wlaunch.flops_exec : plan->launch.flops_exec;
}

void dlb_plan_destroy(dlb_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->device);
  for (void* ptr : plan->owned) cudaFree(ptr);
  for (void* ptr : {(void*)plan->w_F, (void*)plan->w_partial, (void*)plan->w_params, (void*)plan->w_out, (void*)plan->w_status, (void*)plan->w_theory, (void*)plan->w_T, (void*)plan->w_pext})
    if (ptr) cudaFree(ptr);
  for (auto& e : plan->ev) if (e) cudaEventDestroy(e);
  if (plan->w_hess) cudaFree(plan->w_hess);
  if (plan->w_oz_slices) cudaFree(plan->w_oz_slices);
  if (plan->w_oz_exps) cudaFree(plan->w_oz_exps);
  if (plan->w_oz_D) cudaFree(plan->w_oz_D);
  for (auto& e : plan->ev_pipe) if (e) cudaEventDestroy(e);
  if (plan->s_in) cudaStreamDestroy(plan->s_in);
  if (plan->s_out) cudaStreamDestroy(plan->s_out);
  if (plan->w_prop) cudaFree(plan->w_prop);
  if (plan->w_logpy) cudaFree(plan->w_logpy);
  delete plan;
}

int dlb_plan_create(const dlb_plan_desc* desc, dlb_plan** out) {
  DLB_REQUIRE(desc && out, "NULL descriptor / output");
  *out = nullptr;
  DLB_REQUIRE(desc->abi_version == DLB_ABI_VERSION, "ABI version mismatch");
  DLB_REQUIRE(desc->d > 0 && desc->ngen > 0 && desc->nblocks > 0 && desc->gens && desc->blocks, "d, ngen, nblocks must be positive");
  DLB_REQUIRE(desc->ns >= 0 && desc->ns <= 64 && desc->nmarg >= 0 && desc->nmarg <= desc->ns, "0 <= nmarg <= ns <= 64 required");
  int ndev = 0;
  DLB_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev == 0) { set_error("no CUDA device: desilike_b200 has no CPU fallback"); return DLB_ERR_CUDA; }
  dlb_plan* plan = new dlb_plan();
  struct Guard { dlb_plan* p; ~Guard() { if (p) dlb_plan_destroy(p); } } guard{plan};
  DLB_CUDA(cudaGetDevice(&plan->device));
  DLB_REQUIRE(desc->nderived >= 0 && desc->nderived <= 64 && (desc->nderived == 0 || (desc->derived_off && desc->derived_code)), "derived parameters: 0 <= nderived <= 64 and tables required");
  DLB_REQUIRE(desc->mlp_slices == 0 || (desc->mlp_slices >= 2 && desc->mlp_slices <= 8), "mlp_slices must be 0 (FP64 DMMA) or 2 .. 8");
  plan->mlp_slices = desc->mlp_slices;
  plan->d_in = desc->d; plan->nder = desc->nderived;
  plan->d = desc->d + desc->nderived; plan->ngen = desc->ngen; plan->ncoef = desc->ncoef; plan->ns = desc->ns; plan->nmarg = desc->nmarg;
  plan->dpad = (plan->d + 1) | 1;   // + the constant-one column; odd stride keeps per-row accesses conflict free
  const int dext = plan->d;         // every index table addresses the extended parameter row
  if (plan->nder) {
    const int nops = desc->derived_off[plan->nder];
    int nconst = 0;
    for (int i = 0; i < plan->nder; ++i) {
      DLB_REQUIRE(desc->derived_off[i] <= desc->derived_off[i + 1], "derived parameters: offsets must be non-decreasing");
      int depth = 0;
      for (int o = desc->derived_off[i]; o < desc->derived_off[i + 1]; ++o) {
        const int op = desc->derived_code[2 * o], arg = desc->derived_code[2 * o + 1];
        DLB_REQUIRE(op >= 0 && op <= 15, "derived parameters: unknown opcode");
        if (op == 0) DLB_REQUIRE(arg >= 0 && arg < desc->d + i, "derived parameters: a program may read the inputs and EARLIER derived parameters only");
        if (op == 1) { DLB_REQUIRE(arg >= 0 && desc->derived_const, "derived parameters: constant index"); nconst = std::max(nconst, arg + 1); }
        depth += op <= 1 ? 1 : (op <= 6 ? -1 : 0);
        DLB_REQUIRE(depth >= 1 && depth <= 16, "derived parameters: malformed program (stack depth)");
      }
      DLB_REQUIRE(depth == 1, "derived parameters: a program must leave exactly one value");
    }
    DLB_TRY(upload(plan, desc->derived_off, (size_t)plan->nder + 1, &plan->d_der_off));
    DLB_TRY(upload(plan, desc->derived_code, (size_t)2 * nops, &plan->d_der_code));
    DLB_TRY(upload(plan, desc->derived_const, (size_t)nconst, &plan->d_der_const));
  }
  DLB_CUDA(cudaDeviceGetAttribute(&plan->nsm, cudaDevAttrMultiProcessorCount, plan->device));
  // generators / feature layout
  int foff = 0;
  std::vector<GenDev> gendev(desc->ngen);
  for (int ig = 0; ig < desc->ngen; ++ig) {
    const dlb_generator_desc& gd = desc->gens[ig];
    DLB_REQUIRE(gd.nparams >= 0 && gd.nparams <= MAXP, "generator: too many parameters");
    DLB_REQUIRE(gd.nfeat > 0, "generator: nfeat must be positive");
    DevGen g;
    std::memset(&g, 0, sizeof(g));
    std::memset(&gendev[ig], 0, sizeof(GenDev));
    g.kind = gd.kind; g.nparams = gd.nparams; g.nfeat = gd.nfeat; g.fpad = (gd.nfeat + 3) & ~3; g.foff = foff;
    for (int j = 0; j < gd.nparams; ++j) {
      DLB_REQUIRE(gd.param_idx[j] >= -1 && gd.param_idx[j] < dext, "generator: parameter index out of range");
      g.pidx[j] = gd.param_idx[j];
      g.pdef[j] = gd.param_default ? gd.param_default[j] : 0.;
    }
    int* di = nullptr; double* dd = nullptr;
    DLB_TRY(upload(plan, gd.idata, (size_t)gd.ni, &di));
    DLB_TRY(upload(plan, gd.ddata, (size_t)gd.nd, &dd));
    g.idata = di; g.ni = gd.ni; g.ddata = dd; g.nd = gd.nd;
    if (gd.kind == DLB_GEN_TAYLOR) {
      DLB_REQUIRE(gd.nfeat > 2 && gd.ni == (gd.nfeat - 2) * gd.nparams && gd.nd == gd.nparams, "taylor generator: nfeat = 2 + T, idata = powers[T, d], ddata = center[d]");
      GenDev& gv = gendev[ig];
      gv.nparams = gd.nparams;
      for (int j = 0; j < gd.nparams; ++j) gv.cval[j] = g.pidx[j] >= 0 ? gd.ddata[j] : 1. - (g.pdef[j] - gd.ddata[j]);
      gv.cval[MAXP] = 0.;
      for (int t = 0; t < gd.nfeat - 2; ++t) {
        int deg = 0;
        for (int j = 0; j < gd.nparams; ++j) { DLB_REQUIRE(gd.idata[t * gd.nparams + j] >= 0, "taylor generator: negative power"); deg += gd.idata[t * gd.nparams + j]; }
        gv.maxdeg = std::max(gv.maxdeg, deg);
      }
      g.aux[6] = gv.maxdeg <= 7 ? 1 : 0;   // in-kernel monomials (factor lists hold 7 entries; higher orders go through the feature kernel)
    }
    plan->mlps.emplace_back();
    if (gd.kind == DLB_GEN_MLP) DLB_TRY(build_mlp(plan, gd, g, plan->mlps.back()));
    if (gd.kind == DLB_GEN_KAISER || gd.kind == DLB_GEN_BAO) {
      DLB_REQUIRE(gd.ni >= 8, "theory generator: idata header missing");
      g.fpad = std::max(g.fpad, (gd.idata[7] + 3) & ~3);   // idata[7] = total segment length incl. scratch
      g.aux[1] = gd.idata[3];                              // nt (template knots), needed on the host for shared-memory sizing
      g.aux[2] = gd.idata[0]; g.aux[3] = gd.idata[1]; g.aux[4] = gd.idata[5];
      if (gd.kind == DLB_GEN_BAO) { DLB_REQUIRE(gd.ni >= 16, "bao generator: idata header"); g.aux[5] = gd.idata[14]; }
      if (gd.kind == DLB_GEN_KAISER) {
        DLB_REQUIRE(gd.ni >= 16 && gd.idata[10] >= 1 && gd.idata[10] <= 8 && gd.idata[2] >= 1 && gd.idata[2] <= 16 && gd.idata[15] >= 4 && gd.idata[15] <= gd.idata[3],
                    "kaiser generator: idata header (1 <= nseg <= 8, 1 <= nmu <= 16, 4 <= nw_max <= nt)");
        g.aux[5] = gd.idata[2] | (gd.idata[10] << 8) | (gd.idata[13] << 16);   // nmu, nseg, interp
        g.aux[7] = gd.idata[15];                                               // largest knot window
      }
    }
    foff += g.fpad;
    plan->gens.push_back(g);
  }
  plan->fstride = foff;
  {   // fused MLP forward when every MLP emulator qualifies (hidden widths <= 64, weights fit in shared memory)
    std::vector<int> mlp_index;
    size_t smem = 0;
    bool ok = true;
    int groups = getenv("DLB_MLP_ONE_GROUP") ? 1 : 2;   // two out-of-phase warp groups per CTA when their activation buffers fit next to the weights
    for (;;) {
      mlp_index.clear(); smem = 0; ok = true;
      for (int ig = 0; ig < desc->ngen; ++ig)
        if (desc->gens[ig].kind == DLB_GEN_MLP) {
          mlp_index.push_back(ig);
          const size_t need = mlp_fused_smem(desc->gens[ig].idata, groups);
          if (!need || need > (size_t)224 * 1024) ok = false;
          smem = std::max(smem, need);
        }
      if (ok || groups == 1) break;
      groups = 1;
    }
    plan->mlp_groups = groups;
    if (ok && !mlp_index.empty() && !getenv("DLB_NO_FUSED_MLP") && plan->mlp_slices == 0) {
      plan->mlp_fused_smem = smem; plan->nmlp = (int)mlp_index.size();
      DLB_TRY(upload(plan, plan->gens.data(), plan->gens.size(), &plan->d_devgens));
      DLB_TRY(upload(plan, mlp_index.data(), mlp_index.size(), &plan->d_mlp_index));
    }
  }
  // coefficients
  std::vector<DevCoef> coefs(std::max(desc->ncoef, 1));
  for (int j = 0; j < desc->ncoef; ++j) {
    const dlb_coef_desc& cd = desc->coefs[j];
    DLB_REQUIRE(cd.nf >= 0 && cd.nf <= DLB_COEF_MAXF, "coefficient: too many factors");
    coefs[j].scale = cd.scale; coefs[j].nf = cd.nf;
    for (int f = 0; f < cd.nf; ++f) {
      DLB_REQUIRE(cd.idx[f] >= 0 && cd.idx[f] < dext, "coefficient: parameter index out of range");
      coefs[j].idx[f] = cd.idx[f]; coefs[j].pow[f] = cd.pow[f];
    }
  }
  DLB_TRY(upload(plan, coefs.data(), coefs.size(), &plan->d_coefs));
  // priors
  DLB_REQUIRE(desc->prior_kind && desc->prior_lo && desc->prior_hi && desc->prior_loc && desc->prior_scale, "prior tables required");
  DLB_TRY(upload(plan, desc->prior_kind, (size_t)dext, &plan->d_prior_kind));
  DLB_TRY(upload(plan, desc->prior_lo, (size_t)dext, &plan->d_prior_lo));
  DLB_TRY(upload(plan, desc->prior_hi, (size_t)dext, &plan->d_prior_hi));
  DLB_TRY(upload(plan, desc->prior_loc, (size_t)dext, &plan->d_prior_loc));
  DLB_TRY(upload(plan, desc->prior_scale, (size_t)dext, &plan->d_prior_scale));
  // blocks: validate, pick the kernel variant (plan-wide) and the A-operand mode
  int max_slots = 1, ast = 4, maxK = 0;
  bool resident = true;
  for (int ib = 0; ib < desc->nblocks; ++ib) {
    const dlb_block_desc& bd = desc->blocks[ib];
    DLB_REQUIRE(bd.ncols > 0 && bd.nslots >= 1 && bd.nslots <= 8, "block: ncols > 0 and 1 <= nslots <= 8 required");
    DLB_REQUIRE(bd.ncols_white == 0 || bd.ncols_white == bd.ncols || (bd.ncols_white > 0 && bd.nslots == 1 && bd.nconst == 0),
                "block: a rank-reduced whitened problem (ncols_white != ncols) requires nslots == 1 and nconst == 0");
    DLB_REQUIRE(bd.nterms > 0 && bd.terms && bd.matrices_white && bd.r0_white, "block: terms / matrices_white / r0_white required");
    max_slots = std::max(max_slots, bd.nslots);
    for (int t = 0; t < bd.nterms; ++t) {
      const dlb_term_desc& td = bd.terms[t];
      DLB_REQUIRE(td.slot >= 0 && td.slot < bd.nslots, "term slot out of range");
      DLB_REQUIRE(td.coef >= -1 && td.coef < desc->ncoef, "term coef out of range");
      DLB_REQUIRE(td.gen >= 0 && td.gen < desc->ngen, "term generator out of range");
      DLB_REQUIRE(td.K > 0 && td.foff >= 0 && (td.foff % 2) == 0 && td.foff + td.K <= plan->gens[td.gen].nfeat, "term feature range invalid (foff must be even)");
      if (td.gen != bd.terms[0].gen) resident = false;   // resident A holds ONE generator segment per block
      ast = std::max(ast, td.foff + chunk_extent(td.K));
      maxK = std::max(maxK, td.K);
    }
    // quadratic-form tables: the epilogue reads whole 8-column feature tiles -- keep them inside the row of the resident A tile (its padding is zero), so
    // that a non-finite feature of one point can never reach a neighbouring point through a read past the row end
    if (bd.q_ncols > 0 && bd.q_K > 0) ast = std::max(ast, bd.q_foff + ((bd.q_K + 7) & ~7));
  }
  ast = round_ast(ast);
  int npart_bound = 1;
  for (int ib = 0; ib < desc->nblocks; ++ib) npart_bound = std::max(npart_bound, desc->blocks[ib].nslots * (desc->blocks[ib].nslots + 1) / 2 + desc->blocks[ib].nslots * desc->blocks[ib].nconst);
  {   // resident A must fit the one-CTA-per-SM form of the slot class with at least 2 TMA stages
    const Variant v1 = max_slots <= 1 ? Variant{1, 2, 7, 4, 2, 28} : (max_slots <= 4 ? Variant{4, 1, 7, 6, 1, 36} : Variant{8, 1, 4, 6, 1, 36});
    if (resident && (contract_smem_bytes(v1, 2, true, ast, plan->dpad, desc->ncoef, npart_bound, desc->ngen) > contract_smem_limit(v1.CPS) || desc->ngen > 32)) {
      // single-slot launches: a 16-warp CTA with the whole SM's shared memory is the roomiest form
      if (max_slots > 1 || contract_smem_bytes(Variant{1, 2, 7, 8, 1, 36}, 2, true, ast, plan->dpad, desc->ncoef, npart_bound, desc->ngen) > contract_smem_limit(1) || desc->ngen > 32)
        resident = false;
    }
  }
  // Multiplied tile area (contraction rows x padded columns) with 112- and with 80-column blocks.  Rank-reduced blocks are lower
  // triangular: column block cb only needs the rows >= cb * BN (the leading zero tiles are skipped), so narrower blocks also skip more.
  long long area7 = 0, area5 = 0;
  for (int ib = 0; ib < desc->nblocks; ++ib) {
    const dlb_block_desc& bd = desc->blocks[ib];
    const int cols = (bd.ncols_white > 0 ? bd.ncols_white : bd.ncols) + bd.nconst;
    const bool reduced = bd.ncols_white > 0 && bd.ncols_white < bd.ncols;
    long long rows = 0;
    for (int t = 0; t < bd.nterms; ++t) rows += bd.terms[t].K;
    for (int cb = 0; cb * 112 < cols; ++cb) area7 += std::max(0LL, reduced ? rows - (long long)cb * 112 : rows) * 112;
    for (int cb = 0; cb * 80 < cols; ++cb) area5 += std::max(0LL, reduced ? rows - (long long)cb * 80 : rows) * 80;
  }
  const bool narrow = area5 * 100 < area7 * 93 || getenv("DLB_FORCE_NARROW");   // the 80-column variant is ~3-4 % less efficient per multiplied tile
  const Variant var = pick_variant(plan, max_slots, resident, resident ? ast : 0, desc->ncoef, npart_bound, maxK, narrow);
  plan->resident = resident;
  if (resident) {   // factor lists of the monomials, one table per generator, indexed by segment column
    plan->d_paths.resize(desc->ngen, nullptr);
    plan->d_mono2.resize(desc->ngen, nullptr); plan->mono_n1.resize(desc->ngen, 0); plan->mono_n2.resize(desc->ngen, 0);
    for (int ig = 0; ig < desc->ngen; ++ig) {
      const dlb_generator_desc& gd = desc->gens[ig];
      if (!plan->gens[ig].aux[6]) continue;
      // per column: word 0 = parameter-tile column of each factor (desc->d = the constant-one column), word 1 = index into cval
      const unsigned long long one = (unsigned long long)dext, none = (unsigned long long)MAXP;
      unsigned long long w0 = 0ull, w1 = 0ull;
      for (int k = 0; k < 7; ++k) { w0 |= one << (8 * k); w1 |= none << (8 * k); }
      std::vector<unsigned long long> paths((size_t)2 * ast);
      for (int c = 0; c < ast; ++c) { paths[2 * c] = w0 | 0xFF00000000000000ull; paths[2 * c + 1] = w1; }   // structural zeros
      paths[0] = w0;                                                                                          // column 0 = 1 (empty product)
      DLB_REQUIRE(dext < 255, "too many parameters for the in-kernel monomial evaluation");
      for (int t = 0; t < gd.nfeat - 2 && 2 + t < ast; ++t) {
        unsigned long long a0 = w0, a1 = w1;
        int k = 0;
        for (int j = 0; j < gd.nparams; ++j)
          for (int e = 0; e < gd.idata[t * gd.nparams + j]; ++e, ++k) {
            const unsigned long long colj = plan->gens[ig].pidx[j] >= 0 ? (unsigned long long)plan->gens[ig].pidx[j] : one;
            a0 = (a0 & ~(0xFFull << (8 * k))) | (colj << (8 * k));
            a1 = (a1 & ~(0xFFull << (8 * k))) | ((unsigned long long)j << (8 * k));
          }
        paths[2 * (2 + t)] = a0; paths[2 * (2 + t) + 1] = a1;
      }
      DLB_TRY(upload(plan, paths.data(), paths.size(), &plan->d_paths[ig]));
      // Two-pass form (Taylor order <= 4, the reference default): the columns of degree <= 2 are evaluated from their (at most two)
      // factors, every column of degree 3 - 4 is then ONE product of two of those columns of the same row -- 1 FP64 multiply per entry
      // instead of 3 multiplies + 4 subtractions queued behind the co-resident CTA's DMMAs in the FP64 pipe.
      if (gendev[ig].maxdeg <= 4 && ast < 65536 && !getenv("DLB_MONO_FACTORS")) {
        const int T = gd.nfeat - 2, np_ = gd.nparams;
        auto key_of = [&](const std::vector<int>& pw) { std::string k; for (int v : pw) k.push_back((char)('0' + v)); return k; };
        std::vector<std::pair<std::string, int>> col_of;   // powers -> column (first occurrence)
        auto find_col = [&](const std::string& k) { for (auto& kv : col_of) if (kv.first == k) return kv.second; return -1; };
        for (int t = 0; t < T && 2 + t < ast; ++t) {
          std::vector<int> pw(gd.idata + (size_t)t * np_, gd.idata + (size_t)(t + 1) * np_);
          if (find_col(key_of(pw)) < 0) col_of.emplace_back(key_of(pw), 2 + t);
        }
        std::vector<unsigned long long> pass1, pass2;
        const unsigned long long one = (unsigned long long)dext, none = (unsigned long long)MAXP;
        auto word1 = [&](int col, unsigned long long ca, unsigned long long cb, unsigned long long ia, unsigned long long ib, bool zero) {
          return ca | (cb << 8) | (ia << 16) | (ib << 24) | ((unsigned long long)col << 32) | (zero ? 0xFF00000000000000ull : 0ull);
        };
        bool ok = true;
        pass1.push_back(word1(0, one, one, none, none, false));
        for (int c = 1; c < ast; ++c) {
          const int t = c - 2;
          if (t < 0 || t >= T) { pass1.push_back(word1(c, one, one, none, none, true)); continue; }
          std::vector<std::pair<unsigned long long, unsigned long long>> fac;   // (parameter-tile column, cval index) per factor
          std::vector<int> jf;
          for (int j = 0; j < np_; ++j)
            for (int e = 0; e < gd.idata[(size_t)t * np_ + j]; ++e) {
              fac.emplace_back(plan->gens[ig].pidx[j] >= 0 ? (unsigned long long)plan->gens[ig].pidx[j] : one, (unsigned long long)j);
              jf.push_back(j);
            }
          if (fac.size() <= 2) {
            while (fac.size() < 2) fac.emplace_back(one, none);
            pass1.push_back(word1(c, fac[0].first, fac[1].first, fac[0].second, fac[1].second, false));
          } else {
            std::vector<int> pl(np_, 0), pr(np_, 0);
            for (size_t f = 0; f < jf.size(); ++f) (f < 2 ? pl : pr)[jf[f]] += 1;
            const int cl = find_col(key_of(pl)), cr = find_col(key_of(pr));
            if (cl < 0 || cr < 0) { ok = false; break; }
            pass2.push_back((unsigned long long)cl | ((unsigned long long)cr << 16) | ((unsigned long long)c << 32));
          }
        }
        if (ok) {
          plan->mono_n1[ig] = (int)pass1.size(); plan->mono_n2[ig] = (int)pass2.size();
          pass1.insert(pass1.end(), pass2.begin(), pass2.end());
          DLB_TRY(upload(plan, pass1.data(), pass1.size(), &plan->d_mono2[ig]));
        }
      }
    }
  }
  DLB_TRY(upload(plan, gendev.data(), gendev.size(), &plan->d_gens));
  plan->blocks.resize(desc->nblocks);
  int coloff = 0; long long poff = 0;
  std::vector<Block*> bptr;
  for (int ib = 0; ib < desc->nblocks; ++ib) {
    Block& blk = plan->blocks[ib];
    int rc = build_block(plan, desc, desc->blocks[ib], var, blk);
    if (rc) return rc;
    blk.coloff = coloff; coloff += blk.ncols;
    blk.poff = poff; poff += blk.npart;     // the kernel sums the Gram partials over column blocks
    blk.fill = plan->gens[blk.gen].aux[6] ? 0 : 1; blk.fbase = plan->gens[blk.gen].foff;
    DLB_REQUIRE(!resident || blk.maxcol <= ast, "internal: resident A tile narrower than a chunk extent");
    bptr.push_back(&blk);
  }
  plan->ntheory = coloff; plan->ptotal = poff;
  {
    bool all = plan->mlp_slices > 0;
    for (const Block& b : plan->blocks) all = all && b.oz.on;
    plan->oz_contract = all;
    if (all)
      for (const Block& b : plan->blocks) { plan->oz_ldD = std::max(plan->oz_ldD, b.oz.ldD); plan->oz_maxk = std::max(plan->oz_maxk, b.oz.K); }
  }
  for (int ib = 0; ib < desc->nblocks; ++ib) plan->two_stage = plan->two_stage || desc->blocks[ib].transform_data != nullptr;
  if (plan->two_stage) {
    // Solved parameters in a two-stage plan: the reference linearises flatdiff = u(t(x)) - data at the defaults (FastFisher, likelihoods/base.py:
    // 152-158: jacfwd through the transform), i.e. Jacobian rows u'(t) * dt/dx_i.  Stage 1 then carries EVERY solved template of the block as a slot
    // (constant templates too: scaled by u'(t(theta)) they depend on theta) and writes [u | u' S_1 | u' S_2 ...] per block into w_T.
    int tcol = 0;
    for (int ib = 0; ib < desc->nblocks; ++ib) {
      const dlb_block_desc& bd = desc->blocks[ib];
      DLB_REQUIRE(bd.transform_data && bd.whitening_t && bd.r0_data_white && bd.matrices_plain && bd.t0_plain && bd.nconst == 0,
                  "observable transform: every block needs transform_data, whitening_t, r0_data_white, matrices_plain, t0_plain (and nconst == 0)");
      DLB_TRY(upload(plan, bd.transform_data, (size_t)bd.ncols, &plan->blocks[ib].d_tdata));
      plan->blocks[ib].tcol = tcol;
      tcol += bd.nslots * ((bd.ncols + 3) & ~3);   // slots start on 32-byte boundaries; k-chunks may read up to the padded length (zeros)
    }
    plan->tstride = tcol;
  }
  DLB_TRY(build_launch(plan, bptr, resident, ast, desc->ncoef, plan->launch));
  {   // quadratic-form evaluation (lowering.py: qform_tables): every block or none
    bool allq = resident && !plan->two_stage && !getenv("DLB_NO_QFORM");
    for (int ib = 0; ib < desc->nblocks; ++ib) {
      const dlb_block_desc& bd = desc->blocks[ib];
      allq = allq && bd.q_ncols > 0 && bd.q_matrix && bd.q_tiles && bd.q_groups && bd.q_dent && bd.q_const && bd.q_ngroups > 0 && (bd.q_ncols % 8) == 0;
    }
    // Candidates (measured within 3 % of each other on C4 / C5 + marginalisation): 2 CTAs / SM x 8 warps with 32-row tiles and two k-chunks of 36, then
    // the 16-warp, 64-row forms: KC = 68 with 112-column blocks (one k-chunk per column block) when the features fit, KC = 76 with 96-column blocks, KC = 36
    int qK = 0;
    for (int ib = 0; ib < desc->nblocks && allq; ++ib) qK = std::max(qK, desc->blocks[ib].q_K);
    std::vector<Variant> qcand;
    if (const char* e = getenv("DLB_QF_VARIANT")) { const int i = atoi(e); qcand.push_back(i == 0 ? Variant{1, 1, 7, 8, 1, 68} : (i == 1 ? Variant{1, 1, 6, 8, 1, 76} : (i == 3 ? Variant{1, 1, 7, 4, 2, 36} : Variant{1, 1, 7, 8, 1, 36}))); }
    else {
      qcand.push_back(Variant{1, 1, 7, 4, 2, 36});   // 2 CTAs / SM x 8 warps, 32-row tiles: one CTA's epilogue overlaps the other's DMMAs (C4 +2 %, C5 + marg. +3 % over the 16-warp forms)
      if (chunk_extent(qK) <= 68) qcand.push_back(Variant{1, 1, 7, 8, 1, 68});
      if (chunk_extent(qK) <= 76) qcand.push_back(Variant{1, 1, 6, 8, 1, 76});
      qcand.push_back(Variant{1, 1, 7, 8, 1, 36});
    }
    Variant qv = qcand.back();
    bool qfits = false;
    for (const Variant& c : qcand)
      if (contract_smem_bytes(c, 2, true, ast, plan->dpad, desc->ncoef, npart_bound, desc->ngen) <= contract_smem_limit(c.CPS)) { qv = c; qfits = true; break; }
    if (allq && qfits) {
      plan->qblocks.resize(desc->nblocks);
      std::vector<Block*> qptr;
      const int BN = variant_bn(qv);
      for (int ib = 0; ib < desc->nblocks; ++ib) {
        const dlb_block_desc& bd = desc->blocks[ib];
        const dlb_term_desc& t0 = bd.terms[0];
        Block& qb = plan->qblocks[ib];
        const Block& ob = plan->blocks[ib];
        DLB_REQUIRE(bd.q_K > 0 && bd.q_foff >= t0.foff && (bd.q_foff % 2) == 0 && bd.q_foff + bd.q_K <= t0.foff + t0.K && bd.q_nw >= 0 && (bd.q_nw == 0 || (bd.q_wconst_val && bd.q_wconst_idx)),
                    "quadratic-form tables: feature range / weighted constants invalid");
        std::vector<TermSpec> terms{TermSpec{0, -1, plan->gens[t0.gen].foff + bd.q_foff, bd.q_foff, bd.q_K, bd.q_matrix, nullptr}};
        std::vector<double> zeros((size_t)bd.q_ncols, 0.);
        DLB_TRY(build_block_core(plan, qv, true, bd.q_ncols, 1, terms, zeros.data(), nullptr, 0, nullptr, qb));
        qb.npart = ob.npart; qb.poff = ob.poff; qb.coloff = ob.coloff; qb.gen = ob.gen; qb.fill = ob.fill; qb.fbase = ob.fbase; qb.d_map = ob.d_map;
        DLB_REQUIRE(qb.maxcol <= ast, "internal: resident A tile narrower than the quadratic-form features");
        const int ntile = qb.ncb * (BN / 8), nt_in = bd.q_ncols / 8;
        for (int g = 0; g < bd.q_ngroups; ++g) {
          const int32_t* q = bd.q_groups + 4 * g;
          DLB_REQUIRE((q[0] == 0 || q[0] == 1) && q[1] >= 0 && q[1] < ob.npart && q[2] >= -1 && q[2] < desc->ncoef && q[3] >= -1 && q[3] < desc->ncoef,
                      "quadratic-form tables: group entry out of range");
          for (int j = 0; j < 8; ++j)
            DLB_REQUIRE(bd.q_dent[24 * g + 3 * j] >= -1 && bd.q_dent[24 * g + 3 * j] < ob.npart && bd.q_dent[24 * g + 3 * j + 1] >= -1 && bd.q_dent[24 * g + 3 * j + 1] < desc->ncoef &&
                        bd.q_dent[24 * g + 3 * j + 2] >= -1 && bd.q_dent[24 * g + 3 * j + 2] < desc->ncoef, "quadratic-form tables: direct entry out of range");
          for (int j = 0; j < 8; ++j)
            for (int i = 0; i < j; ++i)
              DLB_REQUIRE(bd.q_dent[24 * g + 3 * j] < 0 || bd.q_dent[24 * g + 3 * j] != bd.q_dent[24 * g + 3 * i], "quadratic-form tables: the columns of a direct tile must feed distinct Gram entries");
        }
        DLB_REQUIRE(bd.q_ngroups < 32768 && desc->ncoef < 65535, "quadratic-form tables: too many groups / coefficients");
        // one 16-byte record per n-tile: { group | valid columns << 16 | kind << 24, first feature column, Gram entry, (coef a + 1) | (coef b + 1) << 16 }
        std::vector<int> tiles((size_t)4 * ntile, 0);
        for (int t = 0; t < ntile; ++t) tiles[4 * t] = 0xFFFF;   // group -1
        for (int t = 0; t < nt_in; ++t) {
          const int gid = bd.q_tiles[3 * t], k0 = bd.q_tiles[3 * t + 1], nv = bd.q_tiles[3 * t + 2];
          DLB_REQUIRE(gid >= -1 && gid < bd.q_ngroups && k0 >= 0 && k0 <= ast && nv >= 0 && nv <= 8, "quadratic-form tables: tile entry out of range");
          if (gid < 0) continue;
          const int32_t* q = bd.q_groups + 4 * gid;
          tiles[4 * t] = gid | (nv << 16) | (q[0] << 24);
          tiles[4 * t + 1] = k0; tiles[4 * t + 2] = q[1];
          tiles[4 * t + 3] = (q[2] + 1) | ((q[3] + 1) << 16);
        }
        DLB_TRY(upload(plan, tiles.data(), tiles.size(), &qb.d_qtile));
        DLB_TRY(upload(plan, bd.q_dent, (size_t)24 * bd.q_ngroups, &qb.d_qdent));
        DLB_TRY(upload(plan, bd.q_const, (size_t)ob.npart, &qb.d_qconst));
        {   // weighted constants, grouped by Gram entry
          std::vector<int> woff((size_t)ob.npart + 1, 0), wcoef((size_t)2 * std::max(bd.q_nw, 1), -1);
          std::vector<double> wval((size_t)std::max(bd.q_nw, 1), 0.);
          int prev = 0;
          for (int w = 0; w < bd.q_nw; ++w) {
            const int e = bd.q_wconst_idx[3 * w], ca = bd.q_wconst_idx[3 * w + 1], cbq = bd.q_wconst_idx[3 * w + 2];
            DLB_REQUIRE(e >= prev && e < ob.npart && ca >= -1 && ca < desc->ncoef && cbq >= -1 && cbq < desc->ncoef, "quadratic-form tables: weighted constants must be sorted by entry, indices in range");
            prev = e;
            woff[e + 1] += 1; wcoef[2 * w] = ca; wcoef[2 * w + 1] = cbq; wval[w] = bd.q_wconst_val[w];
          }
          for (int e = 0; e < ob.npart; ++e) woff[e + 1] += woff[e];
          DLB_TRY(upload(plan, woff.data(), woff.size(), &qb.d_qw_off));
          DLB_TRY(upload(plan, wcoef.data(), wcoef.size(), &qb.d_qw_coef));
          DLB_TRY(upload(plan, wval.data(), wval.size(), &qb.d_qw_val));
        }
        qptr.push_back(&qb);
      }
      DLB_TRY(build_launch(plan, qptr, true, ast, desc->ncoef, plan->qlaunch));
      plan->qform = true;
    }
  }
  if (plan->two_stage) {
    int maxn = 0;
    for (int ib = 0; ib < desc->nblocks; ++ib) maxn = std::max(maxn, desc->blocks[ib].ncols);
    const Variant wvar = pick_variant(plan, max_slots, false, 0, 0, npart_bound, maxn);
    plan->wblocks.resize(desc->nblocks);
    std::vector<Block*> wptr;
    for (int ib = 0; ib < desc->nblocks; ++ib) {
      const dlb_block_desc& bd = desc->blocks[ib];
      Block& wb = plan->wblocks[ib];
      const int cpad = (bd.ncols + 3) & ~3;
      std::vector<TermSpec> terms;
      for (int a = 0; a < bd.nslots; ++a) terms.push_back(TermSpec{a, -1, plan->blocks[ib].tcol + a * cpad, plan->blocks[ib].tcol + a * cpad, bd.ncols, bd.whitening_t, nullptr});
      DLB_TRY(build_block_core(plan, wvar, false, bd.ncols, bd.nslots, terms, bd.r0_data_white, nullptr, 0, nullptr, wb));
      wb.poff = plan->blocks[ib].poff; wb.coloff = plan->blocks[ib].coloff; wb.gen = -1; wb.d_map = plan->blocks[ib].d_map;
      wptr.push_back(&wb);
    }
    DLB_TRY(build_launch(plan, wptr, false, 0, 0, plan->wlaunch));
  }
  {
    size_t maxdepth = 0;
    for (MlpLayers& mlp : plan->mlps) maxdepth = std::max(maxdepth, mlp.layers.size());
    for (size_t l = 0; l < maxdepth; ++l)
      for (int act = 0; act <= 3; ++act)
        for (int res = 0; res < 2; ++res) {
          std::vector<Block*> group;
          int gast = 4;
          for (MlpLayers& mlp : plan->mlps)
            if (l < mlp.layers.size() && mlp.act == act && (mlp.layers[l].maxcol >= 0) == (res == 1)) {
              group.push_back(&mlp.layers[l]);
              if (res) gast = std::max(gast, round_ast(mlp.layers[l].maxcol));
            }
          if (group.empty()) continue;
          plan->mlp_launches.emplace_back();
          plan->mlp_launches.back().second = act;
          DLB_TRY(build_launch(plan, group, res == 1, res ? gast : 0, 0, plan->mlp_launches.back().first));
        }
  }
  std::vector<SolveBlock> sb(desc->nblocks);
  for (int ib = 0; ib < desc->nblocks; ++ib) sb[ib] = SolveBlock{plan->blocks[ib].poff, 1, plan->blocks[ib].npart, plan->blocks[ib].d_map};
  DLB_TRY(upload(plan, sb.data(), sb.size(), &plan->d_solve_blocks));
  // solved parameters: M0 = diag(prior precision) + constant-row Gram, g0 = prior gradient at x0
  const int ns = desc->ns;
  if (ns > 0) {
    DLB_REQUIRE(desc->x0 && desc->xprior_kind && desc->xprior_loc && desc->xprior_scale && desc->xprior_lo && desc->xprior_hi, "solved-parameter tables required");
    std::vector<double> M0((size_t)ns * ns, 0.), g0(ns, 0.), xprec(ns, 0.);
    for (int i = 0; i < ns; ++i) {
      if (desc->xprior_kind[i] == 2) {
        DLB_REQUIRE(desc->xprior_scale[i] > 0., "solved-parameter prior scale must be positive");
        xprec[i] = std::isinf(desc->xprior_scale[i]) ? 0. : 1. / (desc->xprior_scale[i] * desc->xprior_scale[i]);
      }
      M0[(size_t)i * ns + i] = xprec[i];
      g0[i] = -(desc->x0[i] - desc->xprior_loc[i]) * xprec[i];
    }
    for (int ib = 0; ib < desc->nblocks; ++ib) {
      const dlb_block_desc& bd = desc->blocks[ib];
      for (int i = 0; i < bd.nconst && !bd.skip_const_gram; ++i)
        for (int j = 0; j < bd.nconst; ++j) {
          double s = 0.;
          for (int c = 0; c < bd.ncols; ++c) s += bd.gconst_white[(size_t)i * bd.ncols + c] * bd.gconst_white[(size_t)j * bd.ncols + c];
          M0[(size_t)bd.const_solved[i] * ns + bd.const_solved[j]] += s;
        }
    }
    if (desc->m0_extra)
      for (size_t i = 0; i < M0.size(); ++i) M0[i] += desc->m0_extra[i];
    DLB_TRY(upload(plan, M0.data(), M0.size(), &plan->d_M0));
    DLB_TRY(upload(plan, g0.data(), g0.size(), &plan->d_g0));
    DLB_TRY(upload(plan, desc->x0, (size_t)ns, &plan->d_x0));
    DLB_TRY(upload(plan, xprec.data(), (size_t)ns, &plan->d_xprec));
    DLB_TRY(upload(plan, desc->xprior_loc, (size_t)ns, &plan->d_xloc));
    DLB_TRY(upload(plan, desc->xprior_lo, (size_t)ns, &plan->d_xlo));
    DLB_TRY(upload(plan, desc->xprior_hi, (size_t)ns, &plan->d_xhi));
    DLB_TRY(upload(plan, desc->xprior_kind, (size_t)ns, &plan->d_xkind));
    // connected components of the solved parameters over the likelihood blocks (parameters sharing a block are coupled;
    // the reference groups likelihoods sharing solved parameters the same way, likelihoods/base.py:60-97)
    std::vector<int> parent(ns);
    for (int i = 0; i < ns; ++i) parent[i] = i;
    auto find = [&](int i) { while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; } return i; };
    for (int ib = 0; ib < desc->nblocks; ++ib) {
      const dlb_block_desc& bd = desc->blocks[ib];
      int first = -1;
      auto join = [&](int i) { if (first < 0) first = i; else parent[find(i)] = find(first); };
      if (bd.entry_kind) {   // explicit maps: parameters coupled by a Hessian entry (or by the plan-constant extra Gram) share a group
        for (int e = 0; e < plan->blocks[ib].npart; ++e)
          if (bd.entry_kind[e] == 2) { join(bd.entry_i[e]); join(bd.entry_j[e]); first = -1; }
        continue;
      }
      for (int a = 0; a < bd.nslots - 1; ++a) join(bd.slot_solved[a]);
      for (int a = 0; a < bd.nconst; ++a) join(bd.const_solved[a]);
    }
    if (desc->m0_extra)
      for (int i = 0; i < ns; ++i)
        for (int j = 0; j < i; ++j)
          if (desc->m0_extra[(size_t)i * ns + j] != 0.) parent[find(i)] = find(j);
    std::vector<int> comp_off{0}, comp_idx, comp_nmarg;
    int maxsize = 0;
    for (int r = 0; r < ns; ++r) {
      if (find(r) != r) continue;
      int nm = 0;
      for (int i = 0; i < ns; ++i) if (find(i) == r && i < desc->nmarg) { comp_idx.push_back(i); ++nm; }   // '.marg' members first
      for (int i = 0; i < ns; ++i) if (find(i) == r && i >= desc->nmarg) comp_idx.push_back(i);
      comp_nmarg.push_back(nm);
      comp_off.push_back((int)comp_idx.size());
      maxsize = std::max(maxsize, comp_off.back() - comp_off[comp_off.size() - 2]);
    }
    {
      std::vector<int> hij, comp_of(ns);
      for (int i = 0; i < ns; ++i) { comp_of[i] = find(i); for (int j = i; j < ns; ++j) hij.push_back((i << 16) | j); }
      DLB_TRY(upload(plan, hij.data(), hij.size(), &plan->d_hess_ij));
      DLB_TRY(upload(plan, comp_of.data(), comp_of.size(), &plan->d_comp_of));
    }
    if (maxsize <= 12) {   // small groups: one lane factorises one group
      plan->ncomp = (int)comp_nmarg.size();
      DLB_TRY(upload(plan, comp_off.data(), comp_off.size(), &plan->d_comp_off));
      DLB_TRY(upload(plan, comp_idx.data(), comp_idx.size(), &plan->d_comp_idx));
      DLB_TRY(upload(plan, comp_nmarg.data(), comp_nmarg.size(), &plan->d_comp_nmarg));
      std::vector<int> pos;
      std::vector<double> m0v;
      for (size_t c = 0; c + 1 < comp_off.size(); ++c)
        for (int a = comp_off[c]; a < comp_off[c + 1]; ++a)
          for (int b = comp_off[c]; b < comp_off[c + 1]; ++b) {
            pos.push_back(comp_idx[a] * (ns + 1) + comp_idx[b]);
            m0v.push_back(M0[(size_t)comp_idx[a] * ns + comp_idx[b]]);
          }
      plan->npos = (int)pos.size();
      DLB_TRY(upload(plan, pos.data(), pos.size(), &plan->d_comp_pos));
      DLB_TRY(upload(plan, m0v.data(), m0v.size(), &plan->d_comp_m0));
      plan->comp_max = maxsize;
      if (maxsize <= 8 && !getenv("DLB_SOLVE_WARP")) {   // tables of the thread-per-point solve kernel (csrc/solve.cu)
        const int nc = plan->ncomp;
        std::vector<int> group_of(ns, -1), local_of(ns, -1), m0_off(nc, 0);
        for (int c = 0, acc = 0; c < nc; ++c) {
          m0_off[c] = acc;
          const int n = comp_off[c + 1] - comp_off[c];
          acc += n * n;
          for (int a = 0; a < n; ++a) { group_of[comp_idx[comp_off[c] + a]] = c; local_of[comp_idx[comp_off[c] + a]] = a; }
        }
        std::vector<std::vector<GroupEntry>> per(nc);
        std::vector<long long> chi_rows;
        for (int ib = 0; ib < desc->nblocks; ++ib) {
          const Block& blk = plan->blocks[ib];
          for (int e = 0; e < blk.npart; ++e) {
            const EntryMap& m = blk.h_map[e];
            const long long prow = blk.poff + e;
            if (m.kind == 0) { chi_rows.push_back(prow); continue; }
            if (m.kind == 3) continue;
            const int c = group_of[m.i];
            if (m.kind == 1) per[c].push_back(GroupEntry{prow, 1, local_of[m.i], 0});
            else {
              DLB_REQUIRE(group_of[m.j] == c, "internal: coupled solved parameters in different groups");
              per[c].push_back(GroupEntry{prow, 2, std::max(local_of[m.i], local_of[m.j]), std::min(local_of[m.i], local_of[m.j])});
            }
          }
        }
        std::vector<GroupEntry> gent;
        std::vector<int> gent_off{0};
        for (int c = 0; c < nc; ++c) { gent.insert(gent.end(), per[c].begin(), per[c].end()); gent_off.push_back((int)gent.size()); }
        plan->nchi = (int)chi_rows.size();
        DLB_TRY(upload(plan, gent.data(), gent.size(), &plan->d_gent));
        DLB_TRY(upload(plan, gent_off.data(), gent_off.size(), &plan->d_gent_off));
        DLB_TRY(upload(plan, chi_rows.data(), chi_rows.size(), &plan->d_chi_rows));
        DLB_TRY(upload(plan, m0_off.data(), m0_off.size(), &plan->d_m0_off));
      }
    }
  }
  for (auto& e : plan->ev) DLB_CUDA(cudaEventCreate(&e));
  guard.p = nullptr;
  *out = plan;
  return DLB_OK;
}

static const long long MAX_CHUNK = 1LL << 20;

// Points per pass of the batched entry points: at most 2^20, fewer when the per-point workspace (feature rows of wide MLP emulators,
// Gram partials, transformed theory vectors) would exceed ~16 GiB -- the workspace stays a small part of the 180 GB of HBM.
static long long chunk_points(const dlb_plan* plan) {
  const long long per_point = 8LL * ((long long)plan->fstride + std::max(plan->ptotal, 1LL) + plan->d + 4 + plan->ns + plan->tstride + plan->oz_ldD) +
                              (long long)plan->oz_maxk * plan->mlp_slices;
  long long n = (16LL << 30) / std::max(per_point, 1LL) / ROW_ALIGN * ROW_ALIGN;
  n = std::max<long long>(ROW_ALIGN * 32, std::min(MAX_CHUNK, n));
  // Feature matrices written by their own kernels (direct Kaiser / BAO theories) are consumed by the contraction right after: a pass whose
  // feature matrix fits in L2 (~64 MB of the 126 MB) never sends it to HBM (ncu, C1: 7.5 MB of DRAM writes per 65 536 points instead of 170 MB).
  // Only when such a pass still holds >= 65 536 points -- smaller passes cost more in kernel efficiency than the traffic saves.
  if (!plan->resident && plan->fstride > 4 && !getenv("DLB_NO_L2_CHUNK")) {
    const long long fit = (64LL << 20) / (8LL * plan->fstride) / ROW_ALIGN * ROW_ALIGN;
    if (fit >= 65536) n = std::min(n, fit);
  }
  if (const char* e = getenv("DLB_MAX_CHUNK")) n = std::max<long long>(ROW_ALIGN, std::min(n, atoll(e) / ROW_ALIGN * ROW_ALIGN));   // tests: force several passes
  return n;
}

int dlb_eval_batch(dlb_plan* plan, const double* params, int64_t N, double* logposterior, double* loglikelihood, double* logprior,
                   double* x_solved, int32_t* status, void* stream_) {
  return dlb_eval_batch_hessian(plan, params, N, logposterior, loglikelihood, logprior, x_solved, nullptr, status, stream_);
}

// Sub-batches of the pipelined host-buffer path  H2D (copy-in stream) | kernels (caller's stream) | D2H (copy-out stream).
// The persistent contraction kernel takes as long as its busiest CTA, so lengths are whole multiples of the WAVE QUANTUM (grid x tile rows) of the
// likelihood contraction (the last sub-batch takes the rest): a split then adds no tile pass.  Candidates: equal pass counts, and a short first / last
// sub-batch (1-2 passes: little H2D before the first kernel, little D2H after the last) around equal middle ones; the three-stage timeline of every
// candidate is simulated with the plan's executed FLOP per point at ~ 55 % of the FP64 tensor peak and 50 GB/s per PCIe direction, and the shortest wins.
// Every launch after the first is charged a fraction of a tile pass for ramp-up / tail imbalance (measured on B200, 131 072 points: C4 0.12 ms of a
// 0.36 ms pass; variants whose second-resident CTA starts with a half tile stream the matrices once more per launch -- C5 0.2 ms) plus a fixed cost
// for the launches, cross-stream events and small copies.  DLB_NSUB forces the count.
static int choose_sub_batches(const dlb_plan* plan, long long n, double bytes_in, double bytes_out, long long* lens) {
  const int forced = getenv("DLB_NSUB") ? atoi(getenv("DLB_NSUB")) : 0;   // (read per call: tests switch it)
  static const double T_SUB = getenv("DLB_SUB_COST_US") ? atof(getenv("DLB_SUB_COST_US")) * 1e-6 : 40e-6;
  // The timeline model rests on the persistent contraction being the step: validated on the plans with a shared-memory-resident A tile (C2, C4, C5,
  // C5 + marginalisation).  Plans led by other kernels (direct-theory generators, per-layer MLP launches, the int8 path, two-stage plans) keep the
  // earlier rule: from 131 072 points up, up to 4 equal sub-batches of >= 65 536 points.
  if (!forced && !(plan->mlp_slices == 0 && !plan->two_stage && plan->mlp_launches.empty() && (plan->qform || plan->resident))) {
    const int c = n >= 131072 ? (int)std::min<long long>(4, n / 65536) : 1;
    const long long len = ((n + c - 1) / c + 63) / 64 * 64;
    for (int i = 0; i < c; ++i) lens[i] = i + 1 < c ? len : n - len * (c - 1);
    return c;
  }
  const Launch& L = plan->two_stage ? plan->wlaunch : (plan->qform ? plan->qlaunch : plan->launch);
  const long long Q = (std::max<long long>(64, (long long)plan->nsm * L.var.CPS * variant_bm(L.var)) + 63) / 64 * 64;
  const double t_pt = std::max(dlb_plan_executed_flops(plan), 0.) / 20e12, BW = 50e9;
  const double t_launch = T_SUB + ((L.var.CPS == 2 && L.var.MT > 1) ? 0.8 : 0.3) * (double)Q * t_pt;
  const long long P = (n + Q - 1) / Q;   // tile passes of the whole batch
  auto simulate = [&](const long long* len, int c) {
    double tin = 0., kend = 0., tout = 0.;
    for (int i = 0; i < c; ++i) {
      tin += (double)len[i] * bytes_in / BW;
      kend = std::max(tin, kend) + (double)len[i] * t_pt + (i ? t_launch : 0.);
      tout = std::max(kend, tout) + (double)len[i] * bytes_out / BW;
    }
    return tout;
  };
  lens[0] = n;
  int best = 1;
  double best_t = simulate(lens, 1);
  bool have_forced = forced == 1;
  long long cand[MAX_SUB];
  auto consider = [&](const long long* passes, int c) {
    long long used = 0;
    for (int i = 0; i + 1 < c; ++i) { cand[i] = passes[i] * Q; used += cand[i]; }
    if (used >= n) return;
    cand[c - 1] = n - used;
    const double t = simulate(cand, c);
    if (forced ? (c == forced && (!have_forced || t < best_t)) : t < best_t) {
      best = c; best_t = t; have_forced = have_forced || forced == c;
      for (int i = 0; i < c; ++i) lens[i] = cand[i];
    }
  };
  long long passes[MAX_SUB];
  for (int c = 2; c <= MAX_SUB && c <= P; ++c) {
    for (int i = 0; i < c; ++i) passes[i] = P / c + (i < P % c ? 1 : 0);       // equal pass counts
    consider(passes, c);
    for (int f = 1; f <= 2; ++f)
      for (int l = 1; l <= 2; ++l) {                                             // short first / last sub-batch
        const long long mid = P - f - l;
        if (c < 3 || mid < c - 2) continue;
        passes[0] = f; passes[c - 1] = l;
        for (int i = 0; i < c - 2; ++i) passes[1 + i] = mid / (c - 2) + (i < mid % (c - 2) ? 1 : 0);
        consider(passes, c);
      }
  }
  return best;
}

static int eval_impl(dlb_plan* plan, const double* params, int64_t N, double* logposterior, double* loglikelihood, double* logprior,
                     double* x_solved, double* hessian, int32_t* status, void* stream_, dlb_gather* gather, int gbuf, long long gstart, long long gtotal,
                     bool gsolved);

int dlb_eval_batch_hessian(dlb_plan* plan, const double* params, int64_t N, double* logposterior, double* loglikelihood, double* logprior,
                           double* x_solved, double* hessian, int32_t* status, void* stream_) {
  return eval_impl(plan, params, N, logposterior, loglikelihood, logprior, x_solved, hessian, status, stream_, nullptr, 0, 0, 0, false);
}

int dlb_eval_batch_gather(dlb_plan* plan, const double* params, int64_t N, dlb_gather* gather, int32_t buf, int64_t slice_start, int64_t total,
                          int32_t with_solved, double* loglikelihood, double* logprior, double* x_solved, int32_t* status, int32_t wait_buf, void* stream_) {
  DLB_REQUIRE(gather, "NULL gather");
  DLB_TRY(eval_impl(plan, params, N, nullptr, loglikelihood, logprior, x_solved, nullptr, status, stream_, gather, buf, slice_start, total, with_solved != 0));
  if (wait_buf >= 0) return dlb_gather_wait(gather, wait_buf, stream_);
  return DLB_OK;
}

static int eval_impl(dlb_plan* plan, const double* params, int64_t N, double* logposterior, double* loglikelihood, double* logprior,
                     double* x_solved, double* hessian, int32_t* status, void* stream_, dlb_gather* gather, int gbuf, long long gstart, long long gtotal,
                     bool gsolved) {
  DLB_REQUIRE(plan && (params || N == 0) && N >= 0, "NULL plan / params or negative N");
  if (gather) {
    gather_next_sequence(gather, gbuf);
    DLB_TRY(gather_before_write(gather, gbuf, (cudaStream_t)stream_));   // an earlier DMA push out of this buffer must have completed (data AND flags)
    if (N == 0) return gather_publish_empty(gather, gbuf, gtotal, (cudaStream_t)stream_);   // a rank with an empty slice still publishes its sequence number
  }
  if (N == 0) return DLB_OK;
  if (plan->ns == 0) hessian = nullptr;
  const int nhess = plan->ns * (plan->ns + 1) / 2;
  const bool hs_dev = is_device_ptr(hessian);
  std::lock_guard<std::mutex> lock(plan->mutex);
  cudaStream_t stream = (cudaStream_t)stream_;
  DLB_CUDA(cudaSetDevice(plan->device));
  const bool in_dev = is_device_ptr(params);
  const bool lp_dev = is_device_ptr(logposterior), ll_dev = is_device_ptr(loglikelihood), pr_dev = is_device_ptr(logprior),
             xs_dev = is_device_ptr(x_solved), st_dev = is_device_ptr(status);
  bool any_host = !in_dev || (logposterior && !lp_dev) || (loglikelihood && !ll_dev) || (logprior && !pr_dev) || (x_solved && !xs_dev) || (status && !st_dev) ||
                  (hessian && !hs_dev);
  const int ns = plan->ns;
  // x_solved to the peers by DMA after the evaluation: large payloads, device-resident buffers only (with host buffers the copy engines are busy with
  // the pipelined H2D / D2H of the sub-batches: measured at 8 GPUs, C4: value 1.61e8 -> 1.68e8 with DMA, but e2e 1.19e8 -> 1.08e8)
  const bool gdma = !any_host && gather_use_dma(gather, N, ns, gsolved);
  const long long CHUNK = chunk_points(plan);
  for (long long start = 0; start < N; start += CHUNK) {
    const long long n = std::min<long long>(CHUNK, N - start);
    const long long npad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    int rc = ensure_workspace(plan, npad);
    if (rc) return rc;
    if (hessian && !hs_dev && plan->hess_cap < n) {
      if (plan->w_hess) cudaFree(plan->w_hess);
      plan->w_hess = nullptr; plan->hess_cap = 0;
      DLB_CUDA(cudaMalloc((void**)&plan->w_hess, (size_t)n * nhess * sizeof(double)));
      plan->hess_cap = n;
    }
    // Host buffers: split the chunk into sub-batches and pipeline  H2D (copy-in stream) | kernels (caller's stream) | D2H (copy-out
    // stream)  so that the PCIe transfers of one sub-batch overlap the kernels of its neighbours.  Device buffers: one pass.
    long long sublens[MAX_SUB] = {n};
    int nsub = 1;
    if (any_host && !plan->timing) {
      const double bytes_out = ((logposterior && !lp_dev) ? 8 : 0) + ((loglikelihood && !ll_dev) ? 8 : 0) + ((logprior && !pr_dev) ? 8 : 0) +
                               ((x_solved && ns && !xs_dev) ? 8 * ns : 0) + ((status && !st_dev) ? 4 : 0) + ((hessian && !hs_dev) ? 8 * nhess : 0);
      nsub = choose_sub_batches(plan, n, in_dev ? 0. : plan->d_in * 8., bytes_out, sublens);
    }
    cudaStream_t s_in = stream, s_out = stream;
    if (nsub > 1) {
      if (!plan->s_in) {
        DLB_CUDA(cudaStreamCreateWithFlags(&plan->s_in, cudaStreamNonBlocking));
        DLB_CUDA(cudaStreamCreateWithFlags(&plan->s_out, cudaStreamNonBlocking));
        for (auto& e : plan->ev_pipe) DLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      }
      s_in = plan->s_in; s_out = plan->s_out;
      DLB_CUDA(cudaEventRecord(plan->ev_pipe[2 * MAX_SUB], stream));          // staging buffers may still be in use by earlier work on `stream`
      DLB_CUDA(cudaStreamWaitEvent(s_in, plan->ev_pipe[2 * MAX_SUB], 0));
      DLB_CUDA(cudaStreamWaitEvent(s_out, plan->ev_pipe[2 * MAX_SUB], 0));
    }
    double* w = plan->w_out;
    long long off = 0;
    for (int c = 0; c < nsub; off += sublens[c], ++c) {
      const long long len = sublens[c];
      const long long npad_s = (len + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
      const double* d_params = params + (start + off) * plan->d_in;
      if (!in_dev) {
        DLB_CUDA(cudaMemcpyAsync(plan->w_params + off * plan->d_in, d_params, (size_t)len * plan->d_in * sizeof(double), cudaMemcpyHostToDevice, s_in));
        d_params = plan->w_params + off * plan->d_in;
        if (nsub > 1) { DLB_CUDA(cudaEventRecord(plan->ev_pipe[c], s_in)); DLB_CUDA(cudaStreamWaitEvent(stream, plan->ev_pipe[c], 0)); }
      }
      if (plan->nder) {   // extended rows: inputs + derived-expression parameters (reference base.py:537)
        DLB_TRY(launch_derived(d_params, plan->d_in, plan->nder, plan->d_der_off, plan->d_der_code, plan->d_der_const, len, plan->w_pext + off * plan->d, stream));
        d_params = plan->w_pext + off * plan->d;
      }
      if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[0], stream));
      rc = run_front(plan, d_params, len, npad_s, stream);
      if (rc) return rc;
      if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[1], stream));
      rc = run_contract(plan, d_params, len, npad_s, 0, nullptr, stream);
      if (rc) return rc;
      if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[2], stream));
      SolveParams sp;
      std::memset(&sp, 0, sizeof(sp));
      sp.params = d_params; sp.d = plan->d; sp.N = len; sp.npad = npad_s;
      sp.prior_kind = plan->d_prior_kind; sp.prior_lo = plan->d_prior_lo; sp.prior_hi = plan->d_prior_hi; sp.prior_loc = plan->d_prior_loc; sp.prior_scale = plan->d_prior_scale;
      sp.partial = plan->w_partial; sp.blocks = plan->d_solve_blocks; sp.nblocks = (int)plan->blocks.size();
      sp.ns = ns; sp.nmarg = plan->nmarg; sp.M0 = plan->d_M0; sp.g0 = plan->d_g0; sp.x0 = plan->d_x0; sp.xprec = plan->d_xprec;
      sp.xloc = plan->d_xloc; sp.xlo = plan->d_xlo; sp.xhi = plan->d_xhi; sp.xkind = plan->d_xkind;
      sp.ncomp = plan->ncomp; sp.comp_off = plan->d_comp_off; sp.comp_idx = plan->d_comp_idx; sp.comp_nmarg = plan->d_comp_nmarg;
      sp.npos = plan->npos; sp.comp_pos = plan->d_comp_pos; sp.comp_m0 = plan->d_comp_m0;
      const long long g0 = start + off;
      sp.logpost = logposterior ? (lp_dev ? logposterior + g0 : w + off) : nullptr;
      sp.loglike = loglikelihood ? (ll_dev ? loglikelihood + g0 : w + npad + off) : nullptr;
      sp.logprior = logprior ? (pr_dev ? logprior + g0 : w + 2 * npad + off) : nullptr;
      sp.xsolved = (x_solved && ns) ? (xs_dev ? x_solved + g0 * ns : w + 3 * npad + off * ns) : nullptr;
      sp.status = status ? (st_dev ? status + g0 : plan->w_status + off) : nullptr;
      sp.hess = hessian ? (hs_dev ? hessian + g0 * nhess : plan->w_hess + off * nhess) : nullptr;
      sp.nhess = nhess; sp.hess_ij = plan->d_hess_ij; sp.comp_of = plan->d_comp_of;
      sp.gent = plan->d_gent; sp.gent_off = plan->d_gent_off; sp.chi_rows = plan->d_chi_rows; sp.nchi = plan->nchi; sp.m0_off = plan->d_m0_off; sp.comp_max = plan->comp_max;
      if (gather) {   // the finish kernel stores this sub-batch's results into every rank's window; the last launch publishes the sequence number
        const bool last = start + CHUNK >= N && c == nsub - 1;
        DLB_TRY(gather_fill(gather, gbuf, gstart + g0, len, gtotal, ns, gsolved, last && !gdma, &sp.gather));
        if (gdma) sp.gather.xs_only = gather_rank(gather);
      }
      rc = launch_solve(sp, stream);
      if (rc) return rc;
      if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[3], stream));
      if (nsub > 1) { DLB_CUDA(cudaEventRecord(plan->ev_pipe[MAX_SUB + c], stream)); DLB_CUDA(cudaStreamWaitEvent(s_out, plan->ev_pipe[MAX_SUB + c], 0)); }
      if (logposterior && !lp_dev) DLB_CUDA(cudaMemcpyAsync(logposterior + g0, sp.logpost, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, s_out));
      if (loglikelihood && !ll_dev) DLB_CUDA(cudaMemcpyAsync(loglikelihood + g0, sp.loglike, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, s_out));
      if (logprior && !pr_dev) DLB_CUDA(cudaMemcpyAsync(logprior + g0, sp.logprior, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, s_out));
      if (x_solved && ns && !xs_dev) DLB_CUDA(cudaMemcpyAsync(x_solved + g0 * ns, sp.xsolved, (size_t)len * ns * sizeof(double), cudaMemcpyDeviceToHost, s_out));
      if (status && !st_dev) DLB_CUDA(cudaMemcpyAsync(status + g0, sp.status, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost, s_out));
      if (hessian && !hs_dev) DLB_CUDA(cudaMemcpyAsync(hessian + g0 * nhess, sp.hess, (size_t)len * nhess * sizeof(double), cudaMemcpyDeviceToHost, s_out));
    }
    if (nsub > 1) {   // the caller's stream (and the host) see the call complete only when the last copy-out has landed
      DLB_CUDA(cudaEventRecord(plan->ev_pipe[2 * MAX_SUB + 1], s_out));
      DLB_CUDA(cudaStreamWaitEvent(stream, plan->ev_pipe[2 * MAX_SUB + 1], 0));
    }
    if (any_host || plan->timing || start + CHUNK < N) DLB_CUDA(cudaStreamSynchronize(stream));
    if (plan->timing) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, plan->ev[0], plan->ev[1]);
      cudaEventElapsedTime(&b, plan->ev[1], plan->ev[2]);
      cudaEventElapsedTime(&c, plan->ev[2], plan->ev[3]);
      if (start == 0) plan->ms_features = plan->ms_contract = plan->ms_solve = 0.;
      plan->ms_features += a; plan->ms_contract += b; plan->ms_solve += c;
    }
  }
  if (gdma) DLB_TRY(gather_push_xs(gather, gbuf, gstart, N, gtotal, ns, stream));
  return DLB_OK;
}

int dlb_theory_batch(dlb_plan* plan, const double* params, int64_t N, double* theory, void* stream_) {
  DLB_REQUIRE(plan && params && theory && N >= 0, "NULL plan / params / theory or negative N");
  if (N == 0) return DLB_OK;
  std::lock_guard<std::mutex> lock(plan->mutex);
  cudaStream_t stream = (cudaStream_t)stream_;
  DLB_CUDA(cudaSetDevice(plan->device));
  const bool in_dev = is_device_ptr(params), out_dev = is_device_ptr(theory);
  const long long chunk = std::min<long long>(MAX_CHUNK, std::max<long long>(ROW_ALIGN, (1LL << 28) / std::max(plan->ntheory, 1) / ROW_ALIGN * ROW_ALIGN));
  for (long long start = 0; start < N; start += chunk) {
    const long long n = std::min<long long>(chunk, N - start);
    const long long npad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    int rc = ensure_workspace(plan, npad);
    if (rc) return rc;
    const double* d_params = params + start * plan->d_in;
    if (!in_dev) {
      DLB_CUDA(cudaMemcpyAsync(plan->w_params, d_params, (size_t)n * plan->d_in * sizeof(double), cudaMemcpyHostToDevice, stream));
      d_params = plan->w_params;
    }
    if (plan->nder) {
      DLB_TRY(launch_derived(d_params, plan->d_in, plan->nder, plan->d_der_off, plan->d_der_code, plan->d_der_const, n, plan->w_pext, stream));
      d_params = plan->w_pext;
    }
    double* d_theory = theory + start * plan->ntheory;
    if (!out_dev) {
      if (plan->theory_cap < n) {
        if (plan->w_theory) cudaFree(plan->w_theory);
        plan->w_theory = nullptr; plan->theory_cap = 0;
        DLB_CUDA(cudaMalloc((void**)&plan->w_theory, (size_t)n * plan->ntheory * sizeof(double)));
        plan->theory_cap = n;
      }
      d_theory = plan->w_theory;
    }
    rc = run_front(plan, d_params, n, npad, stream);
    if (rc) return rc;
    rc = run_contract(plan, d_params, n, npad, 1, d_theory, stream);
    if (rc) return rc;
    if (!out_dev) DLB_CUDA(cudaMemcpyAsync(theory + start * plan->ntheory, d_theory, (size_t)n * plan->ntheory * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (!out_dev || !in_dev || start + chunk < N) DLB_CUDA(cudaStreamSynchronize(stream));
  }
  return DLB_OK;
}

int dlb_features_batch(dlb_plan* plan, int32_t gen, const double* params, int64_t N, double* features, void* stream_) {
  DLB_REQUIRE(plan && params && features && N >= 0 && gen >= 0 && gen < plan->ngen, "bad arguments");
  if (N == 0) return DLB_OK;
  std::lock_guard<std::mutex> lock(plan->mutex);
  cudaStream_t stream = (cudaStream_t)stream_;
  DLB_CUDA(cudaSetDevice(plan->device));
  const bool in_dev = is_device_ptr(params), out_dev = is_device_ptr(features);
  const DevGen& g = plan->gens[gen];
  for (long long start = 0; start < N; start += MAX_CHUNK) {
    const long long n = std::min<long long>(MAX_CHUNK, N - start);
    const long long npad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    int rc = ensure_workspace(plan, npad);
    if (rc) return rc;
    const double* d_params = params + start * plan->d_in;
    if (!in_dev) {
      DLB_CUDA(cudaMemcpyAsync(plan->w_params, d_params, (size_t)n * plan->d_in * sizeof(double), cudaMemcpyHostToDevice, stream));
      d_params = plan->w_params;
    }
    if (plan->nder) {
      DLB_TRY(launch_derived(d_params, plan->d_in, plan->nder, plan->d_der_off, plan->d_der_code, plan->d_der_const, n, plan->w_pext, stream));
      d_params = plan->w_pext;
    }
    rc = run_front(plan, d_params, n, npad, stream, true);
    if (rc) return rc;
    DLB_CUDA(cudaMemcpy2DAsync(features + start * g.nfeat, (size_t)g.nfeat * sizeof(double), plan->w_F + g.foff, (size_t)plan->fstride * sizeof(double),
                               (size_t)g.nfeat * sizeof(double), (size_t)n, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream));
    DLB_CUDA(cudaStreamSynchronize(stream));
  }
  return DLB_OK;
}

int dlb_set_timing(dlb_plan* plan, int32_t enabled) {
  DLB_REQUIRE(plan, "NULL plan");
  plan->timing = enabled != 0;
  g_timing.store(plan->timing);
  return DLB_OK;
}

int dlb_last_kernel_ms(const dlb_plan* plan, double* features_ms, double* contract_ms, double* solve_ms) {
  DLB_REQUIRE(plan, "NULL plan");
  if (features_ms) *features_ms = plan->ms_features;
  if (contract_ms) *contract_ms = plan->ms_contract;
  if (solve_ms) *solve_ms = plan->ms_solve;
  return DLB_OK;
}

int dlb_ensemble_step(dlb_plan* plan, double* walkers, double* logposterior, int64_t nwalkers, uint64_t seed, uint64_t iteration, double a,
                      int64_t rank_offset, uint64_t* naccept, void* stream_) {
  DLB_REQUIRE(plan && walkers && logposterior && nwalkers >= 2 && a > 1., "NULL plan / walkers / logposterior, fewer than 2 walkers or a <= 1");
  DLB_REQUIRE(is_device_ptr(walkers) && is_device_ptr(logposterior) && (!naccept || is_device_ptr(naccept)),
              "ensemble step: walkers, logposterior (and naccept) must be device memory");
  std::lock_guard<std::mutex> lock(plan->sampler_mutex);
  cudaStream_t stream = (cudaStream_t)stream_;
  DLB_CUDA(cudaSetDevice(plan->device));
  const long long n0 = nwalkers / 2, n1 = nwalkers - n0;
  if (n1 > plan->prop_cap) {
    if (plan->w_prop) cudaFree(plan->w_prop);
    if (plan->w_logpy) cudaFree(plan->w_logpy);
    plan->w_prop = plan->w_logpy = nullptr; plan->prop_cap = 0;
    DLB_CUDA(cudaMalloc((void**)&plan->w_prop, (size_t)n1 * plan->d_in * sizeof(double)));
    DLB_CUDA(cudaMalloc((void**)&plan->w_logpy, (size_t)n1 * sizeof(double)));
    plan->prop_cap = n1;
  }
  for (int half = 0; half < 2; ++half) {
    const long long na = half == 0 ? n0 : n1;
    DLB_TRY(launch_stretch_propose(walkers, plan->d_in, n0, n1, half, seed, iteration, a, rank_offset, plan->w_prop, stream));
    DLB_TRY(dlb_eval_batch(plan, plan->w_prop, na, plan->w_logpy, nullptr, nullptr, nullptr, nullptr, stream_));
    DLB_TRY(launch_stretch_accept(walkers, logposterior, plan->w_prop, plan->w_logpy, plan->d_in, n0, n1, half, seed, iteration, a, rank_offset,
                                  reinterpret_cast<unsigned long long*>(naccept), stream));
  }
  return DLB_OK;
}

int dlb_ozaki_matmul(const double* X, int64_t rows, int32_t K, const double* W, const double* bias, int32_t N, int32_t slices, int32_t act, double* out) {
  DLB_REQUIRE(X && W && out && rows > 0 && K > 0 && N > 0, "ozaki matmul: NULL buffer or empty problem");
  int dev = 0, nsm = 0;
  DLB_CUDA(cudaGetDevice(&dev));
  DLB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  std::vector<int8_t> bsl; std::vector<int> beb;
  ozaki_slice_weights(W, K, N, slices, bsl, beb);
  const long long rows_pad = ozaki_rows_pad(rows);
  double *dX = nullptr, *dO = nullptr, *dB = nullptr, *dSb = nullptr; int8_t *dBs = nullptr, *dAs = nullptr; int* dEa = nullptr;
  std::vector<double> sb(beb.size());
  for (size_t i = 0; i < beb.size(); ++i) sb[i] = std::ldexp(1., beb[i]);
  struct Free { std::vector<void*> v; ~Free() { for (void* q : v) if (q) cudaFree(q); } } fr;
  auto alloc = [&](void** ptr, size_t bytes) { cudaError_t e = cudaMalloc(ptr, std::max<size_t>(bytes, 16)); if (e == cudaSuccess) fr.v.push_back(*ptr); return e; };
  DLB_CUDA(alloc((void**)&dX, (size_t)rows * K * 8)); DLB_CUDA(alloc((void**)&dO, (size_t)rows * N * 8)); DLB_CUDA(alloc((void**)&dB, (size_t)N * 8));
  DLB_CUDA(alloc((void**)&dBs, bsl.size())); DLB_CUDA(alloc((void**)&dSb, sb.size() * 8));
  DLB_CUDA(alloc((void**)&dAs, ozaki_slice_bytes(rows_pad, K, slices))); DLB_CUDA(alloc((void**)&dEa, (size_t)rows_pad * 4));
  DLB_CUDA(cudaMemcpy(dX, X, (size_t)rows * K * 8, cudaMemcpyHostToDevice));
  if (bias) DLB_CUDA(cudaMemcpy(dB, bias, (size_t)N * 8, cudaMemcpyHostToDevice));
  DLB_CUDA(cudaMemcpy(dBs, bsl.data(), bsl.size(), cudaMemcpyHostToDevice));
  DLB_CUDA(cudaMemcpy(dSb, sb.data(), sb.size() * 8, cudaMemcpyHostToDevice));
  DLB_TRY(launch_ozaki_layer(dX, rows, K, K, dBs, dSb, bias ? dB : nullptr, N, slices, act, dO, N, dAs, dEa, nsm, nullptr));
  DLB_CUDA(cudaDeviceSynchronize());
  DLB_CUDA(cudaMemcpy(out, dO, (size_t)rows * N * 8, cudaMemcpyDeviceToHost));
  return DLB_OK;
}

int dlb_last_fftlog_ms(double* ms) {
  DLB_REQUIRE(ms, "NULL output");
  return last_fftlog_ms(ms);
}

int dlb_measure_fp64_peak(int32_t kind, double* tflops) {
  DLB_REQUIRE(tflops, "NULL output");
  return measure_fp64_peak(kind, tflops);
}

}  // extern "C"

// C-ABI layer: plan construction (constant upload + pre-tiling of the term matrices into the contraction kernel's
// shared-memory image), workspace management and the batched entry points declared in include/desilike_b200.h.
#include "common.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace dlb {

static thread_local std::string g_error;
std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

struct Block {
  Variant var;
  int ncols = 0, nslots = 1, nconst = 0, ncb = 0, nchunks = 0, npart = 0, nstages = 2, coloff = 0;
  long long poff = 0;
  Chunk* d_chunks = nullptr;
  double *d_tiles_white = nullptr, *d_tiles_plain = nullptr, *d_r0_white = nullptr, *d_t0_plain = nullptr, *d_gconst = nullptr;
  EntryMap* d_map = nullptr;
};

struct MlpLayers { int act = 0; std::vector<Block> layers; };

}  // namespace dlb

struct dlb_plan {
  int device = 0;
  int d = 0, ngen = 0, ncoef = 0, cstride = 1, fstride = 4, ns = 0, nmarg = 0, ntheory = 0;
  std::vector<dlb::DevGen> gens;
  std::vector<dlb::Block> blocks;
  std::vector<dlb::MlpLayers> mlps;   // one per generator (empty unless kind == MLP)
  dlb::DevCoef* d_coefs = nullptr;
  int* d_prior_kind = nullptr;
  double *d_prior_lo = nullptr, *d_prior_hi = nullptr, *d_prior_loc = nullptr, *d_prior_scale = nullptr;
  double *d_M0 = nullptr, *d_g0 = nullptr, *d_x0 = nullptr, *d_xprec = nullptr, *d_xloc = nullptr, *d_xlo = nullptr, *d_xhi = nullptr;
  int* d_xkind = nullptr;
  dlb::SolveBlock* d_solve_blocks = nullptr;
  long long ptotal = 0;  // partial rows (in units of npad doubles)
  std::vector<void*> owned;
  // workspace
  long long cap = 0;
  double *w_F = nullptr, *w_C = nullptr, *w_partial = nullptr, *w_params = nullptr, *w_out = nullptr, *w_theory = nullptr;
  int* w_status = nullptr;
  long long theory_cap = 0;
  std::mutex mutex;
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

static Variant choose_variant(int nslots) {
  if (nslots <= 1) return Variant{1, 4, 7};
  if (nslots <= 4) return Variant{4, 2, 4};
  return Variant{8, 1, 4};
}

struct TermSpec { int slot, coef, fcol, K; const double* white; const double* plain; };  // matrices [K, ncols] row-major

// Chunk the terms, pre-tile their matrices into the contraction kernel's shared-memory image and upload everything.
static int build_block_core(dlb_plan* plan, int ncols, int nslots, const std::vector<TermSpec>& terms, const double* r0_white,
                            const double* t0_plain, int nconst, const double* gconst_white, Block& blk) {
  blk.ncols = ncols; blk.nslots = nslots; blk.nconst = nconst;
  blk.var = choose_variant(nslots);
  const int BN = variant_bn(blk.var);
  blk.ncb = (ncols + BN - 1) / BN;
  std::vector<Chunk> chunks;
  struct Src { int term; int k0; };
  std::vector<Src> srcs;
  for (size_t t = 0; t < terms.size(); ++t) {
    const TermSpec& td = terms[t];
    const int Kpad = (td.K + 3) & ~3;
    const int nch = (Kpad + KC - 1) / KC;
    int kc = (Kpad + nch - 1) / nch;
    kc = (kc + 3) & ~3;
    for (int k0 = 0; k0 < Kpad; k0 += kc) {
      Chunk ch;
      ch.slot = td.slot; ch.coef = td.coef; ch.fcol = td.fcol + k0; ch.kc = std::min(kc, Kpad - k0);
      chunks.push_back(ch);
      srcs.push_back(Src{(int)t, k0});
    }
  }
  blk.nchunks = (int)chunks.size();
  blk.npart = nslots * (nslots + 1) / 2 + nslots * nconst;
  blk.nstages = 0;
  for (int s = 4; s >= 2; --s)
    if (contract_smem_bytes(blk.var, s, blk.nchunks, blk.npart) <= 227 * 1024) { blk.nstages = s; break; }
  DLB_REQUIRE(blk.nstages >= 2, "block needs too much shared memory (too many terms or constant rows)");
  DLB_TRY(upload(plan, chunks.data(), chunks.size(), &blk.d_chunks));
  // pre-tile: [ncb][nchunks][BN][KC], n-major rows, zero padded
  const size_t tile = (size_t)BN * KC;
  std::vector<double> tiles((size_t)blk.ncb * blk.nchunks * tile);
  for (int pass = 0; pass < 2; ++pass) {
    bool any = false;
    for (const TermSpec& td : terms) any = any || (pass == 0 ? td.white : td.plain) != nullptr;
    if (!any) continue;
    std::fill(tiles.begin(), tiles.end(), 0.);
    for (int cb = 0; cb < blk.ncb; ++cb)
      for (int c = 0; c < blk.nchunks; ++c) {
        double* dst = tiles.data() + ((size_t)cb * blk.nchunks + c) * tile;
        const TermSpec& td = terms[srcs[c].term];
        const double* mat = pass == 0 ? td.white : td.plain;
        if (!mat) continue;
        for (int n = 0; n < BN; ++n) {
          const int col = cb * BN + n;
          if (col >= ncols) break;
          for (int k = 0; k < chunks[c].kc; ++k) {
            const int kk = srcs[c].k0 + k;
            if (kk < td.K) dst[(size_t)n * KC + k] = mat[(long long)kk * ncols + col];
          }
        }
      }
    DLB_TRY(upload(plan, tiles.data(), tiles.size(), pass == 0 ? &blk.d_tiles_white : &blk.d_tiles_plain));
  }
  std::vector<double> pad((size_t)blk.ncb * BN, 0.);
  if (r0_white) {
    std::copy(r0_white, r0_white + ncols, pad.begin());
    DLB_TRY(upload(plan, pad.data(), pad.size(), &blk.d_r0_white));
  }
  if (t0_plain) {
    std::copy(t0_plain, t0_plain + ncols, pad.begin());
    DLB_TRY(upload(plan, pad.data(), pad.size(), &blk.d_t0_plain));
  }
  std::vector<double> gpad((size_t)std::max(nconst, 1) * blk.ncb * BN, 0.);
  for (int i = 0; i < nconst; ++i) std::copy(gconst_white + (size_t)i * ncols, gconst_white + (size_t)(i + 1) * ncols, gpad.begin() + (size_t)i * blk.ncb * BN);
  DLB_TRY(upload(plan, gpad.data(), gpad.size(), &blk.d_gconst));
  return DLB_OK;
}

static int build_block(dlb_plan* plan, const dlb_plan_desc* desc, const dlb_block_desc& bd, Block& blk) {
  DLB_REQUIRE(bd.ncols > 0 && bd.nslots >= 1 && bd.nslots <= 8, "block: ncols > 0 and 1 <= nslots <= 8 required");
  DLB_REQUIRE(bd.nterms > 0 && bd.terms && bd.matrices_white && bd.r0_white, "block: terms / matrices_white / r0_white required");
  std::vector<TermSpec> terms;
  for (int t = 0; t < bd.nterms; ++t) {
    const dlb_term_desc& td = bd.terms[t];
    DLB_REQUIRE(td.slot >= 0 && td.slot < bd.nslots, "term slot out of range");
    DLB_REQUIRE(td.coef >= -1 && td.coef < desc->ncoef, "term coef out of range");
    DLB_REQUIRE(td.gen >= 0 && td.gen < desc->ngen, "term generator out of range");
    DLB_REQUIRE(td.K > 0 && td.foff >= 0 && (td.foff % 2) == 0 && td.foff + td.K <= plan->gens[td.gen].nfeat, "term feature range invalid (foff must be even)");
    terms.push_back(TermSpec{td.slot, td.coef, plan->gens[td.gen].foff + td.foff, td.K, bd.matrices_white + td.moff,
                             bd.matrices_plain ? bd.matrices_plain + td.moff : nullptr});
  }
  DLB_TRY(build_block_core(plan, bd.ncols, bd.nslots, terms, bd.r0_white, bd.t0_plain, bd.nconst, bd.gconst_white, blk));
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
  for (const EntryMap& m : map)
    DLB_REQUIRE(m.kind == 0 || (m.i >= 0 && m.i < desc->ns && (m.kind == 1 || (m.j >= 0 && m.j < desc->ns))), "solved-parameter index out of range");
  DLB_TRY(upload(plan, map.data(), map.size(), &blk.d_map));
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
  for (int l = 0; l < nh; ++l) {
    Block blk;
    std::vector<TermSpec> terms{TermSpec{0, -1, g.foff + off[l], width[l], nullptr, w}};
    const double* bias = w + (size_t)width[l] * width[l + 1];
    DLB_TRY(build_block_core(plan, width[l + 1], 1, terms, nullptr, bias, 0, nullptr, blk));
    blk.coloff = g.foff + off[l + 1];
    mlp.layers.push_back(blk);
    w = bias + width[l + 1];
  }
  return DLB_OK;
}

static int ensure_workspace(dlb_plan* plan, long long npad) {
  if (npad <= plan->cap) return DLB_OK;
  for (void* ptr : {(void*)plan->w_F, (void*)plan->w_C, (void*)plan->w_partial, (void*)plan->w_params, (void*)plan->w_out, (void*)plan->w_status})
    if (ptr) cudaFree(ptr);
  plan->w_F = plan->w_C = plan->w_partial = plan->w_params = plan->w_out = nullptr; plan->w_status = nullptr;
  plan->cap = 0;
  DLB_CUDA(cudaMalloc((void**)&plan->w_F, ((size_t)npad * plan->fstride + 64) * sizeof(double)));
  DLB_CUDA(cudaMemset(plan->w_F, 0, ((size_t)npad * plan->fstride + 64) * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_C, (size_t)npad * plan->cstride * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_partial, (size_t)std::max(plan->ptotal, 1LL) * npad * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_params, (size_t)npad * std::max(plan->d, 1) * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_out, (size_t)npad * (3 + std::max(plan->ns, 0)) * sizeof(double)));
  DLB_CUDA(cudaMalloc((void**)&plan->w_status, (size_t)npad * sizeof(int)));
  plan->cap = npad;
  return DLB_OK;
}

static int run_front(dlb_plan* plan, const double* d_params, long long N, long long npad, cudaStream_t stream) {
  if (plan->ncoef) {
    int rc = launch_coefs(plan->d_coefs, plan->ncoef, d_params, plan->d, N, npad, plan->w_C, plan->cstride, stream);
    if (rc) return rc;
  }
  for (size_t ig = 0; ig < plan->gens.size(); ++ig) {
    const DevGen& g = plan->gens[ig];
    if (g.kind == DLB_GEN_MLP) {
      DLB_TRY(launch_mlp_input(g, d_params, plan->d, N, npad, plan->w_F, plan->fstride, stream));
      for (const Block& blk : plan->mlps[ig].layers) {
        ContractParams cp;
        std::memset(&cp, 0, sizeof(cp));
        cp.F = plan->w_F; cp.fstride = plan->fstride; cp.C = plan->w_C; cp.cstride = plan->cstride;
        cp.chunks = blk.d_chunks; cp.nchunks = blk.nchunks; cp.tiles = blk.d_tiles_plain; cp.r0 = blk.d_t0_plain;
        cp.gconst = blk.d_gconst; cp.nslots = 1; cp.npad = npad;
        cp.out = plan->w_F; cp.ostride = plan->fstride; cp.ocol0 = blk.coloff; cp.ncols = blk.ncols; cp.out_rows = npad;
        cp.act = plan->mlps[ig].act; cp.mode = 1; cp.nstages = blk.nstages;
        const int BM = variant_bm(blk.var);
        DLB_TRY(launch_contract(blk.var, cp, (int)((npad + BM - 1) / BM), blk.ncb, stream));
      }
      continue;
    }
    DLB_TRY(launch_features(g, d_params, plan->d, N, npad, plan->w_F, plan->fstride, stream));
  }
  return DLB_OK;
}

static int run_contract(dlb_plan* plan, long long N, long long npad, int mode, double* d_theory, cudaStream_t stream) {
  for (const Block& blk : plan->blocks) {
    ContractParams cp;
    cp.F = plan->w_F; cp.fstride = plan->fstride; cp.C = plan->w_C; cp.cstride = plan->cstride;
    cp.chunks = blk.d_chunks; cp.nchunks = blk.nchunks;
    cp.tiles = mode == 0 ? blk.d_tiles_white : blk.d_tiles_plain;
    cp.r0 = mode == 0 ? blk.d_r0_white : blk.d_t0_plain;
    cp.gconst = blk.d_gconst; cp.nconst = blk.nconst; cp.nslots = blk.nslots; cp.npart = blk.npart;
    cp.partial = plan->w_partial; cp.npad = npad; cp.poff = blk.poff;
    cp.out = d_theory; cp.ostride = plan->ntheory; cp.ocol0 = blk.coloff; cp.ncols = blk.ncols;
    cp.out_rows = N; cp.act = 0; cp.mode = mode; cp.nstages = blk.nstages;
    if (mode == 1 && (!cp.tiles || !cp.r0)) { set_error("plan was built without un-whitened matrices: theory_batch unavailable"); return DLB_ERR_ARG; }
    const int BM = variant_bm(blk.var);
    int rc = launch_contract(blk.var, cp, (int)((N + BM - 1) / BM), blk.ncb, stream);
    if (rc) return rc;
  }
  return DLB_OK;
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

void dlb_plan_destroy(dlb_plan* plan) {
  if (!plan) return;
  cudaSetDevice(plan->device);
  for (void* ptr : plan->owned) cudaFree(ptr);
  for (void* ptr : {(void*)plan->w_F, (void*)plan->w_C, (void*)plan->w_partial, (void*)plan->w_params, (void*)plan->w_out, (void*)plan->w_status, (void*)plan->w_theory})
    if (ptr) cudaFree(ptr);
  for (auto& e : plan->ev) if (e) cudaEventDestroy(e);
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
  plan->d = desc->d; plan->ngen = desc->ngen; plan->ncoef = desc->ncoef; plan->ns = desc->ns; plan->nmarg = desc->nmarg;
  plan->cstride = std::max(desc->ncoef, 1);
  // generators / feature layout
  int foff = 0;
  for (int ig = 0; ig < desc->ngen; ++ig) {
    const dlb_generator_desc& gd = desc->gens[ig];
    DLB_REQUIRE(gd.nparams >= 0 && gd.nparams <= MAXP, "generator: too many parameters");
    DLB_REQUIRE(gd.nfeat > 0, "generator: nfeat must be positive");
    DevGen g;
    std::memset(&g, 0, sizeof(g));
    g.kind = gd.kind; g.nparams = gd.nparams; g.nfeat = gd.nfeat; g.fpad = (gd.nfeat + 3) & ~3; g.foff = foff;
    for (int j = 0; j < gd.nparams; ++j) {
      DLB_REQUIRE(gd.param_idx[j] >= -1 && gd.param_idx[j] < desc->d, "generator: parameter index out of range");
      g.pidx[j] = gd.param_idx[j];
      g.pdef[j] = gd.param_default ? gd.param_default[j] : 0.;
    }
    int* di = nullptr; double* dd = nullptr;
    DLB_TRY(upload(plan, gd.idata, (size_t)gd.ni, &di));
    DLB_TRY(upload(plan, gd.ddata, (size_t)gd.nd, &dd));
    g.idata = di; g.ni = gd.ni; g.ddata = dd; g.nd = gd.nd;
    if (gd.kind == DLB_GEN_TAYLOR) DLB_REQUIRE(gd.nfeat > 2 && gd.ni == (gd.nfeat - 2) * gd.nparams && gd.nd == gd.nparams, "taylor generator: nfeat = 2 + T, idata = powers[T, d], ddata = center[d]");
    plan->mlps.emplace_back();
    if (gd.kind == DLB_GEN_MLP) DLB_TRY(build_mlp(plan, gd, g, plan->mlps.back()));
    if (gd.kind == DLB_GEN_KAISER || gd.kind == DLB_GEN_BAO) {
      DLB_REQUIRE(gd.ni >= 8, "theory generator: idata header missing");
      g.fpad = std::max(g.fpad, (gd.idata[7] + 3) & ~3);   // idata[7] = total segment length incl. scratch
      g.aux[1] = gd.idata[3];                              // nt (template knots), needed on the host for shared-memory sizing
      g.aux[2] = gd.idata[0]; g.aux[3] = gd.idata[1]; g.aux[4] = gd.idata[5];
      if (gd.kind == DLB_GEN_BAO) { DLB_REQUIRE(gd.ni >= 16, "bao generator: idata header"); g.aux[5] = gd.idata[14]; }
    }
    foff += g.fpad;
    plan->gens.push_back(g);
  }
  plan->fstride = foff;
  // coefficients
  std::vector<DevCoef> coefs(std::max(desc->ncoef, 1));
  for (int j = 0; j < desc->ncoef; ++j) {
    const dlb_coef_desc& cd = desc->coefs[j];
    DLB_REQUIRE(cd.nf >= 0 && cd.nf <= DLB_COEF_MAXF, "coefficient: too many factors");
    coefs[j].scale = cd.scale; coefs[j].nf = cd.nf;
    for (int f = 0; f < cd.nf; ++f) {
      DLB_REQUIRE(cd.idx[f] >= 0 && cd.idx[f] < desc->d, "coefficient: parameter index out of range");
      coefs[j].idx[f] = cd.idx[f]; coefs[j].pow[f] = cd.pow[f];
    }
  }
  DLB_TRY(upload(plan, coefs.data(), coefs.size(), &plan->d_coefs));
  // priors
  DLB_REQUIRE(desc->prior_kind && desc->prior_lo && desc->prior_hi && desc->prior_loc && desc->prior_scale, "prior tables required");
  DLB_TRY(upload(plan, desc->prior_kind, (size_t)desc->d, &plan->d_prior_kind));
  DLB_TRY(upload(plan, desc->prior_lo, (size_t)desc->d, &plan->d_prior_lo));
  DLB_TRY(upload(plan, desc->prior_hi, (size_t)desc->d, &plan->d_prior_hi));
  DLB_TRY(upload(plan, desc->prior_loc, (size_t)desc->d, &plan->d_prior_loc));
  DLB_TRY(upload(plan, desc->prior_scale, (size_t)desc->d, &plan->d_prior_scale));
  // blocks
  plan->blocks.resize(desc->nblocks);
  int coloff = 0; long long poff = 0;
  for (int ib = 0; ib < desc->nblocks; ++ib) {
    Block& blk = plan->blocks[ib];
    int rc = build_block(plan, desc, desc->blocks[ib], blk);
    if (rc) return rc;
    blk.coloff = coloff; coloff += blk.ncols;
    blk.poff = poff; poff += (long long)blk.ncb * blk.npart;
  }
  plan->ntheory = coloff; plan->ptotal = poff;
  std::vector<SolveBlock> sb(desc->nblocks);
  for (int ib = 0; ib < desc->nblocks; ++ib) sb[ib] = SolveBlock{plan->blocks[ib].poff, plan->blocks[ib].ncb, plan->blocks[ib].npart, plan->blocks[ib].d_map};
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
      for (int i = 0; i < bd.nconst; ++i)
        for (int j = 0; j < bd.nconst; ++j) {
          double s = 0.;
          for (int c = 0; c < bd.ncols; ++c) s += bd.gconst_white[(size_t)i * bd.ncols + c] * bd.gconst_white[(size_t)j * bd.ncols + c];
          M0[(size_t)bd.const_solved[i] * ns + bd.const_solved[j]] += s;
        }
    }
    DLB_TRY(upload(plan, M0.data(), M0.size(), &plan->d_M0));
    DLB_TRY(upload(plan, g0.data(), g0.size(), &plan->d_g0));
    DLB_TRY(upload(plan, desc->x0, (size_t)ns, &plan->d_x0));
    DLB_TRY(upload(plan, xprec.data(), (size_t)ns, &plan->d_xprec));
    DLB_TRY(upload(plan, desc->xprior_loc, (size_t)ns, &plan->d_xloc));
    DLB_TRY(upload(plan, desc->xprior_lo, (size_t)ns, &plan->d_xlo));
    DLB_TRY(upload(plan, desc->xprior_hi, (size_t)ns, &plan->d_xhi));
    DLB_TRY(upload(plan, desc->xprior_kind, (size_t)ns, &plan->d_xkind));
  }
  for (auto& e : plan->ev) DLB_CUDA(cudaEventCreate(&e));
  guard.p = nullptr;
  *out = plan;
  return DLB_OK;
}

static const long long MAX_CHUNK = 1LL << 20;

int dlb_eval_batch(dlb_plan* plan, const double* params, int64_t N, double* logposterior, double* loglikelihood, double* logprior,
                   double* x_solved, int32_t* status, void* stream_) {
  DLB_REQUIRE(plan && params && N >= 0, "NULL plan / params or negative N");
  if (N == 0) return DLB_OK;
  std::lock_guard<std::mutex> lock(plan->mutex);
  cudaStream_t stream = (cudaStream_t)stream_;
  DLB_CUDA(cudaSetDevice(plan->device));
  const bool in_dev = is_device_ptr(params);
  const bool lp_dev = is_device_ptr(logposterior), ll_dev = is_device_ptr(loglikelihood), pr_dev = is_device_ptr(logprior),
             xs_dev = is_device_ptr(x_solved), st_dev = is_device_ptr(status);
  bool any_host = !in_dev || (logposterior && !lp_dev) || (loglikelihood && !ll_dev) || (logprior && !pr_dev) || (x_solved && !xs_dev) || (status && !st_dev);
  const int ns = plan->ns;
  for (long long start = 0; start < N; start += MAX_CHUNK) {
    const long long n = std::min<long long>(MAX_CHUNK, N - start);
    const long long npad = (n + ROW_ALIGN - 1) / ROW_ALIGN * ROW_ALIGN;
    int rc = ensure_workspace(plan, npad);
    if (rc) return rc;
    const double* d_params = params + start * plan->d;
    if (!in_dev) {
      DLB_CUDA(cudaMemcpyAsync(plan->w_params, d_params, (size_t)n * plan->d * sizeof(double), cudaMemcpyHostToDevice, stream));
      d_params = plan->w_params;
    }
    if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[0], stream));
    rc = run_front(plan, d_params, n, npad, stream);
    if (rc) return rc;
    if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[1], stream));
    rc = run_contract(plan, n, npad, 0, nullptr, stream);
    if (rc) return rc;
    if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[2], stream));
    SolveParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.params = d_params; sp.d = plan->d; sp.N = n; sp.npad = npad;
    sp.prior_kind = plan->d_prior_kind; sp.prior_lo = plan->d_prior_lo; sp.prior_hi = plan->d_prior_hi; sp.prior_loc = plan->d_prior_loc; sp.prior_scale = plan->d_prior_scale;
    sp.partial = plan->w_partial; sp.blocks = plan->d_solve_blocks; sp.nblocks = (int)plan->blocks.size();
    sp.ns = ns; sp.nmarg = plan->nmarg; sp.M0 = plan->d_M0; sp.g0 = plan->d_g0; sp.x0 = plan->d_x0; sp.xprec = plan->d_xprec;
    sp.xloc = plan->d_xloc; sp.xlo = plan->d_xlo; sp.xhi = plan->d_xhi; sp.xkind = plan->d_xkind;
    double* w = plan->w_out;
    sp.logpost = logposterior ? (lp_dev ? logposterior + start : w) : nullptr;
    sp.loglike = loglikelihood ? (ll_dev ? loglikelihood + start : w + npad) : nullptr;
    sp.logprior = logprior ? (pr_dev ? logprior + start : w + 2 * npad) : nullptr;
    sp.xsolved = (x_solved && ns) ? (xs_dev ? x_solved + start * ns : w + 3 * npad) : nullptr;
    sp.status = status ? (st_dev ? status + start : plan->w_status) : nullptr;
    rc = launch_solve(sp, stream);
    if (rc) return rc;
    if (plan->timing) DLB_CUDA(cudaEventRecord(plan->ev[3], stream));
    if (logposterior && !lp_dev) DLB_CUDA(cudaMemcpyAsync(logposterior + start, sp.logpost, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (loglikelihood && !ll_dev) DLB_CUDA(cudaMemcpyAsync(loglikelihood + start, sp.loglike, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (logprior && !pr_dev) DLB_CUDA(cudaMemcpyAsync(logprior + start, sp.logprior, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (x_solved && ns && !xs_dev) DLB_CUDA(cudaMemcpyAsync(x_solved + start * ns, sp.xsolved, (size_t)n * ns * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (status && !st_dev) DLB_CUDA(cudaMemcpyAsync(status + start, sp.status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (any_host || plan->timing || start + MAX_CHUNK < N) DLB_CUDA(cudaStreamSynchronize(stream));
    if (plan->timing) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, plan->ev[0], plan->ev[1]);
      cudaEventElapsedTime(&b, plan->ev[1], plan->ev[2]);
      cudaEventElapsedTime(&c, plan->ev[2], plan->ev[3]);
      if (start == 0) plan->ms_features = plan->ms_contract = plan->ms_solve = 0.;
      plan->ms_features += a; plan->ms_contract += b; plan->ms_solve += c;
    }
  }
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
    const double* d_params = params + start * plan->d;
    if (!in_dev) {
      DLB_CUDA(cudaMemcpyAsync(plan->w_params, d_params, (size_t)n * plan->d * sizeof(double), cudaMemcpyHostToDevice, stream));
      d_params = plan->w_params;
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
    rc = run_contract(plan, n, npad, 1, d_theory, stream);
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
    const double* d_params = params + start * plan->d;
    if (!in_dev) {
      DLB_CUDA(cudaMemcpyAsync(plan->w_params, d_params, (size_t)n * plan->d * sizeof(double), cudaMemcpyHostToDevice, stream));
      d_params = plan->w_params;
    }
    rc = launch_features(g, d_params, plan->d, n, npad, plan->w_F, plan->fstride, stream);
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
  return DLB_OK;
}

int dlb_last_kernel_ms(const dlb_plan* plan, double* features_ms, double* contract_ms, double* solve_ms) {
  DLB_REQUIRE(plan, "NULL plan");
  if (features_ms) *features_ms = plan->ms_features;
  if (contract_ms) *contract_ms = plan->ms_contract;
  if (solve_ms) *solve_ms = plan->ms_solve;
  return DLB_OK;
}

int dlb_measure_fp64_peak(int32_t kind, double* tflops) {
  DLB_REQUIRE(tflops, "NULL output");
  return measure_fp64_peak(kind, tflops);
}

}  // extern "C"

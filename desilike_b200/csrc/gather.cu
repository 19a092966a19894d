// Multi-GPU result exchange over NVLink peer memory (one process per GPU).
//
// The path shards by parameter point (reference desilike/base.py:310-316 scatters the points, mpi.py:145-150 fixes the partition); its only
// exchange is returning the per-point results to every rank (reference base.py:335 pickled reduce + bcast; samplers/base.py:196-200).
// Here every rank owns a *window* in its HBM -- `nbuf` result buffers + one flag word per (buffer, rank) -- exported with CUDA IPC and
// mapped by every peer.  The finish kernels of an evaluation (csrc/solve.cu: store_logpost / store_xsolved / gather_tail) store their
// results straight into every rank's window while they compute (plain st.global over NVLink / NVSwitch: no staging copy, no collective
// kernel competing for SMs with the persistent contraction kernel), and the last thread block of the launch publishes the evaluation's
// sequence number to every rank's flag word with a system-scope release.  A consumer enqueues dlb_gather_wait: one small thread block that
// acquires all ranks' flags (bounded spin; a rank that never arrives is reported, the GPU is not hung).
#include "common.h"

#include <cstring>

struct dlb_gather {
  int device = 0, rank = 0, world = 1, nbuf = 2;
  long long capacity = 0;                 // doubles per buffer
  size_t bytes = 0;
  unsigned char* window = nullptr;        // this rank's window (cudaMalloc)
  unsigned char* peer[dlb::DLB_MAX_RANKS] = {};   // every rank's window as mapped here (peer[rank] == window)
  bool opened[dlb::DLB_MAX_RANKS] = {};
  bool connected = false;
  unsigned* counter = nullptr;            // device: finished thread blocks of the publishing launch
  int* h_err = nullptr; int* d_err = nullptr;   // mapped pinned: first rank that timed out in a wait (1-based), 0 = none
  unsigned long long seq[8] = {};         // evaluations published into each buffer so far
  double timeout_s = 20.;
  cudaStream_t s_push = nullptr;          // side stream of the DMA pushes
  cudaEvent_t ev_done = nullptr, ev_push[8] = {};
  bool pushed[8] = {};
};

namespace dlb {

size_t gather_flag_offset(const dlb_gather* g) { return ((size_t)g->nbuf * g->capacity * sizeof(double) + 255) / 256 * 256; }

double* gather_peer_buffer(const dlb_gather* g, int r, int buf) { return reinterpret_cast<double*>(g->peer[r]) + (size_t)buf * g->capacity; }

unsigned long long* gather_peer_flag(const dlb_gather* g, int r, int buf, int from) {
  return reinterpret_cast<unsigned long long*>(g->peer[r] + gather_flag_offset(g)) + (size_t)buf * DLB_MAX_RANKS + from;
}

// Fill the device-side descriptor for an evaluation of `n` points starting at point `start` of a gathered batch of `total` points into buffer
// `buf`: logposterior [total], then (with_solved) x_solved [total, ns].
int gather_fill(dlb_gather* g, int buf, long long start, long long n, long long total, int ns, bool with_solved, bool signal, GatherDev* out) {
  std::memset(out, 0, sizeof(*out));
  if (!g) return DLB_OK;
  DLB_REQUIRE(g->connected, "gather: dlb_gather_connect has not been called");
  DLB_REQUIRE(buf >= 0 && buf < g->nbuf, "gather: buffer index out of range");
  const long long need = total * (with_solved ? 1 + ns : 1);
  DLB_REQUIRE(start >= 0 && n >= 0 && start + n <= total && need <= g->capacity, "gather: slice outside the batch or batch larger than the window capacity");
  out->world = g->world; out->signal = signal ? 1 : 0; out->with_solved = (with_solved && ns > 0) ? 1 : 0; out->xs_only = -1;
  for (int r = 0; r < g->world; ++r) {
    double* base = gather_peer_buffer(g, r, buf);
    out->lp[r] = base + start;
    out->xs[r] = out->with_solved ? base + total + start * ns : nullptr;
    out->flag[r] = gather_peer_flag(g, r, buf, g->rank);
  }
  out->seq = g->seq[buf];
  out->counter = g->counter;
  return DLB_OK;
}

void gather_next_sequence(dlb_gather* g, int buf) { g->seq[buf] += 1; }

int gather_rank(const dlb_gather* g) { return g->rank; }

// DMA pushes pay ~2 (world - 1) copy launches per evaluation: worthwhile for large x_solved payloads only (default threshold 8 MB; DLB_GATHER_DMA=0
// disables, =1 forces)
bool gather_use_dma(const dlb_gather* g, long long n, int ns, bool with_solved) {
  if (!g || g->world <= 1 || !with_solved || ns <= 0 || n <= 0) return false;
  static const int mode = getenv("DLB_GATHER_DMA") ? atoi(getenv("DLB_GATHER_DMA")) : -1;
  if (mode == 0) return false;
  if (mode == 1) return true;
  return (size_t)n * ns * sizeof(double) >= ((size_t)8 << 20);
}

int gather_before_write(dlb_gather* g, int buf, cudaStream_t stream) {
  if (g->pushed[buf]) DLB_CUDA(cudaStreamWaitEvent(stream, g->ev_push[buf], 0));   // the previous push out of this buffer has read the own slice
  return DLB_OK;
}

namespace {

__global__ void gather_publish_kernel(const double* __restrict__ src, long long count, GatherDev gd) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const double v = src[i];
    for (int r = 0; r < gd.world; ++r) gd.lp[r][i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(gd.counter, 1u);
    if (prev == gridDim.x - 1) {
      *gd.counter = 0u;
      __threadfence_system();
      for (int r = 0; r < gd.world; ++r) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(gd.flag[r]), "l"(gd.seq) : "memory");
    }
  }
}

// sequence number to every rank's flag word (behind DMA copies on the same stream; its own kernel: the block counter of gather_publish_kernel / the
// finish kernels belongs to the launches of the main stream, which may run concurrently)
__global__ void gather_flag_kernel(GatherDev gd) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < gd.world; ++r) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(gd.flag[r]), "l"(gd.seq) : "memory");
  }
}

__global__ void gather_wait_kernel(const unsigned long long* flags, int world, unsigned long long seq, int* err, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r >= world) return;
  unsigned long long t0, t1, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
    if (v >= seq) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) { atomicCAS(err, 0, r + 1); break; }
    __nanosleep(200);
  }
}

}  // namespace

int gather_publish_empty(dlb_gather* g, int buf, long long total, cudaStream_t stream) {
  GatherDev gd;
  DLB_TRY(gather_fill(g, buf, 0, 0, total, 0, false, true, &gd));
  gather_publish_kernel<<<1, 32, 0, stream>>>(nullptr, 0, gd);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

int gather_push_xs(dlb_gather* g, int buf, long long start, long long n, long long total, int ns, cudaStream_t stream) {
  if (!g->s_push) {
    DLB_CUDA(cudaStreamCreateWithFlags(&g->s_push, cudaStreamNonBlocking));
    DLB_CUDA(cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
    for (int b = 0; b < g->nbuf; ++b) DLB_CUDA(cudaEventCreateWithFlags(&g->ev_push[b], cudaEventDisableTiming));
  }
  DLB_CUDA(cudaEventRecord(g->ev_done, stream));            // the finish kernels of this evaluation have filled the own window
  DLB_CUDA(cudaStreamWaitEvent(g->s_push, g->ev_done, 0));
  const size_t off = (size_t)total + (size_t)start * ns, bytes = (size_t)n * ns * sizeof(double);
  const double* src = gather_peer_buffer(g, g->rank, buf) + off;
  for (int r = 0; r < g->world; ++r)
    if (r != g->rank) DLB_CUDA(cudaMemcpyAsync(gather_peer_buffer(g, r, buf) + off, src, bytes, cudaMemcpyDeviceToDevice, g->s_push));
  GatherDev gd;
  DLB_TRY(gather_fill(g, buf, 0, 0, total, 0, false, true, &gd));
  gather_flag_kernel<<<1, 32, 0, g->s_push>>>(gd);   // behind the copies: sequence number to every rank's flag word
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  DLB_CUDA(cudaEventRecord(g->ev_push[buf], g->s_push));
  g->pushed[buf] = true;
  return DLB_OK;
}

}  // namespace dlb

using namespace dlb;

extern "C" {

int dlb_gather_create(int32_t rank, int32_t world, int64_t capacity, int32_t nbuf, dlb_gather** out) {
  DLB_REQUIRE(out, "NULL output");
  *out = nullptr;
  DLB_REQUIRE(world >= 1 && world <= DLB_MAX_RANKS && rank >= 0 && rank < world, "gather: 1 <= world <= 16 and 0 <= rank < world required");
  DLB_REQUIRE(capacity > 0 && nbuf >= 1 && nbuf <= 8, "gather: capacity > 0 and 1 <= nbuf <= 8 required");
  dlb_gather* g = new dlb_gather();
  struct Guard { dlb_gather* p; ~Guard() { if (p) dlb_gather_destroy(p); } } guard{g};
  DLB_CUDA(cudaGetDevice(&g->device));
  g->rank = rank; g->world = world; g->nbuf = nbuf; g->capacity = capacity;
  g->bytes = gather_flag_offset(g) + (size_t)nbuf * DLB_MAX_RANKS * sizeof(unsigned long long);
  DLB_CUDA(cudaMalloc((void**)&g->window, g->bytes));
  DLB_CUDA(cudaMemset(g->window, 0, g->bytes));
  DLB_CUDA(cudaMalloc((void**)&g->counter, sizeof(unsigned)));
  DLB_CUDA(cudaMemset(g->counter, 0, sizeof(unsigned)));
  DLB_CUDA(cudaHostAlloc((void**)&g->h_err, sizeof(int), cudaHostAllocMapped));
  *g->h_err = 0;
  DLB_CUDA(cudaHostGetDevicePointer((void**)&g->d_err, g->h_err, 0));
  DLB_CUDA(cudaDeviceSynchronize());
  g->peer[rank] = g->window;
  if (world == 1) g->connected = true;
  guard.p = nullptr;
  *out = g;
  return DLB_OK;
}

int dlb_gather_ipc_handle(const dlb_gather* g, void* handle) {
  DLB_REQUIRE(g && handle, "NULL gather / handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == DLB_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  DLB_CUDA(cudaIpcGetMemHandle(&h, g->window));
  std::memcpy(handle, &h, sizeof(h));
  return DLB_OK;
}

int dlb_gather_connect(dlb_gather* g, const void* handles) {
  DLB_REQUIRE(g && handles, "NULL gather / handles");
  DLB_CUDA(cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank || g->opened[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * DLB_IPC_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    DLB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer[r] = static_cast<unsigned char*>(ptr);
    g->opened[r] = true;
  }
  g->connected = true;
  return DLB_OK;
}

double* dlb_gather_buffer(dlb_gather* g, int32_t buf) {
  if (!g || buf < 0 || buf >= g->nbuf) return nullptr;
  return gather_peer_buffer(g, g->rank, buf);
}

int dlb_gather_publish(dlb_gather* g, int32_t buf, const double* src, int64_t count, int64_t offset, void* stream_) {
  DLB_REQUIRE(g && (src || count == 0) && count >= 0 && offset >= 0, "gather publish: NULL gather / src or negative count / offset");
  DLB_REQUIRE(buf >= 0 && buf < g->nbuf && offset + count <= g->capacity, "gather publish: buffer index / range outside the window");
  DLB_CUDA(cudaSetDevice(g->device));
  // a DMA push of an evaluation into this buffer publishes its (older) sequence number from the side stream: this publish must come after it, so that
  // the flags stay monotonic and a wait for the newer number also covers the pushed data
  DLB_TRY(gather_before_write(g, buf, (cudaStream_t)stream_));
  g->seq[buf] += 1;
  GatherDev gd;
  DLB_TRY(gather_fill(g, buf, offset, count, offset + count, 0, false, true, &gd));
  const int threads = 256;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((count + threads - 1) / threads, 8 * 148));
  gather_publish_kernel<<<blocks, threads, 0, (cudaStream_t)stream_>>>(src, count, gd);
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

int dlb_gather_wait(dlb_gather* g, int32_t buf, void* stream_) {
  DLB_REQUIRE(g && buf >= 0 && buf < g->nbuf, "gather wait: NULL gather or buffer index out of range");
  DLB_CUDA(cudaSetDevice(g->device));
  if (*g->h_err) { set_error("gather: rank " + std::to_string(*g->h_err - 1) + " did not publish its results within the timeout"); return DLB_ERR_CUDA; }
  gather_wait_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(gather_peer_flag(g, g->rank, buf, 0), g->world, g->seq[buf], g->d_err,
                                                          (unsigned long long)(g->timeout_s * 1e9));
  g_launches++;
  DLB_CUDA(cudaGetLastError());
  return DLB_OK;
}

int dlb_gather_status(const dlb_gather* g) {
  if (!g) return DLB_ERR_ARG;
  return *g->h_err;
}

int dlb_gather_set_timeout(dlb_gather* g, double seconds) {
  DLB_REQUIRE(g && seconds > 0., "gather: positive timeout required");
  g->timeout_s = seconds;
  return DLB_OK;
}

void dlb_gather_destroy(dlb_gather* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < g->world; ++r)
    if (g->opened[r] && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
  if (g->window) cudaFree(g->window);
  if (g->counter) cudaFree(g->counter);
  if (g->h_err) cudaFreeHost(g->h_err);
  if (g->s_push) cudaStreamDestroy(g->s_push);
  if (g->ev_done) cudaEventDestroy(g->ev_done);
  for (auto& e : g->ev_push) if (e) cudaEventDestroy(e);
  delete g;
}

}  // extern "C"

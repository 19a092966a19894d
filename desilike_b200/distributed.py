"""Multi-GPU sharding of the batch by parameter point (one process per GPU, ``torch.distributed`` for the plumbing).

The reference's only parallelism on this path is data-parallel over points: ``vmap(backend='mpi')`` scatters chunks with
``mpi.scatter`` (desilike/base.py:310-316, partition rule ``mpi.local_size`` desilike/mpi.py:145-150), each rank evaluates its
slice, and results come back by a pickled ``reduce`` + ``bcast`` (base.py:335; samplers/base.py:196-200).  Here: contiguous
slices by the same rule, each rank runs its slice through its own device-resident plan, and the results reach every rank
through NVLink peer memory: the finish kernel of the evaluation stores logposterior / x_solved straight into every rank's
*gather window* (``PeerGather`` -> ``dlb_eval_batch_gather``, csrc/gather.cu) -- no collective kernel, no host hop.  When CUDA IPC
is unavailable the same API falls back to one NCCL all-gather; gloo (CPU) serves the host-logic tests.
No exchange happens inside the evaluation.
"""
import ctypes

import numpy as np

from . import _capi


def local_size(size, rank, world):
    """Number of points of rank ``rank`` (reference desilike/mpi.py:145-150: the first ``size % world`` ranks take one more)."""
    return size // world + (1 if rank < size % world else 0)


def local_slice(size, rank, world):
    start = sum(local_size(size, r, world) for r in range(rank))
    return slice(start, start + local_size(size, rank, world))


class _DevicePointer(object):
    """``__cuda_array_interface__`` holder: lets torch wrap a window buffer owned by the library without copying."""

    def __init__(self, ptr, count, owner):
        self.__cuda_array_interface__ = dict(shape=(int(count),), typestr='<f8', data=(int(ptr), False), version=2, strides=None)
        self._owner = owner


class PeerGather(object):
    """
    ``nbuf`` gathered-result buffers of ``capacity`` doubles in every rank's HBM, mapped by every peer through CUDA IPC.

    ``Plan.eval_gather`` / ``publish`` store this rank's slice into every rank's copy of a buffer and publish a sequence number;
    ``wait`` enqueues the acquire of all ranks' sequence numbers on the current stream.  All ranks must issue the same sequence of
    publishing calls per buffer.  ``comm=None``: single process (world 1).
    """

    def __init__(self, comm, capacity, nbuf=2):
        import torch
        self._lib = _capi.load()
        self.rank, self.world = (comm.rank, comm.size) if comm is not None else (0, 1)
        self.capacity, self.nbuf = int(capacity), int(nbuf)
        handle = ctypes.c_void_p()
        _capi.check(self._lib.dlb_gather_create(self.rank, self.world, self.capacity, self.nbuf, ctypes.byref(handle)))
        self._handle = handle
        self.device = torch.device('cuda', torch.cuda.current_device())
        if self.world > 1:
            mine = ctypes.create_string_buffer(64)
            _capi.check(self._lib.dlb_gather_ipc_handle(self._handle, mine))
            handles = [None] * self.world
            comm.dist.all_gather_object(handles, bytes(mine.raw), group=comm.group)
            rc = self._lib.dlb_gather_connect(self._handle, b''.join(handles))
            flags = [None] * self.world   # every rank learns whether every rank could map its peers (collective decision)
            comm.dist.all_gather_object(flags, int(rc), group=comm.group)
            if any(flags):
                msg = self._lib.dlb_last_error().decode() if rc else 'a peer rank could not map the windows'
                self.close()
                raise _capi.DesilikeB200Error('peer gather unavailable: ' + msg)

    def buffer(self, buf, count=None):
        """This rank's copy of buffer ``buf`` as a float64 CUDA tensor (a view: it is overwritten when the buffer is published into again)."""
        import torch
        ptr = self._lib.dlb_gather_buffer(self._handle, int(buf))
        return torch.as_tensor(_DevicePointer(ptr, self.capacity if count is None else count, self), device=self.device)

    def publish(self, buf, src, offset):
        """Store the contiguous float64 CUDA tensor ``src`` at ``offset`` of buffer ``buf`` in every rank's window."""
        import torch
        if not (src.is_cuda and src.dtype == torch.float64 and src.is_contiguous()):
            raise ValueError('publish expects a contiguous float64 CUDA tensor')
        stream = torch.cuda.current_stream(src.device).cuda_stream
        _capi.check(self._lib.dlb_gather_publish(self._handle, int(buf), src.data_ptr(), src.numel(), int(offset), ctypes.c_void_p(stream)))

    def wait(self, buf):
        import torch
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _capi.check(self._lib.dlb_gather_wait(self._handle, int(buf), ctypes.c_void_p(stream)))

    def check(self):
        """Raise if a wait timed out on some rank (call after a synchronize)."""
        st = self._lib.dlb_gather_status(self._handle)
        if st:
            raise _capi.DesilikeB200Error('peer gather: rank {:d} did not publish its results within the timeout'.format(st - 1))

    def close(self):
        if getattr(self, '_handle', None):
            self._lib.dlb_gather_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try: self.close()
        except Exception: pass


class Communicator(object):
    """Thin wrapper over an initialised ``torch.distributed`` process group."""

    def __init__(self, group=None, peer_gather=True):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised (launch with torchrun, then dist.init_process_group)')
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self._want_peer = bool(peer_gather) and self.backend == 'nccl'
        self._gather, self._turn = None, 0
        self.transport = None   # 'peer-store' | 'nccl' | 'gloo', set by the first sharded evaluation

    # ---- transports --------------------------------------------------------------------------------------------------
    def peer_gather(self, capacity):
        """The (cached, grown on demand) gather window, or None when CUDA IPC peer mapping is unavailable."""
        if not self._want_peer:
            return None
        if self._gather is None or self._gather.capacity < capacity:
            if self._gather is not None:
                import torch
                torch.cuda.synchronize()
                self.dist.barrier(group=self.group)   # no peer may still be storing into the window that is about to be freed
                self._gather.close()
                self._gather = None
            try:
                self._gather = PeerGather(self, capacity)
            except _capi.DesilikeB200Error:
                self._want_peer = False
                return None
        return self._gather

    def allgather_ragged(self, local, total, device=None):
        """All-gather 1-D / 2-D chunks of unequal length (padded to the largest chunk): the library-collective transport."""
        import torch
        local = torch.as_tensor(local)
        trailing = tuple(local.shape[1:])
        nmax = local_size(total, 0, self.size)
        buf = torch.zeros((nmax,) + trailing, dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
        out = torch.empty((self.size * nmax,) + trailing, dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.reshape((self.size, nmax) + trailing)
        return torch.cat([out[r, :local_size(total, r, self.size)] for r in range(self.size)], dim=0)

    # ---- the sharded batch entry -------------------------------------------------------------------------------------
    def eval_sharded(self, plan, X, evaluate=None, return_solved=True, return_all=True):
        """
        Every rank passes the FULL batch ``X[N, d]`` (NumPy array or float64 CUDA tensor); every rank gets the full result dict.

        NumPy in -> NumPy out; CUDA tensor in -> CUDA tensors out (views of the gather window when the peer-store transport is used:
        valid until the second-next sharded evaluation).  ``return_all=False`` gathers only ``logposterior`` (and ``x_solved``), which is
        what samplers need (reference samplers/base.py:196-200); otherwise ``loglikelihood``, ``logprior`` and ``status`` follow too.
        ``evaluate``: replacement for ``plan.eval`` on host arrays (CPU tests of the host logic).
        """
        import torch
        is_cuda = hasattr(X, 'is_cuda') and X.is_cuda
        N = int(X.shape[0])
        sl = local_slice(N, self.rank, self.size)
        if evaluate is not None or self.backend != 'nccl':
            return self._eval_sharded_host(plan, X, sl, N, evaluate)
        dev = torch.device('cuda', torch.cuda.current_device())
        Xl = (X[sl] if is_cuda else torch.as_tensor(np.ascontiguousarray(X[sl], dtype=np.float64), device=dev)).contiguous()
        ns = plan.ns if return_solved else 0
        nl = Xl.shape[0]
        gather = self.peer_gather(N * (1 + ns + (3 if return_all else 0)))
        out = {}
        if gather is not None:
            self.transport = 'peer-store'
            buf = self._turn % gather.nbuf
            self._turn += 1
            extra = {}
            if return_all:
                extra = dict(loglikelihood=torch.empty(nl, dtype=torch.float64, device=dev), logprior=torch.empty(nl, dtype=torch.float64, device=dev),
                             status=torch.empty(nl, dtype=torch.int32, device=dev))
            plan.eval_gather(Xl, gather, buf, sl.start, N, with_solved=ns > 0, out=extra)
            off = N * (1 + ns)
            for i, key in enumerate(['loglikelihood', 'logprior', 'status'] if return_all else []):
                gather.publish(buf, extra[key].to(torch.float64) if key == 'status' else extra[key], off + i * N + sl.start)
            gather.wait(buf)
            window = gather.buffer(buf)
            out['logposterior'] = window[:N]
            if plan.ns and return_solved:
                xs = window[N:N + N * ns].reshape(N, ns)
                out['x_solved'] = xs if plan._identity_perm else xs[:, torch.as_tensor(plan._unperm, device=dev)]
            if return_all:
                out['loglikelihood'], out['logprior'] = window[off:off + N], window[off + N:off + 2 * N]
                out['status'] = window[off + 2 * N:off + 3 * N].to(torch.int32)
        else:
            self.transport = 'nccl'
            res = plan.eval(Xl, return_solved=return_solved)
            for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
                if key not in res or (not return_all and key not in ('logposterior', 'x_solved')): continue
                out[key] = self.allgather_ragged(res[key], N)
        if plan.ns == 0 and return_solved:
            out['x_solved'] = torch.zeros((N, 0), dtype=torch.float64, device=dev)
        if not is_cuda:
            torch.cuda.synchronize()
            if gather is not None: gather.check()
            out = {key: value.cpu().numpy() for key, value in out.items()}
        return out

    def _eval_sharded_host(self, plan, X, sl, N, evaluate):
        import torch
        X = np.ascontiguousarray(X.cpu().numpy() if hasattr(X, 'cpu') else X, dtype=np.float64)
        res = (evaluate or plan.eval)(X[sl])
        device = torch.device('cuda', torch.cuda.current_device()) if self.backend == 'nccl' else torch.device('cpu')
        self.transport = 'nccl' if self.backend == 'nccl' else 'gloo'
        out = {}
        for key in ['logposterior', 'loglikelihood', 'logprior', 'x_solved', 'status']:
            if key not in res: continue
            local = torch.as_tensor(np.asarray(res[key]), device=device)
            if local.ndim == 2 and local.shape[1] == 0:
                out[key] = np.zeros((N, 0)); continue
            out[key] = self.allgather_ragged(local, N).cpu().numpy()
        return out

    def close(self):
        if self._gather is not None:
            self._gather.close()
            self._gather = None

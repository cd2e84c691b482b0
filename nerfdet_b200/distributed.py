"""View-sharded lifting across the GPUs of one box.

Rank r owns a contiguous slice of the views (the backbone would have produced them
there).  Every rank accumulates per-voxel ``S1 = sum f``, ``S2 = sum f^2`` and the valid-view
count over ITS views into one flat fp32 buffer, the buffers are summed with ONE
all-reduce (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests), and every
rank finalises mean / all-view variance with the GLOBAL number of views -- the
reference's variance runs over all views, invalid ones included (nerfdet.py:179), so
``n_views_total`` is part of the formula, not just a normaliser.

The reference has no collective on this path (it lifts one scene per GPU); this is the
multi-GPU extension BASELINE.json's north_star asks for.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def view_shard(n_views: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the views owned by ``rank``: contiguous, sizes differ by at most 1,
    earlier ranks take the remainder."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, rem = divmod(n_views, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _cuda_accumulate(features, points, projection):
    from . import ops
    return ops.lift_accumulate_planned(features, points, projection)


def _cuda_finalize(acc, n_views_total, channels, n_voxels, alpha, want_cov):
    from . import ops
    return ops.direct.lift_finalize(acc, n_views_total, channels, n_voxels, alpha, want_cov)


def lift_mean_var_view_sharded(features_local: torch.Tensor, points: torch.Tensor,
                               projection_local: torch.Tensor, n_views_total: Optional[int] = None,
                               group=None, alpha: Optional[torch.Tensor] = None, want_cov: bool = True,
                               accumulate_fn: Callable = _cuda_accumulate,
                               finalize_fn: Callable = _cuda_finalize):
    """Each rank passes ITS views (``features_local [nv_r, C, H, W]``, ``projection_local
    [nv_r, 3, 4]``); returns the same ``(volume_mean [C,X,Y,Z], volume_cov, valid [1,X,Y,Z])``
    on every rank as the single-GPU ``lifting.lift_mean_var`` over all views would.

    ``accumulate_fn`` / ``finalize_fn`` default to the CUDA ops; the CPU (gloo) tests inject
    oracle-backed stand-ins to exercise the sharding / collective logic without a GPU."""
    c = features_local.shape[1]
    gx, gy, gz = points.shape[-3:]
    n = gx * gy * gz
    acc = accumulate_fn(features_local, points, projection_local)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if n_views_total is None:
        nv = torch.tensor([features_local.shape[0]], dtype=torch.int64, device=acc.device)
        if world > 1:
            dist.all_reduce(nv, group=group)
        n_views_total = int(nv.item())
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    mean, cov, count = finalize_fn(acc, n_views_total, c, n,
                                   alpha.reshape(-1) if alpha is not None else None, want_cov)
    return (mean.view(c, gx, gy, gz), cov.view(c, gx, gy, gz) if want_cov else None,
            count.view(1, gx, gy, gz))


def channel_shard(channels: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the channels rank ``rank`` reduces and finalises in the peer-memory exchange
    (same split as ``view_shard``; csrc/peer.cu computes it the same way)."""
    return view_shard(channels, rank, world_size)


class _DevicePtr:
    """``__cuda_array_interface__`` shim: lets torch address a raw device allocation without owning it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


_last_exchange = {}      # device index -> event of the most recent exchange issued by this process


class PeerLift:
    """View-sharded lift whose exchange step runs over NVLink peer memory instead of NCCL (csrc/peer.cu).

    Every rank owns one peer-mapped segment ``[flags | S1 S2 count | mean | cov]``.  A call accumulates this rank's
    views into the local segment (``nd_lift_accumulate``) and then launches ``nd_lift_finalize_peers``: the kernel
    loads this rank's channel slice of every rank's partial sums (P2P reads), finalises it with the global view count
    and stores the rows into every rank's mean / cov (P2P writes).  Compared with ``lift_mean_var_view_sharded`` the
    52.5 MB accumulators cross the links once instead of twice, no reduced accumulator is written back, the finalise
    pass is split over the ranks, and no NCCL kernel is launched on the data path.

    Exchange kernels wait inside the kernel for their peers.  When one process drives several PeerLift objects from
    different streams, each exchange is therefore made to wait (event) for the previous one issued on the device, so
    that exchanges EXECUTE in the order they were issued -- the same order on every rank -- and a waiting kernel can
    never hold the SMs that the kernel it waits for needs (the rule NCCL imposes on concurrent collectives).

    The returned tensors are views of the segment: they are valid until the next call on ANY rank reaches its
    exchange step, i.e. consume (or copy) them on the same stream before calling again.
    """

    def __init__(self, channels: int, n_voxels: int, device=None, group=None, want_cov: bool = True,
                 transport: str = 'ipc', overlap_sms: int = 0, timeout_ms: int = 0, _local_group=None):
        """``transport='ipc'``: cudaMalloc segments shared with CUDA IPC handles, per-peer P2P loads / stores.
        ``transport='multicast'``: symmetric-memory segments (``torch.distributed._symmetric_memory``, plumbing only)
        bound to an NVLS multicast object; the kernel then reduces in the NVSwitch (``multimem.ld_reduce``) and
        broadcasts its rows with ``multimem.st``.  Raises if the box has no multicast support.

        ``overlap_sms`` > 0 (scenes in flight on two streams, one PeerLift each): the exchange kernel runs as that many
        one-per-SM CTAs and the accumulate keeps off that many SMs, so that the exchange of scene i (bound by the links,
        not by the SMs) runs beside the accumulate of scene i + 1."""
        from . import _lib
        import ctypes
        self._lib = _lib
        self._ct = ctypes
        self.lib = _lib.load()
        self.channels, self.n_voxels, self.want_cov = int(channels), int(n_voxels), bool(want_cov)
        self.overlap_sms = max(int(overlap_sms), 0)
        self.timeout_ms = max(int(timeout_ms), 0)                  # 0 = the library default (4 s)
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else torch.zeros(1, dtype=torch.int32)
        self.device = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        self.group = group
        if _local_group is not None:
            self.rank, self.world = _local_group
        elif dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if self.world > _lib.ND_MAX_PEERS:
            raise ValueError(f'PeerLift supports up to {_lib.ND_MAX_PEERS} ranks on one box, got {self.world}')
        cn = self.channels * self.n_voxels
        al = lambda b: (b + 255) // 256 * 256
        self._n_acc = ((2 if self.want_cov else 1) * cn + self.n_voxels)    # [S1 | S2 | count], or [S1 | count] without cov
        self._off_acc = al(4 * _lib.ND_PEER_FLAG_WORDS)
        self._off_mean = self._off_acc + al(4 * self._n_acc)
        self._off_cov = self._off_mean + al(4 * cn)
        self._bytes = self._off_cov + (al(4 * cn) if self.want_cov else 0)
        self.epoch = 0
        self._ordered = _local_group is None     # in-process test ranks must run concurrently, see local_group
        self._peer_bases = None
        self._opened = []
        self._mc_base = 0
        self._symm = None
        if transport not in ('ipc', 'multicast'):
            raise ValueError(f"transport must be 'ipc' or 'multicast', got {transport!r}")
        self.transport = transport
        if transport == 'multicast':
            if _local_group is not None or not dist.is_initialized() or self.world < 2:
                raise RuntimeError('the multicast transport needs an initialised process group with at least 2 ranks')
            if self.n_voxels % 4 != 0 or not self.want_cov:
                raise RuntimeError('the multicast transport needs a voxel count that is a multiple of 4 and want_cov=True')
            import torch.distributed._symmetric_memory as symm_mem
            seg = symm_mem.empty(self._bytes, dtype=torch.uint8, device=self.device)
            seg.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm_mem.rendezvous(seg, group if group is not None else dist.group.WORLD)
            if not hdl.has_multicast_support or int(hdl.multicast_ptr) == 0:
                raise RuntimeError('symmetric memory on this box has no NVLS multicast support')
            self._symm = hdl
            self._base = int(seg.data_ptr())
            self._handle = None
            bases = [int(p) for p in hdl.buffer_ptrs]
            if bases[self.rank] != self._base:
                raise RuntimeError('symmetric-memory rendezvous returned a local pointer that is not the segment')
            self._mc_base = int(hdl.multicast_ptr)
        else:
            try:
                with torch.cuda.device(self.device):
                    base = ctypes.c_void_p()
                    handle = (ctypes.c_uint8 * 64)()
                    _lib.check(self.lib.nd_peer_alloc(self._bytes, ctypes.byref(base), handle), 'nd_peer_alloc')
                self._base = int(base.value)
                self._handle = bytes(handle)
                seg = torch.as_tensor(_DevicePtr(self._base, self._bytes), device=self.device)
            except Exception as e:
                if self.world == 1 or _local_group is not None or not dist.is_initialized():
                    raise
                # a rank that cannot allocate / export still joins the collectives of _connect, which then fails on
                # every rank together (its handle is None)
                self._base, self._handle = None, None
                self._alloc_error = e
                self._connect()
                raise AssertionError('unreachable: _connect raises when a handle is missing')
        self._seg = seg
        self.flags = seg[:4 * _lib.ND_PEER_FLAG_WORDS].view(torch.int32)
        self.acc = seg[self._off_acc:self._off_acc + 4 * self._n_acc].view(torch.float32)
        self.mean = seg[self._off_mean:self._off_mean + 4 * cn].view(torch.float32).view(self.channels, self.n_voxels)
        self.cov = (seg[self._off_cov:self._off_cov + 4 * cn].view(torch.float32).view(self.channels, self.n_voxels)
                    if self.want_cov else None)
        self.count = torch.empty((self.n_voxels,), dtype=torch.int64, device=self.device)
        if transport == 'multicast':
            self._set_peers(bases)
            dist.barrier(group=self.group)
        elif _local_group is None:
            self._connect()

    # -- wiring ---------------------------------------------------------------------------------
    def _connect(self):
        ctypes = self._ct
        if self.world == 1:
            self._set_peers([self._base])
            return
        infos = [None] * self.world
        dist.all_gather_object(infos, (self._handle, self._bytes), group=self.group)
        bases, failure = [], None
        try:
            missing = [g for g, (h, _) in enumerate(infos) if h is None]
            if missing:
                raise RuntimeError(f'ranks {missing} could not allocate or export their segment'
                                   + (f' ({self._alloc_error})' if self._handle is None else ''))
            with torch.cuda.device(self.device):
                for g, (h, nbytes) in enumerate(infos):
                    if nbytes != self._bytes:
                        raise RuntimeError(f'rank {g} built a segment of {nbytes} bytes, this rank {self._bytes}: shapes differ')
                    if g == self.rank:
                        bases.append(self._base)
                        continue
                    p = ctypes.c_void_p()
                    hb = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                    self._lib.check(self.lib.nd_peer_open(hb, ctypes.byref(p)), f'nd_peer_open (segment of rank {g})')
                    self._opened.append(int(p.value))
                    bases.append(int(p.value))
        except Exception as e:                  # keep the ranks in step: everybody learns about it in the reduction below
            failure = e
        # every rank reaches this collective whether or not its mappings succeeded; it is also the barrier that
        # guarantees every segment is zeroed and mapped before the first epoch
        ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            with torch.cuda.device(self.device):
                for p in self._opened:
                    self.lib.nd_peer_close(ctypes.c_void_p(p))
                self._opened = []
                dist.barrier(group=self.group)  # nobody frees a segment a peer still has mapped
                self.flags = self.acc = self.mean = self.cov = self._seg = None
                if self._base is not None:
                    self.lib.nd_peer_free(ctypes.c_void_p(self._base))
            self._base = None
            raise RuntimeError('PeerLift: peer-memory segments could not be mapped on every rank'
                               + (f' (this rank: {failure})' if failure is not None else ' (another rank failed)'))
        self._set_peers(bases)

    def _set_peers(self, bases):
        ctypes = self._ct
        arr = ctypes.c_void_p * len(bases)
        self._peer_bases = list(bases)
        self._p_flags = arr(*[b for b in bases])
        self._p_acc = arr(*[b + self._off_acc for b in bases])
        self._p_mean = arr(*[b + self._off_mean for b in bases])
        self._p_cov = arr(*[b + self._off_cov for b in bases])

    @classmethod
    def local_group(cls, world: int, channels: int, n_voxels: int, device=None, want_cov: bool = True,
                    overlap_sms: int = 0, timeout_ms: int = 0):
        """``world`` ranks inside ONE process on one device (their segments are addressed directly, no IPC): the
        single-GPU test of the multi-rank protocol -- run each rank's call on its own stream.  Small shapes only:
        the waiting CTAs of all ranks must fit the device together."""
        ranks = [cls(channels, n_voxels, device, want_cov=want_cov, overlap_sms=overlap_sms, timeout_ms=timeout_ms,
                     _local_group=(r, world))
                 for r in range(world)]
        bases = [r._base for r in ranks]
        for r in ranks:
            r._set_peers(bases)
        return ranks

    # -- the op ---------------------------------------------------------------------------------
    def exchange(self, n_views_total: int, alpha: Optional[torch.Tensor] = None, owner: int = -1):
        """Reduce + finalise the accumulators already in ``self.acc`` (epoch advances by one).  ``owner`` >= 0: only that
        rank receives the result (returns ``(None, None, count)`` elsewhere); every rank must pass the same value."""
        from . import ops
        ctypes = self._ct
        if alpha is not None:
            alpha = alpha.reshape(-1).contiguous()
        # the error word of the previous steps, copied to pinned memory behind each of them: no synchronisation here
        if int(self._err_host[0]) != 0:
            raise RuntimeError('PeerLift: a peer did not reach an earlier exchange step within the time-out; the results of '
                               'that step are invalid (NaN rows) and the segments must be rebuilt')
        self.epoch += 1
        mc = self._mc_base
        stream = torch.cuda.current_stream(self.device)
        prev = _last_exchange.get(self.device.index) if self._ordered else None
        if prev is not None:
            stream.wait_event(prev)
        self._lib.check(self.lib.nd_lift_finalize_peers(
            self._p_acc, self._p_mean, self._p_cov if self.want_cov else None, self._p_flags, self.world, self.rank,
            self.epoch, int(n_views_total), self.channels, self.n_voxels,
            ctypes.c_void_p(alpha.data_ptr()) if alpha is not None else None,
            ctypes.c_void_p(self.count.data_ptr()),
            ctypes.c_void_p(mc + self._off_acc) if mc else None, ctypes.c_void_p(mc + self._off_mean) if mc else None,
            ctypes.c_void_p(mc + self._off_cov) if mc and self.want_cov else None,
            self.overlap_sms, self.timeout_ms, int(owner), stream.cuda_stream), 'nd_lift_finalize_peers')
        with torch.cuda.stream(stream):
            w = 2 * self._lib.ND_MAX_PEERS + 1
            self._err_host.copy_(self.flags[w:w + 1], non_blocking=True)
        if self._ordered:
            ev = torch.cuda.Event()
            ev.record(stream)
            _last_exchange[self.device.index] = ev
        if owner >= 0 and owner != self.rank:
            return None, None, self.count
        return self.mean, (self.cov if self.want_cov else None), self.count

    def __call__(self, features_local: torch.Tensor, points: torch.Tensor, projection_local: torch.Tensor,
                 n_views_total: int, alpha: Optional[torch.Tensor] = None, owner: int = -1):
        """Same results on every rank as ``lifting.lift_mean_var`` over all views (see class docstring for lifetime).
        ``owner`` >= 0 (same value on every rank): the scene belongs to that rank -- a data-parallel detector runs the neck
        and heads of a scene on one GPU -- and only it receives mean / cov (``None`` elsewhere); the result rows then cross
        the links once instead of G - 1 times."""
        from . import ops
        if features_local.shape[1] != self.channels or points[0].numel() != self.n_voxels:
            raise ValueError('PeerLift was built for another shape')
        sm_limit = 0
        if self.overlap_sms:
            sm_limit = max(torch.cuda.get_device_properties(self.device).multi_processor_count - self.overlap_sms, 1)
        ops.lift_accumulate_planned(features_local, points, projection_local, self.acc, sm_limit, self.want_cov)
        mean, cov, count = self.exchange(n_views_total, alpha, owner)
        if mean is None:
            return None, None, None
        shape = tuple(points.shape[1:]) if points.dim() == 4 else (self.n_voxels,)
        return (mean.view(self.channels, *shape), cov.view(self.channels, *shape) if cov is not None else None,
                count.view(1, *shape))

    def check(self):
        """Host-synchronising health check: raises if a peer never arrived at some exchange step."""
        torch.cuda.synchronize(self.device)
        if int(self.flags[2 * self._lib.ND_MAX_PEERS + 1].item()) != 0:
            raise RuntimeError('PeerLift: a peer did not reach the exchange step within the time-out')

    def close(self):
        """Unmaps the peers' segments and frees the local one (collective when world > 1)."""
        if self._base is None:
            return
        multi = self.world > 1 and dist.is_initialized() and (bool(self._opened) or self._symm is not None)
        torch.cuda.synchronize(self.device)
        if multi:
            dist.barrier(group=self.group)      # nobody unmaps while a peer may still be inside an exchange
        with torch.cuda.device(self.device):
            for p in self._opened:
                self.lib.nd_peer_close(self._ct.c_void_p(p))
            self._opened = []
        if multi:
            dist.barrier(group=self.group)      # nobody frees a segment a peer still has mapped
        self.flags = self.acc = self.mean = self.cov = self._seg = None
        if self._symm is not None:
            self._symm = None                   # the symmetric-memory allocator owns that segment
        else:
            with torch.cuda.device(self.device):
                self.lib.nd_peer_free(self._ct.c_void_p(self._base))
        self._base = None

"""Sharding of one denoise step over the GPUs of a node: one process per GPU, `torch.distributed` for plumbing.

The reference has no model parallelism (Lightning DDP over prompts only, main.py:63; `predict` uses no collective).
What shards naturally: the two classifier-free-guidance halves are independent until the CFG combine
(PanoGenerator.py:253-262), and perspective views are independent everywhere except inside EPPA, where the panorama
queries attend to the keys/values of ALL views (models/pano/modules.py:44-48). Layout used here:

    world = batch_shards x view_shards,   rank -> (bs, vs) = divmod(rank, view_shards)
    rank owns CFG/batch elements [bs*b/B .. ) and views [vs*m/V .. ); the panorama branch runs once per batch shard
    (replicated over the view shards of that batch shard).

Collectives: (1) per EPPA block, ONE all-gather of the projected K|V of the local views inside the batch shard's
view group (skipped when view_shards == 1, e.g. N = 2 = pure CFG split); (2) per forward, one all-gather of the tiny
eps outputs over the world so every rank ends the step with identical full latents.

Transport. Default: `DeviceAllGather` — receive buffers shared between the processes of the node through CUDA IPC
(`torch` storage sharing), filled by the peers' `pf_allgather_views` kernels with plain stores over NVLink and
synchronised with flag words in the same mappings. The collective is an ordinary kernel on the compute stream, so the
whole step stays ONE CUDA graph. `PF_DEVICE_GATHER=0` selects NCCL (`torch.distributed.all_gather_into_tensor`), which is
issued from the host between graph segments (`GraphSegments`).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor


DEVICE_GATHER = __import__("os").environ.get("PF_DEVICE_GATHER", "1") != "0"


class TransportUnavailable(RuntimeError):
    """The IPC / peer-access set-up of the device-side all-gather failed on at least one rank of the group."""


class _Site:
    __slots__ = ("recv", "ctrl_ptr", "peer_data", "peer_flags", "base", "opened")


class _RawCuda:
    """Zero-copy torch view of a raw device allocation (through __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(ptr, False), version=3)


_TYPESTR = {torch.float32: "<f4", torch.float16: "<f2", torch.bfloat16: "<u2", torch.int32: "<i4", torch.uint8: "|u1"}


class DeviceAllGather:
    """All-gather inside `group` through IPC-mapped receive buffers and the `pf_allgather_views` kernel.

    A call site is identified by `key` (slot, index of the collective inside the step, shape): its buffer — cudaMalloc'ed
    by the library, [S slices | S flag words | 2 state words] — is created on first use with an eager, host-synchronising
    exchange of CUDA IPC handles, so the first step of every slot must run eagerly (the sampler's warm-up step does), and
    reused afterwards. `slot` separates the buffers of consecutive steps: a site is rewritten only after every peer has
    passed at least one later collective of the SAME slot sequence (see csrc/comm.cu), which the sampler guarantees by
    cycling through >= 2 slots."""

    def __init__(self, group, agree_group=None):
        self.group = group
        # the success of a site's set-up is agreed over `agree_group` (ViewParallel passes its whole group: all its ranks
        # create their n-th site at the same point of the step, each inside its own view group, and must switch transport
        # together or not at all)
        self.agree_group = agree_group if agree_group is not None else group
        self.S = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sites: dict = {}

    def _create(self, key, x: Tensor) -> _Site:
        import ctypes as C
        from . import _lib
        lib = _lib.lib()
        S = self.S
        nbytes = x.numel() * x.element_size()
        data_bytes = (S * nbytes + 255) // 256 * 256
        site = _Site()
        # Every rank runs the SAME sequence of host collectives whatever happens locally, then the group agrees (MIN over
        # a success flag) on whether the mappings exist everywhere; a failure anywhere makes every rank raise
        # TransportUnavailable, and ViewParallel falls back to NCCL for the rest of the run (all ranks together).
        err = None
        site.base, handle = 0, None
        try:
            if __import__("os").environ.get("PF_FORCE_IPC_FAIL", "0") != "0":
                raise RuntimeError("PF_FORCE_IPC_FAIL is set (test hook)")
            base = C.c_void_p()
            _lib.check(lib.pf_comm_alloc(C.c_longlong(data_bytes + 256), C.byref(base)))
            site.base = base.value
            hbuf = (C.c_ubyte * 64)()
            _lib.check(lib.pf_ipc_export(C.c_void_p(site.base), hbuf))
            handle = bytes(hbuf)
        except Exception as e:  # noqa: BLE001 - reported below, after the group has agreed
            err = e
        site.ctrl_ptr = site.base + data_bytes
        everyone = [None] * S
        dist.all_gather_object(everyone, handle, group=self.group)
        data_ptrs, flag_ptrs, site.opened = [], [], []
        if err is None and all(h is not None for h in everyone):
            try:
                for r, h in enumerate(everyone):
                    if r == self.rank:
                        peer = site.base
                    else:
                        out = C.c_void_p()
                        _lib.check(lib.pf_ipc_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(out)))
                        peer = out.value
                        site.opened.append(peer)
                    data_ptrs.append(peer)
                    flag_ptrs.append(peer + data_bytes)
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer could not export its receive buffer")
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=x.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.agree_group)
        if int(ok.item()) == 0:
            for ptr in site.opened:
                lib.pf_ipc_close(C.c_void_p(ptr))
            if site.base:
                lib.pf_comm_free(C.c_void_p(site.base))
            raise TransportUnavailable(f"device all-gather set-up failed on a rank of the group ({err})")
        site.peer_data = torch.tensor(data_ptrs, dtype=torch.int64, device=x.device)
        site.peer_flags = torch.tensor(flag_ptrs, dtype=torch.int64, device=x.device)
        if x.dtype not in _TYPESTR:
            raise TypeError(f"DeviceAllGather: unsupported dtype {x.dtype}")
        raw = torch.as_tensor(_RawCuda(site.base, (S * x.numel(),), _TYPESTR[x.dtype]), device=x.device)
        site.recv = (raw.view(torch.bfloat16) if x.dtype == torch.bfloat16 else raw).view(S, *x.shape)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # every rank's flag words are zeroed and mapped before anyone pushes
        self.sites[key] = site
        return site

    def all_gather(self, key, x: Tensor) -> Tensor:
        """x contiguous, identical shape on every rank -> [S, *x.shape] (rank order), valid in stream order."""
        import ctypes as C
        from . import _lib, ops
        assert x.is_contiguous()
        key = (key, tuple(x.shape), x.dtype)
        site = self.sites.get(key)
        if site is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("DeviceAllGather: a new call site cannot be created while a CUDA graph is being captured "
                                   "(run one eager step first)")
            site = self._create(key, x)
        nbytes = x.numel() * x.element_size()
        ctrl = site.ctrl_ptr
        ops._count(1)
        _lib.check(_lib.lib().pf_allgather_views(
            C.c_void_p(x.data_ptr()), C.c_longlong(nbytes), C.c_void_p(site.peer_data.data_ptr()),
            C.c_void_p(site.peer_flags.data_ptr()), C.c_void_p(ctrl), C.c_void_p(ctrl + 4 * self.S), self.rank, self.S,
            C.c_void_p(_lib.stream_ptr())))
        return site.recv


def pick_layout(world: int, b: int, m: int) -> tuple[int, int]:
    """(batch_shards, view_shards): split the CFG batch first (halves the panorama branch per rank), views next."""
    batch_shards = 2 if (world % 2 == 0 and b % 2 == 0) else 1
    view_shards = world // batch_shards
    if b % batch_shards or m % view_shards:
        raise ValueError(f"cannot shard b={b} x m={m} over {world} ranks as {batch_shards} x {view_shards}")
    return batch_shards, view_shards


class GraphSegments:
    """One denoise step as [graph, collective, graph, collective, ..., graph].

    NCCL collectives are not captured: every time the forward reaches one (`collective(fn)`), the CUDA graph under
    capture is closed, `fn` runs eagerly on the same stream and is remembered, and a new capture begins (same memory
    pool, so buffers keep their addresses). `replay()` walks the list. At the break points the panorama side stream
    has always been joined (EPPA and the output gather run on the main stream), so ending the capture is legal.
    """

    def __init__(self, pool=None):
        self.items = []  # CUDAGraph | callable
        self.pool = pool if pool is not None else torch.cuda.graph_pool_handle()
        self.stream = torch.cuda.Stream()
        self._cur = None

    def _begin(self):
        self._cur = torch.cuda.CUDAGraph()
        self._cur.capture_begin(pool=self.pool)

    def _end(self):
        self._cur.capture_end()
        self.items.append(self._cur)
        self._cur = None

    def capture(self, body) -> None:
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self._begin()
            try:
                body()
            except BaseException:
                # leave capture mode (otherwise the stream stays unusable) and drop the partial segments
                if self._cur is not None:
                    try:
                        self._cur.capture_end()
                    except Exception:
                        pass
                    self._cur = None
                self.items.clear()
                raise
            self._end()
        torch.cuda.current_stream().wait_stream(self.stream)

    def collective(self, fn):
        self._end()
        fn()
        self.items.append(fn)
        self._begin()

    def replay(self) -> None:
        for it in self.items:
            if isinstance(it, torch.cuda.CUDAGraph):
                it.replay()
            else:
                it()


class ViewParallel:
    def __init__(self, group=None, batch_shards: Optional[int] = None, view_shards: Optional[int] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.batch_shards, self.view_shards = batch_shards, view_shards
        self._view_groups = None
        self.segments: Optional[GraphSegments] = None  # set by the sampler while it captures a step (NCCL transport)
        self.device_gather = DEVICE_GATHER
        self._dev_view: Optional[DeviceAllGather] = None
        self._dev_world: Optional[DeviceAllGather] = None
        self._slot, self._site = 0, 0

    def _fall_back(self, why: Exception) -> None:
        """Every rank of the group raised together (DeviceAllGather._create agrees collectively): use NCCL from now on."""
        import warnings
        warnings.warn(f"panfusion_b200: {why}; the sharded step uses NCCL all-gathers between CUDA-graph segments instead",
                      stacklevel=3)
        self.device_gather = False

    def begin_step(self, slot: int = 0) -> None:
        """Called at the start of every forward: `slot` selects the set of receive buffers (see DeviceAllGather)."""
        self._slot, self._site = int(slot), 0

    def _next_key(self):
        self._site += 1
        return (self._slot, self._site)

    def _run(self, fn) -> None:
        if self.segments is not None:
            self.segments.collective(fn)
        else:
            fn()

    def configure(self, b: int, m: int) -> None:
        if self.batch_shards is None or self.view_shards is None:
            self.batch_shards, self.view_shards = pick_layout(self.world, b, m)
        if self.batch_shards * self.view_shards != self.world:
            raise ValueError("batch_shards * view_shards must equal the world size")
        if self._view_groups is None:
            # every rank must create every subgroup, in the same order
            self._view_groups = []
            for bs in range(self.batch_shards):
                ranks = [bs * self.view_shards + v for v in range(self.view_shards)]
                self._view_groups.append(dist.new_group(ranks) if self.view_shards > 1 else None)
        self.bs, self.vs = divmod(self.rank, self.view_shards)

    @property
    def view_group(self):
        return self._view_groups[self.bs]

    def slices(self, b: int, m: int) -> tuple[slice, slice]:
        bl, ml = b // self.batch_shards, m // self.view_shards
        return slice(self.bs * bl, (self.bs + 1) * bl), slice(self.vs * ml, (self.vs + 1) * ml)

    def gather_views(self, x: Tensor) -> Tensor:
        """x: [b_loc, L_loc, C] contiguous -> [b_loc, view_shards * L_loc, C] in view order."""
        if self.view_shards == 1:
            return x
        bl, L, C = x.shape
        if self.device_gather and x.is_cuda:  # CPU tensors (the gloo plumbing tests) take the torch.distributed path
            if self._dev_view is None:
                self._dev_view = DeviceAllGather(self.view_group, agree_group=self.group)
            try:
                out = self._dev_view.all_gather(self._next_key(), x.contiguous())
            except TransportUnavailable as e:
                self._fall_back(e)
                return self.gather_views(x)
        else:
            out = torch.empty((self.view_shards * bl, L, C), dtype=x.dtype, device=x.device)  # concat along dim 0
            xc, grp = x.contiguous(), self.view_group
            self._run(lambda: dist.all_gather_into_tensor(out, xc, group=grp))
        out = out.reshape(self.view_shards, bl, L, C)
        if bl == 1:
            return out.reshape(1, self.view_shards * L, C)
        return out.permute(1, 0, 2, 3).reshape(bl, self.view_shards * L, C)

    def gather_outputs(self, sample_loc: Optional[Tensor], pano_loc: Tensor, b: int, m: int):
        """sample_loc [b_loc, m_loc, ...], pano_loc [b_loc, 1, ...] -> full [b, m, ...], [b, 1, ...] on every rank."""
        bl, ml = b // self.batch_shards, m // self.view_shards
        pc, grp = pano_loc.contiguous(), self.group
        s_all, sc = None, None
        if self.device_gather and pc.is_cuda:
            if self._dev_world is None:
                self._dev_world = DeviceAllGather(self.group)
            try:
                pano_all = self._dev_world.all_gather(self._next_key(), pc)
                if sample_loc is not None:
                    sc = sample_loc.contiguous()
                    s_all = self._dev_world.all_gather(self._next_key(), sc)
            except TransportUnavailable as e:
                self._fall_back(e)
                return self.gather_outputs(sample_loc, pano_loc, b, m)
        else:
            pano_all = torch.empty((self.world * pano_loc.shape[0], *pano_loc.shape[1:]), dtype=pano_loc.dtype,
                                   device=pano_loc.device)
            if sample_loc is not None:
                sc = sample_loc.contiguous()
                s_all = torch.empty((self.world * sc.shape[0], *sc.shape[1:]), dtype=sc.dtype, device=sc.device)

            def both():
                dist.all_gather_into_tensor(pano_all, pc, group=grp)
                if s_all is not None:
                    dist.all_gather_into_tensor(s_all, sc, group=grp)

            self._run(both)
        pano = pano_all.reshape(self.batch_shards, self.view_shards, *pano_loc.shape)[:, 0].reshape(b, *pano_loc.shape[1:])
        sample = None
        if sample_loc is not None:
            tail = sample_loc.shape[2:]
            s = s_all.reshape(self.batch_shards, self.view_shards, bl, ml, *tail)
            sample = s.permute(0, 2, 1, 3, *range(4, 4 + len(tail))).reshape(b, m, *tail)
        return sample, pano

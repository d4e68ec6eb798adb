"""Node-sharded nearest-neighbour queries over the GPUs of one box (BASELINE config 4).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Node i of the logical tree lives on
shard i mod G at local slot i div G, so shards stay balanced under incremental appends and the reference's
"lowest index wins" tie rule survives: global slot = local * G + shard.  Every rank holds the same queries,
answers them on its slice with the ordinary CUDA kernels (gc_nn1_dev / gc_knn_dev / gc_radius_dev), the
per-shard lists are exchanged with ONE all-gather, and every rank merges them on its device
(gc_merge_shard_lists_dev) into the list a single store would have returned: ascending (dist, global slot).

The exchange step is the only collective; bytes per rank = nq * cap * 12 + nq * 4.

exchange="p2p": no collective at all -- every shard writes its lists into a buffer of torch symmetric memory
(peer-mapped over NVLink) and gc_merge_shard_lists_p2p_dev loads the G lists of a query straight from the G
GPUs inside the merge kernel; the only synchronisation is the symmetric-memory barrier (a few microseconds on
the device) before the merge and before the buffers are reused.

`engine` / `merge` are injectable so that the partition + gather + index logic can be exercised on CPU
with the gloo backend (tests/test_sharded_cpu.py); the product path uses CudaShardEngine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi

NO_NODE = 0xFFFFFFFF


class PeerBuffers:
    """Result buffers of one shard in torch symmetric memory: idx [nq_max][cap_max] u32 | dist f64 | count u32,
    peer-mapped on every rank of `group`."""

    def __init__(self, group, device, nq_max: int, cap_max: int):
        import torch.distributed._symmetric_memory as symm_mem
        self.nq_max, self.cap_max = nq_max, cap_max
        n = nq_max * cap_max
        self.off_dist = (4 * n + 255) // 256 * 256
        self.off_cnt = self.off_dist + 8 * n
        nbytes = self.off_cnt + 4 * nq_max + 256
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        g = group if group is not None else dist.group.WORLD
        self.hdl = symm_mem.rendezvous(self.buf, g)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def views(self, nq, cap):
        assert nq <= self.nq_max and cap <= self.cap_max and nq * cap <= self.nq_max * self.cap_max
        idx = self.buf[:4 * nq * cap].view(torch.int32).view(nq, cap)
        dd = self.buf[self.off_dist:self.off_dist + 8 * nq * cap].view(torch.float64).view(nq, cap)
        cnt = self.buf[self.off_cnt:self.off_cnt + 4 * nq].view(torch.int32)
        return idx, dd, cnt

    def barrier(self):
        self.hdl.barrier(channel=0)


class CudaShardEngine:
    """The local slice: a NodeStore driven through the device-pointer C ABI with torch-owned buffers."""

    def __init__(self, dim: int, capacity: int, device: int):
        self.dim = dim
        self.device = device
        self.torch_device = torch.device("cuda", device)
        self.store = capi.NodeStore(dim, capacity, 1, device)
        self.lib = capi.load()
        self.store.set_stream(torch.cuda.current_stream(self.torch_device).cuda_stream)
        self.peer = None  # PeerBuffers when the p2p exchange is enabled
        self._bufs = {}   # (nq, cap) -> result buffers; a result is valid until the next query of the same shape
        self._flags = None
        self._last = None  # packed buffer of the lists produced by the latest query

    def merge_gathered_packed(self, gathered: torch.Tensor, world: int, nq: int, cap: int, k_out: int, out_cap: int):
        """Merge after ONE all-gather of the packed result buffers: gathered is [world][nbytes] uint8, shard g's arrays
        sit at the packed offsets inside row g (gc_merge_shard_lists_p2p_dev takes one pointer per shard and array)."""
        _, off_idx, off_cnt, nbytes = self._last[:4]
        base = gathered.data_ptr()
        arr = C.c_void_p * world
        idx_p = arr(*[C.c_void_p(base + g * nbytes + off_idx) for g in range(world)])
        dist_p = arr(*[C.c_void_p(base + g * nbytes) for g in range(world)])
        cnt_p = arr(*[C.c_void_p(base + g * nbytes + off_cnt) for g in range(world)])
        key = ("out", nq, out_cap)
        if key not in self._bufs:
            self._bufs[key] = (torch.empty((nq, out_cap), dtype=torch.int32, device=self.torch_device),
                               torch.empty((nq, out_cap), dtype=torch.float64, device=self.torch_device),
                               torch.empty(nq, dtype=torch.int32, device=self.torch_device))
        oi, od, oc = self._bufs[key]
        stream = torch.cuda.current_stream(self.torch_device).cuda_stream
        capi._check(self.lib.gc_merge_shard_lists_p2p_dev(self.device, stream, world, nq, cap, idx_p, dist_p, cnt_p, k_out,
                                                          out_cap, oi.data_ptr(), od.data_ptr(), oc.data_ptr()))
        return oi, od, oc

    def enable_p2p(self, group, nq_max: int, cap_max: int):
        self.peer = PeerBuffers(group, self.torch_device, nq_max, cap_max)

    def merge_p2p(self, nq, cap, k_out, out_cap):
        """Merge the lists that every shard left in its PeerBuffers (same nq, cap everywhere)."""
        pb = self.peer
        G = len(pb.ptrs)
        arr = C.c_void_p * G
        idx_p = arr(*[C.c_void_p(p) for p in pb.ptrs])
        dist_p = arr(*[C.c_void_p(p + pb.off_dist) for p in pb.ptrs])
        cnt_p = arr(*[C.c_void_p(p + pb.off_cnt) for p in pb.ptrs])
        oi = torch.full((nq, out_cap), -1, dtype=torch.int32, device=self.torch_device)
        od = torch.full((nq, out_cap), float("inf"), dtype=torch.float64, device=self.torch_device)
        oc = torch.zeros(nq, dtype=torch.int32, device=self.torch_device)
        stream = torch.cuda.current_stream(self.torch_device).cuda_stream
        pb.barrier()  # every shard's lists are complete
        capi._check(self.lib.gc_merge_shard_lists_p2p_dev(self.device, stream, G, nq, cap, idx_p, dist_p, cnt_p, k_out,
                                                          out_cap, oi.data_ptr(), od.data_ptr(), oc.data_ptr()))
        pb.barrier()  # nobody overwrites its lists while a peer still reads them
        return oi, od, oc

    def append(self, rows: np.ndarray):
        if len(rows):
            self.store.append(rows)

    def _q(self, q) -> torch.Tensor:
        """Queries on the device: a float64 CUDA tensor is used as it is (resident queries), anything else is copied."""
        if isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float64:
            return q.contiguous().view(-1, self.dim)
        return torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dim)).to(self.torch_device)

    def _out(self, nq, cap, fill=True):
        """Result buffers (kept per shape: a query loop does not allocate).  fill=False: the kernel writes every entry
        the caller reads (k-nearest and nearest write counts / sentinels themselves)."""
        if self.peer is not None:
            idx, dd, cnt = self.peer.views(nq, cap)
        else:
            key = (nq, cap)
            if key not in self._bufs:
                if len(self._bufs) > 32:
                    self._bufs.clear()
                # ONE buffer per shape -- dist f64 | idx u32 | count u32 -- so that the NCCL exchange is a single
                # all-gather of it and the merge kernel reads every shard's three arrays through per-shard pointers
                n = nq * cap
                off_idx = 8 * n
                off_cnt = off_idx + 4 * n
                nbytes = (off_cnt + 4 * nq + 15) // 16 * 16
                buf = torch.empty(nbytes, dtype=torch.uint8, device=self.torch_device)
                self._bufs[key] = (buf, off_idx, off_cnt, nbytes,
                                   buf[off_idx:off_idx + 4 * n].view(torch.int32).view(nq, cap),
                                   buf[:8 * n].view(torch.float64).view(nq, cap),
                                   buf[off_cnt:off_cnt + 4 * nq].view(torch.int32))
            self._last = self._bufs[key]
            idx, dd, cnt = self._last[4:7]
        if fill:
            idx.fill_(-1)  # 0xFFFFFFFF = no node
            dd.fill_(float("inf"))
            cnt.zero_()
        return idx, dd, cnt

    def nearest_lists(self, q, want_counts=True):
        dq = self._q(q)
        nq = dq.shape[0]
        idx, dd, cnt = self._out(nq, 1, fill=False)  # gc_nn1_dev writes (NO_NODE, inf) for an empty tree
        capi._check(self.lib.gc_nn1_dev(self.store.handle, dq.data_ptr(), None, nq, idx.data_ptr(), dd.data_ptr()))
        if want_counts:
            torch.ne(idx[:, 0], -1, out=self._flag(nq))
            cnt.copy_(self._flag(nq))
        return idx, dd, cnt

    def _flag(self, nq):
        if self._flags is None or self._flags.shape[0] != nq:
            self._flags = torch.empty(nq, dtype=torch.bool, device=self.torch_device)
        return self._flags

    def knn_lists(self, q, k):
        dq = self._q(q)
        nq = dq.shape[0]
        idx, dd, cnt = self._out(nq, k, fill=False)  # entries [0, count) and the counts are written by the kernels
        capi._check(self.lib.gc_knn_dev(self.store.handle, dq.data_ptr(), None, nq, k, k, idx.data_ptr(), dd.data_ptr(),
                                        cnt.data_ptr()))
        return idx, dd, cnt

    def near_lists(self, q, radius, cap):
        dq = self._q(q)
        nq = dq.shape[0]
        r = torch.full((nq,), float(radius), dtype=torch.float64, device=self.torch_device)
        idx, dd, cnt = self._out(nq, cap)
        capi._check(self.lib.gc_radius_dev(self.store.handle, dq.data_ptr(), None, r.data_ptr(), nq, cap,
                                           idx.data_ptr(), dd.data_ptr(), cnt.data_ptr()))
        return idx, dd, cnt

    def merge(self, idx, dd, cnt, k_out, out_cap):
        G, nq, cap = idx.shape
        oi = torch.full((nq, out_cap), -1, dtype=torch.int32, device=self.torch_device)
        od = torch.full((nq, out_cap), float("inf"), dtype=torch.float64, device=self.torch_device)
        oc = torch.zeros(nq, dtype=torch.int32, device=self.torch_device)
        stream = torch.cuda.current_stream(self.torch_device).cuda_stream
        capi._check(self.lib.gc_merge_shard_lists_dev(self.device, stream, G, nq, cap, idx.data_ptr(), dd.data_ptr(),
                                                      cnt.data_ptr(), k_out, out_cap, oi.data_ptr(), od.data_ptr(),
                                                      oc.data_ptr()))
        return oi, od, oc


class ShardedNodeStore:
    """Logical tree of N nodes sharded round-robin over the ranks of `group`."""

    def __init__(self, dim: int, capacity_total: int, engine=None, group=None, device: int | None = None,
                 exchange: str = "nccl", nq_max: int = 4096, cap_max: int = 1024):
        self.group = group
        self.exchange = exchange
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dim = dim
        self.n_global = 0
        if engine is None:
            if device is None:
                device = torch.cuda.current_device()
            engine = CudaShardEngine(dim, (capacity_total + self.world - 1) // self.world + 1, device)
        self.engine = engine
        if exchange == "p2p" and self.world > 1:
            engine.enable_p2p(group, nq_max, cap_max)

    # ---- partition -------------------------------------------------------------------------------
    def owner_of(self, global_slot):
        return global_slot % self.world

    def append(self, rows):
        """Every rank passes the same rows; each keeps the ones it owns.  Returns the first global slot."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.dim)
        first = self.n_global
        g = np.arange(first, first + len(rows))
        self.engine.append(rows[g % self.world == self.rank])
        self.n_global += len(rows)
        return first

    # ---- exchange + merge --------------------------------------------------------------------------
    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t.unsqueeze(0)
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)  # rank-major concatenation along dim 0
        return out.view((self.world,) + tuple(t.shape))

    def _exchange(self, lists, k_out, out_cap):
        idx, dd, cnt = lists
        if self.world == 1 and idx.shape[1] == out_cap:
            # one shard: its ordered list IS the answer (local slot == global slot), nothing to exchange or merge
            return idx, dd, torch.clamp(cnt, max=k_out)
        if self.exchange == "p2p" and self.world > 1:
            return self.engine.merge_p2p(idx.shape[0], idx.shape[1], k_out, out_cap)
        if self.world > 1 and isinstance(self.engine, CudaShardEngine) and self.engine._last is not None and \
                self.engine._last[4].data_ptr() == idx.data_ptr():
            buf, nbytes = self.engine._last[0], self.engine._last[3]
            key = ("gather", nbytes)
            if key not in self.engine._bufs:
                self.engine._bufs[key] = torch.empty(self.world * nbytes, dtype=torch.uint8, device=buf.device)
            out = self.engine._bufs[key]
            dist.all_gather_into_tensor(out, buf, group=self.group)
            return self.engine.merge_gathered_packed(out, self.world, idx.shape[0], idx.shape[1], k_out, out_cap)
        return self.engine.merge(self._gather(idx), self._gather(dd), self._gather(cnt), k_out, out_cap)

    def nearest(self, q):
        """(global slot [nq], dist [nq]); NO_NODE / inf when the tree is empty."""
        if self.world == 1 and isinstance(self.engine, CudaShardEngine):
            oi, od, _ = self.engine.nearest_lists(q, want_counts=False)  # one shard: its answer is the answer
            return oi[:, 0], od[:, 0]
        oi, od, oc = self._exchange(self.engine.nearest_lists(q), 1, 1)
        return oi[:, 0], od[:, 0]

    def knn(self, q, k: int):
        """(global slots [nq,k], dists [nq,k], counts [nq]) ascending (dist, global slot)."""
        return self._exchange(self.engine.knn_lists(q, k), k, k)

    def near(self, q, radius: float, cap: int = 256):
        """Radius-near, complete: when some shard found more than `cap` neighbours of a query the per-shard lists are
        taken again with a capacity that holds them (all ranks agree on it through a MAX all-reduce), as
        CudaNearestNeighbors::near does for the unsharded store -- a truncated shard list never reaches the merge."""
        while True:
            lists = self.engine.near_lists(q, radius, cap)
            mx = lists[2].max().to(torch.int64) if lists[2].numel() else torch.zeros((), dtype=torch.int64,
                                                                                       device=lists[2].device)
            if self.world > 1:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            need = int(mx.item())
            if need <= cap:
                break
            while cap < need:
                cap *= 2
        return self._exchange(lists, self.world * cap, self.world * cap)

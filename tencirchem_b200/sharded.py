"""Alpha-row sharded UCC engine: one CI vector spread over `world` GPUs (the large-N regime).

Every rank holds the rows of the [na, nb] CI matrix whose occupation pattern on t = log2(world) alpha orbitals T
equals its rank.  T is chosen per batch among the orbitals that are not active on the alpha side of the batch
(include/tcc_b200.h, multi-GPU section), so all batch kernels run on local rows; only when T changes do rows move
between ranks (one all-to-all per vector).  H c has two modes: "replicated" works on row blocks of the gathered
vector (`tcc_apply_h_rows`), which needs the full vector on every GPU (fine up to (18e,18o), 18.9 GB);
"distributed" keeps the vector sharded (three alpha-beta passes in layouts with disjoint T plus a column transpose
for the alpha-alpha terms, `tcc_shard_sigma_*`), which is what the 273 GB (20e,20o) vector needs.

The same driver runs
* one rank per process over `torch.distributed` (NCCL; `TorchComm`), and
* all ranks inside one process on one GPU (`LocalComm`): the parity tests exercise every layout change, ownership
  map and partial-gradient sum without needing a multi-GPU box.

Mirrors get_energy_and_grad_civector / get_civector (tencirchem/static/evolve_civector.py:180-218).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .plan import Plan
from .parallel import shard_bounds


class PhaseTimer:
    """Per-phase device time (CUDA events on the current stream) and host time of the sharded driver.  Off by default;
    `bench.py` turns it on for `phases_rank0`.  The device time of a collective is the span the compute stream waited
    for it (torch's NCCL calls make the current stream wait for the communication stream)."""

    def __init__(self, torch_mod, enabled=False):
        self.t = torch_mod
        self.enabled = enabled
        self.reset()

    def reset(self):
        self._spans = []  # (name, ev0, ev1)
        self.host_s = {}
        self.bytes = {}

    class _Span:
        def __init__(self, owner, name, nbytes):
            self.o, self.name, self.nbytes = owner, name, nbytes

        def __enter__(self):
            o = self.o
            if o.enabled:
                import time

                self.ev0 = o.t.cuda.Event(enable_timing=True)
                self.ev0.record()
                self.t0 = time.perf_counter()
            return self

        def __exit__(self, *exc):
            o = self.o
            if o.enabled:
                import time

                ev1 = o.t.cuda.Event(enable_timing=True)
                ev1.record()
                o._spans.append((self.name, self.ev0, ev1))
                o.host_s[self.name] = o.host_s.get(self.name, 0.0) + time.perf_counter() - self.t0
                o.bytes[self.name] = o.bytes.get(self.name, 0) + int(self.nbytes)
            return False

    def phase(self, name, nbytes=0):
        return PhaseTimer._Span(self, name, nbytes)

    def summary(self):
        """{phase: {"device_ms", "host_ms", "calls", "bytes"}} accumulated since reset()"""
        if not self.enabled:
            return {}
        self.t.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._spans:
            d = out.setdefault(name, {"device_ms": 0.0, "host_ms": 0.0, "calls": 0, "bytes": 0})
            d["device_ms"] += e0.elapsed_time(e1)
            d["calls"] += 1
        for name, d in out.items():
            d["host_ms"] = 1e3 * self.host_s.get(name, 0.0)
            d["bytes"] = self.bytes.get(name, 0)
        return out


class LocalComm:
    """All ranks live in this process (virtual ranks on one device): collectives are local copies."""

    def __init__(self, world: int, peer_memory: bool = True):
        """peer_memory=False makes the row moves take the pack / exchange / unpack path of a transport without peer
        mapping (test coverage of that path on one GPU)"""
        self.world = int(world)
        self.local_ranks = list(range(self.world))
        self.peer_memory = bool(peer_memory)

    def exchange(self, send: List[List["torch.Tensor"]], recv_counts=None) -> List[List["torch.Tensor"]]:
        # send[i][d]: rows that local rank i sends to rank d  ->  recv[i][s]: rows rank i receives from rank s
        return [[send[s][d] for s in range(self.world)] for d in range(self.world)]

    def exchange_packed(self, send_bufs, send_counts, recv_counts):
        import torch

        parts = [list(torch.split(b, list(c))) for b, c in zip(send_bufs, send_counts)]
        return [torch.cat([parts[s][d] for s in range(self.world)]) for d in range(self.world)]

    def all_reduce_sum(self, tensors: List["torch.Tensor"]):
        total = tensors[0].clone()
        for t in tensors[1:]:
            total += t
        for t in tensors:
            t.copy_(total)

    def alloc_pool(self, n_doubles, device):
        """-> (one flat float64 buffer per local rank, per local rank a device tensor [world] with the base address of
        EVERY rank's buffer as seen from that rank, or None when peers are not addressable).  All virtual ranks share a
        device: every buffer is directly addressable."""
        import torch

        bufs = [torch.empty(n_doubles, dtype=torch.float64, device=device) for _ in range(self.world)]
        if not self.peer_memory:
            return bufs, [None for _ in range(self.world)]
        ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=device)
        return bufs, [ptrs for _ in range(self.world)]

    def stream_barrier(self):
        """all ranks' streams have reached this point (virtual ranks share one stream: nothing to do)"""

    def all_gather_padded(self, blocks, max_rows):
        """blocks[i]: [rows_i, nb] of local rank i -> per local rank a [world, max_rows, nb] tensor whose slab s holds
        the rows of rank s (rows beyond rows_s are undefined)"""
        import torch

        out = torch.empty((self.world, max_rows) + tuple(blocks[0].shape[1:]), dtype=blocks[0].dtype, device=blocks[0].device)
        for s_, b in enumerate(blocks):
            out[s_, : b.shape[0]] = b
        return [out for _ in blocks]

    def reduce_scatter_sum(self, stacked):
        """stacked[i]: [world, max_rows, nb] of local rank i (slab d = its contribution to rank d) -> per local rank the
        [max_rows, nb] sum over all ranks of their slab for it"""
        outs = []
        for d in range(self.world):
            acc = stacked[0][d].clone()
            for s_ in range(1, self.world):
                acc += stacked[s_][d]
            outs.append(acc)
        return outs


class TorchComm:
    """One rank per process over torch.distributed (NCCL on GPUs; gloo in the CPU tests)."""

    def __init__(self, peer_memory="auto"):
        """peer_memory: "auto" maps the shard pool of every rank into every process (CUDA symmetric memory over NVLink /
        NVSwitch) when the backend is NCCL and the mapping succeeds, so that rows move by direct peer stores
        (tcc_push_rows); "off" (and gloo, and a failed mapping) moves rows through NCCL all-to-all."""
        import torch.distributed as dist

        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.local_ranks = [self.rank]
        self.peer_memory = peer_memory
        self._symm = None
        self.peer_memory_error = None

    def alloc_pool(self, n_doubles, device):
        import os

        import torch

        want = self.peer_memory != "off" and self.dist.get_backend() == "nccl" and os.environ.get("TCC_B200_PEER_MEMORY", "1") != "0"
        buf = hdl = None
        if want:
            try:
                import torch.distributed._symmetric_memory as symm

                buf = symm.empty(n_doubles, dtype=torch.float64, device=device)
                hdl = symm.rendezvous(buf, self.dist.group.WORLD)
            except Exception as exc:  # no peer mapping on this system: NCCL transport
                self.peer_memory_error = repr(exc)
                buf = hdl = None
            # every rank must take the same path: if the mapping failed anywhere, nobody uses it
            flag = torch.tensor([1.0 if hdl is not None else 0.0], device=device)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            if float(flag[0]) < 1.0:
                buf = hdl = None
        if hdl is not None:
            ptrs = torch.tensor([int(x) for x in hdl.buffer_ptrs], dtype=torch.int64, device=device)
            self._symm = hdl
            return [buf], [ptrs]
        self._symm = None
        return [torch.empty(n_doubles, dtype=torch.float64, device=device)], [None]

    def stream_barrier(self):
        if self._symm is not None:
            self._symm.barrier(channel=0)
        else:
            self.dist.barrier()

    def exchange(self, send, recv_counts):
        import torch

        (mine,) = send
        (counts,) = recv_counts  # rows to expect from every peer (both sides derive them from the layouts)
        recv = [None] * self.world
        ops = []
        for peer in range(self.world):
            if peer == self.rank:
                recv[peer] = mine[peer]
                continue
            recv[peer] = torch.empty((counts[peer],) + tuple(mine[peer].shape[1:]), dtype=mine[peer].dtype, device=mine[peer].device)
            if mine[peer].shape[0]:
                ops.append(self.dist.P2POp(self.dist.isend, mine[peer].contiguous(), peer))
            if counts[peer]:
                ops.append(self.dist.P2POp(self.dist.irecv, recv[peer], peer))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        return [recv]

    def exchange_packed(self, send_bufs, send_counts, recv_counts):
        """send_bufs[0]: rows grouped by destination rank ([sum(send_counts), nb]); returns the rows grouped by source.
        One NCCL all-to-all; the gloo backend of the CPU tests has no all-to-all and takes the isend/irecv path."""
        import torch

        (buf,), (sc,), (rc,) = send_bufs, send_counts, recv_counts
        out = torch.empty((sum(rc),) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
        if self.dist.get_backend() == "nccl":
            self.dist.all_to_all_single(out, buf, output_split_sizes=list(rc), input_split_sizes=list(sc))
            return [out]
        parts = list(torch.split(buf, list(sc)))
        got = self.exchange([parts], [list(rc)])[0]
        return [torch.cat(got)] if got else [out]

    def all_reduce_sum(self, tensors):
        (t,) = tensors
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)

    def all_gather_padded(self, blocks, max_rows):
        import torch

        (b,) = blocks
        inp = torch.empty((max_rows,) + tuple(b.shape[1:]), dtype=b.dtype, device=b.device)
        inp[: b.shape[0]] = b
        out = torch.empty((self.world, max_rows) + tuple(b.shape[1:]), dtype=b.dtype, device=b.device)
        self.dist.all_gather_into_tensor(out.view(self.world * max_rows, -1), inp.view(max_rows, -1))
        return [out]

    def reduce_scatter_sum(self, stacked):
        import torch

        (x,) = stacked
        if self.dist.get_backend() == "nccl":
            out = torch.empty(tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
            self.dist.reduce_scatter_tensor(out, x, op=self.dist.ReduceOp.SUM)
            return [out]
        self.dist.all_reduce(x, op=self.dist.ReduceOp.SUM)  # gloo (CPU tests) has no reduce-scatter
        return [x[self.rank].clone()]


class ShardPool:
    """Fixed pool of shard-sized buffers: every [rows, nb]-sized block of the sharded driver lives in one of `n_slots`
    slots of ONE allocation per rank, made once.  Nothing shard-sized is allocated during an evaluation (at (20e,20o) on
    8 GPUs four 33.5 GiB slots are all the memory there is), and because the allocation is symmetric -- same size and
    slot offsets on every rank, mapped into every process when the transport allows it -- a row move writes straight
    into slot s of the peers.  All ranks run the same program, so slot numbers agree across ranks."""

    def __init__(self, comm, device, slot_doubles: int, n_slots: int):
        import torch

        self.torch = torch
        self.comm = comm
        self.slot = int(slot_doubles)
        self.n_slots = int(n_slots)
        self.buffers, self.peer_ptrs = comm.alloc_pool(self.slot * self.n_slots, device)
        self.p2p = self.peer_ptrs[0] is not None
        self._used = [False] * self.n_slots
        self._owner = {}  # data_ptr of local rank 0's view -> (first slot, number of slots)
        # stream barriers seen so far, and for every slot the count at its last release: a slot released BEFORE the last
        # barrier is idle on every rank (all ranks run the same program), so peers may write into it without another one
        self.barriers = 0
        self._released_at = [-1] * self.n_slots

    def reset(self):
        self._used = [False] * self.n_slots
        self._owner.clear()
        self._released_at = [self.barriers] * self.n_slots

    def idle_everywhere(self, first: int, n: int) -> bool:
        return all(self._released_at[j] < self.barriers for j in range(first, first + n))

    def _take(self, n: int) -> int:
        run = 0
        for i in range(self.n_slots):
            run = run + 1 if not self._used[i] else 0
            if run == n:
                first = i - n + 1
                for j in range(first, i + 1):
                    self._used[j] = True
                return first
        raise MemoryError(f"shard pool exhausted: {n} contiguous slot(s) of {self.n_slots} requested, in use {self._used}")

    def empty(self, rows_per_rank, trailing):
        """one block [rows_k, *trailing] per local rank, all in the same slot(s); `trailing` = (nb,) or (nb, 2) ..."""
        per_row = int(np.prod(trailing))
        need = max(int(r) for r in rows_per_rank) * per_row if len(rows_per_rank) else 0
        n = max(1, -(-need // self.slot))
        first = self._take(n)
        out = []
        for buf, r in zip(self.buffers, rows_per_rank):
            out.append(buf[first * self.slot: first * self.slot + int(r) * per_row].view((int(r),) + tuple(trailing)))
        self._owner[first * self.slot] = (first, n)
        return out

    def zeros(self, rows_per_rank, trailing):
        out = self.empty(rows_per_rank, trailing)
        for v in out:
            v.zero_()
        return out

    def _key(self, vecs):
        """storage offset of the block inside the pool buffer (None for blocks that live elsewhere)"""
        if not vecs:
            return None
        v = vecs[0]
        if v.untyped_storage().data_ptr() != self.buffers[0].untyped_storage().data_ptr():
            return None
        return int(v.storage_offset()) - int(self.buffers[0].storage_offset())

    def slot_of(self, vecs):
        return self._owner.get(self._key(vecs))

    def release(self, vecs):
        """give the slot(s) of a block back (blocks that do not live in the pool are ignored)"""
        got = self._owner.pop(self._key(vecs), None)
        if got is not None:
            for j in range(got[0], got[0] + got[1]):
                self._used[j] = False
                self._released_at[j] = self.barriers


class RowMover:
    """Moves the rows of local [rows, nb] matrices between the row layouts of a sharded plan.  Host logic only
    (index maps from `tcc_shard_layout`); the tensors may live on the GPU (NCCL) or on the CPU (gloo tests)."""

    def __init__(self, plan: Plan, comm, device, chunk_bytes: int = 4 << 30):
        import torch

        self.torch = torch
        self.chunk_bytes = int(chunk_bytes)
        self.comm = comm
        self.world = comm.world
        self.device = device
        self.nb = plan.n_strings
        info = plan.shard_info()
        self.owner = [plan.shard_layout(l)[0] for l in range(info["n_layouts"])]
        # rows[l][r]: storage rows of rank r in layout l, ascending (= its local row order)
        self.rows = [[np.nonzero(own == r)[0] for r in range(self.world)] for own in self.owner]
        self._moves = {}
        self._push = {}
        self._rows_dev = {}
        self.reshards = 0
        self.timer = PhaseTimer(torch, False)
        self.pool = None  # ShardPool, attached by the driver
        self.lib = plan.lib
        self.max_rows = max(len(r) for lay in self.rows for r in lay)

    def rows_dev(self, layout: int, rank: int):
        """Storage rows of `rank` in `layout` as an index tensor on the device (cached)."""
        key = (layout, rank)
        if key not in self._rows_dev:
            t = self.torch
            self._rows_dev[key] = t.as_tensor(self.rows[layout][rank], dtype=t.int64, device=self.device)
        return self._rows_dev[key]

    def _move_plan(self, l_from: int, l_to: int, rank: int, row_elems: int):
        key = (l_from, l_to, rank, row_elems)
        if key not in self._moves:
            t = self.torch
            mine_from = self.rows[l_from][rank]
            mine_to = self.rows[l_to][rank]
            n_max = max(max(len(r) for r in self.rows[l_from]), 1)
            # the move runs in `pieces` rounds so that the staging buffers stay below chunk_bytes (every rank derives
            # the same number of rounds and the same split of every (source, destination) row list)
            pieces = max(1, -(-(n_max * row_elems * 8) // self.chunk_bytes))
            send_idx, recv_idx = [], []
            for peer in range(self.world):
                out_rows = np.intersect1d(mine_from, self.rows[l_to][peer], assume_unique=True)
                in_rows = np.intersect1d(self.rows[l_from][peer], mine_to, assume_unique=True)
                send_idx.append(np.array_split(np.searchsorted(mine_from, out_rows), pieces))
                recv_idx.append(np.array_split(np.searchsorted(mine_to, in_rows), pieces))
            rounds = []
            for j in range(pieces):
                s_j = [send_idx[peer][j] for peer in range(self.world)]
                r_j = [recv_idx[peer][j] for peer in range(self.world)]
                rounds.append((t.as_tensor(np.concatenate(s_j), dtype=t.int64, device=self.device),
                               t.as_tensor(np.concatenate(r_j), dtype=t.int64, device=self.device),
                               [len(x) for x in s_j], [len(x) for x in r_j]))
            self._moves[key] = (rounds, len(mine_to))
        return self._moves[key]

    def _push_plan(self, l_from: int, l_to: int, rank: int):
        """index arrays of the peer-memory move of `rank`: for every row it holds in l_from, the local source row, the
        owner in l_to and the row's position among that owner's rows"""
        key = (l_from, l_to, rank)
        if key not in self._push:
            t = self.torch
            mine = self.rows[l_from][rank]
            dst_rank = self.owner[l_to][mine].astype(np.int32)
            dst_row = np.empty(len(mine), dtype=np.int64)
            for q in range(self.world):
                sel = dst_rank == q
                if sel.any():
                    dst_row[sel] = np.searchsorted(self.rows[l_to][q], mine[sel])
            self._push[key] = (t.arange(len(mine), dtype=t.int64, device=self.device),
                               t.as_tensor(dst_row, dtype=t.int64, device=self.device),
                               t.as_tensor(dst_rank, dtype=t.int32, device=self.device))
        return self._push[key]

    def _stream(self):
        import ctypes as C

        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def reshard(self, vecs: List["torch.Tensor"], l_from: int, l_to: int, keep_src: bool = False) -> List["torch.Tensor"]:
        """vecs[i]: local [rows, ...] block of the i-th local rank in layout l_from -> the same in layout l_to, in a
        fresh slot of the pool; the source slot is released unless keep_src.  With a peer-mapped pool every row is
        written straight into the destination slot of its new owner (tcc_push_rows, one pass over NVLink), followed by
        a stream barrier (all rows have arrived) and preceded by one only when the destination slot was released since the
        last barrier (otherwise every rank is known to have finished with whatever lived there).  Otherwise: pack, NCCL all-to-all, unpack in rounds of chunk_bytes."""
        import ctypes as C

        from . import _lib

        t = self.torch
        if l_from == l_to:
            return vecs
        self.reshards += 1
        trailing = tuple(vecs[0].shape[1:])
        row_elems = int(np.prod(trailing))
        new_rows = [len(self.rows[l_to][r]) for r in self.comm.local_ranks]
        pool = self.pool
        if pool is None:  # CPU tests of the transport: plain tensors
            out = [t.empty((n,) + trailing, dtype=vecs[0].dtype, device=self.device) for n in new_rows]
        else:
            out = pool.empty(new_rows, trailing)
        if pool is not None and pool.p2p:
            first, n_dst = pool.slot_of(out)
            nbytes = sum(int(v.numel()) * 8 for v in vecs)
            if not pool.idle_everywhere(first, n_dst):
                with self.timer.phase("move_barrier"):
                    self.comm.stream_barrier()
                pool.barriers += 1
            with self.timer.phase("move_push", nbytes):
                for k, r in enumerate(self.comm.local_ranks):
                    src_row, dst_row, dst_rank = self._push_plan(l_from, l_to, r)
                    if vecs[k].shape[0]:
                        _lib.check(self.lib.tcc_push_rows(
                            int(self.device.index), C.c_void_p(vecs[k].data_ptr()), row_elems, int(vecs[k].shape[0]),
                            C.c_void_p(src_row.data_ptr()), C.c_void_p(dst_row.data_ptr()), C.c_void_p(dst_rank.data_ptr()),
                            C.c_void_p(pool.peer_ptrs[k].data_ptr()), first * pool.slot, self._stream()))
            with self.timer.phase("move_barrier"):
                self.comm.stream_barrier()
            pool.barriers += 1
        else:
            plans = [self._move_plan(l_from, l_to, r, row_elems) for r in self.comm.local_ranks]
            for j in range(len(plans[0][0])):
                rnd = [mp[0][j] for mp in plans]
                # pack the rows by destination (one gather), one all-to-all, unpack by local row (one scatter)
                with self.timer.phase("move_pack"):
                    send = [v.index_select(0, r_[0]) for v, r_ in zip(vecs, rnd)]
                with self.timer.phase("move_all_to_all", sum(int(x.numel()) * 8 for x in send)):
                    recv = self.comm.exchange_packed(send, [r_[2] for r_ in rnd], [r_[3] for r_ in rnd])
                del send
                with self.timer.phase("move_unpack"):
                    for new, r_, got in zip(out, rnd, recv):
                        if r_[1].numel():
                            new.index_copy_(0, r_[1], got)
                del recv
        if pool is not None and not keep_src:
            pool.release(vecs)
        return out

    # ---- row-sharded [rows, nb]  <->  column-sharded [na, cols] (all rows, a block of columns) ----------------
    def col_bounds(self, rank: int, piece: int = 0, pieces: int = 1):
        """Columns of `rank` (contiguous, balanced); with pieces > 1 the `piece`-th part of that block."""
        lo, hi = shard_bounds(self.nb, rank, self.world)
        plo, phi = shard_bounds(hi - lo, piece, pieces)
        return lo + plo, lo + phi

    def column_pieces(self, layout: int) -> int:
        n_max = max(max(len(r) for r in self.rows[layout]), 1)
        return max(1, -(-(n_max * self.nb * 8) // self.chunk_bytes))

    def to_columns(self, vecs, layout: int, piece: int = 0, pieces: int = 1):
        """vecs[i]: [rows of local rank i in `layout`, nb] -> [na, (piece of the) columns of that rank], storage rows."""
        t = self.torch
        na = len(self.owner[layout])
        send, sc, rc = [], [], []
        for v, r in zip(vecs, self.comm.local_ranks):
            parts = [v[:, lo:hi].reshape(-1) for lo, hi in (self.col_bounds(d, piece, pieces) for d in range(self.world))]
            send.append(t.cat(parts))
            sc.append([int(x.numel()) for x in parts])
            lo, hi = self.col_bounds(r, piece, pieces)
            rc.append([len(self.rows[layout][s_]) * (hi - lo) for s_ in range(self.world)])
        recv = self.comm.exchange_packed(send, sc, rc)
        out = []
        for r, got, counts in zip(self.comm.local_ranks, recv, rc):
            lo, hi = self.col_bounds(r, piece, pieces)
            full = t.empty((na, hi - lo), dtype=got.dtype, device=self.device)
            for s_, blk in enumerate(t.split(got, counts)):
                if blk.numel():
                    full.index_copy_(0, self.rows_dev(layout, s_), blk.reshape(-1, hi - lo))
            out.append(full)
        return out

    def to_rows(self, cols, layout: int, piece: int = 0, pieces: int = 1, add_into=None):
        """Inverse of to_columns.  Returns new [rows, nb] matrices (pieces == 1), or ADDS the received column blocks
        into add_into[i] (the accumulation of the alpha-alpha sigma, piece by piece)."""
        t = self.torch
        send, sc, rc = [], [], []
        for c, r in zip(cols, self.comm.local_ranks):
            parts = [c.index_select(0, self.rows_dev(layout, d)).reshape(-1) for d in range(self.world)]
            send.append(t.cat(parts))
            sc.append([int(x.numel()) for x in parts])
            n_rows = len(self.rows[layout][r])
            rc.append([n_rows * (hi - lo) for lo, hi in (self.col_bounds(s_, piece, pieces) for s_ in range(self.world))])
        recv = self.comm.exchange_packed(send, sc, rc)
        out = []
        for i, (r, got, counts) in enumerate(zip(self.comm.local_ranks, recv, rc)):
            n_rows = len(self.rows[layout][r])
            new = add_into[i] if add_into is not None else t.zeros((n_rows, self.nb), dtype=got.dtype, device=self.device)
            for s_, blk in enumerate(t.split(got, counts)):
                lo, hi = self.col_bounds(s_, piece, pieces)
                if blk.numel():
                    new[:, lo:hi] += blk.reshape(n_rows, hi - lo)
            out.append(new)
        return out


class _RankState:
    def __init__(self, plan: Plan, rank: int):
        self.plan = plan
        self.rank = rank


class ShardedUCC:
    """energy / energy_and_grad / civector of one ansatz with the CI vector sharded over `comm.world` ranks."""

    def __init__(self, n_qubits, n_elec, ex_ops, param_ids, comm, device=0, sigma_mode="auto", chunk_bytes=4 << 30,
                 fuse_reverse_moves="auto"):
        """sigma_mode: "distributed" keeps the vector sharded through H c (three alpha-beta passes + a column
        transpose for the alpha-alpha terms); "replicated" all-reduces the full vector (2 full vectors per GPU; faster:
        one pass of the opposite-spin kernel); "auto" = replicated while two full vectors take less than a quarter of
        the GPU's memory, distributed beyond ((20e,20o): 273 GB per vector)."""
        import torch

        if sigma_mode not in ("auto", "distributed", "replicated"):
            raise ValueError("sigma_mode must be 'auto', 'distributed' or 'replicated'")
        if sigma_mode == "auto":
            from math import comb

            n = n_qubits // 2
            full_bytes = 8 * comb(n, n_elec // 2) ** 2
            total = torch.cuda.get_device_properties(int(device)).total_memory
            sigma_mode = "replicated" if 2 * full_bytes < total // 4 else "distributed"
        self.sigma_mode = sigma_mode if comm.world > 1 else "replicated"
        # reverse sweep: ket and bra rows interleaved in one [rows, 2, nb] block move with ONE all-to-all per segment
        # instead of two; costs two extra shards while the block is built ("auto": only while that is < 1/8 of HBM)
        self.fuse_reverse_moves = fuse_reverse_moves
        self.torch = torch
        self.comm = comm
        self.world = comm.world
        self.device = torch.device("cuda", int(device))
        self.states = [_RankState(Plan(n_qubits, n_elec, ex_ops, param_ids, device=int(device), shard=(r, self.world)), r)
                       for r in comm.local_ranks]
        p0 = self.states[0].plan
        self.n_params, self.na = p0.n_params, p0.n_strings
        self.nb = p0.n_strings
        self.segments = p0.shard_segments()
        self.mover = RowMover(p0, comm, self.device, chunk_bytes)
        self.rows = self.mover.rows
        self.timer = self.mover.timer  # one PhaseTimer for the driver and its row mover (off unless enabled)
        # every shard-sized block lives in a fixed pool of slots (ShardPool): 4 cover the deepest point of the path (ket,
        # sigma, the moved copy of a sigma pass and its partial result); the fused reverse sweep moves (ket, bra) as one
        # double-size block and gets 6 so that two free neighbouring slots always exist
        slot = max(self.mover.max_rows, 1) * self.nb
        total = torch.cuda.get_device_properties(self.device).total_memory
        if fuse_reverse_moves == "auto":
            fuse_reverse_moves = self.world > 1 and 2 * 8 * slot < total // 8
        self.fuse_reverse_moves = bool(fuse_reverse_moves)
        self.pool = ShardPool(comm, self.device, slot, 6 if self.fuse_reverse_moves else 4)
        self.mover.pool = self.pool
        self.ref2int = p0.storage_order()  # storage address of every per-spin string, by reference address
        self.hf_row = int(self.ref2int[0])  # HF determinant = reference address 0 of both spins

    @property
    def reshards(self) -> int:
        return self.mover.reshards

    def set_hamiltonian(self, int1e, int2e):
        for st in self.states:
            st.plan.set_hamiltonian(int1e, int2e)

    def _reshard(self, vecs, l_from, l_to, keep_src=False):
        return self.mover.reshard(vecs, l_from, l_to, keep_src)

    def _rows_of(self, layout):
        return [len(self.rows[layout][st.rank]) for st in self.states]

    # ---- the path -----------------------------------------------------------------------------------
    def _forward(self, params, init_sparse=None):
        """init_sparse = (alpha reference addresses, beta reference addresses, amplitudes) of the non-zeros of an
        initial state (the dense `init_state` of get_civector, evolve_civector.py:189-193, does not exist at the sizes
        this engine is for); None = the HF determinant."""
        t = self.torch
        first = self.segments[0][0] if self.segments else 0
        if init_sparse is None:
            ia, ib, iv = np.array([self.hf_row]), np.array([self.hf_row]), np.array([1.0])
        else:
            ia = self.ref2int[np.asarray(init_sparse[0], dtype=np.int64)].astype(np.int64)
            ib = self.ref2int[np.asarray(init_sparse[1], dtype=np.int64)].astype(np.int64)
            iv = np.asarray(init_sparse[2], dtype=np.float64)
        self.pool.reset()  # a new evaluation owns the whole pool
        kets = self.pool.zeros(self._rows_of(first), (self.nb,))
        for ket, st in zip(kets, self.states):
            st.plan.shard_set_params(params)
            rows = self.rows[first][st.rank]
            pos = np.searchsorted(rows, ia)
            pos[pos >= len(rows)] = 0
            mine = (rows[pos] == ia) if len(rows) else np.zeros(len(ia), dtype=bool)
            if mine.any():
                flat = t.as_tensor(pos[mine] * self.nb + ib[mine], dtype=t.int64, device=self.device)
                ket.view(-1).index_put_((flat,), t.as_tensor(iv[mine], dtype=t.float64, device=self.device))
        layout = first
        for seg, (l, _, _) in enumerate(self.segments):
            kets = self._reshard(kets, layout, l)
            layout = l
            with self.timer.phase("forward_kernels"):
                for k, st in enumerate(self.states):
                    if kets[k].numel():  # a rank may own no row in a layout of a tiny active space
                        st.plan.shard_forward(seg, kets[k])
        return kets, layout

    def _gather_full(self, vecs, layout):
        """Full [na, nb] storage-order matrix on every rank: one all-gather of the (padded) row blocks, then the rows
        of every source rank are copied to their storage positions."""
        t = self.torch
        max_rows = max(len(r) for r in self.rows[layout])
        with self.timer.phase("sigma_all_gather", 8 * max_rows * self.nb * (self.world - 1)):
            gathered = self.comm.all_gather_padded(vecs, max_rows)
        fulls = []
        with self.timer.phase("sigma_unpack"):
            for g in gathered:
                full = t.empty((self.na, self.nb), dtype=t.float64, device=self.device)
                for s_ in range(self.world):
                    n_s = len(self.rows[layout][s_])
                    if n_s:
                        full.index_copy_(0, self.mover.rows_dev(layout, s_), g[s_, :n_s])
                fulls.append(full)
                del full
        return fulls

    def _sigma_distributed(self, kets, layout):
        """sigma = H c with the vector kept sharded (include/tcc_b200.h, sharded sigma): kets in `layout` -> bras in
        the same layout.  Needs no full vector on any rank."""
        t = self.torch
        l3 = self.states[0].plan.shard_sigma_layouts()
        with self.timer.phase("sigma_kernels"):
            sig = self.pool.zeros(self._rows_of(layout), (self.nb,))
            for st, ket, s_loc in zip(self.states, kets, sig):
                if ket.numel():
                    st.plan.shard_sigma_bb(ket, s_loc)
        # alpha-beta: pass i in layout l3[i].  The passes live in helper methods: a tensor that a loop variable of this
        # frame still names would stay allocated through the next pass (a fifth shard at (20e,20o) is an OOM).
        for i, l in enumerate(l3):
            self._sigma_ab_pass(i, l, kets, layout, sig)
        # alpha-alpha: on column blocks of all rows (in pieces, so that the transposed copies stay small)
        pieces = self.mover.column_pieces(layout)
        for j in range(pieces):
            self._sigma_aa_piece(j, pieces, kets, layout, sig)
        return sig

    def _sigma_replicated(self, kets, layout):
        """sigma = H c on the gathered full vector: every rank contracts a block of source rows, the partial results are
        packed by destination rank and summed with ONE reduce-scatter (each rank receives only its rows).  Two to three
        full vectors per GPU."""
        t = self.torch
        c_full = self._gather_full(kets, layout)
        max_rows = max(len(r) for r in self.rows[layout])
        stacked = []
        for k, st in enumerate(self.states):
            with self.timer.phase("sigma_kernels"):
                sig = t.zeros_like(c_full[k])
                lo, hi = shard_bounds(self.na, st.rank, self.world)
                st.plan.apply_h_rows(c_full[k], sig, lo, hi - lo)
            with self.timer.phase("sigma_pack"):
                buf = t.empty((self.world, max_rows, self.nb), dtype=t.float64, device=self.device)
                for d in range(self.world):
                    n_d = len(self.rows[layout][d])
                    if n_d:
                        t.index_select(sig, 0, self.mover.rows_dev(layout, d), out=buf[d, :n_d])
                    if n_d < max_rows:
                        buf[d, n_d:].zero_()
            stacked.append(buf)
            del sig, buf
        del c_full
        with self.timer.phase("sigma_reduce_scatter", 8 * max_rows * self.nb * (self.world - 1)):
            mine = self.comm.reduce_scatter_sum(stacked)
        del stacked
        return [mine[k][: len(self.rows[layout][st.rank])] for k, st in enumerate(self.states)]

    def _sigma_ab_pass(self, i, l, kets, layout, sig):
        """sig += pass i of the opposite-spin terms.  Peak: kets + sig + the moved copy + its partial sigma = 4 slots."""
        c_i = self._reshard(kets, layout, l, keep_src=True)
        with self.timer.phase("sigma_kernels"):
            part = self.pool.zeros(self._rows_of(l), (self.nb,))
            for k in range(len(self.states)):
                if part[k].numel():
                    self.states[k].plan.shard_sigma_ab(i, c_i[k], part[k])
        if c_i is not kets:
            self.pool.release(c_i)
        del c_i
        part = self._reshard(part, l, layout)
        with self.timer.phase("sigma_accumulate"):
            for k in range(len(sig)):
                sig[k] += part[k]
        self.pool.release(part)

    def _sigma_aa_piece(self, j, pieces, kets, layout, sig):
        t = self.torch
        with self.timer.phase("sigma_aa_to_columns"):
            cols = self.mover.to_columns(kets, layout, j, pieces)
        s_cols = []
        for k in range(len(self.states)):
            with self.timer.phase("sigma_kernels"):
                out = t.empty_like(cols[k])
                if out.numel():
                    self.states[k].plan.shard_sigma_aa(cols[k], out)
            s_cols.append(out)
            del out
        del cols
        with self.timer.phase("sigma_aa_to_rows"):
            self.mover.to_rows(s_cols, layout, j, pieces, add_into=sig)

    def _dot(self, x, y):
        """<x|y> of two local blocks as a 1-element device tensor, without a temporary; in slices of at most 2^30
        elements (the BLAS dot behind torch.dot takes 32-bit lengths; a (20e,20o) shard has 4.5e9 elements)."""
        t = self.torch
        xf, yf = x.reshape(-1), y.reshape(-1)
        acc = t.zeros(1, dtype=t.float64, device=self.device)
        step = 1 << 30
        for lo in range(0, xf.numel(), step):
            acc += t.dot(xf[lo:lo + step], yf[lo:lo + step])
        return acc

    def civector(self, params, init_sparse=None) -> np.ndarray:
        """Reference-order CI vector on the host (small systems / tests: gathers the full vector)."""
        t = self.torch
        kets, layout = self._forward(np.asarray(params, dtype=np.float64), init_sparse)
        full = self._gather_full(kets, layout)[0].reshape(-1)
        out = t.empty_like(full)
        self.states[0].plan.from_storage_order_dev(full, out)
        t.cuda.synchronize()
        return out.cpu().numpy()

    def energy_and_grad(self, params, want_grad: bool = True, init_sparse=None):
        t = self.torch
        params = np.asarray(params, dtype=np.float64)
        kets, layout = self._forward(params, init_sparse)
        if self.sigma_mode == "distributed":
            bras = self._sigma_distributed(kets, layout)
        else:
            bras = self._sigma_replicated(kets, layout)
        with self.timer.phase("energy_dot_allreduce"):
            parts = [self._dot(bra, ket) for bra, ket in zip(bras, kets)]
            self.comm.all_reduce_sum(parts)
            energy = float(parts[0][0])
        if not want_grad:
            return energy, None
        if self.fuse_reverse_moves:
            # (ket, bra) as ONE block: a segment boundary costs one row move instead of two.  Element-interleaved
            # [rows, nb, 2] when every batch has lean streams (16-byte elements in the kernel too), else row-interleaved
            # [rows, 2, nb] (two vectors with a row stride)
            il = all(st.plan.shard_can_interleave() for st in self.states)
            with self.timer.phase("reverse_interleave"):
                both = self.pool.empty(self._rows_of(layout), (self.nb, 2) if il else (2, self.nb))
                for k in range(len(both)):
                    if il:
                        both[k][:, :, 0] = kets[k]
                        both[k][:, :, 1] = bras[k]
                    else:
                        both[k][:, 0] = kets[k]
                        both[k][:, 1] = bras[k]
            self.pool.release(kets)
            self.pool.release(bras)
            del kets, bras
            for seg in range(len(self.segments) - 1, -1, -1):
                l = self.segments[seg][0]
                both = self._reshard(both, layout, l)
                layout = l
                with self.timer.phase("reverse_kernels"):
                    for k, st in enumerate(self.states):
                        if not both[k].numel():
                            continue
                        if il:
                            st.plan.shard_reverse_interleaved(seg, both[k])
                        else:
                            st.plan.shard_reverse_strided(seg, both[k][:, 0], both[k][:, 1], 2 * self.nb)
        else:
            for seg in range(len(self.segments) - 1, -1, -1):
                l = self.segments[seg][0]
                kets = self._reshard(kets, layout, l)
                bras = self._reshard(bras, layout, l)
                layout = l
                with self.timer.phase("reverse_kernels"):
                    for k, st in enumerate(self.states):
                        if kets[k].numel():
                            st.plan.shard_reverse(seg, kets[k], bras[k])
        with self.timer.phase("gradient_allreduce"):
            grads = [t.as_tensor(st.plan.shard_gradient(), dtype=t.float64, device=self.device) for st in self.states]
            self.comm.all_reduce_sum(grads)
            out = grads[0].cpu().numpy()
        return energy, out

    def energy(self, params) -> float:
        return self.energy_and_grad(params, want_grad=False)[0]

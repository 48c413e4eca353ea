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

    def __init__(self, world: int):
        self.world = int(world)
        self.local_ranks = list(range(self.world))

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

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.local_ranks = [self.rank]

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
        self._rows_dev = {}
        self.reshards = 0
        self.timer = PhaseTimer(torch, False)

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

    def reshard(self, vecs: List["torch.Tensor"], l_from: int, l_to: int) -> List["torch.Tensor"]:
        """vecs[i]: local [rows, nb] matrix of the i-th local rank in layout l_from -> the same in layout l_to."""
        t = self.torch
        if l_from == l_to:
            return vecs
        self.reshards += 1
        row_elems = int(vecs[0][0].numel()) if vecs[0].shape[0] else int(np.prod(vecs[0].shape[1:]))
        plans = [self._move_plan(l_from, l_to, r, row_elems) for r in self.comm.local_ranks]
        out = [t.empty((mp[1],) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device) for v, mp in zip(vecs, plans)]
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
        self.ref2int = p0.storage_order()  # storage address of every per-spin string, by reference address
        self.hf_row = int(self.ref2int[0])  # HF determinant = reference address 0 of both spins

    @property
    def reshards(self) -> int:
        return self.mover.reshards

    def set_hamiltonian(self, int1e, int2e):
        for st in self.states:
            st.plan.set_hamiltonian(int1e, int2e)

    def _reshard(self, vecs, l_from, l_to):
        return self.mover.reshard(vecs, l_from, l_to)

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
        kets = []
        for st in self.states:
            st.plan.shard_set_params(params)
            rows = self.rows[first][st.rank]
            ket = t.zeros((len(rows), self.nb), dtype=t.float64, device=self.device)
            pos = np.searchsorted(rows, ia)
            pos[pos >= len(rows)] = 0
            mine = (rows[pos] == ia) if len(rows) else np.zeros(len(ia), dtype=bool)
            if mine.any():
                flat = t.as_tensor(pos[mine] * self.nb + ib[mine], dtype=t.int64, device=self.device)
                ket.view(-1).index_put_((flat,), t.as_tensor(iv[mine], dtype=t.float64, device=self.device))
            kets.append(ket)
            del ket
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
        sig = []
        for st, ket in zip(self.states, kets):
            with self.timer.phase("sigma_kernels"):
                s_loc = t.zeros_like(ket)
                if ket.numel():
                    st.plan.shard_sigma_bb(ket, s_loc)
            sig.append(s_loc)
            del s_loc
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
        """sig += pass i of the opposite-spin terms.  Peak: kets + sig + the moved copy + its partial sigma."""
        t = self.torch
        c_i = self._reshard(kets, layout, l)
        part = []
        for k in range(len(self.states)):
            with self.timer.phase("sigma_kernels"):
                s_i = t.zeros_like(c_i[k])
                if s_i.numel():
                    self.states[k].plan.shard_sigma_ab(i, c_i[k], s_i)
            part.append(s_i)
            del s_i
        del c_i
        part = self._reshard(part, l, layout)
        with self.timer.phase("sigma_accumulate"):
            for k in range(len(sig)):
                sig[k] += part[k]

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
        fuse = self.fuse_reverse_moves
        if fuse == "auto":
            total = t.cuda.get_device_properties(self.device).total_memory
            fuse = self.world > 1 and 2 * 8 * max(k_.numel() for k_ in kets) < total // 8
        if fuse:
            # element-interleaved block [rows, nb, 2] when every batch has lean streams (16-byte elements in the kernel
            # too), else row-interleaved [rows, 2, nb] (two vectors with a row stride)
            il = all(st.plan.shard_can_interleave() for st in self.states)
            with self.timer.phase("reverse_interleave"):
                both = [t.stack((kets[k], bras[k]), dim=2 if il else 1) for k in range(len(kets))]
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
                    for k, st in enumerate(self.states):  # by index: no loop variable keeps a shard of the old layout alive
                        if kets[k].numel():
                            st.plan.shard_reverse(seg, kets[k], bras[k])
        with self.timer.phase("gradient_allreduce"):
            grads = [t.as_tensor(st.plan.shard_gradient(), dtype=t.float64, device=self.device) for st in self.states]
            self.comm.all_reduce_sum(grads)
            out = grads[0].cpu().numpy()
        return energy, out

    def energy(self, params) -> float:
        return self.energy_and_grad(params, want_grad=False)[0]

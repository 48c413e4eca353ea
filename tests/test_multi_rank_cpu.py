"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU benchmark path (small-N /
multi-theta regime): every rank evaluates its own parameter sets, no data-path collective, timings are reduced
with MAX over ranks and the per-rank results are gathered.  No GPU work is done here."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tencirchem_b200.parallel import reduce_timing_max, shard_parameter_sets, gather_results

    thetas = np.arange(10 * 3, dtype=np.float64).reshape(10, 3)
    mine = shard_parameter_sets(thetas, rank, world)
    # stand-in for the GPU evaluation: a deterministic function of theta
    energies = mine.sum(axis=1)
    grads = 2.0 * mine
    t = reduce_timing_max(10.0 + rank)
    e_all, g_all = gather_results(energies, grads, n_total=len(thetas), rank=rank, world=world)
    if rank == 0:
        out.put((t, e_all.tolist(), g_all.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_parameter_sets_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, e_all, g_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    thetas = np.arange(30, dtype=np.float64).reshape(10, 3)
    assert t == 11.0  # MAX over ranks
    np.testing.assert_array_equal(np.array(e_all), thetas.sum(axis=1))
    np.testing.assert_array_equal(np.array(g_all), 2.0 * thetas)


def test_shard_sizes_are_balanced():
    sys.path.insert(0, ROOT)
    from tencirchem_b200.parallel import shard_bounds

    for n, w in [(10, 2), (7, 4), (3, 8), (1024, 8), (0, 2)]:
        bounds = [shard_bounds(n, r, w) for r in range(w)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        sizes = [b - a for a, b in bounds]
        assert max(sizes) - min(sizes) <= 1
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(w - 1))


def _reshard_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ["TCC_B200_TILE_ELEMS"] = "400"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pools
    from tencirchem_b200.plan import Plan
    from tencirchem_b200.sharded import RowMover, TorchComm

    ex_ops, param_ids = pools.uccsd_ex_ops(3, 3)[:2]
    plan = Plan(12, 6, ex_ops, param_ids, device=None, shard=(rank, world))  # host-only: schedule + layouts
    mover = RowMover(plan, TorchComm(), torch.device("cpu"), chunk_bytes=512)  # several rounds per move
    na = plan.n_strings
    full = torch.arange(na * na, dtype=torch.float64).reshape(na, na)  # full[a, b] identifies its element
    segs = plan.shard_segments()
    layout = segs[0][0]
    local = [full[torch.as_tensor(mover.rows[layout][rank])].clone()]
    ok = True
    # walk the forward sweep's layout sequence and back: after every move the rank must hold exactly its rows
    for l in [s[0] for s in segs] + [s[0] for s in reversed(segs)]:
        local = mover.reshard(local, layout, l)
        layout = l
        ok = ok and torch.equal(local[0], full[torch.as_tensor(mover.rows[l][rank])])
    # row-sharded <-> column-sharded (the transposes of the sharded alpha-alpha sigma)
    lo, hi = mover.col_bounds(rank)
    cols = mover.to_columns(local, layout)
    ok = ok and torch.equal(cols[0], full[:, lo:hi])
    back = mover.to_rows(cols, layout)
    ok = ok and torch.equal(back[0], local[0])
    # the collectives of the replicated H c: all-gather of padded row blocks, reduce-scatter of packed partial sums
    comm = mover.comm
    max_rows = max(len(r) for r in mover.rows[layout])
    gathered = comm.all_gather_padded(local, max_rows)[0]
    for s_ in range(world):
        n_s = len(mover.rows[layout][s_])
        ok = ok and torch.equal(gathered[s_, :n_s], full[torch.as_tensor(mover.rows[layout][s_])])
    stacked = torch.zeros((world, max_rows, na), dtype=torch.float64)
    for d in range(world):
        n_d = len(mover.rows[layout][d])
        stacked[d, :n_d] = (rank + 1) * full[torch.as_tensor(mover.rows[layout][d])]
    mine = comm.reduce_scatter_sum([stacked])[0]
    n_me = len(mover.rows[layout][rank])
    ok = ok and torch.equal(mine[:n_me], sum(range(1, world + 1)) * full[torch.as_tensor(mover.rows[layout][rank])])
    # ownership is a partition of the rows in every layout
    for own in mover.owner:
        ok = ok and own.min() >= 0 and own.max() < world and len(own) == na
    covered = sorted(b for s in segs for b in range(s[1], s[2]))
    ok = ok and covered == list(range(plan.n_batches))
    t = torch.tensor([1.0 if ok else 0.0, float(mover.reshards)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put((t[0].item(), t[1].item(), len(segs)))
    dist.barrier()
    dist.destroy_process_group()


def test_alpha_row_reshard_world2():
    """Alpha-row sharding, host side: the rows move between the layouts of consecutive segments through
    isend/irecv (gloo here, NCCL on GPUs) and every rank always holds exactly the rows its layout assigns."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reshard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, reshards, n_segs = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == 1.0
    assert n_segs > 2 and reshards >= n_segs - 1

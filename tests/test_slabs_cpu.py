"""CPU (gloo, world_size 2) tests of the host-side slab logic: partition/assemble round trip and a
halo-exchange emulation -- each rank applies the oracle's stencil to its slab with neighbour rows
received from the ring neighbours, the result must equal the slab of the global apply."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mpbp_oracle as O
    from mp_block_preconditioners_b200.parallel import assemble_global, gather_slabs, scatter_slab, slab_rows
    rng = np.random.default_rng(3)
    x = rng.standard_normal(5 * n * n)
    xs = scatter_slab(x, n, 5, rank, world)
    r0, rows = slab_rows(n, rank, world)
    assert xs.shape == (5 * rows * n,)
    # round trip through the collective
    xg = gather_slabs(torch.from_numpy(xs), n, 5, world).numpy()
    ok_rt = np.array_equal(xg, x)
    # halo exchange with the ring neighbours (same send/recv order as the CUDA plan: last row -> next,
    # first row -> prev; top <- prev, bot <- next), then a padded periodic-in-x stencil apply
    f = xs.reshape(5, rows, n)
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    top, bot = np.empty((5, n)), np.empty((5, n))
    reqs = [dist.isend(torch.from_numpy(f[:, -1, :].copy()), nxt, tag=1), dist.isend(torch.from_numpy(f[:, 0, :].copy()), prev, tag=2)]
    ttop, tbot = torch.empty(5, n, dtype=torch.float64), torch.empty(5, n, dtype=torch.float64)
    dist.recv(ttop, prev, tag=1)
    dist.recv(tbot, nxt, tag=2)
    for r in reqs:
        r.wait()
    padded = np.concatenate([ttop.numpy()[:, None, :], f, tbot.numpy()[:, None, :]], axis=1)  # rows r0-1 .. r0+rows
    # reference: the same padded block cut from the global field
    g = x.reshape(5, n, n)
    idx = [(r0 + k) % n for k in range(-1, rows + 1)]
    ok_halo = np.array_equal(padded, g[:, idx, :])
    # slab of the global operator apply == operator apply restricted to slab rows (uses the oracle)
    ops = O.Operators(n, 1.0, 10.0, 1.0, 1.0, -1.0)
    y = ops.A @ x
    ys = scatter_slab(y, n, 5, rank, world)
    parts = [torch.empty(ys.shape[0], dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(ys))
    ok_asm = np.array_equal(assemble_global([p.numpy() for p in parts], n, 5), y)
    q.put((rank, ok_rt, ok_halo, ok_asm))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 16])
def test_slab_partition_and_halo_gloo(n):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b, c in res:
        assert a and b and c, (rank, a, b, c)


def test_slab_rows_errors():
    sys.path.insert(0, ROOT)
    from mp_block_preconditioners_b200.parallel import slab_rows
    assert slab_rows(4096, 3, 8) == (1536, 512)
    with pytest.raises(ValueError):
        slab_rows(10, 0, 4)

"""The 2-D block decomposition's host logic on CPU: grid algebra, and the panel exchange over the
gloo backend with world_size 2 (the compute call itself needs CUDA and is covered by bench.py --gpus)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dmod():
    sys.path.insert(0, ROOT)
    import importlib
    import gemmul8_b200  # noqa: F401
    return importlib.import_module("gemmul8_b200.distributed")


def test_grid_shapes():
    d = _dmod()
    assert [d.grid_shape(w) for w in (1, 2, 4, 8, 6, 16)] == [(1, 1), (1, 2), (2, 2), (2, 4), (2, 3), (4, 4)]
    for world in (1, 2, 4, 8):
        seen = set()
        for r in range(world):
            gr = d.BlockGrid(world=world, rank=r)
            seen.add((gr.p, gr.q))
            assert 0 <= gr.p < gr.P and 0 <= gr.q < gr.Q
        assert len(seen) == world


def test_row_pieces_cover_the_block_on_tile_boundaries():
    d = _dmod()
    for m_loc in (1, 255, 256, 257, 300, 777, 1000, 4096, 16384, 32768, 33000):
        for want in (1, 2, 4, 8):
            pcs = d.row_pieces(m_loc, want)
            assert pcs[0][0] == 0 and pcs[-1][1] == m_loc and len(pcs) <= want
            assert all(a[1] == b[0] for a, b in zip(pcs, pcs[1:]))
            assert all(r0 % 256 == 0 and r1 > r0 for r0, r1 in pcs)


def test_ownership_partitions_cover_panels():
    d = _dmod()
    m, n, k = 64, 96, 48
    for world in (2, 4, 8):
        P, Q = d.grid_shape(world)
        for p in range(P):
            cover = []
            for q in range(Q):
                gr = d.BlockGrid(world=world, rank=p * Q + q)
                cover.append(gr.a_slice_k(k))
            assert cover == [(i * k // Q, (i + 1) * k // Q) for i in range(Q)]
        for q in range(Q):
            gr0 = d.BlockGrid(world=world, rank=q)
            n_loc = gr0.block_dims(m, n)[1]
            cover = [d.BlockGrid(world=world, rank=p * Q + q).b_slice_cols(n_loc) for p in range(P)]
            assert cover == [(i * n_loc // P, (i + 1) * n_loc // P) for i in range(P)]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = _dmod()
    grid = d.BlockGrid()
    m, n, k = 8, 12, 10
    torch.manual_seed(0)
    A = torch.randn(k, m, dtype=torch.float64)        # column-major m x k (global), identical on every rank
    B = torch.randn(n, k, dtype=torch.float64)        # column-major k x n
    m_loc, n_loc = grid.block_dims(m, n)
    klo, khi = grid.a_slice_k(k)
    clo, chi = grid.b_slice_cols(n_loc)
    a_panel_true = A[:, grid.p * m_loc:(grid.p + 1) * m_loc].contiguous()
    b_panel_true = B[grid.q * n_loc:(grid.q + 1) * n_loc].contiguous()
    a_slice = a_panel_true[klo:khi].contiguous()
    b_slice = b_panel_true[clo:chi].contiguous()
    a_panel = grid.gather_a_panel(a_slice, m_loc, k)
    b_panel = grid.gather_b_panel(b_slice, n_loc, k)
    ok = torch.equal(a_panel, a_panel_true) and torch.equal(b_panel, b_panel_true)
    # the C block this rank would compute equals the corresponding block of the global product
    c_block = b_panel @ a_panel
    c_true = (B @ A)[grid.q * n_loc:(grid.q + 1) * n_loc, grid.p * m_loc:(grid.p + 1) * m_loc]
    ok = ok and torch.allclose(c_block, c_true)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_panel_exchange_gloo(world):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))

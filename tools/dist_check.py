#!/usr/bin/env python3
"""Correctness of the multi-GPU path under torchrun: the pipelined panel exchange of distributed.pgemm (A in row
pieces read in place, products started while later pieces are in flight) must give the bits of the plain path
(gather both panels, then one gemm call) on every rank.  usage: torchrun --nproc-per-node N tools/dist_check.py [size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gemmul8_b200 as g
from importlib import import_module


def run_checks(grid, dmod, S=2048, verbose=True, pgemm=None):
    """The three comparisons, on every rank; returns a dict (ok, per-case details).  `pgemm` defaults to
    distributed.pgemm (bench.py passes the entry point it timed)."""
    pgemm = pgemm or dmod.pgemm
    rank = dist.get_rank()
    ok = True
    details = []
    for (m, n, k, N) in ((S * grid.P, S * grid.Q, S, 14), (grid.P * 1280, grid.Q * 768, 4 * 640, 9)):
        m_loc, n_loc = grid.block_dims(m, n)
        klo, khi = grid.a_slice_k(k)
        clo, chi = grid.b_slice_cols(n_loc)
        a_slice = g.phi_matrix(m_loc, khi - klo, 0.5, torch.float64, seed=100 + 17 * rank)
        b_slice = g.phi_matrix(k, chi - clo, 0.5, torch.float64, seed=200 + 17 * rank)
        work = torch.empty(g.workSize(m_loc, n_loc, k, N), dtype=torch.uint8, device="cuda")
        C1 = torch.zeros((n_loc, m_loc), dtype=torch.float64, device="cuda")
        C2 = torch.full_like(C1, 7.0)
        pgemm(grid, g, m, n, k, 1.0, a_slice, b_slice, 0.0, C1, N, True, work)                            # pipelined
        dmod.pgemm(grid, g, m, n, k, 1.0, a_slice, b_slice, 0.0, C2, N, True, work, flags=g.FLAG_TIMERS)  # plain
        torch.cuda.synchronize()
        same = torch.equal(C1, C2) and bool(C1.abs().sum() > 0)
        # and against the mathematics: a sampled double-double product of the gathered panels
        a_panel = grid.gather_a_panel(a_slice, m_loc, k)
        b_panel = grid.gather_b_panel(b_slice, n_loc, k)
        rows = torch.arange(0, m_loc, max(1, m_loc // 64), dtype=torch.int32, device="cuda")
        cols = torch.arange(0, n_loc, max(1, n_loc // 64), dtype=torch.int32, device="cuda")
        T1, T2 = g.dd_gemm(m_loc, n_loc, k, a_panel, m_loc, b_panel, k, rows=rows, cols=cols)
        err = (((C1[cols.long()][:, rows.long()] - T1) - T2) / T1).abs().max().item()
        good = same and err < (1e-6 if N >= 14 else 1e-3)      # 9 moduli carry ~1e-5 by construction
        if verbose:
            print(f"rank {rank} grid {grid.P}x{grid.Q} block ({grid.p},{grid.q}) m={m} n={n} k={k} N={N}: identical={same} relerr_max={err:.3e}", flush=True)
        ok = ok and good
        # accurate mode: the partitioned call (bound maxima combined over the grid row / column) against the
        # UNPARTITIONED accurate product of the full matrices, computed here on every rank: its block, bit for bit
        C3 = torch.zeros_like(C1)
        pgemm(grid, g, m, n, k, 1.0, a_slice, b_slice, 0.0, C3, N, False, work)
        if grid.P > 1:
            parts = [torch.empty_like(a_panel) for _ in range(grid.P)]
            dist.all_gather(parts, a_panel.contiguous(), group=grid.col_group)
            A_full = torch.cat(parts, dim=1).contiguous()              # (k, m): column-major m x k
        else:
            A_full = a_panel
        if grid.Q > 1:
            parts = [torch.empty_like(b_panel) for _ in range(grid.Q)]
            dist.all_gather(parts, b_panel.contiguous(), group=grid.row_group)
            B_full = torch.cat(parts, dim=0).contiguous()              # (n, k): column-major k x n
        else:
            B_full = b_panel
        Cf = torch.zeros((n, m), dtype=torch.float64, device="cuda")
        wf = torch.empty(g.workSize(m, n, k, N), dtype=torch.uint8, device="cuda")
        g.gemm(None, 0, 0, m, n, k, 1.0, A_full, m, B_full, k, 0.0, Cf, m, N, False, wf)
        torch.cuda.synchronize()
        blk = Cf[grid.q * n_loc:(grid.q + 1) * n_loc, grid.p * m_loc:(grid.p + 1) * m_loc]
        acc_same = torch.equal(C3, blk) and bool(C3.abs().sum() > 0)
        if verbose:
            print(f"rank {rank} accurate mode, block of the unpartitioned product: identical={acc_same}", flush=True)
        ok = ok and acc_same
        details.append({"m": m, "n": n, "k": k, "moduli": N, "pipelined_equals_plain": bool(same), "relerr_max_vs_dd": err,
                        "accurate_block_equals_unpartitioned": bool(acc_same)})
        del Cf, wf, A_full, B_full, work
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    return {"ok": flag.item() == 0, "ranks": dist.get_world_size(), "cases_rank0": details}


def cpp_pgemm(mpgrid):
    """distributed.pgemm's signature on top of the C ABI (gemmul8_b200_pgemm): what bench.py times for N > 1."""
    def f(grid, pkg, m, n, k, alpha, a_slice, b_slice, beta, Cm, N, fast, work, flags=0):
        m_loc = m // grid.P
        return mpgrid.pgemm(m, n, k, alpha, a_slice, m_loc, b_slice, k, beta, Cm, m_loc, N, fast, work, flags)
    return f


def complex_check(grid, mg):
    """ZGEMM through the C ABI pgemm (fast mode, Karatsuba): this rank's block against one gemm call on the gathered panels."""
    KARA = 3
    m, n, k, N = grid.P * 512, grid.Q * grid.P * 256, grid.Q * 384, 14
    m_loc, n_loc = grid.block_dims(m, n)
    klo, khi = grid.a_slice_k(k)
    clo, chi = grid.b_slice_cols(n_loc)
    rank = dist.get_rank()
    a_slice = g.phi_matrix(m_loc, khi - klo, 0.5, torch.complex128, seed=300 + 17 * rank)
    b_slice = g.phi_matrix(k, chi - clo, 0.5, torch.complex128, seed=400 + 17 * rank)
    work = torch.empty(g.workSize(m_loc, n_loc, k, N, KARA), dtype=torch.uint8, device="cuda")
    C1 = torch.zeros((n_loc, m_loc), dtype=torch.complex128, device="cuda")
    mg.pgemm(m, n, k, 1.0, a_slice, m_loc, b_slice, k, 0.0, C1, m_loc, N, True, work, computeType=KARA)
    a_panel = grid.gather_a_panel(a_slice, m_loc, k)
    b_panel = grid.gather_b_panel(b_slice, n_loc, k)
    C2 = torch.zeros_like(C1)
    g.gemm(None, 0, 0, m_loc, n_loc, k, 1.0, a_panel, m_loc, b_panel, k, 0.0, C2, m_loc, N, True, work, computeType=KARA)
    torch.cuda.synchronize()
    same = torch.equal(torch.view_as_real(C1), torch.view_as_real(C2)) and bool(C1.abs().sum() > 0)
    flag = torch.tensor([0 if same else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(f"complex128 Karatsuba through pgemm: identical={flag.item() == 0}", flush=True)
    return flag.item() == 0


def make_mp_grid(grid, a_bytes, b_bytes, want="copy"):
    """The C++ grid with the copy-engine exchange, or (if CUDA IPC / peer access is unavailable, or asked for) NCCL."""
    mp = import_module("gemmul8_b200.mp")
    if want == "copy":
        ok = torch.tensor([1], device="cuda")
        try:
            mg = mp.Grid(grid.P, grid.Q, a_bytes, b_bytes, mp.EXCHANGE_COPY)
        except Exception as e:       # every rank must take the same decision
            mg, ok[0] = None, 0
            print(f"rank {dist.get_rank()}: copy-engine exchange unavailable ({e})", flush=True)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 1:
            return mg, "copy engines over CUDA IPC peer memory (cudaMemcpy2DAsync pushes + stream memory-op flags)"
        if mg is not None:
            mg.close()
    return mp.Grid(grid.P, grid.Q, a_bytes, b_bytes, mp.EXCHANGE_NCCL), "NCCL all-gather / broadcast of FP64 panel pieces"


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    which = sys.argv[2] if len(sys.argv) > 2 else "all"        # python | nccl | copy | all
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dmod = import_module("gemmul8_b200.distributed")
    grid = dmod.BlockGrid()
    ok = True
    for name in (["python", "nccl", "copy"] if which == "all" else [which]):
        if name == "python":
            res = run_checks(grid, dmod, S)
        else:
            cap = 16 * max(S * S, 1280 * 4 * 640, 4 * 640 * 768)
            mg, how = make_mp_grid(grid, cap, cap, name)
            res = run_checks(grid, dmod, S, pgemm=cpp_pgemm(mg))
            # the same grid again: epochs, acknowledgements and buffer reuse across calls
            res2 = run_checks(grid, dmod, S // 2, verbose=False, pgemm=cpp_pgemm(mg))
            res["ok"] = res["ok"] and res2["ok"] and complex_check(grid, mg)
            mg.close()
            if dist.get_rank() == 0:
                print(f"C ABI pgemm, exchange: {how}", flush=True)
        if dist.get_rank() == 0:
            print(f"DIST {name}: " + ("OK" if res["ok"] else "FAILED"), flush=True)
        ok = ok and res["ok"]
    if dist.get_rank() == 0:
        print("DIST OK" if ok else "DIST FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

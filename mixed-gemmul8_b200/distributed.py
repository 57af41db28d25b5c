"""2-D block decomposition of C across the GPUs of one node (NEW relative to the reference, which is
single-GPU).

The emulation shards with no data-path reduction: shifts are per row of op(A) over all k and per
column of op(B) over all k, residues and the CRT are element-wise, so rank (p, q) of a P x Q grid
only needs row panel p of A and column panel q of B, and K is never split.  The only exchange is
the panel distribution: every rank starts with a 1/Q slice of its A row panel and a 1/P slice of
its B column panel (HPL-like ownership), and the panels are assembled with NCCL all-gathers over
NVLink inside grid rows / grid columns.  Each rank then runs the single-GPU path on its C block
(FP64 panels travel at 8 B/element; shipping 14 int8 slices would cost 14 B/element, and the
redundant encode is ~1% of the GEMM time).

`BlockGrid` works on the gloo backend too (CPU tests exercise the index algebra and the collectives
with world_size 2); the compute call itself needs CUDA.
"""
import torch
import torch.distributed as dist


def grid_shape(world):
    """P x Q with P <= Q and P*Q == world (1x1, 1x2, 2x2, 2x4, ...)."""
    p = 1
    while (p * 2) * (p * 2) <= world:
        p *= 2
    while world % p:
        p -= 1
    return p, world // p


class BlockGrid:
    def __init__(self, world=None, rank=None):
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        self.P, self.Q = grid_shape(self.world)
        self.p, self.q = divmod(self.rank, self.Q)
        self.row_group = self.col_group = None
        if dist.is_initialized() and self.world > 1:
            # every rank must create every group, in the same order
            for pp in range(self.P):
                grp = dist.new_group([pp * self.Q + qq for qq in range(self.Q)])
                if pp == self.p:
                    self.row_group = grp
            for qq in range(self.Q):
                grp = dist.new_group([pp * self.Q + qq for pp in range(self.P)])
                if qq == self.q:
                    self.col_group = grp

    # -- ownership -------------------------------------------------------------------------
    def block_dims(self, m, n):
        """Rows / columns of the C block of this rank (m, n must divide evenly)."""
        assert m % self.P == 0 and n % self.Q == 0
        return m // self.P, n // self.Q

    def a_slice_k(self, k):
        """k-range [lo, hi) of A's row panel p that this rank owns before the exchange."""
        assert k % self.Q == 0
        w = k // self.Q
        return self.q * w, (self.q + 1) * w

    def b_slice_cols(self, n_loc):
        """column range [lo, hi) of B's column panel q (n_loc columns) owned before the exchange."""
        assert n_loc % self.P == 0
        w = n_loc // self.P
        return self.p * w, (self.p + 1) * w

    # -- panel exchange --------------------------------------------------------------------
    def gather_a_panel(self, a_slice, m_loc, k, async_op=False):
        """a_slice: (k/Q, m_loc) tensor = column-major m_loc x k/Q block.  Returns the (k, m_loc) panel
        (and the NCCL work handle when async_op)."""
        if self.Q == 1:
            return (a_slice, None) if async_op else a_slice
        out = torch.empty((k, m_loc), dtype=a_slice.dtype, device=a_slice.device)
        w = dist.all_gather_into_tensor(out, a_slice.contiguous(), group=self.row_group, async_op=async_op)
        return (out, w) if async_op else out

    def gather_b_panel(self, b_slice, n_loc, k, async_op=False):
        """b_slice: (n_loc/P, k) tensor = column-major k x n_loc/P block.  Returns the (n_loc, k) panel."""
        if self.P == 1:
            return (b_slice, None) if async_op else b_slice
        out = torch.empty((n_loc, k), dtype=b_slice.dtype, device=b_slice.device)
        w = dist.all_gather_into_tensor(out, b_slice.contiguous(), group=self.col_group, async_op=async_op)
        return (out, w) if async_op else out


def pgemm(grid, pkg, m, n, k, alpha, a_slice, b_slice, beta, c_block, num_moduli, fastmode, work, flags=0):
    """C_block(p,q) = alpha * A_panel(p) * B_panel(q) + beta * C_block, ops N/N, all column-major.

    a_slice / b_slice are this rank's pre-exchange pieces (see BlockGrid); returns the phase timers."""
    m_loc, n_loc = grid.block_dims(m, n)
    overlap = fastmode and a_slice.is_cuda and (flags & ~pkg.FLAG_PHASE_LOG) == 0 and not a_slice.is_complex()
    if not overlap:
        a_panel = grid.gather_a_panel(a_slice, m_loc, k)
        b_panel = grid.gather_b_panel(b_slice, n_loc, k)
        if not fastmode and not a_slice.is_complex() and a_slice.is_cuda and grid.world > 1 and k > 0:
            # Accurate mode: the shift of a row of A comes from the maximum of its bound-product row over ALL n columns,
            # i.e. over the Q blocks of this grid row (and the shift of a column of B over the P blocks of its grid
            # column).  Bound product of the own block, one int32 max-all-reduce per direction, then the rest of the
            # call: same shifts, same bits of C as the unpartitioned product.
            pkg.gemm(None, 0, 0, m_loc, n_loc, k, alpha, a_panel, m_loc, b_panel, k, beta, c_block, m_loc,
                     num_moduli, False, work, flags=pkg.FLAG_ONLY_BOUND)
            L = pkg.work_layout(m_loc, n_loc, k, num_moduli)
            rowmax = work[L.off_A8i + L.sizeA:L.off_A8i + L.sizeA + 4 * m_loc].view(torch.int32)
            colmax = work[L.off_B8i + L.sizeB:L.off_B8i + L.sizeB + 4 * n_loc].view(torch.int32)
            if grid.Q > 1:
                dist.all_reduce(rowmax, op=dist.ReduceOp.MAX, group=grid.row_group)
            if grid.P > 1:
                dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=grid.col_group)
            return pkg.gemm(None, 0, 0, m_loc, n_loc, k, alpha, a_panel, m_loc, b_panel, k, beta, c_block, m_loc,
                            num_moduli, False, work, flags=flags | pkg.FLAG_SKIP_BOUND)
        return pkg.gemm(None, 0, 0, m_loc, n_loc, k, alpha, a_panel, m_loc, b_panel, k, beta, c_block, m_loc,
                        num_moduli, fastmode, work, flags=flags)
    # The row and the column communicator are different NCCL communicators with their own streams, so the
    # A and B exchanges run at the same time; A travels as up to four row pieces, and everything that depends
    # only on what has arrived is computed while the rest is still in flight:
    #     A rows of piece 0, B panel  ->  scale A_0, scale B, C[piece 0, :]      (later pieces still arriving)
    #     A rows of piece i           ->  scale A_i, C[piece i, :]
    # A piece is gathered into its own (k, rows) tensor -- a column-major rows x k matrix with leading dimension
    # `rows` -- and the block-wise entry reads it in place: the args of piece i carry lda = rows and an A pointer
    # moved back by r0 elements, so that "row r of the full panel" (r0 <= r < r1, all that SCALE_A touches)
    # addresses row r - r0 of the piece.  No copy into a contiguous panel.
    import ctypes
    pieces = row_pieces(m_loc, 4 if grid.Q > 1 else 1)
    es = a_slice.element_size()
    pending = []
    if grid.Q == 1:
        pending.append((0, m_loc, None, a_slice))
    else:
        for (r0, r1) in pieces:
            piece = torch.empty((k, r1 - r0), dtype=a_slice.dtype, device=a_slice.device)
            w = dist.all_gather_into_tensor(piece, a_slice[:, r0:r1].contiguous(), group=grid.row_group, async_op=True)
            pending.append((r0, r1, w, piece))
    b_panel, wb = grid.gather_b_panel(b_slice, n_loc, k, async_op=True)
    first = True
    for (r0, r1, w, piece) in pending:
        args = pkg.make_args(0, 0, m_loc, n_loc, k, alpha, piece, r1 - r0, b_panel, k, beta, c_block, m_loc, num_moduli, fastmode, work,
                             flags=flags)
        args.A = ctypes.c_void_p(piece.data_ptr() - r0 * es)
        if w is not None:
            w.wait()                                          # the compute stream waits; the host does not block
        pkg.gemm_part(args, pkg.PART_SCALE_A, r0, r1, 0, 0)
        if first:
            if wb is not None:
                wb.wait()
            pkg.gemm_part(args, pkg.PART_SCALE_B, 0, 0, 0, n_loc)
            first = False
        pkg.gemm_part(args, pkg.PART_PRODUCT, r0, r1, 0, n_loc)
    return [0.0, 0.0, 0.0, 0.0]


def row_pieces(m_loc, want):
    """Up to `want` row ranges of [0, m_loc) whose starts are multiples of 256 (whole GEMM tiles)."""
    tiles = (m_loc + 255) // 256
    n = max(1, min(want, tiles))
    cuts = sorted({min(m_loc, (tiles * i // n) * 256) for i in range(n)} | {m_loc})
    return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]

"""Tensor-parallel sharding of the FP4 GEMM the way Llama-3.3-70B TP shards it.

The reference has no distributed code: TP lives in its callers,
which shard the checkpoint, repack the local shard and all-reduce row-parallel
outputs.  Its benchmark list already contains the TP=8 shard shapes
(tools/benchmarks/matmul.py:18-25).  This module is that caller-side logic:

* column-parallel (qkv, gate_up): split N -> independent GEMMs, no collective;
* row-parallel (o_proj, down_proj): split K (on scale-group boundaries) -> partial
  [M, N] outputs summed by ONE all-reduce (NCCL over NVLink on GPUs).

The slicing helpers are pure tensor indexing (usable on CPU, e.g. under gloo in the
tests); the Linear classes call the CUDA ops of ``petit_kernel`` and nothing else.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

# Llama-3.3-70B decoder-layer GEMMs: name -> (N, K, parallel kind)
LLAMA70B_LAYER = {
    "qkv": (10240, 8192, "column"),
    "o": (8192, 8192, "row"),
    "gate_up": (57344, 8192, "column"),
    "down": (8192, 28672, "row"),
}


def shard_shape(n: int, k: int, kind: str, tp: int) -> tuple[int, int]:
    """Per-rank (N, K) of a layer; N splits stay multiples of 64, K splits of 256."""
    if kind == "column":
        assert n % (tp * 64) == 0, "N shard must be a multiple of 64"
        return n // tp, k
    assert k % (tp * 256) == 0, "K shard must be a multiple of 256"
    return n, k // tp


def column_shard(q_u8: torch.Tensor, scales: torch.Tensor, tp: int, rank: int):
    """N-split: rows [rank*N/tp, (rank+1)*N/tp) of weights [N, K/2] and scales [N, K/g]."""
    n = q_u8.shape[0]
    lo, hi = rank * n // tp, (rank + 1) * n // tp
    return q_u8[lo:hi].contiguous(), scales[lo:hi].contiguous()


def row_shard(q_u8: torch.Tensor, scales: torch.Tensor, tp: int, rank: int):
    """K-split: k in [rank*K/tp, (rank+1)*K/tp); the cut falls on scale-group
    boundaries because K/tp is a multiple of 256."""
    kb, kg = q_u8.shape[1], scales.shape[1]
    return (q_u8[:, rank * kb // tp:(rank + 1) * kb // tp].contiguous(),
            scales[:, rank * kg // tp:(rank + 1) * kg // tp].contiguous())


def row_shard_activation(a: torch.Tensor, tp: int, rank: int) -> torch.Tensor:
    k = a.shape[1]
    return a[:, rank * k // tp:(rank + 1) * k // tp].contiguous()


def all_reduce_sum(c: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the row-parallel layers."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return c


class SymmAllReduce:
    """One-shot all-reduce of small row-parallel outputs over NVLink peer memory
    (torch symmetric memory).  The GEMM writes its partial [M, N] straight into a
    symmetric buffer and `torch.ops.symm_mem.one_shot_all_reduce` reads all peers'
    buffers in one kernel -- ~3x lower latency than ncclAllReduce at these sizes
    (16 KiB .. 1 MiB).  Falls back to NCCL when symmetric memory is unavailable."""

    def __init__(self, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.bufs = {}
        self.ok = True
        try:
            import torch.distributed._symmetric_memory as symm_mem

            self.symm_mem = symm_mem
        except Exception:  # pragma: no cover
            self.ok = False

    def buffer(self, m: int, n: int, dtype, device, slot: int = 0) -> torch.Tensor:
        key = (m, n, dtype, slot)
        if key not in self.bufs:
            buf = self.symm_mem.empty((m, n), dtype=dtype, device=device)
            self.symm_mem.rendezvous(buf, self.group.group_name)
            self.bufs[key] = buf
        return self.bufs[key]

    def reduce(self, buf: torch.Tensor) -> torch.Tensor:
        return torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", self.group.group_name)


class PeerAllReduce:
    """One-shot all-reduce of small row-parallel outputs with this package's own kernel
    (csrc/allreduce.cu) over torch symmetric memory: the GEMM writes its partial [M, N]
    into a peer-mapped buffer (`buffer()`), `reduce()` launches ONE kernel that exchanges
    per-CTA flags with the peers, sums all ranks' buffers through NVLink in rank order
    (bit-identical on every rank) and releases the buffer.  ~5 us of host time per call
    (no NCCL, no dispatcher), chained to the GEMMs with programmatic dependent launch.

    Invariant of the default (``end_barrier=False``) mode: the two buffers of a slot
    alternate per ``buffer()`` call, and at least one ``reduce()`` on a buffer's pad lies
    between a read of that buffer and its next overwrite.  That holds when every
    ``buffer()`` is followed by its ``reduce()`` and the calls run eagerly.  Under CUDA-graph
    capture the buffer choice is frozen into the graph: construct with ``end_barrier=True``
    (or pass it to ``reduce``) when a captured graph holds fewer than two reduces per slot.
    ``fenced=True`` (or env PETIT_AR_FENCED=1) selects the release/acquire flag protocol."""

    def __init__(self, group=None, end_barrier: bool = False, fenced: bool = False):
        import petit_kernel as pk  # CUDA extension; no fallback
        import torch.distributed._symmetric_memory as symm_mem

        self.pk = pk
        self.symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.slots = {}
        self.by_ptr = {}
        self.end_barrier = end_barrier
        self.fenced = fenced

    def status(self) -> int:
        """0 if every reduce so far met all its peers; 1 + missing rank otherwise
        (synchronises the current stream)."""
        worst = 0
        for entry in self.by_ptr.values():
            worst = max(worst, int(self.pk.ops.allreduce_status(entry[3])))
        return worst

    def _alloc(self, m: int, n: int, dtype, device):
        pad_elems = self.pk.ops.allreduce_pad_bytes() // 2
        numel = m * n
        data_elems = (numel + 127) // 128 * 128  # keep the pad 256-byte aligned
        raw = self.symm_mem.empty((data_elems + pad_elems,), dtype=dtype, device=device)
        raw.zero_()
        hdl = self.symm_mem.rendezvous(raw, self.group.group_name)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every pad is zero before anyone signals
        esz = raw.element_size()
        bufs = [int(hdl.buffer_ptrs[r]) for r in range(self.world)]
        pads = [b + data_elems * esz for b in bufs]
        epoch = torch.zeros(self.pk.ops.allreduce_epoch_bytes() // 4, dtype=torch.int32,
                            device=device)
        view = raw[:numel].view(m, n)
        self.by_ptr[view.data_ptr()] = (view, bufs, pads, epoch, raw, hdl)
        return view

    def buffer(self, m: int, n: int, dtype, device, slot: int = 0) -> torch.Tensor:
        """The buffer the next GEMM of this (shape, slot) must write.  Two buffers alternate
        per call, which is what lets `reduce` skip the closing peer barrier: call it once
        per GEMM, on every rank in the same order."""
        key = (m, n, dtype, slot)
        if key not in self.slots:
            self.slots[key] = [[self._alloc(m, n, dtype, device) for _ in range(2)], 0]
        pair = self.slots[key]
        pair[1] ^= 1
        return pair[0][pair[1]]

    def reduce(self, buf: torch.Tensor, out: "torch.Tensor | None" = None,
               end_barrier: "bool | None" = None) -> torch.Tensor:
        entry = self.by_ptr.get(buf.data_ptr())
        if entry is None:
            raise ValueError("buf was not allocated by PeerAllReduce.buffer()")
        view, bufs, pads, epoch = entry[:4]
        if out is None:
            out = torch.empty_like(view)
        eb = self.end_barrier if end_barrier is None else end_barrier
        if torch.cuda.is_current_stream_capturing():
            eb = True  # the buffer parity is frozen into the graph
        return self.pk.ops.allreduce_oneshot(out, bufs, pads, epoch, self.rank, view.numel(),
                                             eb, self.fenced)


class FusedAllReduce:
    """Row-parallel GEMM fused with the all-reduce of its output (csrc/fp4_gemm.cu, "fused
    all-reduce"): `matmul()` launches ONE kernel per rank; the CTA that finishes an output tile
    pushes its 16-bit partial into every peer's receive buffer over NVLink and sums the
    peers' packets of that tile in rank order -- no collective launch, no barrier kernel.
    One receive buffer (torch symmetric memory) per output width n; m <= 64."""

    MAX_TOKENS = 64

    def __init__(self, group=None):
        import petit_kernel as pk  # CUDA extension; no fallback
        import torch.distributed._symmetric_memory as symm_mem

        self.pk = pk
        self.symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.ctx = {}

    def _context(self, n: int, device, slot: int):
        key = (n, slot)
        if key not in self.ctx:
            nbytes = self.pk.ops.fused_allreduce_recv_bytes(n)
            raw = self.symm_mem.empty((nbytes,), dtype=torch.uint8, device=device)
            raw.zero_()
            hdl = self.symm_mem.rendezvous(raw, self.group.group_name)
            state = torch.zeros(self.pk.ops.fused_allreduce_state_bytes() // 4, dtype=torch.int32,
                                device=device)
            torch.cuda.synchronize(device)
            dist.barrier(self.group)  # every receive buffer is zero before anyone sends
            ptrs = [int(hdl.buffer_ptrs[r]) for r in range(self.world)]
            self.ctx[key] = (ptrs, state, raw, hdl)
        return self.ctx[key]

    def matmul(self, a, b, s, global_scale, n: int, k: int, fmt: str = "nvfp4",
               out: "torch.Tensor | None" = None, slot: int = 0, solution_id: int = -1):
        """sum over ranks of a_r @ dequant(b_r)^T * global_scale; `slot` separates layers of the
        same width that are in flight in the same step."""
        m = a.shape[0]
        assert m <= self.MAX_TOKENS, "fused all-reduce handles at most 64 tokens"
        ptrs, state = self._context(n, a.device, slot)[:2]
        if out is None:
            out = torch.empty((m, n), dtype=a.dtype, device=a.device)
        return self.pk.ops.mul_fp4_a16_allreduce_out(out, a, b, s, global_scale, m, n, k,
                                                     solution_id, fmt == "mxfp4", ptrs, state,
                                                     self.rank)

    def status(self) -> int:
        worst = 0
        for entry in self.ctx.values():
            worst = max(worst, int(self.pk.ops.fused_allreduce_status(entry[1])))
        return worst


@dataclass
class PackedLinear:
    """One FP4 linear layer resident on the current CUDA device."""
    b: torch.Tensor
    s: torch.Tensor
    global_scale: torch.Tensor
    n: int
    k: int
    fmt: str          # "nvfp4" | "mxfp4"
    kind: str         # "column" | "row"

    @classmethod
    def from_native(cls, q_u8, scales, global_scale, fmt: str, kind: str):
        import petit_kernel as pk  # CUDA extension; no fallback

        n, k = q_u8.shape[0], q_u8.shape[1] * 2
        qw = q_u8.cuda().contiguous().view(torch.int32)
        if fmt == "nvfp4":
            b, s = pk.repack_nvfp4(qw, n, k), pk.process_nvfp4_scales(scales.cuda(), n, k)
        else:
            b, s = pk.repack_mxfp4(qw, n, k), pk.process_mxfp4_scales(scales.cuda(), n, k)
        return cls(b, s, global_scale.cuda(), n, k, fmt, kind)

    def forward(self, a: torch.Tensor, group=None, reduce: bool = True,
                symm: "SymmAllReduce | None" = None, slot: int = 0) -> torch.Tensor:
        import petit_kernel as pk

        m = a.shape[0]
        if self.kind == "row" and reduce and isinstance(symm, FusedAllReduce):
            return symm.matmul(a, self.b, self.s, self.global_scale, self.n, self.k, self.fmt,
                               slot=slot)
        if self.kind == "row" and reduce and isinstance(symm, PeerAllReduce):
            out = symm.buffer(m, self.n, a.dtype, a.device, slot)
            mul_out = (pk.ops.mul_nvfp4_a16_out if self.fmt == "nvfp4"
                       else pk.ops.mul_mxfp4_a16_out)
            mul_out(out, a, self.b, self.s, self.global_scale, m, self.n, self.k, -1)
            return symm.reduce(out)
        if self.kind == "row" and reduce and symm is not None and symm.ok:
            out = symm.buffer(m, self.n, a.dtype, a.device, slot)
            mul_out = (pk.ops.mul_nvfp4_a16_out if self.fmt == "nvfp4"
                       else pk.ops.mul_mxfp4_a16_out)
            mul_out(out, a, self.b, self.s, self.global_scale, m, self.n, self.k, -1)
            return symm.reduce(out)
        mul = pk.mul_nvfp4_a16 if self.fmt == "nvfp4" else pk.mul_mxfp4_a16
        c = mul(a, self.b, self.s, self.global_scale, m, self.n, self.k, -1)
        if self.kind == "row" and reduce:
            all_reduce_sum(c, group)
        return c

"""Solution tuning for the FP4 GEMM: the `-algo tune` loop of the reference's benchmark
(`tools/benchmarks/matmul/main.cc:269-325`) as a library call, plus the tuned-solution
table that feeds its result back to `solution_id=-1` (`petit.h: petit_tune_table_*`).

The reference tells callers to autotune (`README.md:35`) and its framework glue carries a
"TODO: use auto-tuning to find the performant solution_id"; here the winner is stored in
the library, so `mul_*_a16(..., solution_id=-1)` picks it up without any caller change:

    import petit_kernel.tuning as tuning
    tuning.tune_gemm(a, b, s, global_scale, m, n, k)        # times every solution, records the best
    tuning.save_table(path) ... PETIT_TUNE_TABLE=path       # next process: loaded on first use

Timing uses CUDA events on the current stream (never the host clock).  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops

_entries: dict[tuple, int] = {}  # mirror of what this process put into the library's table


def _key(size_m, size_n, size_k, a_type, mx):
    return ("mxfp4" if mx else "nvfp4", "bf16" if a_type == torch.bfloat16 else "fp16",
            int(size_m), int(size_n), int(size_k))


def solution_hex(solution_id: int) -> str:
    """The id as `bench_matmul` prints it: its 8 bytes, little endian."""
    return (int(solution_id) & (2**64 - 1)).to_bytes(8, "little").hex()


def default_solution(size_m: int, size_n: int, size_k: int, a_type: torch.dtype,
                     mx: bool = False) -> int:
    """What `solution_id=-1` resolves to (table entry if present, else the built-in rule)."""
    return ops.get_default_solution(size_m, size_n, size_k, a_type, mx)


def set_solution(size_m: int, size_n: int, size_k: int, a_type: torch.dtype, mx: bool,
                 solution_id: int) -> None:
    """Pin the solution of one problem (-1 removes the entry)."""
    ops.tune_table_set(size_m, size_n, size_k, a_type, mx, solution_id)
    if solution_id == -1:
        _entries.pop(_key(size_m, size_n, size_k, a_type, mx), None)
    else:
        _entries[_key(size_m, size_n, size_k, a_type, mx)] = int(solution_id)


def load_table(path: str) -> int:
    """Read a table written by `save_table` or `bench_matmul -algo tune -table`."""
    return ops.tune_table_load(path)


def clear_table() -> None:
    ops.tune_table_clear()
    _entries.clear()


def save_table(path: str) -> int:
    """Write the entries recorded by this process (tune_gemm / set_solution)."""
    with open(path, "w") as f:
        f.write("# petit tuned solutions: <btype> <atype> m n k <solution id, 8 bytes little endian>\n")
        for (bt, at, m, n, k), sol in sorted(_entries.items()):
            f.write(f"{bt} {at} {m} {n} {k} {solution_hex(sol)}\n")
    return len(_entries)


def time_solution(a, b, s, global_scale, size_m, size_n, size_k, solution_id, mx=False,
                  warmup=3, repeat=20) -> float:
    """Microseconds per call of one solution, CUDA events on the current stream."""
    mul = ops.mul_mxfp4_a16_out if mx else ops.mul_nvfp4_a16_out
    out = torch.empty((size_m, size_n), dtype=a.dtype, device=a.device)
    for _ in range(warmup):
        mul(out, a, b, s, global_scale, size_m, size_n, size_k, solution_id)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeat):
        mul(out, a, b, s, global_scale, size_m, size_n, size_k, solution_id)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / repeat


def tune_gemm(a, b, s, global_scale, size_m, size_n, size_k, mx=False, warmup=3, repeat=20,
              record=True):
    """Time every solution `get_fp4_solutions` lists for this problem on the caller's own
    tensors; returns [(us, solution_id), ...] fastest first and (record=True) makes the
    fastest one the default for this exact (types, m, n, k)."""
    if not a.is_cuda:
        raise RuntimeError("tune_gemm needs CUDA tensors (there is no CPU path)")
    b_type = ops.CDataType.kDataTypeMxFp4e2m1 if mx else ops.CDataType.kDataTypeFp4e2m1
    sols = ops.get_fp4_solutions(size_m, size_n, size_k, a.dtype, a.dtype, b_type=int(b_type))
    results = sorted((time_solution(a, b, s, global_scale, size_m, size_n, size_k, sol, mx,
                                    warmup, repeat), int(sol)) for sol in sols)
    if record and results:
        set_solution(size_m, size_n, size_k, a.dtype, mx, results[0][1])
    return results

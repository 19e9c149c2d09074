"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/causalflow/petit/petit.h declares, the host-only entry points behave
like the reference's, and the Python surface has the reference's signatures.
No compute call needs a GPU here."""
import ctypes
import inspect
import os
import re
import subprocess

import pytest
import torch

from helpers import HEADER, ROOT, ensure_built


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(ensure_built())


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(petit_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for s in ["petit_gemm_nvfp4_a16", "petit_gemm_mxfp4_a16", "petit_get_solutions",
              "petit_repack_fp4_weights", "petit_repack_nvfp4_scales",
              "petit_repack_mxfp4_scales", "petit_dequant_nvfp4", "petit_dequant_mxfp4",
              "petit_dequant_packed_nvfp4", "petit_dequant_packed_mxfp4",
              "petit_hal_malloc", "petit_hal_synchronize"]:
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in petit.h but not exported"


class Hints(ctypes.Structure):
    _fields_ = [("a_type", ctypes.c_int32), ("b_type", ctypes.c_int32),
                ("c_type", ctypes.c_int32), ("require_high_precision", ctypes.c_int32)]


FP16, BF16, FP4, MXFP4, INT4 = 4, 5, 3, 7, 0


def solutions(lib, hints, m, n, k):
    cnt = ctypes.c_uint(0)
    rc = lib.petit_get_solutions(ctypes.byref(hints), m, n, k, None, ctypes.byref(cnt))
    if rc != 0:
        return rc, []
    buf = (ctypes.c_uint64 * max(cnt.value, 1))()
    cap = ctypes.c_uint(cnt.value)
    rc = lib.petit_get_solutions(ctypes.byref(hints), m, n, k, buf, ctypes.byref(cap))
    return rc, list(buf)[:cap.value]


def test_get_solutions_two_call_protocol(lib):
    # fp4/algo_chooser.cc:14-62
    rc, sols = solutions(lib, Hints(BF16, FP4, BF16, 0), 16, 4096, 4096)
    assert rc == 0 and len(sols) == 5 and len(set(sols)) == 5
    lib.petit_solution_name.restype = ctypes.c_char_p
    names = [lib.petit_solution_name(ctypes.c_uint64(s)).decode() for s in sols]
    assert all(n.startswith("sm100_streamk_nvfp4_bf16_tok") for n in names)
    # SolutionId bit layout (gemm.h:33-105): features = Grid, element_b = NvFp4, mfma = bf16
    for s in sols:
        assert (s >> 24) & 0xF == 1 and (s >> 28) & 0xF == 1 and (s >> 32) & 0xF == 1
    rc, sols_mx = solutions(lib, Hints(BF16, MXFP4, BF16, 0), 16, 4096, 4096)
    assert rc == 0 and all((s >> 28) & 0xF == 2 for s in sols_mx)
    # MXFP4 has no fp16 kernels (gemm_fp4_fp16_grid.cc:60-63)
    assert solutions(lib, Hints(FP16, MXFP4, FP16, 0), 16, 4096, 4096) == (0, [])
    # shapes the layout cannot hold enumerate nothing
    assert solutions(lib, Hints(BF16, FP4, BF16, 0), 16, 4096, 4096 + 128) == (0, [])
    assert solutions(lib, Hints(BF16, FP4, BF16, 0), 16, 4096 + 8, 4096) == (0, [])
    # unsupported b_type -> -1 (algo_chooser.cc:20-23)
    assert solutions(lib, Hints(BF16, INT4, BF16, 0), 16, 4096, 4096)[0] == -1
    # capacity smaller than the count truncates but still reports the count
    buf = (ctypes.c_uint64 * 2)()
    cap = ctypes.c_uint(2)
    h = Hints(BF16, FP4, BF16, 0)
    assert lib.petit_get_solutions(ctypes.byref(h), 1, 256, 256, buf, ctypes.byref(cap)) == 0
    assert cap.value == 5


def test_gemm_zero_sized_problem_is_a_noop_without_a_device(lib):
    # gemm_fp4_fp16_grid.cc:42-44 -- returns before touching CUDA
    h = Hints(BF16, FP4, BF16, 0)
    for m, n, k in [(0, 128, 256), (16, 0, 256), (16, 128, 0)]:
        rc = lib.petit_gemm_nvfp4_a16(None, None, None, None, None, m, n, k, ctypes.byref(h),
                                      ctypes.c_uint64(2**64 - 1), None)
        assert rc == 0


def test_error_codes_on_host_side_checks(lib):
    h = Hints(BF16, FP4, BF16, 0)
    auto = ctypes.c_uint64(2**64 - 1)
    # k not a multiple of 256 / n not a multiple of 16 -> kErrorProblemShape (1)
    assert lib.petit_gemm_nvfp4_a16(None, None, None, None, None, 16, 128, 384,
                                    ctypes.byref(h), auto, None) == 1
    assert lib.petit_gemm_nvfp4_a16(None, None, None, None, None, 16, 120, 256,
                                    ctypes.byref(h), auto, None) == 1
    # unknown solution id -> kErrorKernelShape (2)
    assert lib.petit_gemm_nvfp4_a16(None, None, None, None, None, 16, 128, 256,
                                    ctypes.byref(h), ctypes.c_uint64(12345), None) == 2
    # MXFP4 with fp16 activations -> kErrorKernelShape (gemm_fp4_fp16_grid.cc:60-63)
    h16 = Hints(FP16, MXFP4, FP16, 0)
    assert lib.petit_gemm_mxfp4_a16(None, None, None, None, None, 16, 128, 256,
                                    ctypes.byref(h16), auto, None) == 2
    # dense hooks: bad type / shape -> -1 (quantization_utils.cu:619-621,642)
    assert lib.petit_dequant_mxfp4(None, None, None, ctypes.c_float(1.0), FP16, 256, 128, None) == -1
    assert lib.petit_dequant_nvfp4(None, None, None, ctypes.c_float(1.0), BF16, 100, 128, None) == -1
    assert lib.petit_repack_fp4_weights(None, None, 100, 128, None) == 1
    assert lib.petit_packed_layout_version() >= 1


def test_python_surface_matches_reference_signatures():
    import petit_kernel as pk

    expect = {
        "repack_nvfp4": ["qw", "size_n", "size_k"],
        "repack_mxfp4": ["qw", "size_n", "size_k"],
        "process_nvfp4_scales": ["scales", "size_n", "size_k"],
        "process_mxfp4_scales": ["scales", "size_n", "size_k"],
        "mul_nvfp4_a16": ["a", "b", "s", "global_scale", "size_m", "size_n", "size_k", "solution_id"],
        "mul_mxfp4_a16": ["a", "b", "s", "global_scale", "size_m", "size_n", "size_k", "solution_id"],
        "get_fp4_solutions": ["size_m", "size_n", "size_k", "a_type", "c_type"],
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(pk, name)).parameters) == params
    assert set(pk.__all__) == set(expect) | {"DataType", "PetitSolutionHints"}
    # reference __init__.py:8-15
    assert {d.name: d.value for d in pk.DataType} == {
        "int4": 0, "float8_e4m3fn": 1, "float4_e2m1": 2, "float16": 3, "bfloat16": 4,
        "float8_e5m2fn": 5, "mxfloat4_e2m1": 6}
    for n in ["repack_nvfp4", "process_nvfp4_scales", "process_mxfp4_scales", "mul_nvfp4_a16",
              "mul_mxfp4_a16", "get_nvfp4_solutions", "get_fp4_solutions", "PetitSolutionHints"]:
        assert hasattr(pk.ops, n)  # pybind.cc:9-25
    h = pk.PetitSolutionHints()
    h.a_type, h.b_type, h.c_type, h.require_high_precision = 5, 3, 5, False
    assert len(pk.ops.get_fp4_solutions(h, 16, 4096, 4096)) == 5
    assert len(pk.get_fp4_solutions(16, 4096, 4096, torch.bfloat16, torch.bfloat16)) == 5


def test_ops_reject_cpu_tensors_no_fallback():
    import petit_kernel as pk

    q = torch.zeros((128, 32), dtype=torch.int32)
    with pytest.raises(RuntimeError, match="not on GPU"):
        pk.repack_nvfp4(q, 128, 256)
    s = torch.zeros((128, 16), dtype=torch.float8_e4m3fn)
    with pytest.raises(RuntimeError, match="not on GPU"):
        pk.process_nvfp4_scales(s, 128, 256)
    a = torch.zeros((4, 256), dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="not on GPU"):
        pk.mul_nvfp4_a16(a, q, s, torch.ones(1), 4, 128, 256, -1)
    with pytest.raises(RuntimeError, match="size_k"):
        pk.repack_nvfp4(q, 128, 200)
    with pytest.raises(RuntimeError, match="groupsize = 16"):
        pk.process_nvfp4_scales(torch.zeros((128, 8), dtype=torch.float8_e4m3fn), 128, 256)


def test_bench_matmul_cli_flag_errors():
    """tools/bench_matmul keeps the reference CLI's flags and messages
    (tools/benchmarks/matmul/main.cc:17-35,336-358); argument errors need no GPU."""
    exe = os.path.join(ROOT, "tools", "bench_matmul")
    if not os.path.exists(exe):
        pytest.skip("bench_matmul not built")
    r = subprocess.run([exe, "-backend", "hipblaslt"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown backend: hipblaslt" in r.stderr
    r = subprocess.run([exe, "-btype", "fp16"], capture_output=True, text=True)
    assert r.returncode == 1 and "Invalid b type for backend 'petit'" in r.stderr
    r = subprocess.run([exe, "-m", "16", "-n", "100", "-k", "256"], capture_output=True, text=True)
    assert r.returncode == 1 and "k % 256" in r.stderr
    r = subprocess.run([exe, "-atype", "bf16", "-ctype", "fp16"], capture_output=True, text=True)
    assert r.returncode == 1

"""CPU oracle for the FP4-weight x 16-bit-activation GEMM path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product (``petit_kernel`` / ``libpetit_b200.so``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module, and only as the checker / the
timed CPU baseline.

What it restates (all citations into /root/reference):

* NVFP4 dequantise + GEMM reference: ``tests/ops/test_fp4_gemm_quark.py:9-24``
  (``_dequant_nvfp4``: LUT order, low nibble = even k, group 16 along K,
  ``scales.float()``; ``_gemm_ref``: fp32 matmul then cast) and ``:51-54``
  (global scale folded into the fp32 weights before the matmul).
* MXFP4 dequantise: the reference test delegates to AMD Quark ``dq_mxfp4``
  (``tests/ops/test_fp4_gemm_quark.py:66-69,83``; dependency ``amd-quark``,
  unpinned in ``pyproject.toml:33-36``, NOT installed here and not vendored).
  We restate the algorithm the reference's own kernels implement:
  ``DequantTraitMxFp4::GetScale`` (``quantization_utils.cu:405-432``) and
  ``DequantizerForE8M0Scale`` e8m0 -> bf16 as ``(s & 0xff) << 7``
  (``dequant.cuh:197-203``), i.e. ``w = LUT[e2m1] * 2**(s-127)`` with group 32
  along K, and ``c = (a @ w.T) * global_scale`` (``test_fp4_gemm_quark.py:87``).
  Tested domain is s in [1, 237] (``quantization_utils_fp4_test.cc:266-278``).
* e4m3 decode used by the exhaustive dequant test:
  ``lib/tests/floating_points.h:21-75`` (fp8_e4m3_t::to_fp32).
* Dense dequant hooks (``DequantizeFp4Kernel``, ``quantization_utils.cu:542-612``):
  16-bit weight times ``Element(global_scale)`` with a 16-bit multiply.
* Synthetic input recipes: ``tests/ops/test_fp4_gemm_quark.py:41-46,71-76`` and
  ``lib/tests/quantization.cc:77-143`` (value ranges; mt19937(42)).
* The C++ GEMM matcher ``IsNearBf16/IsNearFp16``
  (``gemm_fp4_fp16_rocm_test.cc:31-67``).

Pinning status: the NVFP4 functions are checked bit-for-bit against the
reference's own Python oracle executed in the build container
(``tests/golden/make_golden.py`` imports it from /root/reference and commits
the vectors).  The MXFP4 functions are pinned to implementations this repository
did not write: AMD Quark itself is absent from the image, so
``tests/golden/make_golden_mx.py`` builds the expectations of the reference test's
recipe (``test_fp4_gemm_quark.py:71-87``) from torch's own OCP e8m0 dtype
(``uint8.view(torch.float8_e8m0fnu)``) and compressed-tensors' e2m1 decoder
(``unpack_fp4_from_uint8``), cross-checked against compressed-tensors'
``decompress_mx_scale``; ``tests/test_oracle.py`` requires this module to reproduce
``tests/golden/mxfp4_independent.npz`` bit for bit (dequantised weights incl. the
exhaustive 16 x 237 table and the reference's scale-mixing pattern) and the GEMM
vectors within one bf16 ulp.  What stays unpinned is only Quark's own code path.
"""
from __future__ import annotations

import numpy as np
import torch

# tests/ops/test_fp4_gemm_quark.py:10-14 and quantization_utils_fp4_test.cc:259-262
E2M1_VALUES = np.array(
    [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0,
     -0.0, -0.5, -1.0, -1.5, -2.0, -3.0, -4.0, -6.0],
    dtype=np.float32,
)

NVFP4_GROUP = 16
MXFP4_GROUP = 32


# --------------------------------------------------------------------------
# scalar format decoders
# --------------------------------------------------------------------------
def e4m3_to_f32(bits: np.ndarray) -> np.ndarray:
    """OCP e4m3fn byte -> float32 (lib/tests/floating_points.h:21-75).

    sign(1) exp(4, bias 7) mant(3); exp==0 is subnormal (mant/8 * 2**-6);
    0x7f/0xff are NaN; there is no infinity.
    """
    bits = np.asarray(bits, dtype=np.uint8).astype(np.int32)
    sign = np.where(bits & 0x80, -1.0, 1.0).astype(np.float32)
    exp = (bits >> 3) & 0xF
    man = bits & 0x7
    normal = np.ldexp((8 + man).astype(np.float32), exp - 7 - 3)
    sub = np.ldexp(man.astype(np.float32), -6 - 3)
    out = np.where(exp == 0, sub, normal).astype(np.float32)
    out = np.where((exp == 0xF) & (man == 0x7), np.float32(np.nan), out)
    return (sign * out).astype(np.float32)


def e8m0_to_f32(bits: np.ndarray) -> np.ndarray:
    """e8m0 byte -> float32 the way the reference decodes it:
    bf16 bit pattern ``(s & 0xff) << 7`` (dequant.cuh:197-203).

    Consequences outside the tested domain [1, 237]: s == 0 gives 0.0 (OCP says
    2**-127) and s == 255 gives +inf (OCP says NaN).
    """
    bits = np.asarray(bits, dtype=np.uint8).astype(np.uint32)
    as_f32 = (bits << np.uint32(7 + 16)).astype(np.uint32)
    return as_f32.view(np.float32)


def unpack_e2m1(q_u8: np.ndarray) -> np.ndarray:
    """[N, K/2] bytes -> [N, K] float32; low nibble is the even k
    (tests/ops/test_fp4_gemm_quark.py:15-19)."""
    q_u8 = np.asarray(q_u8, dtype=np.uint8)
    n, kh = q_u8.shape
    out = np.empty((n, kh * 2), dtype=np.float32)
    out[:, 0::2] = E2M1_VALUES[q_u8 & 0x0F]
    out[:, 1::2] = E2M1_VALUES[q_u8 >> 4]
    return out


# --------------------------------------------------------------------------
# dequantise
# --------------------------------------------------------------------------
def dequant_nvfp4(q_u8: np.ndarray, scales_e4m3: np.ndarray) -> np.ndarray:
    """[N,K/2] u8 + [N,K/16] e4m3 bytes -> [N,K] float32
    (tests/ops/test_fp4_gemm_quark.py:9-20)."""
    w = unpack_e2m1(q_u8)
    n, k = w.shape
    s = e4m3_to_f32(scales_e4m3).reshape(n, k // NVFP4_GROUP, 1)
    return (w.reshape(n, -1, NVFP4_GROUP) * s).reshape(n, k).astype(np.float32)


def dequant_mxfp4(q_u8: np.ndarray, scales_e8m0: np.ndarray) -> np.ndarray:
    """[N,K/2] u8 + [N,K/32] e8m0 bytes -> [N,K] float32 =
    LUT * 2**(s-127) (quantization_utils.cu:405-432, dequant.cuh:197-203)."""
    w = unpack_e2m1(q_u8)
    n, k = w.shape
    s = e8m0_to_f32(scales_e8m0).reshape(n, k // MXFP4_GROUP, 1)
    with np.errstate(invalid="ignore", over="ignore"):
        return (w.reshape(n, -1, MXFP4_GROUP) * s).reshape(n, k).astype(np.float32)


def _torch_dtype(name_or_dtype) -> torch.dtype:
    if isinstance(name_or_dtype, torch.dtype):
        return name_or_dtype
    return {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
            "fp16": torch.float16, "float16": torch.float16}[name_or_dtype]


def dense16(w_f32: np.ndarray, global_scale: float, dtype) -> torch.Tensor:
    """What the reference's dense dequant hooks return
    (DequantizeFp4Kernel, quantization_utils.cu:563-585): the exact 16-bit
    weight multiplied by ``Element(global_scale)`` with a 16-bit multiply."""
    dt = _torch_dtype(dtype)
    w16 = torch.from_numpy(np.ascontiguousarray(w_f32)).to(dt)
    gs16 = torch.tensor(global_scale, dtype=torch.float32).to(dt)
    return (w16.float() * gs16.float()).to(dt)


# --------------------------------------------------------------------------
# GEMM references
# --------------------------------------------------------------------------
def gemm_ref(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tests/ops/test_fp4_gemm_quark.py:23-24."""
    return (a.float() @ b.t().float()).to(a.dtype)


def nvfp4_gemm_ref(a: torch.Tensor, q_u8: torch.Tensor, scales_e4m3: torch.Tensor,
                   global_scale: torch.Tensor) -> torch.Tensor:
    """tests/ops/test_fp4_gemm_quark.py:51-53: global scale is folded into the
    fp32 weights, then fp32 matmul, then cast to a.dtype."""
    w = dequant_nvfp4(q_u8.cpu().numpy(), scales_e4m3.cpu().view(torch.uint8).numpy())
    b_ref = torch.from_numpy(w) * float(global_scale.item())
    return gemm_ref(a.cpu(), b_ref)


def nvfp4_gemm_ref_torch(a: torch.Tensor, q_u8: torch.Tensor, scales_e4m3: torch.Tensor,
                         global_scale: torch.Tensor) -> torch.Tensor:
    """Same computation as nvfp4_gemm_ref with torch ops only (multi-threaded on the
    host), statement for statement what tests/ops/test_fp4_gemm_quark.py:9-24,51-53
    does; used as the timed CPU baseline."""
    lut = torch.from_numpy(E2M1_VALUES)
    q_u8 = q_u8.cpu()
    lo = q_u8 & 0x0F
    hi = q_u8 >> 4
    deq = torch.empty((q_u8.size(0), q_u8.size(1) * 2), dtype=torch.float32)
    deq[:, 0::2] = lut[lo.long()]
    deq[:, 1::2] = lut[hi.long()]
    w = (deq.view(q_u8.size(0), -1, NVFP4_GROUP) * scales_e4m3.cpu().float().unsqueeze(-1)
         ).view(q_u8.size(0), -1)
    b_ref = w * float(global_scale.item())
    return gemm_ref(a.cpu(), b_ref)


def mxfp4_gemm_ref(a: torch.Tensor, q_u8: torch.Tensor, scales_e8m0: torch.Tensor,
                   global_scale: torch.Tensor) -> torch.Tensor:
    """tests/ops/test_fp4_gemm_quark.py:83-87: dequantise to bf16, fp32 matmul,
    multiply by global_scale, cast."""
    w = dequant_mxfp4(q_u8.cpu().numpy(), scales_e8m0.cpu().numpy())
    b = torch.from_numpy(w).to(torch.bfloat16)
    a = a.cpu()
    return ((a.float() @ b.t().float()) * float(global_scale.item())).to(a.dtype)


# --------------------------------------------------------------------------
# matchers
# --------------------------------------------------------------------------
def is_near_cpp(out: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """Element-wise IsNearBf16 / IsNearFp16
    (gemm_fp4_fp16_rocm_test.cc:31-67): |a-b| < max(1e-2, 0.01*|b|)."""
    a, b = out.float().cpu(), ref.float().cpu()
    return (a - b).abs() < torch.clamp(b.abs() * 0.01, min=1e-2)


def max_rel_err(out: torch.Tensor, ref_f32: torch.Tensor) -> float:
    """north_star tolerance: max |c - ref| / max |ref| against the fp32
    accumulation of the dequantised weights."""
    a, b = out.float().cpu(), ref_f32.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def bits_equal_pm0(x: torch.Tensor, y: torch.Tensor) -> bool:
    """Bit-exact comparison that treats +0 and -0 as equal and NaN as unequal
    (lib/tests/floating_points.h:184-202)."""
    assert x.dtype == y.dtype and x.shape == y.shape
    xi = x.cpu().contiguous().view(torch.int16)
    yi = y.cpu().contiguous().view(torch.int16)
    same = xi == yi
    both_zero = ((xi & 0x7FFF) == 0) & ((yi & 0x7FFF) == 0)
    nan = torch.isnan(x.cpu().float()) | torch.isnan(y.cpu().float())
    return bool(((same | both_zero) & ~nan).all())


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def make_nvfp4_case(m: int, n: int, k: int, seed: int, dtype=torch.bfloat16):
    """Recipe of tests/ops/test_fp4_gemm_quark.py:41-46, generated with the CPU
    generator so that the same tensors exist on every machine (the reference
    draws them with the device generator, whose stream is backend-specific)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn((m, k), generator=g, dtype=torch.float32).to(dtype)
    q = torch.randint(0, 256, (n, k // 2), generator=g, dtype=torch.uint8)
    s = (torch.rand((n, k // NVFP4_GROUP), generator=g) * 3.5 + 0.25).to(torch.float8_e4m3fn)
    gs = torch.rand((1,), generator=g, dtype=torch.float32) * 1.5 + 0.5
    return a, q, s, gs


def make_mxfp4_case(m: int, n: int, k: int, seed: int, dtype=torch.bfloat16):
    """Recipe of tests/ops/test_fp4_gemm_quark.py:71-76 (scales in [1, 237])."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn((m, k), generator=g, dtype=torch.float32).to(dtype)
    q = torch.randint(0, 256, (n, k // 2), generator=g, dtype=torch.uint8)
    s = torch.randint(1, 238, (n, k // MXFP4_GROUP), generator=g, dtype=torch.uint8)
    gs = torch.rand((1,), generator=g, dtype=torch.float32) * 1.5 + 0.5
    return a, q, s, gs


def make_gtest_style_case(m: int, n: int, k: int, fmt: str = "nvfp4",
                          dtype=torch.bfloat16, seed: int = 42):
    """Distributions of lib/tests/quantization.cc:77-143 (mt19937(seed)):
    A ~ U(-2,2) truncated to bf16 (U(-1,1) for fp16), every u32 of weights
    uniformly random, NV scales uniform over the positive e4m3 bit patterns
    0x01..0x7E, MX scales uniform in [1, 237], global_scale = 1.
    (The C++ stream itself is libstdc++-specific and is not reproduced.)"""
    rs = np.random.RandomState(seed)
    if dtype == torch.bfloat16:
        af = rs.uniform(-2.0, 2.0, size=(m, k)).astype(np.float32)
        a = torch.from_numpy((af.view(np.uint32) >> 16).astype(np.uint16).view(np.int16)
                             ).view(torch.bfloat16)
    else:
        a = torch.from_numpy(rs.uniform(-1.0, 1.0, size=(m, k)).astype(np.float32)).to(dtype)
    q = torch.from_numpy(rs.randint(0, 2 ** 32, size=(n, k // 8), dtype=np.uint32
                                    ).view(np.uint8).reshape(n, k // 2).copy())
    if fmt == "nvfp4":
        s = torch.from_numpy(rs.randint(0x01, 0x7F, size=(n, k // NVFP4_GROUP)
                                        ).astype(np.uint8)).view(torch.float8_e4m3fn)
    else:
        s = torch.from_numpy(rs.randint(1, 238, size=(n, k // MXFP4_GROUP)).astype(np.uint8))
    gs = torch.ones((1,), dtype=torch.float32)
    return a, q, s, gs

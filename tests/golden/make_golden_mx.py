"""MXFP4 golden vectors from implementations this repository did NOT write.

The reference's MXFP4 test delegates to AMD Quark's dq_mxfp4
(tests/ops/test_fp4_gemm_quark.py:57-88), which is not installed anywhere in this image
(vLLM 0.22's own dequant_mxfp4 wraps the same missing package).  What the image does carry:

  * torch 2.11's OCP e8m0 dtype: uint8.view(torch.float8_e8m0fnu).float() = 2^(s-127)
    (0 -> 2^-127, 255 -> NaN) -- the scale decode;
  * compressed-tensors 0.15 (the checkpoint tooling vLLM loads MXFP4/NVFP4 models with):
    compressors/nvfp4/helpers.py::unpack_fp4_from_uint8 -- e2m1 decode, low nibble first --
    and compressors/mx_utils.py::decompress_mx_scale -- 2^(s-127) in bf16.

This script builds the expectations from those two, in the reference test's own recipe
(group 32 along K, b_dequant in bf16, c_ref = ((a.float() @ b.t().float()) * gs).to(a.dtype),
:83-87), cross-checks the two scale decoders against each other on the reference's tested
domain s in [1, 237] (quantization_utils_fp4_test.cc:266-278) and writes
tests/golden/mxfp4_independent.npz.  tests/test_oracle.py then requires oracle/petit_oracle.py
to reproduce every vector bit for bit; the GPU tests compare the CUDA path against the same file.

    python tests/golden/make_golden_mx.py        (CPU only; needs compressed_tensors)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import petit_oracle as orc  # noqa: E402  (input generator only: make_mxfp4_case)

from compressed_tensors.compressors.mx_utils import decompress_mx_scale  # noqa: E402
from compressed_tensors.compressors.nvfp4.helpers import unpack_fp4_from_uint8  # noqa: E402

MXFP4_CASES = [(64, 128, 256, 1234), (96, 96, 512, 2026)]  # test_fp4_gemm_quark.py:32-35


def dq_mxfp4_independent(q_u8: torch.Tensor, s_u8: torch.Tensor) -> torch.Tensor:
    """[N, K/2] packed e2m1 + [N, K/32] e8m0 -> [N, K] bf16, nothing from oracle/."""
    n, kb = q_u8.shape
    vals = unpack_fp4_from_uint8(q_u8.contiguous(), n, kb * 2, dtype=torch.float32)
    scale = s_u8.contiguous().view(torch.float8_e8m0fnu).float()            # torch's decoder
    w = (vals.view(n, -1, 32) * scale.unsqueeze(-1)).view(n, kb * 2)
    return w.to(torch.bfloat16)


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def main():
    # the two independent scale decoders agree on the reference's tested domain
    sb = torch.arange(1, 238, dtype=torch.uint8)
    a_dec = sb.view(torch.float8_e8m0fnu).float()
    b_dec = decompress_mx_scale(sb).float()
    assert torch.equal(a_dec, b_dec), "torch e8m0 and compressed-tensors disagree on [1, 237]"

    out = {}
    for (m, n, k, seed) in MXFP4_CASES:
        a, q, s, gs = orc.make_mxfp4_case(m, n, k, seed)
        w = dq_mxfp4_independent(q, s)
        c_ref = ((a.float() @ w.t().float()) * gs.item()).to(a.dtype)
        tag = f"m{m}_n{n}_k{k}_s{seed}"
        out[f"{tag}_c"] = bits16(c_ref)
        out[f"{tag}_w"] = bits16(w)
    # exhaustive: all 16 codes x e8m0 1..237 (MxFp4DequantTest, quantization_utils_fp4_test.cc:273-278)
    qn = torch.from_numpy(np.repeat((np.arange(16, dtype=np.uint8) * 0x11)[:, None], 237 * 16, axis=1).copy())
    sn = torch.from_numpy(np.tile(np.arange(1, 238, dtype=np.uint8), (16, 1)).copy())
    w = dq_mxfp4_independent(qn, sn).view(16, 237, 32)[:, :, 0]
    out["exhaustive_bf16_bits"] = bits16(w)
    # the reference's scale mixing pattern (col + 29 * row) % 237 + 1 on a 64 x 512 slab
    rows, cols = 64, 512
    g = torch.Generator().manual_seed(4242)
    qm = torch.randint(0, 256, (rows, cols // 2), generator=g, dtype=torch.uint8)
    gi = torch.arange(cols // 32).unsqueeze(0) + 29 * torch.arange(rows).unsqueeze(1)
    sm = (gi % 237 + 1).to(torch.uint8)
    out["mix_q"] = qm.numpy().copy()
    out["mix_s"] = sm.numpy().copy()
    out["mix_w_bits"] = bits16(dq_mxfp4_independent(qm, sm))
    # outside the tested domain: what the independent decoders say about s = 0 and 255
    edge = torch.tensor([0, 255], dtype=torch.uint8).view(torch.float8_e8m0fnu).float()
    out["edge_scale_f32_bits"] = edge.view(torch.int32).numpy().view(np.uint32).copy()
    np.savez_compressed(os.path.join(HERE, "mxfp4_independent.npz"), **out)
    print("wrote mxfp4_independent.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE'S OWN
ORACLE (tests/ops/test_fp4_gemm_quark.py:9-24) in the build container.

Run once in the container that has /root/reference mounted:
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests only read the committed .npz files.

What is pinned by reference code executed here:
  * nvfp4_gemm_cases.npz  -- c_ref of the two NVFP4 cases (:27-30) on the inputs of
    oracle.make_nvfp4_case (CPU generator), via the reference's _dequant_nvfp4 /
    _gemm_ref, plus the fp32 dequantised weights' checksum.
  * nvfp4_exhaustive.npz  -- all 16 codes x e4m3 0x01..0x7E through the reference's
    _dequant_nvfp4 (the table ExhaustiveFp4DequantTest checks,
    quantization_utils_fp4_test.cc:240-264,344-365).
What is NOT pinned by runnable reference code (AMD Quark is absent): the MXFP4
vectors in mxfp4_cases.npz come from oracle.dequant_mxfp4 / mxfp4_gemm_ref, i.e.
from the restatement itself; they guard against regressions only.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import petit_oracle as orc  # noqa: E402

REF_TEST = "/root/reference/tests/ops/test_fp4_gemm_quark.py"


def load_reference_oracle():
    stub = types.ModuleType("petit_kernel")
    sys.modules["petit_kernel"] = stub
    ns = {"__name__": "ref_test"}
    with open(REF_TEST) as f:
        exec(compile(f.read(), REF_TEST, "exec"), ns)
    del sys.modules["petit_kernel"]
    return ns


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def main():
    ref = load_reference_oracle()
    out = {}
    for (m, n, k, seed) in ref["NVFP4_CASES"]:
        a, q, s, gs = orc.make_nvfp4_case(m, n, k, seed)
        b_ref = ref["_dequant_nvfp4"](q, s) * gs.item()
        c_ref = ref["_gemm_ref"](a, b_ref)
        w = ref["_dequant_nvfp4"](q, s)
        tag = f"m{m}_n{n}_k{k}_s{seed}"
        out[f"{tag}_c"] = bits16(c_ref)
        out[f"{tag}_wsum"] = np.array([w.double().sum().item(), w.double().abs().sum().item()])
        out[f"{tag}_w_row0"] = w[0].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "nvfp4_gemm_cases.npz"), **out)

    # exhaustive: row = code (byte = code | code << 4), one group per scale value
    scale_bits = np.arange(0x01, 0x7F, dtype=np.uint8)
    q = torch.from_numpy(np.repeat((np.arange(16, dtype=np.uint8) * 0x11)[:, None],
                                   len(scale_bits) * 8, axis=1).copy())
    s = torch.from_numpy(np.tile(scale_bits, (16, 1)).copy()).view(torch.float8_e4m3fn)
    w = ref["_dequant_nvfp4"](q, s).view(16, len(scale_bits), 16)[:, :, 0]
    np.savez_compressed(os.path.join(HERE, "nvfp4_exhaustive.npz"),
                        table=w.numpy().copy(), scale_bits=scale_bits)

    out = {}
    for (m, n, k, seed) in ref["MXFP4_CASES"]:
        a, q, s, gs = orc.make_mxfp4_case(m, n, k, seed)
        c_ref = orc.mxfp4_gemm_ref(a, q, s, gs)
        out[f"m{m}_n{n}_k{k}_s{seed}_c"] = bits16(c_ref)
    sb = np.arange(1, 238, dtype=np.uint8)
    qn = np.repeat((np.arange(16, dtype=np.uint8) * 0x11)[:, None], len(sb) * 16, axis=1)
    w = orc.dequant_mxfp4(qn, np.tile(sb, (16, 1))).reshape(16, len(sb), 32)[:, :, 0]
    out["exhaustive_table"] = w
    np.savez_compressed(os.path.join(HERE, "mxfp4_cases.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

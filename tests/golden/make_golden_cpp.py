"""Generate tests/golden/nvfp4_exhaustive_cpp.npz by RUNNING THE REFERENCE'S OWN C++ HOST CODE:
oracle/_ref/ref_numeric is oracle/ref_numeric_driver.cc compiled against
/root/reference/lib/tests/floating_points.h (+ lib/gemm/cpu/half_float.h) -- `make -C oracle ref`.

It holds exactly what the reference's ExhaustiveFp4DequantTest expects
(quantization_utils_fp4_test.cc:240-264,344-365,388-394): for all 16 e2m1 codes x e4m3 bits
0x01..0x7E, `Element::from_fp32(fp8_e4m3_t(s).to_fp32() * fp4_values[q])` as bf16 and fp16 bits,
plus the fp32 product and the reference's e4m3 -> fp32 table for all 256 bytes.

Run once in the container that has /root/reference mounted:
    make -C oracle ref && python tests/golden/make_golden_cpp.py
The GPU box has no /root/reference; tests only read the committed .npz.
"""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_numeric")


def run_reference_driver(exe=EXE):
    """Parse the driver's output into arrays (also used by tests/test_oracle.py)."""
    text = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    e4m3 = np.zeros(256, dtype=np.uint32)
    lo = hi = None
    deq = {}
    eq = None
    for line in text.splitlines():
        f = line.split()
        if f[0] == "range":
            lo, hi = int(f[1]), int(f[2])
        elif f[0] == "e4m3":
            e4m3[int(f[1])] = int(f[2], 16)
        elif f[0] == "deq":
            deq[(int(f[1]), int(f[2]))] = (int(f[3], 16), int(f[4], 16), int(f[5], 16))
        elif f[0] == "eq":
            eq = [int(x) for x in f[1:]]
    n = hi - lo + 1
    f32 = np.zeros((16, n), dtype=np.uint32)
    bf16 = np.zeros((16, n), dtype=np.uint16)
    fp16 = np.zeros((16, n), dtype=np.uint16)
    for (q, s), (a, b, c) in deq.items():
        f32[q, s - lo], bf16[q, s - lo], fp16[q, s - lo] = a, b, c
    return {"scale_bits": np.arange(lo, hi + 1, dtype=np.uint8), "e4m3_f32_bits": e4m3,
            "product_f32_bits": f32, "bf16_bits": bf16, "fp16_bits": fp16,
            "eq_pm0_nan_same": np.array(eq, dtype=np.int32)}


if __name__ == "__main__":
    out = run_reference_driver()
    np.savez_compressed(os.path.join(HERE, "nvfp4_exhaustive_cpp.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})

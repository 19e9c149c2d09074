"""Grouped (MoE) GEMM: one launch over all experts against the experts issued back to back.
Mixtral-8x7B-like expert shapes (w13: 28672 x 4096, w2: 4096 x 14336), NVFP4 x bf16, a decode
step's tokens spread over the experts.  CUDA events, weights > 2x L2, 50 calls per variant."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "petit-kernel_b200"))
import petit_kernel as pk  # noqa: E402


def run(name, n, k, counts, reps=50):
    e = len(counts)
    g = torch.Generator(device="cuda").manual_seed(1)
    b = torch.randint(-2**31, 2**31 - 1, (e, n // 16, 2 * k), dtype=torch.int32, device="cuda", generator=g)
    s = torch.randint(0x30, 0x50, (e, n, k // 16), dtype=torch.uint8, device="cuda", generator=g).view(torch.float8_e4m3fn)
    gs = torch.ones(e, dtype=torch.float32, device="cuda")
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    a = torch.randn(offsets[-1], k, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(offsets[-1], n, dtype=torch.bfloat16, device="cuda")
    live = sum(1 for c in counts if c)
    bytes_ = live * (n * k // 2 + n * k // 16) + offsets[-1] * (k + n) * 2
    res = {"case": name, "n": n, "k": k, "counts": counts}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for single in ("1", "0"):
        os.environ["PETIT_GROUPED_SINGLE"] = single
        for _ in range(3):
            pk.ops.mul_fp4_a16_grouped_out(out, a, b, s, gs, offsets, n, k, -1, False)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for i in range(reps):
            flush.zero_()  # > L2 between calls
            ev[i][0].record()
            pk.ops.mul_fp4_a16_grouped_out(out, a, b, s, gs, offsets, n, k, -1, False)
            ev[i][1].record()
        torch.cuda.synchronize()
        us = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)[reps // 2]
        res["single_launch" if single == "1" else "back_to_back"] = {
            "us": round(us, 2), "gbs": round(bytes_ / us / 1e3, 1), "frac_hbm": round(bytes_ / us / 1e3 / 6535.7, 3)}
    os.environ.pop("PETIT_GROUPED_SINGLE", None)
    print(json.dumps(res))


if __name__ == "__main__":
    run("w13 8 experts, 16 tokens top-2", 28672, 4096, [4, 5, 3, 4, 6, 2, 4, 4])
    run("w2 8 experts, 16 tokens top-2", 4096, 14336, [4, 5, 3, 4, 6, 2, 4, 4])
    run("w13 8 experts, 1 token top-2", 28672, 4096, [0, 1, 0, 0, 0, 1, 0, 0])
    run("w13 8 experts, 64 tokens top-2", 28672, 4096, [16, 20, 12, 16, 24, 8, 16, 16])
    run("64 small experts (2048 x 2048), 32 tokens top-4", 2048, 2048, [2] * 64)

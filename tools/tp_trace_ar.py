"""Per-CTA timeline of the fused GEMM + all-reduce (run under torchrun, one GPU per rank):
where an output tile's epilogue spends its time -- partial ready -> sent -> peers' packets
received -> finished tile sent / received.  Uses the library's trace hook (globaltimer stamps,
[grid][16] u64).  tools/tp_trace_ar.py <layer: o|down> [calls]"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "petit-kernel_b200"))
import petit_kernel as pk  # noqa: E402
import petit_tp  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    layer = sys.argv[1] if len(sys.argv) > 1 else "o"
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    n, k_full = (8192, 8192) if layer == "o" else (8192, 28672)
    k, m = k_full // world, 16
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    packs = []
    for _ in range(4):
        q = torch.randint(0, 256, (n, k // 2), generator=g, dtype=torch.uint8, device=dev)
        s = (torch.rand((n, k // 16), generator=g, device=dev) * 3.5 + 0.25).to(torch.float8_e4m3fn)
        packs.append((pk.repack_nvfp4(q.view(torch.int32), n, k), pk.process_nvfp4_scales(s, n, k)))
    a = torch.randn((m, k), generator=g, device=dev).to(torch.bfloat16)
    gs = torch.ones(1, device=dev)
    far = petit_tp.FusedAllReduce()
    lib = ctypes.CDLL(os.path.join(ROOT, "petit-kernel_b200", "petit_kernel", "libpetit_b200.so"))
    trace = torch.zeros(160 * 16 + 64 * 8 + 160, dtype=torch.int64, device=dev)
    for i in range(5):
        far.matmul(a, packs[i % 4][0], packs[i % 4][1], gs, n, k)
    torch.cuda.synchronize()
    dist.barrier()
    lib.petit_debug_set_trace.argtypes = [ctypes.c_void_p]
    lib.petit_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(calls):
        far.matmul(a, packs[i % 4][0], packs[i % 4][1], gs, n, k)
    e1.record()
    torch.cuda.synchronize()
    lib.petit_debug_set_trace(ctypes.c_void_p(0))
    t = trace[:160 * 16].view(160, 16).cpu().double() / 1e3  # us
    grid = int((t[:, 0] > 0).sum())
    t = t[:grid]
    own = t[:, 13] > 0          # CTAs that reduced a tile in the last call
    other = (t[:, 14] > 0) & ~own
    def stat(x):
        return f"{x.min():6.2f} {x.mean():6.2f} {x.max():6.2f}" if len(x) else "   -"
    t0 = t[:, 0].min()
    msg = [f"rank {rank}: {e0.elapsed_time(e1) / calls * 1e3:.2f} us/call, grid {grid}, "
           f"{int(own.sum())} reducing + {int(other.sum())} receiving CTAs (last call; us min avg max)",
           f"  kernel span (first entry -> last exit)      {(t[:, 8].max() - t0):6.2f}",
           f"  partial ready, since first entry            {stat(t[own | other, 11] - t0)}",
           f"  reducer: partial ready -> all partials in   {stat(t[own, 13] - t[own, 11])}",
           f"  reducer: -> finished tile sent              {stat(t[own, 14] - t[own, 13])}",
           f"  others : partial sent -> finished tile in   {stat(t[other, 14] - t[other, 12])}",
           f"  exit since first entry                      {stat(t[:, 8] - t0)}"]
    for r in range(world):
        dist.barrier()
        if r == rank and rank in (0, world - 1):
            print("\n".join(msg), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

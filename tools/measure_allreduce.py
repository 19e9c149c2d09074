"""All-reduce cost at TP-N, alone and behind a row-parallel GEMM (run under torchrun):
    torchrun --nproc-per-node 2 tools/measure_allreduce.py"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, "petit-kernel_b200"); sys.path.insert(0, ".")
import petit_kernel as pk, petit_tp
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
m, n, k = 16, 8192, 8192 // world
q = torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device=dev)
s = (torch.rand((n, k // 16), device=dev) * 3 + 0.25).to(torch.float8_e4m3fn)
b = pk.repack_nvfp4(q.view(torch.int32), n, k); sp = pk.process_nvfp4_scales(s, n, k)
gs = torch.ones(1, device=dev); a = torch.randn(m, k, device=dev).bfloat16()
par = petit_tp.PeerAllReduce(); buf = par.buffer(m, n, torch.bfloat16, dev, 0); buf.zero_()
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16); plain = torch.empty_like(out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def bench(fn, iters=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
res = {}
res["peer AR alone"] = bench(lambda: par.reduce(buf, out=out))
res["nccl AR alone"] = bench(lambda: dist.all_reduce(plain))
res["gemm -> plain"] = bench(lambda: pk.ops.mul_nvfp4_a16_out(plain, a, b, sp, gs, m, n, k, -1))
res["gemm -> symm buf"] = bench(lambda: pk.ops.mul_nvfp4_a16_out(buf, a, b, sp, gs, m, n, k, -1))
def g_ar():
    bb = par.buffer(m, n, torch.bfloat16, dev, 0)
    pk.ops.mul_nvfp4_a16_out(bb, a, b, sp, gs, m, n, k, -1); par.reduce(bb, out=out)
def g_nccl():
    pk.ops.mul_nvfp4_a16_out(plain, a, b, sp, gs, m, n, k, -1); dist.all_reduce(plain)
res["gemm + peer AR"] = bench(g_ar)
res["gemm + nccl AR"] = bench(g_nccl)
if rank == 0:
    for k_, v in res.items(): print(f"{k_:20s} {v:8.2f} us")
dist.barrier(); dist.destroy_process_group()

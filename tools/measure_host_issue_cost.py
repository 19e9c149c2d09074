"""Host-side issue cost of one GEMM call through the Python API (run on a GPU box):
    python tools/measure_host_issue_cost.py"""
import sys, time, torch
sys.path.insert(0, "petit-kernel_b200"); sys.path.insert(0, ".")
import petit_kernel as pk
torch.cuda.set_device(0)
m, n, k = 16, 1024, 1024
q = torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, device="cuda")
s = (torch.rand((n, k // 16), device="cuda") * 3 + 0.25).to(torch.float8_e4m3fn)
b = pk.repack_nvfp4(q.view(torch.int32), n, k); sp = pk.process_nvfp4_scales(s, n, k)
gs = torch.ones(1, device="cuda"); a = torch.randn(m, k, device="cuda").bfloat16()
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
def bench(fn, iters=3000):
    for _ in range(200): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(iters): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / iters * 1e6, (t2 - t0) / iters * 1e6
print("mul_nvfp4_a16      cpu-issue us/call %.2f  incl. drain %.2f" % bench(lambda: pk.mul_nvfp4_a16(a, b, sp, gs, m, n, k, -1)))
print("mul_nvfp4_a16_out  cpu-issue us/call %.2f  incl. drain %.2f" % bench(lambda: pk.ops.mul_nvfp4_a16_out(out, a, b, sp, gs, m, n, k, -1)))
print("torch.add (ref)    cpu-issue us/call %.2f  incl. drain %.2f" % bench(lambda: torch.add(out, out, out=out)))

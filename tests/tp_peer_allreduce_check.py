"""Run under torchrun (world >= 2, one GPU per rank): petit_tp.PeerAllReduce against
ncclAllReduce on a row-parallel NVFP4 layer.  Launched by tests/test_gpu_parity.py."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "petit-kernel_b200"))
sys.path.insert(0, ROOT)

import petit_kernel as pk  # noqa: E402
import petit_tp  # noqa: E402
from oracle import petit_oracle as orc  # noqa: E402  (checker only)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    m, n, k = 16, 1024, 2048
    a, q, s, gs = orc.make_nvfp4_case(m, n, k, 77)
    # K-split (row parallel): every rank owns k / world columns
    qs, ss = petit_tp.row_shard(q, s, world, rank)
    a_s = petit_tp.row_shard_activation(a, world, rank).cuda()
    ks = k // world
    b = pk.repack_nvfp4(qs.cuda().contiguous().view(torch.int32), n, ks)
    sp = pk.process_nvfp4_scales(ss.cuda().contiguous(), n, ks)
    gsc = gs.cuda()
    par = petit_tp.PeerAllReduce()
    for it in range(5):
        ref = pk.mul_nvfp4_a16(a_s, b, sp, gsc, m, n, ks, -1)
        mine_in = par.buffer(m, n, torch.bfloat16, a_s.device, 0)
        pk.ops.mul_nvfp4_a16_out(mine_in, a_s, b, sp, gsc, m, n, ks, -1)
        got = par.reduce(mine_in)
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        # same partials, both summed in fp32?  NCCL rounds per hop; allow one bf16 ulp of the max
        err = (got.float() - ref.float()).abs().max().item()
        scale = ref.float().abs().max().item()
        assert err <= scale * 2 ** -7, (it, err, scale)
        # identical on every rank
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], x) for x in gathered)
    full = orc.nvfp4_gemm_ref_torch(a, q, s, gs)
    assert orc.max_rel_err(got.float().cpu(), full) <= 1e-2
    dist.barrier()
    if rank == 0:
        print("PEER_ALLREDUCE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host-side tensor-parallel logic on CPU with the gloo backend, world_size = 2:
N-split shards concatenate to the full result with no collective; K-split partials
summed by one all-reduce equal the full result.  The per-shard
GEMMs here are the ORACLE's (this is a test of the sharding + collective plumbing;
the GPU equivalent is tests/test_gpu_parity.py::test_tp_shards_match_single_gpu)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc


def _worker(rank, world, port, fmt, ret):
    import petit_tp as tp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, n, k = 8, 256, 1024
        if fmt == "nvfp4":
            a, q, s, gs = orc.make_nvfp4_case(m, n, k, 7)
            deq, sb = orc.dequant_nvfp4, s.view(torch.uint8)
        else:
            a, q, s, gs = orc.make_mxfp4_case(m, n, k, 7)
            s = (s % 20 + 118).to(torch.uint8)  # keep partial sums comparable in fp32
            deq, sb = orc.dequant_mxfp4, s
        a = a.float()
        full = (a @ torch.from_numpy(deq(q.numpy(), sb.numpy())).t()) * gs.item()

        # column parallel: no collective, gather only to check
        qs, ss = tp.column_shard(q, sb, world, rank)
        assert qs.shape == (n // world, k // 2)
        c_col = (a @ torch.from_numpy(deq(qs.numpy(), ss.numpy())).t()) * gs.item()
        parts = [torch.empty_like(c_col) for _ in range(world)]
        dist.all_gather(parts, c_col)
        assert torch.equal(torch.cat(parts, dim=1), full)

        # row parallel: one all-reduce
        qs, ss = tp.row_shard(q, sb, world, rank)
        a_s = tp.row_shard_activation(a, world, rank)
        assert qs.shape == (n, k // 2 // world) and a_s.shape == (m, k // world)
        c_row = (a_s @ torch.from_numpy(deq(qs.numpy(), ss.numpy())).t()) * gs.item()
        tp.all_reduce_sum(c_row)
        torch.testing.assert_close(c_row, full, rtol=1e-5, atol=1e-3 * full.abs().max().item())
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt", ["nvfp4", "mxfp4"])
def test_tp2_sharding_with_gloo(fmt):
    world = 2
    port = 29500 + os.getpid() % 1000 + (0 if fmt == "nvfp4" else 1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fmt, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_shapes_follow_llama70b_tp():
    import petit_tp as tp

    # per-rank shapes listed in tools/benchmarks/matmul.py:18-25 for TP=8
    expect8 = {"qkv": (1280, 8192), "gate_up": (7168, 8192), "o": (8192, 1024), "down": (8192, 3584)}
    for name, (n, k, kind) in tp.LLAMA70B_LAYER.items():
        assert tp.shard_shape(n, k, kind, 8) == expect8[name]
        for t in (2, 4, 8):
            sn, sk = tp.shard_shape(n, k, kind, t)
            assert sn % 64 == 0 and sk % 256 == 0

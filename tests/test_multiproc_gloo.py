"""CPU, world_size > 1 over gloo: the host-side multi-rank logic (rank plumbing used by bench.py,
process-group layout used by the FSDP drivers).  The CUDA kernels themselves are single-GPU; the
data path has no collective of its own."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ucf_vit_b200.utils import dist_utils
from ucf_vit_b200.utils.misc import init_par_groups, par_group_layout


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, ret):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = dist_utils.init_distributed("gloo")
    try:
        ret[rank] = fn(r, w)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _run(world, fn):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def _timing_and_sharding(rank, world):
    lo, hi = dist_utils.shard_range(10, rank, world)
    mx = dist_utils.max_over_ranks([float(rank + 1), 5.0 - rank])
    sm = dist_utils.sum_over_ranks([float(hi - lo)])
    # data-parallel gradient averaging as DDP performs it (the only collective next to the path)
    g = torch.full((4,), float(rank))
    dist.all_reduce(g)
    return (lo, hi), mx, sm, (g / world).tolist()


def test_rank_plumbing_world2():
    out = _run(2, _timing_and_sharding)
    assert out[0][0] == (0, 5) and out[1][0] == (5, 10)
    assert out[0][1] == out[1][1] == [2.0, 5.0]          # max over ranks
    assert out[0][2] == out[1][2] == [10.0]              # shards cover the global batch exactly
    assert out[0][3] == out[1][3] == [0.5] * 4


def _groups(rank, world):
    seq, ddp, tp, ort, fsdp, sddp = init_par_groups(rank, data_par_size=4, tensor_par_size=1, seq_par_size=1,
                                                    fsdp_size=2, simple_ddp_size=2)
    def members(g):
        t = torch.zeros(world)
        t[rank] = 1
        dist.all_reduce(t, group=g)
        return [i for i in range(world) if t[i] > 0]
    return members(ddp), members(fsdp), members(sddp), members(tp)


def test_process_groups_world4_match_reference_layout():
    out = _run(4, _groups)
    for r in range(4):
        ddp, fsdp, sddp, tp = out[r]
        assert ddp == [0, 1, 2, 3]
        assert fsdp == ([0, 1] if r < 2 else [2, 3])             # contiguous FSDP shards
        assert sddp == ([0, 2] if r % 2 == 0 else [1, 3])        # strided replicas
        assert tp == [r]


def test_group_layout_pure():
    lay = par_group_layout(data_par_size=2, tensor_par_size=2, seq_par_size=1, fsdp_size=2, simple_ddp_size=1)
    assert [r for k, r in lay if k == "tensor"] == [[0, 1], [2, 3]]
    assert [r for k, r in lay if k == "ddp"] == [[0, 2], [1, 3]]
    assert [r for k, r in lay if k == "fsdp"] == [[0, 2], [1, 3]]
    assert [r for k, r in lay if k == "data_seq_ort"] == [[0, 2], [1, 3]]
    assert dist_utils.shard_range(7, 0, 3) == (0, 3) and dist_utils.shard_range(7, 2, 3) == (5, 7)

"""GPU (-m gpu): the modules inside torch FSDP exactly as the reference's FSDP driver wraps them
(/root/reference/training_scripts/train_masked_fsdp.py:361-396): auto-wrap on {Block, Sequential},
bf16 MixedPrecision (parameters arrive in the kernels as bf16 views of the flat parameter),
activation checkpointing on Block (forward kernels re-run inside backward), AdamW.
world_size = min(2, #GPUs): FULL_SHARD over NCCL when two GPUs are visible, a single-rank group
otherwise (same code path through FSDP's flat parameters)."""
import functools
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(seed=0):
    from ucf_vit_b200.fsdp.arch import MAE
    torch.manual_seed(seed)
    return MAE(img_size=[64, 64], patch_size=8, in_chans=3, embed_dim=128, depth=3, num_heads=2,
               decoder_embed_dim=64, decoder_depth=1, decoder_num_heads=2, mlp_ratio=4, mlp_ratio_decoder=4,
               mask_ratio=0.75, linear_decoder=False, class_token=False, weight_init="skip", twoD=True,
               default_vars=["r", "g", "b"], adaptive_patching=False, tensor_par_size=1, tensor_par_group=None)


def _worker(rank, world, port, ret):
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import (apply_activation_checkpointing,
                                                                             checkpoint_wrapper)
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp import MixedPrecision, ShardingStrategy
    from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy
    from torch.nn import Sequential
    from ucf_vit_b200.fsdp.building_blocks import Block
    from ucf_vit_b200.utils.misc import configure_optimizer, patchify
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        model = _build().cuda()
        ref_first = None
        policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={Block, Sequential})
        mp_policy = MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.bfloat16, buffer_dtype=torch.bfloat16)
        model = FSDP(model, device_id=rank, sync_module_states=True,
                     sharding_strategy=ShardingStrategy.FULL_SHARD if world > 1 else ShardingStrategy.NO_SHARD,
                     auto_wrap_policy=policy, mixed_precision=mp_policy, forward_prefetch=True, limit_all_gathers=False)
        apply_activation_checkpointing(model, checkpoint_wrapper_fn=checkpoint_wrapper,
                                       check_fn=lambda m: isinstance(m, Block))
        opt = configure_optimizer(model, 1e-3, 0.9, 0.95, 1e-5)
        xs = [torch.randn(8, 3, 64, 64, generator=torch.Generator().manual_seed(100 + r)).cuda().to(torch.bfloat16)
              for r in range(world)]
        x = xs[rank]
        target = patchify(x.float(), 8, True)
        # the random masking must be the same in the sharded and the unsharded run: the device generator is re-seeded
        # per step and rank right before the forward pass
        losses = []
        for st in range(6):
            torch.manual_seed(1000 * st + rank)           # MAE.random_masking draws torch.rand on the device
            pred, mask = model(x, ["r", "g", "b"])
            loss = torch.nn.functional.mse_loss(pred.float(), target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            l = loss.detach().clone()
            dist.all_reduce(l)
            losses.append(l.item() / world)
        # the same training run WITHOUT FSDP / sharding / checkpointing: the plain product model on the global batch
        # (equal per-rank batches: the mean of the per-rank losses is the global-batch loss, and FSDP averages gradients)
        plain = _build().cuda().train()
        opt_p = configure_optimizer(plain, 1e-3, 0.9, 0.95, 1e-5)
        xg = torch.cat(xs)
        tg = patchify(xg.float(), 8, True)
        losses_plain = []
        for st in range(6):
            preds = []
            for r in range(world):                          # same per-rank masking draws as above
                torch.manual_seed(1000 * st + r)
                preds.append(plain(xs[r], ["r", "g", "b"])[0])
            loss = torch.nn.functional.mse_loss(torch.cat(preds).float(), tg)
            opt_p.zero_grad(set_to_none=True)
            loss.backward()
            opt_p.step()
            losses_plain.append(loss.item())
        ret[rank] = (losses, losses_plain)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_fsdp_mixed_precision_activation_checkpointing_trains():
    world = min(2, torch.cuda.device_count())
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        losses, losses_plain = ret[r]
        assert all(l == l and abs(l) < 1e4 for l in losses), losses          # finite
        assert losses[-1] < losses[0], losses                                # it learns the fixed batch
        # sharded (bf16 parameters / bf16 gradient reduction / recompute) vs unsharded product: same trajectory
        for a, b in zip(losses, losses_plain):
            assert abs(a - b) <= 1e-2 * abs(b), (losses, losses_plain)

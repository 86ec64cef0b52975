"""Workloads of bench.py: one per BASELINE.json config, each driven the way the reference's own training script
drives it (file:line cited per class), through this package's module API.

A workload knows how to build its model / optimizer on one rank, what one pinned HOST batch looks like, how to run
one training step on a device batch, how many FLOPs a sample costs (SURVEY.md §8(d) formulas), and how to time the
oracle's CPU restatement of the same step on a bounded sample.
"""
import functools
import os
import statistics
import time

import torch

VARS3 = ["red", "green", "blue"]


def f_block(N, D, mlp_ratio=4):
    """forward FLOPs of one pre-norm block on N tokens of width D: qkv 6ND^2, proj 2ND^2,
    fc1+fc2 4*mlp*ND^2, attention core 4N^2D"""
    return (8 + 4 * mlp_ratio) * N * D * D + 4 * N * N * D


def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def _time_cpu_steps(step_fn, steps, warm):
    times = []
    for it in range(warm + steps):
        t0 = time.perf_counter()
        step_fn()
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    return statistics.median(times)


class Workload:
    name = ""
    workload = ""
    unit = "images/s"
    batch = 1                   # samples per GPU and step
    cpu_batch = 2               # samples per step of the bounded CPU sample
    uses_fsdp = False
    cpu_kind = "port"           # "reference" when the CPU arm ran the reference's own modules from baseline/_ref
    find_unused = False         # the reference's SAP / UNETR drivers wrap with find_unused_parameters=True
    l2_policy = "per-step activations exceed the 126 MB L2"

    def parallelism(self, world):
        return f"dp{world}" + (" (DDP, bf16 gradient all-reduce)" if world > 1 else "")

    # -- GPU side
    def build(self, dev, world, local, rank, args):
        raise NotImplementedError

    def host_batch(self, rank):
        raise NotImplementedError

    def step(self, *batch):
        raise NotImplementedError

    def flops(self):
        """(train FLOPs per sample, of which transformer blocks) -- train = 3 x forward"""
        raise NotImplementedError

    def extras(self, dev, args):
        return {}

    # -- CPU side (oracle restatement; the one place outside tests/ and smoke() that may execute oracle/)
    def cpu_step_fn(self):
        """-> (callable running ONE fp32 training step of cpu_batch samples on the host, description)"""
        raise NotImplementedError

    def cpu_rate(self, steps=2, warm=1):
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fn, desc = self.cpu_step_fn()
        sec = _time_cpu_steps(fn, steps, warm)
        return self.cpu_batch / sec, cores, sec, f"{steps} timed + {warm} warm-up steps of batch {self.cpu_batch}: {desc}"

    def _wrap_ddp(self, model, world, local, args):
        if world <= 1:
            return model
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=64, find_unused_parameters=self.find_unused)
        if not getattr(args, "fp32_allreduce", False):
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            net.register_comm_hook(None, default_hooks.bf16_compress_hook)
        return net


def _opt_kind(args, torch_default):
    if args.optimizer != "ucf":
        return torch_default
    return "ucf_capturable" if getattr(args, "cuda_graph", False) else "ucf"


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref (installed by baseline/install_ref.sh; git-ignored, travels to
    the GPU box) behind the third-party stand-ins of oracle/shims.  Returns the `UCF_VIT.simple.arch` module or None."""
    import importlib
    import sys
    root = os.path.dirname(os.path.abspath(__file__))
    ref = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "UCF_VIT")):
        return None
    for p_ in (ref, os.path.join(root, "oracle", "shims")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    try:
        return importlib.import_module("UCF_VIT.simple.arch")
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] reference import failed, using the oracle port: {ex}", file=sys.stderr)
        return None


def _cpu_state(model):
    return {k: v.detach().clone().float().requires_grad_(True) for k, v in model.state_dict().items()
            if not k.startswith("token_embeds.") and v.dtype.is_floating_point}


# ------------------------------------------------------------------------------------------------
# configs[0] / configs[1]: VIT classification, training_scripts/train_class_simple.py:37-46,343-357
# ------------------------------------------------------------------------------------------------
class VitClassification(Workload):
    def __init__(self, name, workload, embed_dim, depth, heads, classes, batch, cpu_batch):
        self.name, self.workload, self.batch, self.cpu_batch = name, workload, batch, cpu_batch
        self.cfg = dict(img_size=[224, 224], patch_size=16, in_chans=3, num_classes=classes, embed_dim=embed_dim,
                        depth=depth, num_heads=heads)
        self.l2_policy = ("per-step activations (>10 GB) exceed the 126 MB L2"
                          if batch >= 128 else "256 MB L2 flush write before every timed step")
        self.flush_l2 = batch < 128
        self.default_cuda_graph = batch < 128      # launch-bound at small batch: replay the captured step (single GPU)

    def _model(self):
        from ucf_vit_b200.simple.arch import VIT
        from ucf_vit_b200.utils.fused_attn import FusedAttn
        return VIT(**self.cfg, mlp_ratio=4, class_token=True, twoD=True, default_vars=VARS3, FusedAttn_option=FusedAttn.FLASH)

    def build(self, dev, world, local, rank, args):
        from ucf_vit_b200.utils.misc import configure_optimizer
        torch.manual_seed(0)
        self.fp32_pixels = bool(getattr(args, "fp32_pixels", False))
        self.model = self._model().to(dev).train()
        self.net = self._wrap_ddp(self.model, world, local, args)
        self.opt = configure_optimizer(self.model, 1e-4, 0.9, 0.95, 1e-5, fused=_opt_kind(args, True))
        self.lossf = torch.nn.CrossEntropyLoss()

    fp32_pixels = False

    def host_batch(self, rank):
        g = torch.Generator().manual_seed(1234 + rank)
        if self.fp32_pixels:      # round-1 form: the reference's loader hands float tensors over (154 MB per 256 images)
            x = torch.rand(self.batch, 3, 224, 224, generator=g) * 255.0
        else:                     # decoded uint8 pixels cross the host link; ucf_patchify widens them to bf16 (exact)
            x = torch.randint(0, 256, (self.batch, 3, 224, 224), generator=g, dtype=torch.uint8)
        y = torch.randint(0, self.cfg["num_classes"], (self.batch,), generator=g)
        return _pin(x), _pin(y)

    def step(self, x, y):
        loss = self.lossf(self.net(x, VARS3).float(), y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    def flops(self):
        c = self.cfg
        D, depth = c["embed_dim"], c["depth"]
        L = (c["img_size"][0] // c["patch_size"]) ** 2
        K = c["in_chans"] * c["patch_size"] ** 2
        blocks = depth * f_block(L + 1, D)
        return 3 * (2 * L * K * D + blocks + 2 * D * c["num_classes"]), 3 * blocks

    def cpu_step_fn(self):
        g = torch.Generator().manual_seed(0)
        x = torch.rand(self.cpu_batch, 3, 224, 224, generator=g) * 255.0
        y = torch.randint(0, self.cfg["num_classes"], (self.cpu_batch,), generator=g)
        ref = import_reference()
        if ref is not None:
            # the reference's own VIT + configure_optimizer + training_step (train_class_simple.py:37-46,343-357), fp32
            from UCF_VIT.utils.fused_attn import FusedAttn as RefFusedAttn
            from UCF_VIT.utils.misc import configure_optimizer as ref_configure_optimizer
            torch.manual_seed(0)
            model = ref.VIT(**self.cfg, mlp_ratio=4, class_token=True, twoD=True, default_vars=VARS3,
                            FusedAttn_option=RefFusedAttn.DEFAULT).train()
            opt = ref_configure_optimizer(model, 1e-4, 0.9, 0.95, 1e-5)
            lossf = torch.nn.CrossEntropyLoss()
            self.cpu_kind = "reference"

            def fn_ref():
                loss = lossf(model.forward(x, VARS3, None), y)
                loss.backward()
                opt.step()
                opt.zero_grad()
            return fn_ref, ("the UNMODIFIED reference (baseline/_ref: UCF_VIT.simple.arch.VIT, SDPA attention, "
                            "utils.misc.configure_optimizer AdamW), fp32 on the host cores")
        from oracle import vit_ref as R
        torch.manual_seed(0)
        sd = _cpu_state(self._model())
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)
        cfg = dict(self.cfg)

        def fn():
            loss = torch.nn.functional.cross_entropy(R.vit_forward(x, sd, cfg), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return fn, "same model, fp32, oracle/vit_ref.py + torch autograd + AdamW on the host cores"


# ------------------------------------------------------------------------------------------------
# FSDP wrapping shared by configs[2] and the diffusion half of configs[4]:
# training_scripts/train_masked_fsdp.py:361-396 (auto-wrap {Block, Sequential}, bf16 MixedPrecision,
# activation checkpointing on Block), :417-419 (ShardedGradScaler(init_scale=8192, growth_interval=100))
# ------------------------------------------------------------------------------------------------
def fsdp_wrap(model, world, local):
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import apply_activation_checkpointing, checkpoint_wrapper
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp import MixedPrecision, ShardingStrategy
    from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy
    from torch.nn import Sequential
    from ucf_vit_b200.fsdp.building_blocks import Block
    policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={Block, Sequential})
    mp_policy = MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.bfloat16, buffer_dtype=torch.bfloat16)
    model = FSDP(model, device_id=local, sync_module_states=True,
                 sharding_strategy=ShardingStrategy.FULL_SHARD if world > 1 else ShardingStrategy.NO_SHARD,
                 auto_wrap_policy=policy, mixed_precision=mp_policy, forward_prefetch=True, limit_all_gathers=False)
    apply_activation_checkpointing(model, checkpoint_wrapper_fn=checkpoint_wrapper, check_fn=lambda m: isinstance(m, Block))
    return model


class _FsdpWorkload(Workload):
    uses_fsdp = True

    def parallelism(self, world):
        return (f"fsdp{world} FULL_SHARD" if world > 1 else "fsdp1 NO_SHARD") + ", bf16 MixedPrecision, activation checkpointing on Block, ShardedGradScaler"

    def _finish_build(self, model, dev, world, local, args):
        from torch.distributed.fsdp.sharded_grad_scaler import ShardedGradScaler
        from ucf_vit_b200.utils.misc import configure_optimizer
        self.net = fsdp_wrap(model.to(dev), world, local).train()
        self.opt = configure_optimizer(self.net, 1e-4, 0.9, 0.95, 1e-5, fused=_opt_kind(args, None))
        self.scaler = ShardedGradScaler(init_scale=8192, growth_interval=100)

    def _backward_and_update(self, loss):
        self.scaler.scale(loss).backward()
        self.scaler.step(self.opt)
        self.scaler.update()
        self.opt.zero_grad(set_to_none=True)


class MaeVitLFsdp(_FsdpWorkload):
    """MAE ViT-L/16, 75 % mask: train_masked_fsdp.py:47-62 (training_step, default full-MSE loss), :590-607 (bf16 batch,
    scaler.scale(loss).backward(); scaler.step; scaler.update)."""
    name = "mae_vitl_fsdp"
    workload = ("MAE ViT-L/16 masked pretraining train step (75 % mask, fwd+MSE+bwd+AdamW), 224x224x3 bf16 images, "
                "batch 512/GPU, FSDP + activation checkpointing + grad scaler")
    batch = 512
    cpu_batch = 4
    cfg = dict(img_size=[224, 224], patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
               decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mask_ratio=0.75, class_token=False)

    def _model(self, fsdp=True):
        if fsdp:
            from ucf_vit_b200.fsdp.arch import MAE
            extra = dict(tensor_par_size=1, tensor_par_group=None)
        else:
            from ucf_vit_b200.simple.arch import MAE
            extra = {}
        c = self.cfg
        return MAE(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, embed_dim=c["embed_dim"], depth=c["depth"],
                   num_heads=c["num_heads"], decoder_embed_dim=c["decoder_embed_dim"], decoder_depth=c["decoder_depth"],
                   decoder_num_heads=c["decoder_num_heads"], mlp_ratio=4, mlp_ratio_decoder=4, mask_ratio=c["mask_ratio"],
                   linear_decoder=False, class_token=False, twoD=True, default_vars=VARS3, adaptive_patching=False, **extra)

    def build(self, dev, world, local, rank, args):
        torch.manual_seed(0)
        self._finish_build(self._model(), dev, world, local, args)

    def host_batch(self, rank):
        g = torch.Generator().manual_seed(1234 + rank)
        return (_pin(torch.randn(self.batch, 3, 224, 224, generator=g).to(torch.bfloat16)),)   # data.to(precision_dt) on the host

    def step(self, x):
        from ucf_vit_b200.utils.metrics import patch_mse
        pred, _ = self.net(x, VARS3, None)
        loss = patch_mse(pred, x, self.cfg["patch_size"], True, None)
        self._backward_and_update(loss)
        return loss

    def flops(self):
        c = self.cfg
        L = (c["img_size"][0] // c["patch_size"]) ** 2
        K = 3 * c["patch_size"] ** 2
        keep = int(L * (1 - c["mask_ratio"]))
        D, Dd = c["embed_dim"], c["decoder_embed_dim"]
        blocks = c["depth"] * f_block(keep, D) + c["decoder_depth"] * f_block(L, Dd)
        other = 2 * L * K * D + 2 * keep * D * Dd + 2 * L * Dd * K
        return 3 * (blocks + other), 3 * blocks

    def cpu_step_fn(self):
        g = torch.Generator().manual_seed(0)
        x = torch.randn(self.cpu_batch, 3, 224, 224, generator=g)
        ref = import_reference()
        if ref is not None:
            # the reference's own MAE and training_step (train_masked_fsdp.py:47-62, default full-MSE loss), fp32, unsharded
            from UCF_VIT.utils.misc import configure_optimizer as ref_configure_optimizer
            from UCF_VIT.utils.misc import patchify as ref_patchify
            c = self.cfg
            torch.manual_seed(0)
            model = ref.MAE(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, embed_dim=c["embed_dim"],
                            depth=c["depth"], num_heads=c["num_heads"], decoder_embed_dim=c["decoder_embed_dim"],
                            decoder_depth=c["decoder_depth"], decoder_num_heads=c["decoder_num_heads"], mlp_ratio=4,
                            mlp_ratio_decoder=4, mask_ratio=c["mask_ratio"], linear_decoder=False, class_token=False,
                            weight_init="skip", twoD=True, default_vars=VARS3, adaptive_patching=False).train()
            opt = ref_configure_optimizer(model, 1e-4, 0.9, 0.95, 1e-5)
            self.cpu_kind = "reference"

            def fn_ref():
                output, _ = model.forward(x, VARS3, None)
                loss = torch.nn.MSELoss()(output, ref_patchify(x, 16, True))
                loss.backward()
                opt.step()
                opt.zero_grad()
            return fn_ref, ("the UNMODIFIED reference (baseline/_ref: UCF_VIT.simple.arch.MAE + utils.misc.patchify / "
                            "configure_optimizer), fp32, unsharded, no recompute, on the host cores")
        from oracle import vit_ref as R
        torch.manual_seed(0)
        sd = _cpu_state(self._model(fsdp=False))
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)
        noise = torch.rand(self.cpu_batch, 196, generator=g)
        cfg = dict(self.cfg, kind="mae")
        target = R.patchify_target(x, 16, True)

        def fn():
            pred, mask = R.mae_forward(x, sd, cfg, noise)
            loss = torch.nn.functional.mse_loss(pred, target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return fn, "same model, fp32, no recompute, oracle/vit_ref.py mae_forward + torch autograd + AdamW on the host cores"


class DiffusionFsdp(_FsdpWorkload):
    """DiffusionVIT (configs/imagenet/diffusion/base_config.yaml): train_diffusion_fsdp.py:37-45 (training_step),
    :508-518 (t, noise, q-sample on the device), FSDP wrapping as for MAE."""
    name = "diffusion_fsdp"
    workload = ("DiffusionVIT-B/16 noise-prediction train step (q-sample + fwd + MSE + bwd + AdamW), 256x256x3 bf16 images, "
                "batch 128/GPU, FSDP + activation checkpointing + grad scaler")
    batch = 128
    cpu_batch = 2
    cfg = dict(img_size=[256, 256], patch_size=16, in_chans=3, embed_dim=768, depth=12, num_heads=12,
               decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, class_token=False, time_steps=1000)

    def _model(self, fsdp=True):
        if fsdp:
            from ucf_vit_b200.fsdp.arch import DiffusionVIT
            extra = dict(tensor_par_size=1, tensor_par_group=None)
        else:
            from ucf_vit_b200.simple.arch import DiffusionVIT
            extra = {}
        c = self.cfg
        return DiffusionVIT(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=3, embed_dim=c["embed_dim"],
                            depth=c["depth"], num_heads=c["num_heads"], decoder_embed_dim=c["decoder_embed_dim"],
                            decoder_depth=c["decoder_depth"], decoder_num_heads=c["decoder_num_heads"], mlp_ratio=4,
                            mlp_ratio_decoder=4, linear_decoder=False, class_token=False, twoD=True, default_vars=VARS3,
                            time_steps=c["time_steps"], **extra)

    def build(self, dev, world, local, rank, args):
        from ucf_vit_b200.ddpm.ddpm import DDPM_Scheduler
        torch.manual_seed(0)
        self.sched = DDPM_Scheduler(num_time_steps=self.cfg["time_steps"])
        self.dev = dev
        self._finish_build(self._model(), dev, world, local, args)

    def host_batch(self, rank):
        g = torch.Generator().manual_seed(1234 + rank)
        x = torch.randn(self.batch, 3, 256, 256, generator=g).to(torch.bfloat16)
        t = torch.randint(0, self.cfg["time_steps"], (self.batch,), generator=g)
        a = self.sched.alpha[t].view(self.batch, 1, 1, 1).to(torch.bfloat16)
        return _pin(x), _pin(t), _pin(a)

    def step(self, x, t, a):
        from ucf_vit_b200.utils.misc import unpatchify
        e = torch.randn_like(x)
        xt = torch.sqrt(a) * x + torch.sqrt(1 - a) * e
        out = unpatchify(self.net(xt, t, VARS3), xt, self.cfg["patch_size"], True)
        loss = torch.nn.functional.mse_loss(out.float(), e.float())
        self._backward_and_update(loss)
        return loss

    def flops(self):
        c = self.cfg
        L = (c["img_size"][0] // c["patch_size"]) ** 2
        K = 3 * c["patch_size"] ** 2
        D, Dd = c["embed_dim"], c["decoder_embed_dim"]
        blocks = c["depth"] * f_block(L, D) + c["decoder_depth"] * f_block(L, Dd)
        other = 2 * L * K * D + 2 * L * D * Dd + 2 * L * Dd * K
        return 3 * (blocks + other), 3 * blocks

    def cpu_step_fn(self):
        from oracle import vit_ref as R
        from ucf_vit_b200.ddpm.ddpm import DDPM_Scheduler
        torch.manual_seed(0)
        m = self._model(fsdp=False)
        sd = _cpu_state(m)
        table = m.temporalEmbeddings.embeddings.detach().float()
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)
        g = torch.Generator().manual_seed(0)
        B = self.cpu_batch
        x = torch.randn(B, 3, 256, 256, generator=g)
        t = torch.randint(0, 1000, (B,), generator=g)
        a = DDPM_Scheduler(1000).alpha[t].view(B, 1, 1, 1)
        cfg = dict(self.cfg, kind="diffusion")

        def fn():
            e = torch.randn_like(x)
            xt = torch.sqrt(a) * x + torch.sqrt(1 - a) * e
            # eval-mode restatement (the time-embedding MLP's dropout(0.5) is not part of the timed arithmetic)
            pred = R.diffusion_forward(xt, t, sd, cfg, table)
            loss = torch.nn.functional.mse_loss(pred, R.patchify_target(e, 16, True))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return fn, "same model, fp32, no recompute, oracle/vit_ref.py diffusion_forward + torch autograd + AdamW on the host cores"


# ------------------------------------------------------------------------------------------------
# configs[3]: UNETR 3-D with variable aggregation, training_scripts/train_unetr_simple.py:34-41,447-452
# ------------------------------------------------------------------------------------------------
class Unetr128(Workload):
    find_unused = True          # train_unetr_simple.py:273
    name = "unetr_128"
    unit = "volumes/s"
    workload = ("UNETR 3-D segmentation train step (fwd+DiceCE+bwd+AdamW), 128^3 volumes, 4 variable channels aggregated by "
                "cross-attention, patch 16, ViT-B encoder + conv decoder (feature_size 16), batch 16/GPU")
    batch = 16
    cpu_batch = 1
    V = 4
    vars_ = ["v0", "v1", "v2", "v3"]
    cfg = dict(img_size=[128] * 3, patch_size=16, in_chans=4, num_classes=4, embed_dim=768, depth=12, num_heads=12,
               use_varemb=True, feature_size=16, class_token=False)

    def _model(self):
        from ucf_vit_b200.simple.arch import UNETR
        c = self.cfg
        return UNETR(img_size=c["img_size"], patch_size=c["patch_size"], in_chans=c["in_chans"], num_classes=c["num_classes"],
                     embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], twoD=False, use_varemb=True,
                     default_vars=self.vars_, feature_size=c["feature_size"], skip_connection=True, linear_decoder=False,
                     class_token=False)

    def build(self, dev, world, local, rank, args):
        from ucf_vit_b200.utils.metrics import DiceCELoss
        from ucf_vit_b200.utils.misc import configure_optimizer
        torch.manual_seed(0)
        self.model = self._model().to(dev).train()
        if getattr(args, "bf16_decoder", False):
            self.model.conv_autocast_dtype = torch.bfloat16
            self.workload = self.workload.replace("conv decoder (feature_size 16)", "conv decoder (feature_size 16) under bf16 autocast")
        if getattr(args, "fp32_decoder", False):
            self.workload = self.workload.replace("conv decoder (feature_size 16)", "fp32 PyTorch / cuDNN conv decoder (feature_size 16)")
        elif not getattr(args, "bf16_decoder", False):
            self.model.use_fused_decoder()
            torch.backends.cudnn.benchmark = True       # cuDNN picks its convolution kernels by timing them in the warm-up steps
            self.workload = self.workload.replace("conv decoder (feature_size 16)", "channels-last bf16 conv decoder (feature_size "
                                                  "16; cuDNN convolutions chosen with cudnn.benchmark, fused InstanceNorm + LeakyReLU kernels)")
        self.net = self._wrap_ddp(self.model, world, local, args)
        self.opt = configure_optimizer(self.model, 1e-5, 0.9, 0.95, 1e-5, fused=_opt_kind(args, True))
        self.lossf = DiceCELoss(to_onehot_y=True, softmax=True, squared_pred=True, smooth_nr=0.0, smooth_dr=1e-6)

    def host_batch(self, rank):
        g = torch.Generator().manual_seed(1234 + rank)
        x = torch.rand(self.batch, self.V, 128, 128, 128, generator=g)
        y = torch.randint(0, self.cfg["num_classes"], (self.batch, 1, 128, 128, 128), generator=g).to(torch.uint8)
        return _pin(x), _pin(y)

    def step(self, x, y):
        loss = self.lossf(self.net(x, self.vars_), y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    def _encoder_flops(self):
        c = self.cfg
        D, V = c["embed_dim"], self.V
        L = (128 // 16) ** 3
        K1 = 16 ** 3
        blocks = c["depth"] * f_block(L, D)
        # per-variable patch embed, kv projection over V*L tokens, q projection of the query, output projection
        agg = 2 * V * L * D * 2 * D + 2 * L * D * D + 4 * L * V * D
        return V * 2 * L * K1 * D + agg + blocks, blocks

    def flops(self):
        enc, blocks = self._encoder_flops()
        return 3 * (enc + getattr(self, "decoder_fwd_flops", 0)), 3 * blocks

    def extras(self, dev, args):
        """conv-decoder FLOPs (torch ops only are visible to the counter: exactly the cuDNN decoder) and the encoder-only
        step time, so the attention/MLP fraction of peak is stated for the part this package's kernels run."""
        from torch.utils.flop_counter import FlopCounterMode
        x, y = [t.to(dev) for t in self.host_batch(0)]
        with torch.no_grad(), FlopCounterMode(display=False) as fc:
            self.model(x, self.vars_)
        self.decoder_fwd_flops = fc.get_total_flops() / self.batch
        # encoder only: tokens -> blocks (+ skip features) forward and backward
        def enc_step():
            feats, inter = self.model.forward_intermediates(x, self.vars_, None, indices=self.model.skip_indices)
            loss = feats.float().mean() + sum(t.float().mean() for t in inter)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
        for _ in range(3):
            enc_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n):
            enc_step()
        e1.record()
        torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1) / n
        enc, blocks = self._encoder_flops()
        return {"encoder_ms_per_step": enc_ms, "encoder_tflops": 3 * enc * self.batch / (enc_ms * 1e-3) / 1e12,
                "encoder_attn_mlp_tflops": 3 * blocks * self.batch / (enc_ms * 1e-3) / 1e12,
                "decoder_fwd_gflops_per_volume": self.decoder_fwd_flops / 1e9,
                "decoder_note": "default decoder: channels-last bf16 with this package's InstanceNorm / LeakyReLU, 3x3x3 weight-gradient and "
                                "1x1x1 convolution kernels, cuDNN for the 3x3x3 forward / data-gradient and transposed convolutions; the "
                                "block structure is MONAI's restated (MONAI absent: parity unpinned)"}

    def cpu_step_fn(self):
        from oracle import vit_ref as R
        torch.manual_seed(0)
        sd = _cpu_state(self._model())
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-5, betas=(0.9, 0.95), weight_decay=1e-5)
        g = torch.Generator().manual_seed(0)
        x = torch.rand(self.cpu_batch, self.V, 128, 128, 128, generator=g)
        y = torch.randint(0, 4, (self.cpu_batch, 128, 128, 128), generator=g)
        cfg = dict(self.cfg, kind="unetr")

        def fn():
            o = R.unetr_forward(x, sd, cfg, var_ids=[0, 1, 2, 3])
            loss = torch.nn.functional.cross_entropy(o, y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return fn, "same model, fp32, oracle/vit_ref.py unetr_forward + CE + torch autograd + AdamW on the host cores"


# ------------------------------------------------------------------------------------------------
# configs[4]: SAP on 4096^2 images, training_scripts/train_sap_simple.py:28-46,417-427 + the adaptive-patching
# front end dataloaders/transform.py:21-54 (Patchify.forward) that produces its sequences
# ------------------------------------------------------------------------------------------------
class Sap4096(Workload):
    find_unused = True          # train_sap_simple.py:255,331
    unit = "images/s"
    cpu_batch = 1

    def __init__(self, L):
        self.L = L
        self.s = int(round(L ** 0.5))
        self.p = 16
        self.side = self.p * self.s
        self.name = f"sap_4096_L{L}"
        self.batch = 16 if L <= 1024 else 8
        self.workload = (f"SAP adaptive-patching ViT-B segmentation train step on 4096x4096x3 uint8 images: quadtree of {L} leaves "
                         f"per image (host C++ build) + device cubic gather to {self.p}x{self.p} patches + fwd + DiceBCE + bwd + AdamW, "
                         f"batch {self.batch}/GPU; the trees come from fixed synthetic edge maps (the edge detector is timed in "
                         f"`extras`; --device-edges puts ucf_gaussian_blur_u8 + ucf_canny_u8 of every image into the step)")
        self.cfg = dict(patch_size=self.p, in_chans=3, num_classes=4, embed_dim=768, depth=12, num_heads=12, fixed_length=L,
                        sqrt_len=self.s, class_token=False, use_adaptive_pos_emb=True)

    def _model(self):
        from ucf_vit_b200.simple.arch import SAP
        c = self.cfg
        return SAP(img_size=[self.side, self.side], patch_size=self.p, in_chans=3, num_classes=c["num_classes"],
                   embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"], twoD=True, default_vars=VARS3,
                   adaptive_patching=True, fixed_length=self.L, sqrt_len=self.s, sqrt_len_method=True,
                   use_adaptive_pos_emb=True, class_token=False)

    def build(self, dev, world, local, rank, args):
        from ucf_vit_b200.utils.metrics import DiceBLoss
        from ucf_vit_b200.utils.misc import configure_optimizer
        torch.manual_seed(0)
        self.dev = dev
        self.model = self._model().to(dev).train()
        self.net = self._wrap_ddp(self.model, world, local, args)
        self.opt = configure_optimizer(self.model, 1e-4, 0.9, 0.95, 1e-5, fused=_opt_kind(args, True))
        self.lossf = DiceBLoss(num_class=self.cfg["num_classes"])
        g = torch.Generator().manual_seed(77 + rank)
        self.edges = [self._edge_map(g) for _ in range(self.batch)]
        self.device_edges = bool(getattr(args, "device_edges", False))
        if self.device_edges:
            from concurrent.futures import ThreadPoolExecutor
            self._edge_pool = ThreadPoolExecutor(max_workers=1)
            self._edge_stream = torch.cuda.Stream(device=dev)
            self.workload = self.workload.split("; the trees come from")[0] + (
                "; every step runs the reference's edge detector on the device for the next batch (cv.GaussianBlur 5x5 + cv.Canny "
                "60/110 as ucf_gaussian_blur_u8 + ucf_canny_u8 on a side stream, byte-identical to OpenCV), copies the edge maps to "
                "the host and builds the trees there, all under the current step's GPU work")
        self.label = (torch.rand(self.batch, self.cfg["num_classes"], self.side, self.side, generator=g) > 0.5).float().to(dev)

    @staticmethod
    def _edge_map(g):
        """Canny-like synthetic edge map: 255 on ~3 % of the pixels, clustered in a few regions so the tree is deep
        somewhere and shallow elsewhere"""
        import numpy as np
        e = torch.zeros(4096, 4096, dtype=torch.uint8)
        for _ in range(24):
            cy, cx = [int(v) for v in torch.randint(256, 3840, (2,), generator=g)]
            r = int(torch.randint(64, 512, (1,), generator=g))
            blk = (torch.rand(2 * r, 2 * r, generator=g) < 0.12).to(torch.uint8) * 255
            y0, x0 = max(0, cy - r), max(0, cx - r)
            e[y0:y0 + 2 * r, x0:x0 + 2 * r] |= blk[:min(2 * r, 4096 - y0), :min(2 * r, 4096 - x0)]
        return np.ascontiguousarray(e.numpy())

    def host_batch(self, rank):
        g = torch.Generator().manual_seed(1234 + rank)
        return (_pin(torch.randint(0, 256, (self.batch, 4096, 4096, 3), generator=g, dtype=torch.uint8)),)

    def _trees_from_device_edges(self, imgs):
        """Future of the batch's trees: blur + Canny of every image on a side stream (worker thread: the hysteresis loop
        synchronises that stream only), edge maps to the host, C++ tree build on host threads."""
        from ucf_vit_b200 import ops
        from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
        ready = torch.cuda.Event()
        ready.record()

        def work():
            with torch.cuda.device(self.dev), torch.cuda.stream(self._edge_stream):
                self._edge_stream.wait_event(ready)
                maps = [ops.canny_u8(ops.gaussian_blur_u8(imgs[i], 5), 60, 110) for i in range(imgs.shape[0])]
                host = [m.cpu().numpy() for m in maps]
            return FixedQuadTree.build_many(host, self.L, device=self.dev)
        return self._edge_pool.submit(work)

    def front_end(self, imgs):
        """Patchify.forward_batch: trees on host threads, gather on the device (edge maps: fixed synthetic ones, or with
        --device-edges the device blur + Canny of the images)."""
        from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
        if self.device_edges:
            # the next batch is on the device one step ahead (DevicePrefetcher); the synthetic batches are identical, so the
            # current images stand in for it
            trees = (getattr(self, "_next_trees", None) or self._trees_from_device_edges(imgs)).result()
            self._next_trees = self._trees_from_device_edges(imgs)
            return self._gather(trees, imgs)
        # the trees of THIS batch were submitted during the previous step (a loader knows batch k+1 while the GPU works on
        # batch k); the trees of the next batch are submitted now and built on a host thread under this step's GPU work
        fut = getattr(self, "_next_trees", None) or FixedQuadTree.build_many_async(self.edges, self.L, device=self.dev)
        trees = fut.result()
        self._next_trees = FixedQuadTree.build_many_async(self.edges, self.L, device=self.dev)
        return self._gather(trees, imgs)

    def _gather(self, trees, imgs):
        p = self.p
        seqs, ps = [], []
        for i, qdt in enumerate(trees):
            seq, size, pos = qdt.serialize_device(imgs[i], size=(p, p, 3))
            seqs.append(seq.reshape(3, -1, p * p))                       # the reference's raw reshape (transform.py:45-48)
            ps.append(torch.cat([size.unsqueeze(-1).float(), pos.float()], dim=-1))
        seq = torch.stack(seqs).reshape(-1, 3, self.side, self.side)      # train_sap_simple.py:32
        return seq, torch.stack(ps)

    def step(self, imgs):
        seq, seq_ps = self.front_end(imgs)
        loss = self.lossf(self.net(seq, VARS3, seq_ps), self.label)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    def flops(self):
        c = self.cfg
        D, L, p = c["embed_dim"], self.L, self.p
        blocks = c["depth"] * f_block(L, D)
        other = 2 * L * (3 * p * p) * D + 2 * L * D * 256 * p * p + 2 * self.side * self.side * 256 * c["num_classes"]
        return 3 * (blocks + other), 3 * blocks

    def extras(self, dev, args):
        """front end alone: host tree build and device gather, timed separately; gather bandwidth against HBM peak"""
        from ucf_vit_b200.dataloaders.quadtree import FixedQuadTree
        imgs = self.host_batch(0)[0].to(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            trees = FixedQuadTree.build_many(self.edges, self.L, device=dev)
        tree_ms = (time.perf_counter() - t0) / 3 * 1e3
        for qdt in trees:
            qdt.serialize_device(imgs[0], size=(self.p, self.p, 3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            trees[0].serialize_device(imgs[0], size=(self.p, self.p, 3))
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1) / n
        # algorithmic bytes of one gather: INTER_CUBIC without antialiasing reads 16 taps (1 B each, uint8) per
        # output sample whatever the leaf size, and writes one fp32: L * p^2 * C * (16 + 4) bytes
        alg = self.L * self.p * self.p * 3 * (16 + 4)
        out = {"tree_build_ms_per_batch_host": tree_ms, "gather_ms_per_image": gather_ms,
               "gather_algorithmic_GBps": alg / (gather_ms * 1e-3) / 1e9}
        out.update(self._edge_front_end(imgs[0]))
        return out

    @staticmethod
    def _edge_front_end(img):
        """Edge detection of one 4096^2 image (the reference's cv.GaussianBlur(5x5) + cv.Canny, transform.py:33-34): device
        kernels (ucf_gaussian_blur_u8 + ucf_canny_u8, wall clock of the call: the hysteresis loop synchronises) against OpenCV
        on the host cores when cv2 is importable.  Not part of the timed step: the step uses fixed synthetic edge maps."""
        from ucf_vit_b200 import ops
        run = lambda: ops.canny_u8(ops.gaussian_blur_u8(img, 5), 60, 110)   # noqa: E731
        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            e = run()
        torch.cuda.synchronize()
        out = {"edges_device_ms_per_image": (time.perf_counter() - t0) / 5 * 1e3}
        try:
            import cv2
            h = img.cpu().numpy()
            t0 = time.perf_counter()
            he = cv2.Canny(cv2.GaussianBlur(h, (5, 5), 0), 60, 110)
            out["edges_opencv_host_ms_per_image"] = (time.perf_counter() - t0) * 1e3
            out["edges_identical_to_opencv"] = bool((torch.as_tensor(he) == e.cpu()).all())
        except ImportError:
            pass
        return out

    CPU_SAMPLE_L = 256

    def cpu_rate(self, steps=1, warm=0):
        """The reference's CPU step at L = 1024 / 4096 takes 6+ minutes per image (its 256-channel transposed-convolution
        neck alone writes 1 GB per image), so the bounded sample is ONE image at L = 256 leaves (same 4096^2 image, same
        model depth / width, tree + per-leaf cubic resize + model step), scaled to this workload by the FLOP ratio."""
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        small = Sap4096(self.CPU_SAMPLE_L)
        fn = small._cpu_fn()
        sec = _time_cpu_steps(fn, max(1, steps), warm)
        scale = small.flops()[0] / self.flops()[0]
        rate = (1.0 / sec) * scale
        return rate, cores, sec / scale, (
            f"{max(1, steps)} timed step(s) of ONE 4096^2 image at L = {self.CPU_SAMPLE_L} leaves ({sec:.1f} s: numpy quadtree build + "
            f"per-leaf cubic resize (oracle/quadtree_np.py) + oracle/vit_ref.py sap_forward fp32 + torch autograd + AdamW), "
            f"scaled by the train-FLOP ratio {1 / scale:.1f} to L = {self.L}")

    def _cpu_fn(self):
        import numpy as np
        from oracle import quadtree_np as Q
        from oracle import vit_ref as R
        torch.manual_seed(0)
        sd = _cpu_state(self._model())
        opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.95), weight_decay=1e-5)
        g = torch.Generator().manual_seed(0)
        img = torch.randint(0, 256, (4096, 4096, 3), generator=g, dtype=torch.uint8).numpy()
        edge = self._edge_map(g)
        label = (torch.rand(1, 4, self.side, self.side, generator=g) > 0.5).float()
        cfg = dict(self.cfg, kind="sap")
        L, p = self.L, self.p

        def fn():
            nodes = Q.build_quadtree(edge, L)
            seq, size, pos = Q.serialize2d(nodes, img, p, L)
            seq = torch.from_numpy(np.asarray(seq, dtype=np.float32).reshape(3, -1, p * p)).reshape(1, 3, self.side, self.side)
            seq_ps = torch.cat([torch.tensor(size, dtype=torch.float32).unsqueeze(-1), torch.tensor(pos, dtype=torch.float32)], -1)[None]
            o = R.sap_forward(seq, sd, cfg, seq_ps)
            loss = ((torch.sigmoid(o) - label) ** 2).mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return fn


def registry():
    w = [
        VitClassification("vit_b16", "ViT-B/16 ImageNet-shape classification train step (fwd+CE+bwd+AdamW), 224x224x3, batch 256/GPU",
                          768, 12, 12, 1000, 256, 16),
        VitClassification("vit_tiny", "ViT-Tiny/16 catsdogs classification train step (fwd+CE+bwd+AdamW), 224x224x3, batch 32/GPU "
                          "(configs/catsdogs batch_size)", 192, 12, 3, 2, 32, 32),
        MaeVitLFsdp(), Unetr128(), Sap4096(4096), Sap4096(1024), DiffusionFsdp(),
    ]
    return {x.name: x for x in w}

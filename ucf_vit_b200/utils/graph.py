"""Whole-step CUDA graphs: capture forward + loss + backward + optimizer update once, replay it every step.

The reference has no counterpart (its step is ~1000 eager torch launches, training_scripts/train_class_simple.py:343-357);
this is the B200-side answer to small-batch configs where the host cannot enqueue kernels as fast as the GPU retires
them (ViT-Tiny, batch 32: 6 ms of Python + launch calls per step against ~2 ms of GPU work).  Everything this package
launches is capture-safe: launchers allocate nothing, tensor maps are encoded on the host and passed by value, the
kernels read no host state.  The optimizer must keep its step count and learning rate on the device
(`configure_optimizer(..., fused="ucf_capturable")`).
"""
import torch


class GraphedTrainStep:
    """`step_fn(*batch) -> loss` captured into one CUDA graph.

    `step_fn` must be a pure device-side step: forward, loss, `optimizer.zero_grad(set_to_none=True)`, backward,
    `optimizer.step()`; no `.item()`, no host-dependent control flow, every tensor shape fixed.  Call the object with a
    batch of the example's shapes (device tensors; they are copied into the static input buffers) -- it returns the
    static loss tensor of the replayed step."""

    def __init__(self, step_fn, example_batch, warmup: int = 3):
        if not all(torch.is_tensor(t) and t.is_cuda for t in example_batch):
            raise RuntimeError("GraphedTrainStep needs CUDA example tensors; there is no CPU path")
        self.static_in = tuple(t.clone() for t in example_batch)
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                   # eager warm-up on a side stream (allocator + lazy kernel attributes)
            for _ in range(max(1, warmup)):
                step_fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = step_fn(*self.static_in)
        self.replays = 0

    def __call__(self, *batch):
        for dst, src in zip(self.static_in, batch):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.static_loss

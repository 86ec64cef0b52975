"""Host -> device batch hand-off (ucf_vit_b200/dataloaders/prefetch.py).
CPU: structure handling and loud failure without CUDA.  GPU: ordering, slot reuse and stream hand-shakes."""
import pytest
import torch

from ucf_vit_b200.dataloaders.prefetch import DevicePrefetcher, _flatten


def test_flatten_rebuilds_nested_batches():
    batch = (torch.arange(3), [torch.ones(2), "name"], {"a": torch.zeros(1), "k": 7})
    leaves, rebuild = _flatten(batch)
    assert len(leaves) == 3
    out = rebuild([t + 1 for t in leaves])
    assert isinstance(out, tuple) and isinstance(out[1], list) and out[1][1] == "name" and out[2]["k"] == 7
    assert torch.equal(out[0], torch.arange(3) + 1) and torch.equal(out[2]["a"], torch.ones(1))


def test_prefetcher_is_cuda_only():
    with pytest.raises(RuntimeError, match="no CPU path"):
        DevicePrefetcher([], "cpu")
    with pytest.raises(ValueError):
        DevicePrefetcher([], "cuda", depth=1)


@pytest.mark.gpu
@pytest.mark.parametrize("depth,pinned", [(2, False), (2, True), (3, False)])
def test_prefetcher_delivers_every_batch_in_order_without_overwriting_live_slots(depth, pinned):
    n = 9
    host = []
    for i in range(n):
        rows = 4 if i < n - 1 else 3                     # ragged last batch: slot buffers are re-made
        x, y = torch.full((rows, 3, 64, 64), float(i)), torch.full((rows,), i, dtype=torch.int64)
        host.append((x.pin_memory(), y.pin_memory()) if pinned else (x, y))
    pf = DevicePrefetcher(host, "cuda", depth=depth)
    seen = []
    for x, y in pf:
        assert x.is_cuda and y.is_cuda and x.dtype == torch.float32 and y.dtype == torch.int64
        torch.cuda._sleep(3_000_000)                     # the consumer's stream is busy: reads below run late
        seen.append((x.sum(), y.sum(), x.shape[0]))      # no host sync inside the loop
    torch.cuda.synchronize()
    assert len(seen) == n
    for i, (sx, sy, rows) in enumerate(seen):
        want = 4 if i < n - 1 else 3
        assert rows == want and sx.item() == i * want * 3 * 64 * 64 and sy.item() == i * want
    assert pf.h2d_bytes == sum(x.numel() * 4 + y.numel() * 8 for x, y in host)


@pytest.mark.gpu
def test_prefetcher_handles_empty_and_single_batch_loaders():
    assert list(DevicePrefetcher([], "cuda")) == []
    (only,) = list(DevicePrefetcher([{"x": torch.ones(2, 2), "tag": "t"}], "cuda"))
    assert only["tag"] == "t" and only["x"].is_cuda and only["x"].sum().item() == 4

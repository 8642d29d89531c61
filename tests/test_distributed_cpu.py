"""N>1 host logic on CPU: gloo, world_size 2 -- batch sharding and the single flat-buffer gradient all-reduce
(scan2cap_b200/distributed.py).  No CUDA needed: the model here is a stand-in with the same parameter plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scan2cap_b200.distributed import FlatGradients, shard_batch
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    flat = FlatGradients(model)
    g = torch.Generator().manual_seed(1)
    data = {"x": torch.randn(8, 6, generator=g), "y": torch.randn(8, 2, generator=g), "num_words": 5}
    shard = shard_batch(data, rank, world)
    assert shard["x"].shape[0] == 4 and shard["num_words"] == 5
    assert torch.equal(shard["x"], data["x"][rank * 4:(rank + 1) * 4])
    flat.zero_()
    # only the first two layers get a gradient on rank 1 (static buffer layout: missing grads stay zero)
    h = model[2](model[1](model[0](shard["x"])))
    loss = ((model[3](h) - shard["y"]) ** 2).mean() if rank == 0 else (h ** 2).mean()
    loss.backward()
    local = flat.flat.clone()
    flat.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / world
    ok = torch.allclose(flat.flat, want, rtol=1e-6, atol=1e-7)
    # every parameter's .grad is a view of the flat buffer
    views = all(p.grad.data_ptr() >= flat.flat.data_ptr() and
                p.grad.data_ptr() < flat.flat.data_ptr() + flat.flat.numel() * 4 for p in model.parameters())
    if rank == 0:
        out.put((ok, views, flat.numel, float(flat.flat.abs().sum())))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok, views, numel, total = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and views
    assert numel == 6 * 5 + 5 + 5 * 3 + 3 + 3 * 2 + 2
    assert total > 0


def test_shard_batch_requires_divisible_batch():
    from scan2cap_b200.distributed import shard_batch
    with pytest.raises(AssertionError):
        shard_batch({"x": torch.zeros(5, 2)}, 0, 2)


def test_release_collect_equals_accumulation_into_the_views():
    """FlatGradients.release() / collect() (no per-parameter accumulation kernels) leaves the flat buffer exactly as
    backward into zeroed .grad views does: twice-used parameters summed, unused parameters zero, .grad = the views."""
    from scan2cap_b200.distributed import FlatGradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 5), torch.nn.Linear(5, 2))
    flat = FlatGradients(model)
    x = torch.randn(7, 6)

    def loss():
        h = model[1](model[0](x))
        return (model[2](model[2](h)) ** 2).mean()   # model[2] used twice, model[3] not at all

    flat.flat.fill_(3.0)                              # stale content must not survive
    flat.zero_()
    loss().backward()
    want = flat.flat.clone()
    flat.flat.fill_(7.0)                              # stale content again: collect() must overwrite every slice
    flat.release()
    assert all(p.grad is None for p in model.parameters())
    loss().backward()
    flat.collect()
    for p, v in zip(flat.params, flat.views):
        assert p.grad is v
    got = torch.cat([v.reshape(-1) for v in flat.views])
    ref = torch.cat([want[o:o + p.numel()] for p, o in zip(flat.params, flat.offsets)])
    assert torch.equal(got, ref)
    assert float(model[3].weight.grad.abs().sum()) == 0.0 and float(model[2].weight.grad.abs().sum()) > 0.0

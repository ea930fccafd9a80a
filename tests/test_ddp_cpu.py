"""Host-side data-parallel logic on CPU (gloo, world_size 2): gradient mean all-reduce and batch sharding."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hallucidet_b200.train import allreduce_mean_, shard_batch
    torch.manual_seed(rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    allreduce_mean_([flat], world)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered) / world
    ok = torch.allclose(flat, want, atol=1e-6)
    # bucketed, overlapped path (grad_bucket_hook): two asynchronous bucket all-reduces in reverse-forward order, then the wait
    from hallucidet_b200.train import HalluciDetTrainer
    t = HalluciDetTrainer.__new__(HalluciDetTrainer)
    torch.nn.Module.__init__(t)
    t.world, t._bucket_works = world, []
    flat2 = mine.clone()
    t._flat_grad = lambda: flat2
    t._allreduce_bucket(flat2[600:])          # "late" bucket first (head, decoder, layer4)
    t._allreduce_bucket(flat2[:600])
    t.allreduce_gradients(scale=True)
    ok = ok and torch.allclose(flat2, want, atol=1e-6) and not t._bucket_works
    flat3 = mine.clone()
    t._flat_grad = lambda: flat3
    t._allreduce_bucket(flat3[600:])
    t._allreduce_bucket(flat3[:600])
    t.allreduce_gradients(scale=False)        # fused optimizer: the sum is kept, 1/world is applied by the Adam kernel
    ok = ok and torch.allclose(flat3, want * world, atol=1e-5)
    sl = shard_batch(16, rank, world)
    ok = ok and (sl.start, sl.stop) == (rank * 8, rank * 8 + 8)
    results[rank] = bool(ok)
    dist.destroy_process_group()


def test_allreduce_mean_and_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert results.get(0) and results.get(1)


def test_shard_batch_errors():
    from hallucidet_b200.train import shard_batch
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
    assert shard_batch(8, 0, 1) == slice(0, 8)

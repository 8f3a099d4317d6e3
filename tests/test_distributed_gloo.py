"""BASELINE.json config #1: ResNet-20/CIFAR-shape, top-k 1% + bloom-index allgather,
world_size=2 on CPU/gloo (plumbing, no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cfg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    import deepreduce_b200 as dr
    from deepreduce_b200.models import resnet20
    from deepreduce_b200.trainer import Trainer
    model = resnet20()
    tr = Trainer(model, cfg, lr=0.05, amp_dtype=None)
    torch.manual_seed(100 + rank)
    x = torch.randn(8, 3, 32, 32)
    y = torch.randint(0, 10, (8,))
    losses = [float(tr.step(x, target=y)) for _ in range(3)]
    # all ranks must hold identical parameters after synchronous steps
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ret["losses"] = losses
        ret["same"] = same
        ret["bytes"] = tr.ddp.wire_bytes_per_step()
        ret["dense"] = tr.ddp.dense_bytes()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [
    {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
     'deepreduce': 'index', 'index': 'bloom'},
    {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
     'deepreduce': 'both'},
    {'compressor': 'none', 'memory': 'none', 'communicator': 'allreduce'},
])
@pytest.mark.timeout(300)
def test_resnet20_world2_gloo(cfg):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), cfg, ret), nprocs=2, join=True)
    assert ret["same"], "ranks diverged"
    assert all(l == l and l < 20 for l in ret["losses"])
    if cfg['compressor'] == 'topk':
        assert ret["bytes"] < 0.05 * ret["dense"]          # < 5 % of the dense volume on the wire


def _fuzz_worker(rank, world, port, n_iter, seed, ret):
    import random
    import warnings
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    warnings.simplefilter("ignore")
    import deepreduce_b200 as dr
    rnd = random.Random(seed)                      # the same configuration stream on every rank
    vals, idxs = ['polyfit', 'qsgd', 'gzip', 'dexp', 'polyfit_cpu'], ['bloom', 'rle', 'huffman', 'integer', 'bloom_cpu']
    done = 0
    for it in range(n_iter):
        d = rnd.choice([7, 999, 1001, 4097, 9001])
        comp, extra = rnd.choice([('topk', {}), ('threshold', {'threshold': 1.0}), ('randomk', {})])
        mode = rnd.choice(['value', 'index', 'both', None])
        cfg = {'compressor': comp, 'memory': rnd.choice(['none', 'residual']), 'communicator': 'allgather',
               'compress_ratio': rnd.choice([0.01, 0.1, 0.5]), 'value': rnd.choice(vals), 'index': rnd.choice(idxs),
               'policy': rnd.choice(['leftmost', 'random', 'p0', 'conflict_sets']), 'min_numel': rnd.choice([0, 1000]),
               **extra}
        if mode:
            cfg['deepreduce'] = mode
        kind = rnd.choice(['randn', 'zero_on_rank1', 'scaled'])
        g = torch.randn(d, generator=torch.Generator().manual_seed(1000 * it + rank))
        if kind == 'zero_on_rank1' and rank == 1:
            g = torch.zeros(d)                     # threshold sparsifier -> an empty selection on one rank only
        if kind == 'scaled':
            g = g * (0.1 if rank == 0 else 3.0)    # very different per-rank selection sizes
        out = dr.deepreduce_from_params(cfg).step(g.clone(), 'w')
        assert out.shape == g.shape and torch.isfinite(out).all(), cfg
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], o) for o in gathered), ("ranks disagree", cfg)
        done += 1
    if rank == 0:
        ret["done"] = done
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_fuzz_allgather_variable_sizes_world2():
    """Per-rank payload sizes differ (threshold sparsifier, P0, lossless codecs), one rank may select nothing:
    the size-gather / pad / slice path of the Allgather communicator must hand every rank the same aggregate."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fuzz_worker, args=(2, _free_port(), 60, 0, ret), nprocs=2, join=True)
    assert ret["done"] == 60


def _randomk_allreduce_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import deepreduce_b200 as dr
    d, ratio = 5000, 0.05
    grc = dr.deepreduce_from_params({'compressor': 'randomk', 'memory': 'none', 'communicator': 'allreduce', 'compress_ratio': ratio})
    ok = True
    for step in range(3):
        g = torch.randn(d, generator=torch.Generator().manual_seed(10 * step + rank))
        out = grc.step(g.clone(), 'w')
        gs = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gs, g)
        mean = sum(gs) / world
        nz = torch.nonzero(out).flatten()
        # shared random indices (same seed on every rank): the aggregate is the mean of the ranks' values THERE — an
        # all-reduce that also summed the index tensors would scatter to W * idx (out of range / wrong places)
        ok = ok and 0 < nz.numel() <= int(d * ratio) and torch.allclose(out[nz], mean[nz], atol=1e-6)
        outs = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(outs, out)
        ok = ok and all(torch.equal(outs[0], o) for o in outs)
    if rank == 0:
        ret["ok"] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_randomk_with_allreduce_world2():
    """GRACE's randomk + allreduce pairing (config.py allows it): only the value component may be all-reduced."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_randomk_allreduce_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["ok"]

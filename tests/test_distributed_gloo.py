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

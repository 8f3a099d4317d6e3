"""CLI, checkpoint/resume of the compression state, gradient accumulation (CPU)."""
import json

import torch

from deepreduce_b200 import cli
from deepreduce_b200.models import resnet20
from deepreduce_b200.trainer import Trainer
from deepreduce_b200.utils.checkpoint import load_checkpoint, save_checkpoint

CFG = {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
       'deepreduce': 'index', 'index': 'bloom'}


def test_cli_runs_and_reports_volume(capsys, tmp_path):
    cli.main(["-a", "resnet20", "--batch-size", "4", "--steps", "2", "--cpu", "--log_volume", "--log_time",
              "--train_dir", str(tmp_path), "--grace_config", str(CFG)])
    out = capsys.readouterr().out.strip().splitlines()
    rep = json.loads(out[-1])
    assert rep["arch"] == "resnet20" and 0 < rep["relative_volume"] < 0.05 and rep["s_per_step"] > 0
    assert (tmp_path / "ckpt.pt").exists()


def _batch(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 10, (4,), generator=g)


def test_checkpoint_resume_restores_residuals(tmp_path):
    torch.manual_seed(0)
    a = Trainer(resnet20(), CFG, lr=0.05, amp_dtype=None)
    for s in range(2):
        x, y = _batch(s)
        a.step(x, target=y)
    save_checkpoint(str(tmp_path / "c.pt"), a)
    torch.manual_seed(1)
    b = Trainer(resnet20(), CFG, lr=0.05, amp_dtype=None)
    load_checkpoint(str(tmp_path / "c.pt"), b)
    x, y = _batch(9)
    la, lb = a.step(x, target=y), b.step(x, target=y)
    assert torch.equal(la, lb)
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(pa, pb)            # identical update => residual memory was restored too


def test_gradient_accumulation_exchanges_every_n():
    torch.manual_seed(0)
    t = Trainer(resnet20(), CFG, lr=0.05, amp_dtype=None, accum_steps=2)
    before = [p.detach().clone() for p in t.model.parameters()]
    x, y = _batch(0)
    t.step(x, target=y)
    assert all(torch.equal(a, b) for a, b in zip(before, t.model.parameters()))      # no update on the 1st micro-step
    x, y = _batch(1)
    t.step(x, target=y)
    assert any(not torch.equal(a, b) for a, b in zip(before, t.model.parameters()))
    assert t.ddp.step_count == 1

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


def test_diagnostic_writers(tmp_path):
    """The remaining writers of compression_utils.hpp / integer_compression.cc / logger.cc."""
    import numpy as np
    import torch
    from deepreduce_b200.utils import (StepLogger, bitstream_str, log_bitstream_compressor, log_bitstream_decompressor,
                                       log_decompressor, log_integer_codec)
    assert bitstream_str(np.array([1, 128], dtype=np.uint8)).split("[")[1].split("]")[0].split() == \
        list("1000000000000001")
    root = str(tmp_path)
    p = log_decompressor(root, 0, 3, 7, N=16, bloom_words=torch.tensor([5], dtype=torch.int32),
                         selected_indices=torch.tensor([1, 4]), values=torch.tensor([0.5, -1.0]),
                         decompressed=torch.zeros(16), policy="leftmost", suffix=1)
    assert p.endswith("decompressor_logs_leftmost_1.txt") and "Indices Chosen: [1, 4]" in open(p).read()
    assert log_decompressor(root, 0, 3, 7, N=16, bloom_words=torch.tensor([5], dtype=torch.int32),
                            selected_indices=torch.tensor([1]), values=torch.tensor([0.5]),
                            decompressed=torch.zeros(16), verbosity=1) is None
    enc = torch.tensor([3, 9, 27], dtype=torch.uint8)
    st = log_bitstream_compressor(root, 0, 3, 8, indices=torch.tensor([2, 3, 9]), encoded=enc, initial_bits=32 * 3,
                                  runs=torch.tensor([2, 2, 5, 1]), verbosity=2)
    assert st == {"initial_bits": 96, "final_bits": 3 * 8 + 32}
    d = tmp_path / "0" / "step_3" / "8"
    assert (d / "stats.txt").read_text().startswith("Initial_Size: 96  Final_Size: 56")
    assert "Lengths:" in (d / "RleCompressor_logs.txt").read_text()
    assert log_bitstream_decompressor(root, 0, 3, 8, encoded=enc, indices=torch.tensor([2, 3, 9]), suffix=2).endswith("_2.txt")
    assert log_integer_codec(root, 5, 1, input_words=torch.arange(8), encoded_words=torch.arange(3), verbosity=2) is None
    st = log_integer_codec(root, 4, 1, input_words=torch.arange(8), encoded_words=torch.arange(3), verbosity=2)
    assert st["initial_bits"] == 256 and st["final_bits"] == 3 * 32 + 32 and abs(st["rate"] - 0.375) < 1e-9
    lg = StepLogger(root, gradient_id=2, rank=1, verbosity_frequency=10)
    assert lg(torch.arange(4.0), torch.tensor([1.0, 2.0]), step=7) is False
    assert lg(torch.arange(4.0), torch.tensor([1.0, 2.0]), step=20) is True
    vals = (tmp_path / "1" / "step_20" / "2" / "values.csv").read_text().split()
    assert [float(v) for v in vals] == [0.0, 1.0, 2.0, 3.0]
    assert StepLogger(root, 2)(torch.zeros(2), torch.zeros(1), step=0) is False      # frequency 0 = never


def test_bench_contract_pieces_on_cpu():
    """bench.py: defaults (N=1, W >= 3), the roofline arithmetic, and the no-GPU answer on stdout as ONE JSON line."""
    import importlib.util
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec_ = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(b)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = b.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1 and a.impl == "ours" and a.config == "bloom" and a.model == "resnet50"
    r = b.exchange_roofline(0.724, 102228128, 1612608, 8)
    assert r["bound"] == "hbm" and abs(r["hbm_bound_ms"] - 4 * 102228128 / (r["hbm_gbs_measured"] * 1e9) * 1e3) < 1e-9
    assert 0.05 < r["frac_of_roofline"] < 0.2 and r["nvlink_bytes_out"] == 2 * 7 * 1612608
    assert b.exchange_roofline(0.5, 102228128, 1612608, 1)["nvlink_bytes_out"] == 0
    import torch
    if not torch.cuda.is_available():
        for impl in ("ours", "reference"):
            p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", impl], capture_output=True,
                               text=True, timeout=300)
            lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
            assert p.returncode == 0 and len(lines) == 1, p.stdout + p.stderr[-500:]
            out = json.loads(lines[0])
            assert out["impl"] == impl and "unavailable" in out

"""Training CLI — the stand-in for the reference's trainers (grace-benchmarks ``trainer_grace.py`` /
``ncf_grace.py`` / ``tf_cnn_benchmarks.py``; reference run_deepreduce.sh:11-107).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m deepreduce_b200.cli \
        -a resnet50 --batch-size 256 --steps 100 --log_volume --log_time \
        --grace_config="{'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather',
                         'compress_ratio': 0.01, 'deepreduce': 'index', 'index': 'bloom'}"

Flags mirror the reference scripts: ``--grace_config`` (python-dict literal), ``--log_volume``, ``--log_time``,
``--micro_benchmark``, ``--weak_scaling``, ``--grads_accumulated``, ``--load_checkpoint_path`` / ``--train_dir``.
Data is synthetic (there is no dataset in this environment); shapes follow the named benchmark.
"""
from __future__ import annotations

import argparse
import ast
import json
import os
import time

import torch
import torch.distributed as dist


def build(arch: str, device):
    from . import models as M
    if arch == "resnet20":
        return M.resnet20().to(device), dict(kind="image", hw=32, classes=10)
    if arch == "resnet50":
        return M.resnet50().to(device), dict(kind="image", hw=224, classes=1000)
    if arch == "vgg16":
        return M.VGG16().to(device), dict(kind="image", hw=32, classes=10)
    if arch == "densenet40":
        return M.DenseNet40().to(device), dict(kind="image", hw=32, classes=10)
    if arch == "mobilenet":
        return M.MobileNet().to(device), dict(kind="image", hw=32, classes=10)
    if arch == "ncf":
        return M.NeuMF().to(device), dict(kind="ncf", users=138493, items=26744)
    if arch == "lstm":
        return M.NextWordLSTM().to(device), dict(kind="lm", vocab=10004, seq=20)
    if arch == "bert_large":
        return M.bert_large().to(device), dict(kind="bert", vocab=30522, seq=128)
    raise SystemExit(f"unknown arch {arch}")


def synth_batch(meta, bs, device, gen):
    k = meta["kind"]
    if k == "image":
        return (torch.randn(bs, 3, meta["hw"], meta["hw"], device=device, generator=gen),), \
            torch.randint(0, meta["classes"], (bs,), device=device, generator=gen)
    if k == "ncf":
        u = torch.randint(0, meta["users"], (bs,), device=device, generator=gen)
        i = torch.randint(0, meta["items"], (bs,), device=device, generator=gen)
        return (u, i), torch.randint(0, 2, (bs,), device=device, generator=gen).float()
    if k == "lm":
        t = torch.randint(0, meta["vocab"], (bs, meta["seq"] + 1), device=device, generator=gen)
        return (t[:, :-1],), t[:, 1:]
    if k == "bert":
        t = torch.randint(0, meta["vocab"], (bs, meta["seq"]), device=device, generator=gen)
        return (t,), t
    raise ValueError(k)


def loss_for(meta):
    k = meta["kind"]
    if k == "ncf":
        return torch.nn.functional.binary_cross_entropy_with_logits
    if k == "lm":
        return lambda out, y: torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), y.reshape(-1))
    if k == "bert":
        return lambda out, y: torch.nn.functional.cross_entropy(
            (out.logits if hasattr(out, "logits") else out).reshape(-1, meta["vocab"]).float(), y.reshape(-1))
    return torch.nn.functional.cross_entropy


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-a", "--arch", default="resnet20")
    ap.add_argument("--batch-size", type=int, default=128, help="per-process batch (weak scaling, like the reference)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--lr", type=float, default=0.05)
    ap.add_argument("--grace_config", default="{'compressor': 'none', 'memory': 'none', 'communicator': 'allreduce'}")
    ap.add_argument("--log_volume", action="store_true")
    ap.add_argument("--log_time", action="store_true")
    ap.add_argument("--micro_benchmark", action="store_true")
    ap.add_argument("--weak_scaling", action="store_true", help="accepted for script compatibility (always weak)")
    ap.add_argument("--grads_accumulated", type=int, default=1)
    ap.add_argument("--train_dir", default=None, help="checkpoint directory (model + optimizer + residual state)")
    ap.add_argument("--load_checkpoint_path", default=None)
    ap.add_argument("--seed", type=int, default=44)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args(argv)

    params = ast.literal_eval(args.grace_config)
    if args.micro_benchmark:
        params["micro-benchmark"] = True
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    use_cuda = torch.cuda.is_available() and not args.cpu
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo")
    device = torch.device("cuda" if use_cuda else "cpu")
    torch.manual_seed(args.seed)
    model, meta = build(args.arch, device)
    from .trainer import Trainer
    from .utils.checkpoint import load_checkpoint, save_checkpoint
    tr = Trainer(model, params, lr=args.lr, amp_dtype=torch.bfloat16 if use_cuda else None,
                 channels_last=meta["kind"] == "image", loss_fn=loss_for(meta),
                 bucket_cap_mb=32.0 if world == 1 else 1e9,      # N > 1: one bucket after backward (bench.py --bucket-mb)
                 accum_steps=args.grads_accumulated)
    if args.load_checkpoint_path:
        load_checkpoint(args.load_checkpoint_path, tr)
    gen = torch.Generator(device=device).manual_seed(1000 + rank)
    t_fb = t_comm = 0.0
    for step in range(args.steps):
        x, y = synth_batch(meta, args.batch_size, device, gen)
        t0 = time.perf_counter()
        loss = tr.step(*x, target=y)
        if args.log_time:
            if use_cuda:
                torch.cuda.synchronize()
            t_fb += time.perf_counter() - t0
        if rank == 0 and (step % 10 == 0 or step == args.steps - 1):
            print(f"step {step} loss {float(loss):.4f}", flush=True)
    if rank == 0:
        out = {"arch": args.arch, "world": world, "steps": args.steps, "params": {k: v for k, v in params.items()}}
        if args.log_volume:
            out["wire_bytes_per_step_per_rank"] = tr.ddp.wire_bytes_per_step()
            out["dense_bytes"] = tr.ddp.dense_bytes()
            out["relative_volume"] = out["wire_bytes_per_step_per_rank"] / out["dense_bytes"]
            if getattr(tr.ddp, "fused", False):          # counters the kernel left in the slot headers (last step)
                st = tr.ddp.exchange_stats()
                out["exchange_stats"] = {k: st[k] for k in ("k", "n_sel", "n_pos", "false_pos", "value_bytes",
                                                            "index_bytes", "header_bytes") if k in st}
        if args.log_time:
            out["s_per_step"] = t_fb / args.steps
        from .utils import METRICS
        out["metrics"] = METRICS.summary()
        print(json.dumps(out), flush=True)
    if args.train_dir:                     # model/optimizer once (rank 0), residual + select state per rank
        os.makedirs(args.train_dir, exist_ok=True)
        save_checkpoint(os.path.join(args.train_dir, "ckpt.pt"), tr)
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

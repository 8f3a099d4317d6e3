#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ResNet-50 images/sec (whole job, device-timed,
max over ranks) with top-k 1 % + bloom-index + residual gradient exchange, synthetic
224² data, random-init weights, bf16 autocast.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # unmodified reference file + GRACE/cupy shims (baseline/)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json config 2 (default) and 3; plus context rows
    "bloom": {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
              'deepreduce': 'index', 'index': 'bloom'},
    "bloom_p0": {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
                 'deepreduce': 'index', 'index': 'bloom', 'policy': 'p0'},
    "both": {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
             'deepreduce': 'both', 'index': 'bloom', 'value': 'polyfit'},
    "topk": {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01},
    # BASELINE.json config 4: NCF top-k 0.1 % + run-length index
    "rle": {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.001,
            'deepreduce': 'index', 'index': 'rle'},
    "dense": {'compressor': 'none', 'memory': 'none', 'communicator': 'allreduce'},
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bloom", choices=sorted(CONFIGS))
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling); default 256 (resnet50), 8 (bert_large), 65536 (ncf)")
    ap.add_argument("--seq", type=int, default=128, help="sequence length for bert_large")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-thread", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=128.0,
                    help="flat bucket size; on a compute-saturated GPU one bucket launched at the end of backward is fastest (profiles/)")
    ap.add_argument("--blocks-per-sm", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also time the exchange kernel alone (extra keys)")
    return ap.parse_args()


def exchange_roofline(exchange_ms, dense_bytes, wire_bytes, world):
    """Achieved fraction of the exchange kernel's roofline = the slower of (a) its unavoidable HBM traffic at the
    MEASURED copy bandwidth (read g, read r, write r, write the dense result: 4 x dense bytes) and (b) the bytes it
    sends over NVLink at link bandwidth (the slot to W-1 peers, plus about as much again for the decoded slices of
    the sharded decode).  Denominators: MEASURED_PEAKS.json (fallback: the profiling recipe's 6 650 GB/s) and the
    guide's 770 GB/s/direction measured peer copy."""
    hbm = 6650.0
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f).get("hbm_gbs", hbm))
    except Exception:
        pass
    nvlink = 770.0
    t_hbm = 4.0 * dense_bytes / (hbm * 1e9) * 1e3
    nv_bytes = 2.0 * (world - 1) * wire_bytes
    t_nv = nv_bytes / (nvlink * 1e9) * 1e3
    return {"hbm_min_bytes": int(4 * dense_bytes), "hbm_gbs_measured": hbm, "hbm_bound_ms": t_hbm,
            "nvlink_bytes_out": int(nv_bytes), "nvlink_gbs_per_dir": nvlink, "nvlink_bound_ms": t_nv,
            "bound": "hbm" if t_hbm >= t_nv else "nvlink", "frac_of_roofline": max(t_hbm, t_nv) / exchange_ms,
            "compressed_allgather_bus_gbs": nv_bytes / (exchange_ms * 1e-3) / 1e9}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def init_dist(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    need = world > 1 or args.impl == "reference"
    if need:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(free_port()))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    return rank, world, local


def build_model(name):
    import torch
    if name == "resnet50":
        from deepreduce_b200.models import resnet50
        return resnet50()
    if name == "resnet20":
        from deepreduce_b200.models import resnet20
        return resnet20()
    raise ValueError(name)


def max_over_ranks(x, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world):
    """K steps bracketed by barrier + synchronize, CUDA events on the launching stream."""
    import torch
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world), max_over_ranks(wall, world)


def model_spec(args):
    """(model, kind, default per-GPU batch, metric unit)"""
    from deepreduce_b200 import models as M
    if args.model == "resnet50":
        return M.resnet50(), "image224", 256, "images/s"
    if args.model == "resnet20":
        return M.resnet20(), "image32", 256, "images/s"
    if args.model == "bert_large":
        return M.bert_large(seq_len=max(args.seq, 128)), "bert", 8, "sequences/s"
    if args.model == "ncf":
        return M.NeuMF(), "ncf", 65536, "samples/s"
    raise ValueError(args.model)


def run_ours(args, rank, world, local):
    import torch
    from deepreduce_b200 import ops
    from deepreduce_b200.trainer import Trainer
    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    model, kind, default_b, unit = model_spec(args)
    model = model.cuda()
    B = args.batch or default_b
    cfg = dict(CONFIGS[args.config])
    amp = torch.bfloat16 if args.dtype == "bf16" else None
    gen = torch.Generator().manual_seed(77 + rank)
    loss_fn = None
    if kind.startswith("image"):
        hw = 224 if kind == "image224" else 32
        ncls = 1000 if kind == "image224" else 10
        pool = [(torch.randint(0, 256, (B, hw, hw, 3), dtype=torch.uint8, generator=gen),) for _ in range(2)]
        tgt = [torch.randint(0, ncls, (B,), generator=gen) for _ in range(2)]
    elif kind == "bert":
        V = 30522
        pool = [(torch.randint(0, V, (B, args.seq), generator=gen),) for _ in range(2)]
        tgt = [p[0].clone() for p in pool]
        loss_fn = lambda out, y: torch.nn.functional.cross_entropy(out.logits.reshape(-1, V).float(), y.reshape(-1))  # noqa: E731
    else:   # ncf
        pool = [(torch.randint(0, 138493, (B,), generator=gen), torch.randint(0, 26744, (B,), generator=gen)) for _ in range(2)]
        tgt = [torch.randint(0, 2, (B,), generator=gen).float() for _ in range(2)]
        loss_fn = torch.nn.functional.binary_cross_entropy_with_logits
    tr = Trainer(model, cfg, lr=0.05 if kind != "bert" else 1e-4, amp_dtype=amp, channels_last=kind.startswith("image"),
                 overlap=not args.no_overlap, bucket_cap_mb=args.bucket_mb, background_thread=not args.no_thread,
                 blocks_per_sm=args.blocks_per_sm, u8_input=kind.startswith("image"), loss_fn=loss_fn)
    dev_x = [tuple(t.cuda() for t in p) for p in pool]
    dev_y = [t.cuda() for t in tgt]

    def step(i):
        tr.step(*dev_x[i & 1], target=dev_y[i & 1])

    for i in range(args.warmup):
        step(i)
    tr.ddp.check()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    ms, wall = timed(step, args.steps, world)
    launches = ops.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    tr.ddp.check()
    value = world * B * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        host_x = [tuple(t.pin_memory() for t in p) for p in pool]
        host_y = [t.pin_memory() for t in tgt]

        def step_e2e(i):
            nxt = (host_x[(i + 1) & 1], host_y[(i + 1) & 1])
            tr.step_host(host_x[i & 1], host_y[i & 1], next_batch=nxt)

        for i in range(2):
            step_e2e(i)
        _, wall_e = timed(step_e2e, args.steps, world)
        e2e = {"value": world * B * args.steps / (wall_e / 1e3), "unit": unit,
               "h2d_bytes_per_step": int(tr.h2d_bytes), "d2h_bytes_per_step": int(tr.d2h_bytes)}

    extra = {}
    if args.breakdown and tr.ddp.engines:
        def ex(i):       # exchange kernel alone on the last gradients (all buckets back to back)
            for e in tr.ddp.engines:
                e.step()
        for i in range(3):
            ex(i)
        ms_ex, _ = timed(ex, 20, world)
        extra["exchange_ms_per_step"] = ms_ex / 20
        extra["engine_grid"] = tr.ddp.engines[0].grid()
        extra["roofline"] = exchange_roofline(extra["exchange_ms_per_step"], tr.ddp.dense_bytes(),
                                              tr.ddp.wire_bytes_per_step(), world)
    wire = tr.ddp.wire_bytes_per_step()
    dense = tr.ddp.dense_bytes()
    names = {"resnet50": "ResNet-50 images/sec (whole job, device-timed, max over ranks)",
             "bert_large": "BERT-large sequences/sec (whole job, device-timed, max over ranks)",
             "ncf": "NCF (MovieLens-20M shapes) samples/sec (whole job, device-timed, max over ranks)"}
    out = {
        "metric": names.get(args.model, f"{args.model} {unit}"),
        "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic (shapes of the named benchmark, random-init weights)", "impl": "ours",
        "config": {"model": args.model, "global_batch": B * world, "per_gpu_batch": B,
                   "seq_len": args.seq if kind == "bert" else None,
                   "parallelism": f"dp{world}", "gradient_exchange": args.config, "params": cfg,
                   "l2": "working set (activations + fp32 gradients + residual) exceeds the 126 MB L2 every step",
                   "overlap": not args.no_overlap, "buckets": len(tr.ddp.flat), "bucket_mb": args.bucket_mb},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "wire_bytes_per_step_per_rank": int(wire), "dense_bytes": int(dense),
        "relative_volume": wire / dense, "wall_ms_per_step": wall / args.steps,
    }
    out.update(extra)
    tr.close()
    return out


def run_reference(args, rank, world, local):
    """UNMODIFIED reference pytorch/deepreduce.py through its documented API (README.md:36-48):
    grace_from_params + IndexCompressor wrapper, grc.step(grad, name) per tensor after backward."""
    import numpy as np
    import torch
    ref_file = os.path.join(ROOT, "baseline", "_ref", "deepreduce_ref", "deepreduce.py")
    if not os.path.exists(ref_file):
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        try:
            import install_reference
            install_reference.install(verbose=False)
        except Exception:
            pass
    if not os.path.exists(ref_file):
        return {"impl": "reference", "unavailable": "reference not installable offline and /root/reference absent on this box"}
    sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    if not hasattr(np, "RankWarning"):
        np.RankWarning = np.exceptions.RankWarning        # numpy>=2 moved it; the reference reads np.RankWarning
    from deepreduce_ref import deepreduce as R
    from grace_dl.dist.helper import grace_from_params
    import torchvision

    torch.manual_seed(1234)
    torch.backends.cudnn.benchmark = True
    assert args.model == "resnet50"
    model = torchvision.models.resnet50(weights=None).cuda().to(memory_format=torch.channels_last)
    cfg = dict(CONFIGS[args.config])
    cfg["world_size"] = world
    grc = grace_from_params(cfg)
    if cfg.get("deepreduce"):
        d_max = max(p.numel() for p in model.parameters())
        g = torch.Generator(device="cuda").manual_seed(18)
        cfg["hash_table"] = torch.randint(0, 2 ** 31 - 1, (d_max, 16), dtype=torch.int32, device="cuda", generator=g)
        wrapper = {'value': R.ValueCompressor, 'index': R.IndexCompressor, 'both': R.DeepReduce}[cfg["deepreduce"]]
        grc.compressor = wrapper(grc.compressor, cfg)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    B = args.batch or 256
    gen = torch.Generator().manual_seed(77 + rank)
    host_x = [torch.randn(B, 3, 224, 224, generator=gen).pin_memory() for _ in range(2)]
    host_y = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(2)]
    dev_x = [x.cuda().contiguous(memory_format=torch.channels_last) for x in host_x]
    dev_y = [y.cuda() for y in host_y]
    amp = args.dtype == "bf16"
    named = [(n, p) for n, p in model.named_parameters()]

    def train(x, y):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = model(x)
        loss = torch.nn.functional.cross_entropy(out.float(), y)
        loss.backward()
        for n, p in named:
            p.grad = grc.step(p.grad, n).view_as(p)
        opt.step()
        return loss

    def step(i):
        train(dev_x[i & 1], dev_y[i & 1])

    for i in range(args.warmup):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall = timed(step, args.steps, world)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        def step_e2e(i):
            x = host_x[i & 1].cuda(non_blocking=True).contiguous(memory_format=torch.channels_last)
            y = host_y[i & 1].cuda(non_blocking=True)
            float(train(x, y).item())
        step_e2e(0)
        n_e = max(2, min(args.steps, 5))
        _, wall_e = timed(step_e2e, n_e, world)
        e2e = {"value": world * B * n_e / (wall_e / 1e3), "unit": "images/s",
               "h2d_bytes_per_step": int(host_x[0].numel() * 4 + host_y[0].numel() * 8), "d2h_bytes_per_step": 4}
    return {
        "metric": "ResNet-50 images/sec (whole job, device-timed, max over ranks)",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic (fp32 224x224x3, random-init weights)", "impl": "reference",
        "config": {"model": "resnet50 (torchvision)", "global_batch": B * world, "per_gpu_batch": B,
                   "parallelism": f"dp{world}", "gradient_exchange": args.config, "params": {k: v for k, v in cfg.items() if isinstance(v, (str, int, float, bool, type(None)))},
                   "harness": "unmodified reference pytorch/deepreduce.py + GRACE/cupy shims (baseline/), per-tensor grc.step after backward"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": 0, "wall_ms_per_step": wall / args.steps,
    }


def main():
    args = parse()
    # stdout carries exactly ONE line (the JSON): keep a private handle to it and point fd 1 at stderr so that
    # library banners (e.g. "NCCL version ...") and stray prints cannot end up next to the result
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "unavailable": "no CUDA device on this box"}), file=result_out, flush=True)
        return 0
    rank, world, local = init_dist(args)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    try:
        out = run_ours(args, rank, world, local) if args.impl == "ours" else run_reference(args, rank, world, local)
    finally:
        pass
    if rank == 0:
        print(json.dumps(out), file=result_out, flush=True)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

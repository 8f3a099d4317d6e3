"""Per-CTA phase timeline of the fused exchange kernel (debug_times facility): for every phase the min / median / max
CTA duration, the spread of the CTAs' start and end times, and the slowest CTAs with what their tile range contains.
    python scripts/cta_timeline.py [hist_shift]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepreduce_b200.models import resnet50  # noqa: E402
from deepreduce_b200.parallel import BucketEngine, BucketPlan  # noqa: E402

PH = ["accum", "fallback", "hist2", "insert", "query", "emit", "apply-items", "apply-barrier", "compact-loop", "rank_exact", "fit", "fix",
      "push", "signal", "expand", "decode", "compact", "push2", "signal2", "scatter"]


def main():
    hs = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    world = int(os.environ.get("WORLD_SIZE", 1)); rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    m = resnet50()
    named = list(reversed([(n, p) for n, p in m.named_parameters()]))
    plan = BucketPlan([p.numel() for _, p in named], [n for n, _ in named], compress_ratio=0.01)
    eng = BucketEngine(plan, device=f"cuda:{local}", hist_shift=hs) if world > 1 else BucketEngine(plan, device="cuda:0", world=1, rank=0, hist_shift=hs)
    G = eng.grid()
    dbg = torch.zeros(21 * G * 2, dtype=torch.int64, device="cuda")   # 20 phases + the %smid row
    eng.ctx.set_debug_times(dbg.data_ptr())
    gen = torch.Generator(device="cuda").manual_seed(rank)
    grads = [torch.randn(plan.total_elems, device="cuda", generator=gen) * 0.01 for _ in range(4)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    for i in range(8):
        eng.grad.copy_(grads[i % 4]); flush.zero_(); dbg.zero_()
        if world > 1:
            dist.barrier()
        eng.step()
    torch.cuda.synchronize()
    if rank != 0:
        eng.close(); dist.destroy_process_group(); return
    t = dbg.cpu().numpy().reshape(21, G, 2).astype(np.int64)
    t0 = t[0, :, 0].min()
    tiles = plan.tile_table().numpy().reshape(-1, 4)
    nt = plan.n_tiles
    ranges = plan.cta_ranges(G, eng.balanced)   # accumulate-class ranges (the other phase classes cut slightly differently)
    if getattr(eng, "cuts", None) is not None:
        c0 = eng.cuts.cpu().numpy()[0]
        ranges = [(int(c0[b]), int(c0[b + 1])) for b in range(G)]
    print(f"world {world} grid {G}, tiles {nt}, balanced {eng.balanced}, kernel span {(t[:, :, 1].max() - t0) / 1e3:.1f} us")
    for ph, name in enumerate(PH):
        s, e = t[ph, :, 0], t[ph, :, 1]
        if s.max() == 0:
            continue
        d = (e - s) / 1e3
        print(f"{name:9s} start {(s.min() - t0) / 1e3:7.1f}..{(s.max() - t0) / 1e3:7.1f} us | end {(e.min() - t0) / 1e3:7.1f}..{(e.max() - t0) / 1e3:7.1f} us | "
              f"dur min {d.min():6.1f} med {np.median(d):6.1f} max {d.max():6.1f}")
        for b in np.argsort(-d)[:4]:
            a, z = ranges[b]
            tens = tiles[a:z, 0]
            print(f"      slow CTA {b:3d}: {d[b]:6.1f} us  tiles [{a},{z})  tensors {len(set(tens.tolist()))}  one-tile tensors {int(((tiles[a:z, 2] >> 31) & 1).sum())}")
    if os.environ.get("DR_TIMELINE_JSON"):          # per-CTA table for fitting the partition's cost model
        import json
        single = (tiles[:, 2] >> 31) & 1
        rows = []
        for b, (a, z) in enumerate(ranges):
            tens = tiles[a:z, 0]
            segs = len(set(tens.tolist()))
            n_single = int(single[a:z].sum())
            rows.append({"cta": b, "tiles": int(z - a), "segments": segs, "single": n_single,
                         "elems": int((tiles[a:z, 2] & 0x7FFFFFFF).sum()),
                         "dur_us": {name: float((t[ph, b, 1] - t[ph, b, 0]) / 1e3) for ph, name in enumerate(PH) if t[ph, :, 0].max() != 0}})
        with open(os.environ["DR_TIMELINE_JSON"], "w") as f:
            json.dump({"world": world, "grid": G, "rows": rows}, f)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

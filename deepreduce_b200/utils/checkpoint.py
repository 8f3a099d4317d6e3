"""Checkpoint / resume including the compression state (the reference's PyTorch path never saves the
residual memory; here model + optimizer + residuals + select history + step counter round-trip).

Model and optimizer state are identical on every rank and are written once (rank 0, ``path``); the compression state
(residuals, select history) is PER RANK and goes to ``path + ".rank{r}"`` — every rank calls ``save_checkpoint`` /
``load_checkpoint``.  A checkpoint written by an older version (everything in one file) still loads.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def save_checkpoint(path: str, trainer) -> None:
    rank, world = _rank_world()
    if rank == 0:
        torch.save({"model": trainer.model.state_dict(), "opt": trainer.opt.state_dict(), "world": world}, path)
    torch.save({"ddp": trainer.ddp.state_dict(), "rank": rank, "world": world}, f"{path}.rank{rank}")
    if world > 1:
        dist.barrier()


def load_checkpoint(path: str, trainer) -> None:
    rank, _ = _rank_world()
    ck = torch.load(path, map_location="cpu", weights_only=True)
    trainer.model.load_state_dict(ck["model"])
    trainer.opt.load_state_dict(ck["opt"])
    mine = f"{path}.rank{rank}"
    if os.path.exists(mine):
        trainer.ddp.load_state_dict(torch.load(mine, map_location="cpu", weights_only=True)["ddp"])
    elif "ddp" in ck:                      # legacy single-file checkpoint: rank 0's compression state only
        trainer.ddp.load_state_dict(ck["ddp"])

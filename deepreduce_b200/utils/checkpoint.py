"""Checkpoint / resume including the compression state (the reference's PyTorch path never saves the
residual memory; here model + optimizer + residuals + select history + step counter round-trip)."""
from __future__ import annotations

import torch


def save_checkpoint(path: str, trainer) -> None:
    torch.save({"model": trainer.model.state_dict(), "opt": trainer.opt.state_dict(),
                "ddp": trainer.ddp.state_dict()}, path)


def load_checkpoint(path: str, trainer) -> None:
    ck = torch.load(path, map_location="cpu", weights_only=True)
    trainer.model.load_state_dict(ck["model"])
    trainer.opt.load_state_dict(ck["opt"])
    trainer.ddp.load_state_dict(ck["ddp"])

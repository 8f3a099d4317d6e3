"""cupy stand-in: the reference only calls asarray / packbits / unpackbits / right_shift
(reference pytorch/deepreduce.py:210-214,244,449,454).  Implemented with torch ops on the
same device so `torch.as_tensor(cupy.xxx(...), device="cuda")` keeps working."""
import torch


def asarray(t):
    return t


def packbits(t):
    t = t.to(torch.uint8).flatten()
    pad = (-t.numel()) % 8
    if pad:
        t = torch.cat([t, t.new_zeros(pad)])
    w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=t.device)
    return (t.view(-1, 8) * w).sum(dim=1).to(torch.uint8)


def unpackbits(t):
    t = t.to(torch.uint8).flatten()
    sh = torch.tensor([7, 6, 5, 4, 3, 2, 1, 0], dtype=torch.uint8, device=t.device)
    return ((t[:, None] >> sh[None, :]) & 1).flatten().to(torch.uint8)


def right_shift(t, n):
    return t >> n

"""grace_from_params / tensor_bits + the GRACE pieces the reference's configs name
(topk, residual, allgather, allreduce) — SURVEY.md Appendix A."""
import torch
import torch.distributed as dist

from . import Communicator, Compressor, Memory


def tensor_bits(tensors):
    total = 0
    for t in tensors:
        total += t.numel() * t.element_size() * 8
    return total


class NoneCompressor(Compressor):
    def compress(self, tensor, name):
        return [tensor], None

    def decompress(self, tensors, ctx):
        return tensors[0]


class TopKCompressor(Compressor):
    def __init__(self, compress_ratio):
        super().__init__()
        self.compress_ratio = compress_ratio

    def compress(self, tensor, name):
        flat = tensor.flatten()
        k = max(1, int(flat.numel() * self.compress_ratio))
        _, indices = torch.topk(flat.abs(), k, sorted=False)
        values = torch.gather(flat, 0, indices)
        return (values, indices), tensor.size()

    def decompress(self, tensors, ctx):
        values, indices = tensors
        out = torch.zeros(ctx.numel(), dtype=values.dtype, layout=values.layout, device=values.device)
        out.scatter_(0, indices.long(), values)
        return out.view(ctx)


class NoneMemory(Memory):
    def compensate(self, tensor, name):
        return tensor


class ResidualMemory(Memory):
    def __init__(self, beta=1.0, gamma=1.0):
        self.residuals = {}
        self.beta = beta
        self.gamma = gamma

    def compensate(self, tensor, name):
        if name in self.residuals:
            tensor = self.beta * self.residuals[name] + self.gamma * tensor
        return tensor

    def update(self, tensor, name, compressor, tensor_compressed, ctx):
        tensor_decompressed = compressor.decompress(tensor_compressed, ctx)
        self.residuals[name] = tensor - tensor_decompressed


class Allgather(Communicator):
    def __init__(self, compressor, memory, world_size):
        super().__init__(compressor, memory)
        self.world_size = world_size

    def send_receive(self, tensors, name, ctx):
        W = self.world_size
        if self.compressor.tensors_size_are_same:
            gathered = []
            for t in tensors:
                out = [torch.empty_like(t) for _ in range(W)]
                dist.all_gather(out, t.contiguous())
                gathered.append(out)
        else:
            sizes = torch.tensor([t.numel() for t in tensors], device=tensors[0].device)
            all_sizes = [torch.empty_like(sizes) for _ in range(W)]
            dist.all_gather(all_sizes, sizes)
            all_sizes = torch.stack(all_sizes).cpu()
            gathered = []
            for c, t in enumerate(tensors):
                flat = t.flatten()
                mx = int(all_sizes[:, c].max())
                if mx > flat.numel():
                    flat = torch.cat([flat, flat.new_zeros(mx - flat.numel())])
                out = [torch.empty_like(flat) for _ in range(W)]
                dist.all_gather(out, flat.contiguous())
                gathered.append([o[: int(all_sizes[r, c])] for r, o in enumerate(out)])
        dense = [self.compressor.decompress([g[r] for g in gathered], ctx) for r in range(W)]
        out = self.compressor.aggregate(dense)
        return out / W if self.compressor.average else out


class Allreduce(Communicator):
    def __init__(self, compressor, memory, world_size):
        super().__init__(compressor, memory)
        self.world_size = world_size

    def send_receive(self, tensors, name, ctx):
        for t in tensors:
            dist.all_reduce(t)
        out = self.compressor.decompress(tensors, ctx)
        return out / self.world_size if self.compressor.average else out


def grace_from_params(params):
    comp = params.get('compressor', 'none')
    world_size = params.get('world_size', dist.get_world_size() if dist.is_initialized() else 1)
    compressor = TopKCompressor(params.get('compress_ratio', 0.01)) if comp == 'topk' else NoneCompressor()
    memory = ResidualMemory() if params.get('memory', 'none') == 'residual' else NoneMemory()
    if params.get('communicator', 'allreduce') == 'allgather':
        return Allgather(compressor, memory, world_size)
    return Allreduce(compressor, memory, world_size)

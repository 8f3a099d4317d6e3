from abc import ABC, abstractmethod

import torch
import torch.distributed as dist


class Memory(ABC):
    @abstractmethod
    def compensate(self, tensor, name):
        raise NotImplementedError

    def update(self, tensor, name, compressor, tensor_compressed, ctx):
        pass


class Compressor(ABC):
    def __init__(self, average=True, tensors_size_are_same=True):
        self.average = average
        self.tensors_size_are_same = tensors_size_are_same

    @abstractmethod
    def compress(self, tensor, name):
        raise NotImplementedError

    @abstractmethod
    def decompress(self, tensors, ctx):
        raise NotImplementedError

    def aggregate(self, tensors):
        return sum(tensors)


class Communicator(ABC):
    def __init__(self, compressor, memory):
        self.compressor = compressor
        self.memory = memory

    @abstractmethod
    def send_receive(self, tensors, name, ctx):
        raise NotImplementedError

    def step(self, tensor, name):
        tensor = self.memory.compensate(tensor, name)
        tensors_compressed, ctx = self.compressor.compress(tensor, name)
        self.memory.update(tensor, name, self.compressor, tensors_compressed, ctx)
        return self.send_receive(tensors_compressed, name, ctx)

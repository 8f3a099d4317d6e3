"""Import-time stand-in: the reference does `from pybloomfilter import BloomFilter` at
module scope (pytorch/deepreduce.py:693) but the benchmarked configs never use bloom_cpu."""


class BloomFilter(object):
    def __init__(self, *a, **k):
        raise RuntimeError("pybloomfilter is not available offline; 'bloom_cpu' is not part of the benchmarked config")

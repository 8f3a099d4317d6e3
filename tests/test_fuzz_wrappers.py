"""Seeded fuzz over the GRACE-compatible wrappers: every (sparsifier, mode, value codec, index codec, policy,
small-tensor gate, gradient shape) combination must round-trip to a finite dense tensor — including the degenerate
selections the reference never guards (K = 0 from the threshold sparsifier on a zero gradient, K = 1, one-signed or
heavily tied values) — and the lossless codecs must reproduce the plain sparsifier bit for bit."""
import random
import warnings

import pytest
import torch

import deepreduce_b200 as dr

VALS = ['polyfit', 'qsgd', 'gzip', 'dexp', 'polyfit_cpu']
IDXS = ['bloom', 'rle', 'huffman', 'integer', 'bloom_cpu']
POLICIES = ['leftmost', 'random', 'p0', 'conflict_sets']
SPARSIFIERS = [('topk', {}), ('threshold', {'threshold': 0.5}), ('threshold', {'threshold': 100.0}), ('randomk', {})]
LOSSLESS_VALUE, LOSSLESS_INDEX = {'gzip'}, {'rle', 'huffman', 'integer'}


def _grad(kind, d, gen):
    return {'randn': lambda: torch.randn(d, generator=gen), 'pos': lambda: torch.rand(d, generator=gen) + 0.1,
            'neg': lambda: -torch.rand(d, generator=gen) - 0.1, 'zeros': lambda: torch.zeros(d),
            'ties': lambda: torch.randint(-2, 3, (d,), generator=gen).float()}[kind]()


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.timeout(600)
def test_fuzz_roundtrip(seed):
    rnd = random.Random(seed)
    gen = torch.Generator().manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(120):
            d = rnd.choice([1, 2, 7, 31, 999, 1000, 1001, 1500, 4097, 9001, 20000])
            comp, extra = rnd.choice(SPARSIFIERS)
            mode = rnd.choice(['value', 'index', 'both'])
            cfg = {'compressor': comp, 'memory': rnd.choice(['none', 'residual']), 'communicator': 'allgather',
                   'compress_ratio': rnd.choice([0.001, 0.01, 0.1, 0.5, 1.0]), 'deepreduce': mode,
                   'value': rnd.choice(VALS), 'index': rnd.choice(IDXS), 'policy': rnd.choice(POLICIES),
                   'min_numel': rnd.choice([0, 100, 1000]), **extra}
            g = _grad(rnd.choice(['randn', 'pos', 'neg', 'zeros', 'ties']), d, gen)
            grc = dr.deepreduce_from_params(cfg)
            for _step in range(2):                              # second step exercises the residual path
                out = grc.step(g.clone(), 'w')
                assert out.shape == g.shape and torch.isfinite(out).all(), cfg
            lossless = ((mode == 'value' and cfg['value'] in LOSSLESS_VALUE)
                        or (mode == 'index' and cfg['index'] in LOSSLESS_INDEX)
                        or (mode == 'both' and cfg['value'] in LOSSLESS_VALUE and cfg['index'] in LOSSLESS_INDEX))
            if lossless and comp != 'randomk':                  # randomk draws a fresh index set per call
                nomem = dict(cfg, memory='none')
                plain = dr.grace_from_params({k: v for k, v in nomem.items() if k != 'deepreduce'}).step(g.clone(), 'w')
                assert torch.equal(dr.deepreduce_from_params(nomem).step(g.clone(), 'w'), plain), cfg

"""GRACE shell + DeepReduce wrappers (CPU): wire sizes of SURVEY §3.6 / Appendix C,
residual invariants, small-tensor bypass, factory."""
import pytest
import torch

import deepreduce_b200 as dr
from deepreduce_b200.grace import tensor_bits

BASE = {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01}


def _grad(d=36864, seed=0):
    torch.manual_seed(seed)
    return torch.randn(d)


def test_factory_readme_usage():
    params = dict(BASE, deepreduce='index', index='bloom')
    grc = dr.deepreduce_from_params(params)
    assert isinstance(grc.compressor, dr.IndexCompressor)
    assert isinstance(grc.memory, dr.ResidualMemory)
    grc2 = dr.grace_from_params(params)
    grc2.compressor = dr.deepreduce_wrapper['index'](grc2.compressor, params)     # README.md:42-48 manual wiring
    g = _grad()
    assert torch.equal(grc.step(g.clone(), 'a'), grc2.step(g.clone(), 'a'))
    assert 'dense_tensor' not in params                                           # no side channel in user params
    with pytest.raises(ValueError):
        dr.deepreduce_from_params(dict(BASE, deepreduce='nope'))


def test_index_bloom_wire_and_residual():
    g = _grad()
    grc = dr.deepreduce_from_params(dict(BASE, deepreduce='index', index='bloom'))
    (vals, words), ctx = grc.compressor.compress(g, 'w')
    assert vals.numel() == 368 and vals.dtype == torch.float32          # f32[K]
    assert words.numel() * 4 in (664, 668)                              # u8[662] rounded to words
    dense = grc.compressor.decompress((vals, words), ctx)
    sel = dense.nonzero().flatten()
    assert sel.numel() == 368 and torch.equal(dense[sel], g[sel])       # exact values at S~
    true_topk = set(torch.topk(g.abs(), 368).indices.tolist())
    assert len(true_topk & set(sel.tolist())) >= 320                    # SURVEY: 334/368 kept with leftmost
    out = grc.step(g.clone(), 'w')
    assert torch.equal(grc.memory.residuals['w'][sel], torch.zeros(368))  # residual exactly 0 on S~
    assert torch.allclose(grc.memory.residuals['w'] + out, g)


def test_small_tensor_bypass():
    g = torch.randn(1000)
    grc = dr.deepreduce_from_params(dict(BASE, deepreduce='index', index='bloom'))
    (vals, idxs), ctx = grc.compressor.compress(g, 'b')
    assert idxs.numel() == 10 and vals.numel() == 10                    # raw pairs for <= 1000 elements (:68,115)
    assert idxs.dtype == torch.int32                                    # 32-bit keys on the wire (paper accounting)
    out = grc.compressor.decompress((vals, idxs), ctx)
    assert torch.equal(out.flatten()[idxs.long()], vals) and int((out != 0).sum()) == 10


def test_both_is_fp_aware_and_packed():
    g = _grad()
    grc = dr.deepreduce_from_params(dict(BASE, deepreduce='both'))
    tensors, ctx = grc.compressor.compress(g, 'w')
    coeffs, words, mapping = tensors
    assert mapping.dtype == torch.uint8 and mapping.numel() == 5 + (368 * 9 + 7) // 8    # ceil(log2 K) = 9 bits
    dense = grc.compressor.decompress(tensors, ctx)
    sel = dense.nonzero().flatten()
    rel = (dense[sel] - g[sel]).norm() / g[sel].norm()
    assert rel < 0.08           # fitted on the right coordinates (FP values near 0 sit at the sign change)
    assert tensor_bits(tensors) / 8 < 2200                               # vs 3806 B in the reference


@pytest.mark.parametrize("cfg", [
    dict(deepreduce='value', value='polyfit'), dict(deepreduce='value', value='qsgd'),
    dict(deepreduce='value', value='gzip'), dict(deepreduce='index', index='rle'),
    dict(deepreduce='index', index='huffman'), dict(deepreduce='index', index='integer'),
    dict(deepreduce='both', value='qsgd', policy='p0'), dict(deepreduce='both', value='qsgd', index='rle'),
    dict(deepreduce='both', value='gzip'), dict(deepreduce='both', value='gzip', index='rle'), dict(deepreduce='value', value='dexp'),
    dict(deepreduce='index', index='bloom', policy='p0'), dict(deepreduce='index', index='bloom', policy='conflict_sets'),
    dict(compressor='threshold', threshold=1.5, memory='none', deepreduce='index', index='bloom', policy='p0', fpr=0.01),
    dict(compressor='threshold', threshold=1.5, memory='none', deepreduce='both', index='bloom', policy='random', fpr=0.01, value='qsgd'),
])
def test_configs_run_and_preserve_mass(cfg):
    g = _grad(seed=2)
    grc = dr.deepreduce_from_params(dict(BASE, **cfg))
    out = grc.step(g.clone(), 'w')
    assert out.shape == g.shape and torch.isfinite(out).all()
    nz = out.nonzero().flatten()
    assert nz.numel() >= 300
    assert torch.nn.functional.cosine_similarity(out[nz], g[nz], dim=0) > 0.9


def test_wire_volumes_follow_the_papers_accounting():
    """32-bit keys, no mapping for order-preserving value codecs, coefficient rows sized by K (profiles/volume_table.md)."""
    from deepreduce_b200.grace import tensor_bits
    g = _grad()                                                            # d = 36 864, K = 368
    d, K = g.numel(), 368

    def rel(**cfg):
        tensors, _ = dr.deepreduce_from_params(dict(BASE, **cfg)).compressor.compress(g.clone(), 'w')
        return tensors, tensor_bits(list(tensors)) / (32.0 * d)

    (coef, idx), v = rel(deepreduce='value', value='polyfit')
    assert idx.dtype == torch.int32 and coef.numel() == 6 * 6 + 1          # 6 rows for K = 368, not 22
    assert v < 0.62 * (64.0 * K) / (32.0 * d)                              # paper: Fit-Poly ~40 % below Top-r
    (q, filt, mapping), v = rel(deepreduce='both', value='qsgd')
    assert mapping.numel() == 0                                            # QSGD keeps the order: no permutation shipped
    assert v < 0.5 * (64.0 * K) / (32.0 * d)
    (_, _, mapping), _ = rel(deepreduce='both', value='polyfit')
    assert mapping.numel() > 0                                             # the fit sorts: its permutation must travel


def test_lossless_modes_equal_plain_topk():
    g = _grad(seed=3)
    plain = dr.grace_from_params(dict(BASE, memory='none')).step(g.clone(), 'w')
    for cfg in (dict(deepreduce='value', value='gzip'), dict(deepreduce='index', index='rle'),
                dict(deepreduce='index', index='huffman'), dict(deepreduce='index', index='integer'),
                dict(deepreduce='both', value='gzip', index='rle')):
        out = dr.deepreduce_from_params(dict(BASE, memory='none', **cfg)).step(g.clone(), 'w')
        assert torch.equal(out, plain), cfg


def test_residual_memory_state_roundtrip():
    g = _grad()
    grc = dr.deepreduce_from_params(dict(BASE, deepreduce='index'))
    grc.step(g.clone(), 'w')
    st = grc.memory.state_dict()
    grc2 = dr.deepreduce_from_params(dict(BASE, deepreduce='index'))
    grc2.memory.load_state_dict(st)
    g2 = _grad(seed=9)
    assert torch.equal(grc.step(g2.clone(), 'w'), grc2.step(g2.clone(), 'w'))


def test_beta_gamma():
    m = dr.ResidualMemory(beta=0.5, gamma=2.0)
    m.residuals['x'] = torch.ones(4)
    assert torch.equal(m.compensate(torch.ones(4), 'x'), torch.full((4,), 2.5))


def test_sparsifiers():
    g = _grad()
    (v, i), ctx = dr.TopKCompressor(0.01).compress(g, 'w')
    assert v.numel() == 368 and torch.equal(g[i], v) and isinstance(ctx, torch.Size)
    (v, i), _ = dr.ThresholdCompressor(2.0).compress(g, 'w')
    assert torch.all(v.abs() > 2.0) and dr.ThresholdCompressor(2.0).tensors_size_are_same is False
    rk = dr.RandomKCompressor(0.01)
    (v1, i1), _ = rk.compress(g, 'w')
    rk2 = dr.RandomKCompressor(0.01)
    (v2, i2), _ = rk2.compress(g, 'w')
    assert torch.equal(i1, i2) and i1.numel() == 368         # same coordinates on every rank for (step, name)


def test_micro_benchmark_metrics(capsys):
    from deepreduce_b200.utils import METRICS
    METRICS.reset()
    grc = dr.deepreduce_from_params(dict(BASE, deepreduce='index', **{'micro-benchmark': True}))
    grc.step(_grad(), 'w')
    out = capsys.readouterr().out
    assert 'idx_compression time' in out and 'idx_relative_volume' in out
    s = METRICS.summary()
    assert s['mean_idx_relative_volume'] > 0 and s['calls']['idx_compression'] == 1


def test_file_loggers(tmp_path):
    from deepreduce_b200.utils import log_compressor, log_values
    g = _grad()
    idx = torch.topk(g.abs(), 368).indices
    r = log_compressor(str(tmp_path), 0, 5, 3, N=36864, K=368, true_indices=idx, selected_indices=idx,
                       bloom_bytes=662, positives=400, verbosity=2, values=g[idx])
    assert r == {"false_positives": 32, "policy_errors": 0}
    d = tmp_path / "0" / "step_5" / "3"
    assert (d / "fpr.txt").read_text().startswith("FalsePositives: 32")
    log_values(str(tmp_path), 0, 5, 3, g[:10], torch.arange(6.0))
    assert len((d / "values.csv").read_text().splitlines()) == 10


def test_config_validation():
    """params dict -> frozen, range-checked config (SURVEY §5 'Config / flag system')."""
    import warnings

    import pytest

    from deepreduce_b200 import deepreduce_from_params
    from deepreduce_b200.config import ConfigError, DeepReduceConfig
    base = {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
            'deepreduce': 'index', 'index': 'bloom'}
    cfg = DeepReduceConfig.from_params(base)
    assert cfg.compress_ratio == 0.01 and cfg.index == 'bloom' and cfg.policy == 'leftmost'
    with pytest.raises(Exception):
        cfg.compress_ratio = 0.5                                  # frozen
    assert DeepReduceConfig.from_params(cfg.to_params()) == cfg   # round trip
    before = dict(base)
    deepreduce_from_params(base)
    assert base == before                                         # never written to
    for bad in ({'compress_ratio': 0.0}, {'compress_ratio': 1.5}, {'fpr': 1.0}, {'policy': 'rightmost'},
                {'deepreduce': 'values'}, {'index': 'bloomm'}, {'communicator': 'allreduce'}, {'poly_degree': 9},
                {'compressor': 'top-k'}, {'memory': 'momentum'}):
        with pytest.raises(ConfigError):
            DeepReduceConfig.from_params({**base, **bad})
    with pytest.raises(NotImplementedError):
        DeepReduceConfig.from_params({**base, 'compressor': 'SKCompressGPU'})
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        DeepReduceConfig.from_params({**base, 'compres_ratio': 0.1})
        assert any('compres_ratio' in str(x.message) for x in w)
    with pytest.raises(ConfigError):
        DeepReduceConfig.from_params({**base, 'compres_ratio': 0.1}, strict=True)
    # randomk with a shared seed may be all-reduced (GRACE); a value/index codec needs a sparsifier
    DeepReduceConfig.from_params({'compressor': 'randomk', 'communicator': 'allreduce', 'memory': 'none'})
    with pytest.raises(ConfigError):
        DeepReduceConfig.from_params({'compressor': 'none', 'communicator': 'allgather', 'deepreduce': 'index'})


def test_fused_path_gating():
    """Which params dicts DeepReduceDDP routes to the fused engine (README 'What runs where')."""
    from deepreduce_b200.parallel.ddp import _fused_supported
    base = {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01}
    yes = [base,
           {**base, 'deepreduce': 'index', 'index': 'bloom'},
           {**base, 'deepreduce': 'index', 'index': 'bloom', 'policy': 'p0'},
           {**base, 'deepreduce': 'index', 'index': 'rle'},
           {**base, 'deepreduce': 'both', 'index': 'bloom', 'value': 'polyfit'},
           {**base, 'deepreduce': 'both', 'index': 'bloom', 'value': 'qsgd', 'quantum_num': 127, 'bucket_size': 512},
           # the reference's NCF recipes (run_deepreduce.sh:66-74): threshold sparsifier, value-only mode, int16 QSGD
           {'compressor': 'threshold', 'memory': 'none', 'communicator': 'allgather', 'threshold': 0.0},
           {'compressor': 'threshold', 'memory': 'none', 'communicator': 'allgather', 'threshold': 0.0,
            'deepreduce': 'index', 'index': 'bloom', 'policy': 'p0', 'fpr': 0.01},
           {**base, 'deepreduce': 'value', 'value': 'polyfit'}, {**base, 'deepreduce': 'value', 'value': 'qsgd', 'quantum_num': 32},
           {**base, 'deepreduce': 'both', 'index': 'bloom', 'value': 'qsgd', 'quantum_num': 255},
           # ... and its last two (run_deepreduce.sh:73-74): policy 'random' (P1) with QSGD values
           {**base, 'deepreduce': 'index', 'index': 'bloom', 'policy': 'random'},
           {'compressor': 'threshold', 'memory': 'none', 'communicator': 'allgather', 'threshold': 0.0,
            'deepreduce': 'both', 'index': 'bloom', 'policy': 'random', 'fpr': 0.01, 'value': 'qsgd'}]
    no = [{**base, 'compressor': 'randomk'},
          {**base, 'deepreduce': 'index', 'index': 'bloom', 'policy': 'conflict_sets'},
          {**base, 'deepreduce': 'index', 'index': 'huffman'}, {**base, 'deepreduce': 'index', 'index': 'integer'},
          {**base, 'deepreduce': 'both', 'index': 'rle', 'value': 'polyfit'},
          {**base, 'deepreduce': 'both', 'index': 'bloom', 'value': 'qsgd', 'bucket_size': 256},
          {**base, 'deepreduce': 'both', 'index': 'bloom', 'value': 'gzip'},
          {'compressor': 'none', 'memory': 'none', 'communicator': 'allreduce'}]
    for p in yes:
        assert _fused_supported(p), p
    for p in no:
        assert not _fused_supported(p), p

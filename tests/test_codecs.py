"""Codec round-trips and error bounds (CPU) — SURVEY §4 "Unit" + §3.6 table."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import deepreduce_b200 as dr
from deepreduce_b200.codecs import bitpack, polyfit, rle, integer, qsgd, dexp, lossless
from deepreduce_b200.codecs.polyfit import get_segments, MAX_SEGMENTS


def _sparse(d=36864, ratio=0.01, seed=0):
    torch.manual_seed(seed)
    g = torch.randn(d)
    K = max(1, int(d * ratio))
    idx = torch.topk(g.abs(), K, sorted=False).indices
    return g, g[idx], idx, torch.Size([d])


def test_get_segments_matches_reference_table():
    # reference pytorch/deepreduce.py:362-377
    seg = get_segments(368, 180)
    assert sum(seg) == 368 and len(seg) <= MAX_SEGMENTS
    assert seg == [36, 144, 151, 37]
    seg = get_segments(23592, 11800)
    assert sum(seg) == 23592
    # fine segments sit at both steep ends, symmetric structure
    assert seg[0] == int(11800 / 300) and seg[-1] == int((23592 - 11800) / 300)
    assert get_segments(10, 0) == [0, 10]


def test_gram_basis_orthogonal():
    for n in (7, 40, 1000):
        P = polyfit.gram_basis(n, 5)
        G = P.T @ P
        off = G - torch.diag(torch.diag(G))
        assert off.abs().max() < 1e-9 * G.diag().max()


def test_polyfit_equals_monomial_least_squares():
    torch.manual_seed(0)
    y = torch.sort(torch.randn(500).abs(), descending=True).values.double()
    c = polyfit.fit_segment_oracle(y, 5)
    fit = polyfit.gram_basis(500, 5) @ c
    x = np.arange(500) / 499.0
    ref = np.polynomial.polynomial.polyval(x, np.polynomial.polynomial.polyfit(x, y.numpy(), 5))
    assert np.allclose(fit.numpy(), ref, atol=1e-8)


def test_polyfit_roundtrip_error():
    g, vals, idx, shape = _sparse()
    w, idx2, _ = polyfit.PolyFit.compress((vals, idx, shape), {})
    # K = 368: only the 1/5 and 1/10 ratios can yield segments > 30 -> at most 2*2+2 rows, a function of K alone
    assert polyfit.seg_rows(idx.numel()) == 6 and w.numel() == 6 * polyfit.seg_rows(idx.numel()) + 1
    assert w.dtype == torch.float32 and polyfit.seg_rows(10 ** 7) == MAX_SEGMENTS
    for p_ in (0, 1, 100, 184, 367, 368):          # never more segments than rows, whatever the sign split
        assert len(polyfit.get_segments(368, p_)) <= 6
    v, idx3, _ = polyfit.PolyFit.decompress((w, idx2, shape), {})
    dense = torch.zeros(shape.numel()); dense[idx3] = v
    ref = torch.zeros(shape.numel()); ref[idx] = vals
    assert (dense - ref).norm() / ref.norm() < 0.02       # 0.0067 for the reference's fit
    assert set(idx3.tolist()) == set(idx.tolist())


def test_polyfit_degenerate_segments():
    shape = torch.Size([5000])
    vals = torch.rand(50) + 0.1                            # all positive -> empty negative segments
    idx = torch.arange(50)
    w, i2, _ = polyfit.PolyFit.compress((vals, idx, shape), {})
    v, _, _ = polyfit.PolyFit.decompress((w, i2, shape), {})
    assert torch.isfinite(v).all() and v.numel() == 50
    w, i2, _ = polyfit.PolyFit.compress((vals[:1], idx[:1], shape), {})
    v, _, _ = polyfit.PolyFit.decompress((w, i2, shape), {})
    assert torch.allclose(v, vals[:1], atol=1e-6)


def test_polyfit_cpu_roundtrip():
    g, vals, idx, shape = _sparse()
    w, i2, _ = polyfit.PolyFitCPU.compress((vals, idx, shape), {})
    v, i3, _ = polyfit.PolyFitCPU.decompress((w, i2, shape), {})
    dense = torch.zeros(shape.numel()); dense[i3] = v
    ref = torch.zeros(shape.numel()); ref[idx] = vals
    assert (dense - ref).norm() / ref.norm() < 0.05


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 2**20 - 1), min_size=0, max_size=300), st.integers(1, 3))
def test_bitpack_roundtrip(vals, extra):
    t = torch.tensor(vals, dtype=torch.int64)
    enc = bitpack.pack(t)
    assert enc.dtype == torch.uint8
    assert torch.equal(bitpack.unpack(enc), t)
    if len(vals):
        bits = max(1, int(max(vals)).bit_length())
        assert enc.numel() == 5 + (len(vals) * bits + 7) // 8


def test_pack21_roundtrip():
    t = torch.randint(0, 2**21, (1001,))
    assert torch.equal(bitpack.unpack_(bitpack.pack_(t)).long(), t)


@settings(max_examples=40, deadline=None)
@given(st.sets(st.integers(0, 4999), min_size=0, max_size=400))
def test_rle_property(s):
    d = 5000
    idx = torch.tensor(sorted(s), dtype=torch.int64)
    runs = rle.runs_from_sorted_oracle(idx, d)
    assert int(runs.sum()) == d or (idx.numel() and int(runs.sum()) == int(idx[-1]) + 1)
    assert torch.equal(rle.indices_from_runs_oracle(runs), idx)


def test_rle_codec_lossless():
    g, vals, idx, shape = _sparse()
    v, enc, _ = rle.RunLength.compress((vals, idx, shape), {})
    v2, idx2, _ = rle.RunLength.decompress((v, enc, shape), {})
    dense = torch.zeros(shape.numel()); dense[idx2] = v2
    ref = torch.zeros(shape.numel()); ref[idx] = vals
    assert torch.equal(dense, ref)


@pytest.mark.parametrize("code", ["copy", "vbyte", "bp32", "bp128", "simple8b", "pfor128", "fastpfor128"])
@pytest.mark.parametrize("delta", [True, False])
def test_integer_codec_lossless(code, delta):
    g, vals, idx, shape = _sparse(d=200000, ratio=0.02, seed=3)
    v, enc, _ = integer.IntegerIndex.compress((vals, idx, shape), {"code": code, "delta": delta})
    v2, idx2, _ = integer.IntegerIndex.decompress((v, enc, shape), {})
    dense = torch.zeros(shape.numel()); dense[idx2] = v2
    ref = torch.zeros(shape.numel()); ref[idx] = vals
    assert torch.equal(dense, ref)
    if delta and code in ("bp128", "pfor128", "simple8b", "vbyte"):
        assert enc.numel() * 4 < idx.numel() * 4 * 0.5      # gaps ~50 -> < 16 bits per index


def test_integer_numpy_fallback_matches_native():
    from deepreduce_b200 import ops
    if not ops.has_cpu_native():
        pytest.skip("no native")
    a = (np.random.RandomState(0).randint(0, 1000, size=1000)).astype(np.uint32)
    for code in ("vbyte", "bp32", "bp128"):
        nat = integer.int_encode(a, code)
        cid = integer.CODECS[code]
        py = {1: integer._np_vbyte_encode, 2: lambda x: integer._np_bp_encode(x, 32), 3: lambda x: integer._np_bp_encode(x, 128)}[cid](a)
        assert np.array_equal(nat, py), code


def test_qsgd_roundtrip_and_format():
    g, vals, idx, shape = _sparse()
    w, _, _ = qsgd.QSGD.compress((vals, idx, shape), {})
    K = vals.numel()
    assert w.dtype == torch.int8 and w.numel() == K + 4 * ((K + 511) // 512)
    v, _, _ = qsgd.QSGD.decompress((w, idx, shape), {})
    assert (v - vals).norm() / vals.norm() < 0.12
    w16, _, _ = qsgd.QSGD.compress((vals, idx, shape), {"quantum_num": 1024})
    assert w16.dtype == torch.int16
    v16, _, _ = qsgd.QSGD.decompress((w16, idx, shape), {"quantum_num": 1024})
    assert (v16 - vals).norm() / vals.norm() < 0.02
    # stochastic rounding is unbiased-ish and deterministic for a fixed seed
    w2, _, _ = qsgd.QSGD.compress((vals, idx, shape), {})
    assert torch.equal(w, w2)


def test_gzip_huffman_lossless():
    g, vals, idx, shape = _sparse()
    w, _, _ = lossless.Gzip.compress((vals, idx, shape), {})
    v, _, _ = lossless.Gzip.decompress((w, idx, shape), {})
    assert torch.equal(v, vals)
    _, enc, _ = lossless.Huffman.compress((vals, idx, shape), {})
    _, idx2, _ = lossless.Huffman.decompress((vals, enc, shape), {})
    assert torch.equal(idx2, idx)


def test_dexp_fit_quality():
    torch.manual_seed(0)
    K = 2000
    x = torch.arange(1, K + 1).double() / K
    y = 0.3 * torch.exp(1.7 * x) + 0.05 * torch.exp(-3.0 * x)
    coef = torch.stack(dexp.double_exponential_fit(y))
    fit = dexp.double_exponential_eval(coef.float(), K)
    assert (fit.double() - y).abs().max() / y.abs().max() < 1e-2
    g, vals, idx, shape = _sparse()
    c, sidx, _ = dexp.DoubleExp.compress((vals, idx, shape), {})
    assert c.numel() == 4 and sidx.dtype == torch.int32
    v, i2, _ = dexp.DoubleExp.decompress((c, sidx, shape), {})
    dense = torch.zeros(shape.numel()); dense[i2] = v
    ref = torch.zeros(shape.numel()); ref[idx] = vals
    assert (dense - ref).norm() / ref.norm() < 0.05


def test_registry_and_custom_codec():
    assert set(["bloom", "polyfit", "bloom_cpu", "polyfit_cpu", "gzip", "huffman", "rle", "qsgd"]) <= set(dr.compressor)

    @dr.register("identity_test")
    class Ident(dr.SparseCompressor):
        @staticmethod
        def compress(st, params):
            return st

        @staticmethod
        def decompress(st, params):
            return st

    grc = dr.deepreduce_from_params({'compressor': 'topk', 'memory': 'none', 'communicator': 'allgather',
                                     'compress_ratio': 0.01, 'deepreduce': 'value', 'value': 'identity_test'})
    t = torch.randn(5000)
    out = grc.step(t, "x")
    assert int((out != 0).sum()) == 50


def test_native_bloomfilter_object_matches_oracle():
    """The C++ filter object (surface of the reference's bloom::OrdinaryBloomFilter) vs the torch oracle."""
    import numpy as np
    import pytest
    import torch
    from deepreduce_b200 import ops, spec
    from deepreduce_b200.codecs.bloom import bloom_insert_oracle, bloom_query_oracle
    if not ops.has_cpu_native():
        pytest.skip("native host extension not built")
    k, m_bits, d = 7, 2048, 20000
    gen = torch.Generator().manual_seed(1)
    idx = torch.sort(torch.randperm(d, generator=gen)[:150]).values
    f = ops.cpu.BloomFilter(k, m_bits, spec.DEFAULT_SEED)
    f.insert(idx.numpy())
    ref = bloom_insert_oracle(idx, k, m_bits, spec.DEFAULT_SEED)
    assert np.array_equal(f.words(), ref.numpy().view(np.uint32))
    assert f.num_bytes() == m_bits // 8 and f.num_hashes() == k and all(f.hash(5, j) < m_bits for j in range(k))
    assert f.query(idx.numpy()).all()                                      # no false negatives
    positives = bloom_query_oracle(ref, d, k, m_bits, spec.DEFAULT_SEED)
    assert f.compute_false_positives(d, idx.numpy()) == positives.numel() - idx.numel()
    g = ops.cpu.BloomFilter.from_words(f.words(), k, m_bits, spec.DEFAULT_SEED)    # receiver side: from raw words
    assert np.array_equal(g.query(np.arange(d)), np.isin(np.arange(d), positives.numpy()))


def test_native_huffman_matches_numpy_reference():
    import numpy as np
    import pytest
    from deepreduce_b200 import ops
    from deepreduce_b200.codecs import lossless as L
    if L._native() is None:
        pytest.skip("native host extension not built")
    rng = np.random.default_rng(3)
    for d in (1000, 36864, 2359296):
        lengths, codes = L._model_for(d)
        for n in (0, 1, 7, 5000):
            data = np.sort(rng.integers(0, d, size=n)).astype(np.int32).view(np.uint8)
            a = L.huffman_encode(data, lengths, codes, native=True)
            b = L.huffman_encode(data, lengths, codes, native=False)
            assert np.array_equal(a, b)
            assert np.array_equal(L.huffman_decode(a, lengths, codes, native=True), data)
            assert np.array_equal(L.huffman_decode(a, lengths, codes, native=False), data)
    # a skewed model (long codes) and corrupted input
    freq = np.ones(256, dtype=np.int64); freq[0] = 10 ** 9; freq[1] = 10 ** 6
    lengths = L._code_lengths(freq); codes = L._canonical(lengths)
    data = rng.integers(0, 256, size=2000).astype(np.uint8)
    enc = L.huffman_encode(data, lengths, codes)
    assert np.array_equal(L.huffman_decode(enc, lengths, codes), data)
    with pytest.raises(RuntimeError):
        ops._cpu_mod.huffman_decode(enc[:20], lengths, codes.astype(np.uint64))      # truncated stream

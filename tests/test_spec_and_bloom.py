"""Spec / bloom / policies (CPU).  Mirrors SURVEY §4: sizing vs reference :495-500,
policies deterministic sender==receiver, FP-aware residual exactly 0 on S~."""
import math

import numpy as np
import pytest
import torch

from deepreduce_b200 import spec
from deepreduce_b200.codecs import bloom as B
from deepreduce_b200.codecs.bloom_cpu import bloom_compress_blob, bloom_decompress_blob
from deepreduce_b200 import ops


def test_hash_scalar_matches_tensor():
    x = torch.tensor([0, 1, 2, 12345, 2**31 - 1, 36863])
    pos = spec.bloom_positions(x, 7, 6624, spec.DEFAULT_SEED)
    for i, xi in enumerate(x.tolist()):
        assert pos[i].tolist() == spec.bloom_positions_int(xi, 7, 6624, spec.DEFAULT_SEED)
    assert int(pos.min()) >= 0 and int(pos.max()) < 6624


def test_get_BFconfig_reference_formula():
    # reference pytorch/deepreduce.py:495-500 with K=368, d=36864 -> fpr=1e-3 -> 10 hashes, 5296 bits (662 B)
    k, bits = spec.get_BFconfig(368, 0.1 * 368 / 36864)
    assert k == 10 and math.ceil(bits / 8) == 662
    nh, m_bits, n_words = spec.bloom_layout(368, 36864)
    assert nh == 10 and m_bits % 32 == 0 and m_bits >= bits and m_bits - bits < 32 and n_words * 32 == m_bits


def test_tf_bloom_configuration():
    m, h = spec.bloom_configuration(368, 1e-3)
    assert m > 0 and h >= 1
    assert m == int((368 * abs(math.log(1e-3))) / (math.log(2) ** 2) / 8) + (1 if int((368 * abs(math.log(1e-3))) / (math.log(2) ** 2) / 8) % 8 else 0)


def test_no_false_negatives_and_fpr():
    torch.manual_seed(0)
    d, K = 36864, 368
    idx = torch.randperm(d)[:K].sort().values
    k, m_bits, _ = spec.bloom_layout(K, d)
    words = B.bloom_insert_oracle(idx, k, m_bits)
    pos = B.bloom_query_oracle(words, d, k, m_bits)
    assert set(idx.tolist()) <= set(pos.tolist())
    fp = pos.numel() - K
    assert fp < 0.004 * d           # design fpr 1e-3 -> ~36 expected (SURVEY §2.7 cheat-sheet)


@pytest.mark.parametrize("policy", ["leftmost", "random", "p0", "conflict_sets"])
def test_policy_sender_equals_receiver(policy):
    torch.manual_seed(1)
    d, K = 20000, 200
    g = torch.randn(d)
    vals, idx = torch.topk(g.abs(), K)
    params = {"policy": policy, "dense_tensor": g, "policy_seed": 7}
    v, words, shape = B.Bloom.compress((g[idx], idx, torch.Size([d])), params)
    v2, idx2, _ = B.Bloom.decompress((v, words, shape), {"policy": policy, "policy_seed": 7})
    assert torch.equal(g[idx2], v2)                         # FP-aware: values are the true dense values at S~
    assert torch.all(idx2[1:] > idx2[:-1])                  # ascending
    if policy == "p0":
        assert set(idx.tolist()) <= set(idx2.tolist())      # lossless w.r.t. the sparsifier
    else:
        assert idx2.numel() == K
    if policy == "conflict_sets":
        # P2 should recover (nearly) all true positives: singleton conflict sets are certain
        assert len(set(idx.tolist()) & set(idx2.tolist())) >= K - 2


def test_conflict_sets_native_matches_python(monkeypatch):
    if not ops.has_cpu_native():
        pytest.skip("native cpu ext not built")
    torch.manual_seed(3)
    d, K = 5000, 60
    idx = torch.randperm(d)[:K].sort().values
    k, m_bits, _ = spec.bloom_layout(K, d, fpr=0.02)
    words = B.bloom_insert_oracle(idx, k, m_bits)
    pos = B.bloom_query_oracle(words, d, k, m_bits)
    nat = B.conflict_sets_oracle(pos, K, k, m_bits, spec.DEFAULT_SEED, 99)
    monkeypatch.setattr(ops, "has_cpu_native", lambda: False)
    py = B.conflict_sets_oracle(pos, K, k, m_bits, spec.DEFAULT_SEED, 99)
    assert torch.equal(nat, py)


def test_native_bloom_matches_oracle():
    if not ops.has_cpu_native():
        pytest.skip("native cpu ext not built")
    torch.manual_seed(4)
    d, K = 50000, 500
    idx = torch.randperm(d)[:K].sort().values
    k, m_bits, _ = spec.bloom_layout(K, d)
    w_o = B.bloom_insert_oracle(idx, k, m_bits)
    w_n = torch.from_numpy(ops.cpu.bloom_insert(idx.numpy(), k, m_bits, spec.DEFAULT_SEED).view(np.int32))
    assert torch.equal(w_o, w_n)
    for pol, pid in (("leftmost", 0), ("random", 1), ("p0", 2)):
        sel_o = B.apply_policy_oracle(B.bloom_query_oracle(w_o, d, k, m_bits), K, pol, 5, k, m_bits)
        sel_n = torch.from_numpy(ops.cpu.bloom_select(w_n.numpy().view(np.uint32), d, K, k, m_bits, spec.DEFAULT_SEED, pid, 5))
        assert torch.equal(sel_o, sel_n), pol


@pytest.mark.parametrize("policy", ["conflict_sets", "leftmostK", "randomK", "policy_zero"])
def test_tf_blob_roundtrip(policy):
    torch.manual_seed(5)
    N, K = 9408, 94
    dense = torch.randn(N)
    idx = torch.topk(dense.abs(), K).indices.sort().values
    blob = bloom_compress_blob(dense[idx], idx, dense, step=3, false_positives_aware=True, policy=policy, fpr=1e-3)
    assert blob.dtype == torch.int8
    out = bloom_decompress_blob(blob, N, step=3, policy=policy)
    nz = out.nonzero().flatten()
    assert torch.equal(out[nz], dense[nz])
    assert nz.numel() >= K - 1

"""TF-side API parity (torch-hosted) and the federated simulator (CPU)."""
import torch

from deepreduce_b200 import tf_compat as T
from deepreduce_b200.parallel.federated import FederatedAveraging


def test_tf_bloom_compressor_memory_cycle():
    torch.manual_seed(0)
    g = torch.randn(64, 3, 7, 7)
    p = dict(compress_ratio=0.01, bloom_fpr=1e-3, bloom_on='topk', bloom_policy='leftmostK',
             bloom_false_positives_aware=True, use_memory=True, beta=1.0, gamma=1.0, horovod_size=2, average=True)
    T.Compressor.residuals.clear()
    c = T.BloomFilterCompressor.memory_compensate(g, p, 'w')
    blob, ctx = T.BloomFilterCompressor.compress(c, p)
    assert blob.dtype == torch.int8 and p['tensors_size_are_same'] is False and p['m'] > 0 and p['k'] >= 1
    T.BloomFilterCompressor.memory_update(g, c, blob, ctx, p, 'w')
    d = T.BloomFilterCompressor.decompress(blob, ctx, p)
    assert torch.allclose(T.BloomFilterCompressor.residuals['w'] + d, g)
    agg = T.Compressor.aggregate([d, d], p)
    assert torch.allclose(agg, d)


def test_tf_helpers_and_tables():
    H = T.Values_Approximation_Helper
    assert H.is_convolutional("resnet50", 9408) and not H.is_convolutional("resnet50", 1000)
    assert H.get_num_of_segments("vgg16", 2359296) == 5 and H.get_breaks("resnet20_v2", 432) == [0, 353, 432]
    X = H.GetInputMatrix_Polynomial(3, torch.arange(5.0))
    assert X.shape == (5, 3) and torch.equal(X[:, 2], torch.arange(5.0).double() ** 2)
    th = H.LeastSquares(X, (1 + 2 * torch.arange(5.0)))
    assert torch.allclose(th.flatten(), torch.tensor([1.0, 2.0, 0.0], dtype=torch.float64), atol=1e-8)
    y = torch.sort(torch.randn(4000).abs()).values
    pts, sizes = H.find_breaks(y, 4)
    assert pts[0] == 0 and pts[-1] == 4000 and sum(sizes) == 4000


def test_tf_value_compressors():
    torch.manual_seed(1)
    g = torch.randn(36864)
    ref = torch.zeros_like(g); i = torch.topk(g.abs(), 368).indices; ref[i] = g[i]
    p = dict(compress_ratio=0.01)
    c, ctx = T.DoubleExpCompressor.compress(g, p)
    assert c[1].numel() == 4                                     # (a, b, p, q)
    assert (T.DoubleExpCompressor.decompress(c, ctx, p) - ref).norm() / ref.norm() < 0.05
    p = dict(compress_ratio=0.01, approximation_mode='topk', polynomial_degree=5, num_of_segments=3)
    c, ctx = T.PolySegCompressor.compress(g, p)
    assert c.dtype == torch.float64 and c.numel() == 3 + 15 + 368          # sizes | coefs | signed indices
    assert (T.PolySegCompressor.decompress(c, ctx, p) - ref).norm() / ref.norm() < 0.02


def test_federated_round_reduces_loss_and_volume():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(20, 64), torch.nn.ReLU(), torch.nn.Linear(64, 4))
    params = {'compressor': 'topk', 'memory': 'residual', 'compress_ratio': 0.1, 'deepreduce': 'index',
              'index': 'bloom', 'policy': 'p0', 'min_numel': 100}
    fed = FederatedAveraging(model, params, n_clients=3, local_steps=2, lr=0.1)
    w = torch.randn(20, 4)
    def batches():
        out = []
        for c in range(3):
            bs = []
            for _ in range(2):
                x = torch.randn(64, 20); bs.append((x, (x @ w).argmax(1)))
            out.append(bs)
        return out
    losses = [fed.round(batches(), torch.nn.functional.cross_entropy) for _ in range(12)]
    assert losses[-1] < losses[0]
    v = fed.volumes()
    tm = fed.timings(rounds=12)
    assert all(tm[k] > 0 for k in ('client_encode_s', 'client_decode_s', 'server_encode_s', 'server_decode_s'))
    assert 0 < v["c2s_relative_volume"] < 0.35 and 0 < v["s2c_relative_volume"] < 0.35


def test_tf_helper_bases_sparsifiers_and_op_sizing():
    """Remaining TF-side helpers (tensorflow/deepreduce.py:146-156,273-298; bloom_filter_compression.cc:85-99)."""
    import math

    import torch

    from deepreduce_b200 import spec
    from deepreduce_b200 import tf_compat as tfc
    from deepreduce_b200.codecs.bloom_cpu import tf_bloom_sizes
    H = tfc.Values_Approximation_Helper
    X = torch.arange(1, 6, dtype=torch.float64)
    assert torch.allclose(H.polynomial_basis(X, 3), X ** 3)
    assert torch.allclose(H.exp_basis(X, 2.0, -0.5), 2.0 * torch.exp(-0.5 * X))
    assert torch.allclose(H.logit_basis(X, 1.5, 5), 1.5 * torch.log(X / (6 - X)))
    B = tfc.BloomFilterCompressor
    g = torch.tensor([0.1, -5.0, 0.3, 4.0, -0.2, 0.0, 2.5, -0.05])
    assert B.topk_indices(g, 3).tolist() == [1, 3, 6]                       # ascending positions of the top-3 |g|
    assert B.threshold_indices(g, {"threshold_val": 1.0}).tolist() == [1, 3, 6]
    assert B.threshold_indices(g, {"threshold_val": 99.0}).tolist() == [1]  # clamped to max |g|
    B.global_step = 0
    a = B.randomk_indices("conv1", 1000, 10)
    B.global_step = 0
    b = B.randomk_indices("conv1", 1000, 10)
    c = B.randomk_indices("conv1", 1000, 10)                                # next step: a different draw
    assert a.tolist() == b.tolist() and a.tolist() != c.tolist()
    assert a.numel() == 10 and a.unique().numel() == 10 and int(a.max()) < 1000 and (a[1:] > a[:-1]).all()
    # Python-side sizing uses true division for h, the C++ op integer division
    for K, fpr in ((368, 0.001), (23592, 0.01), (10, 0.5)):
        m, h = spec.bloom_configuration(K, fpr)
        m2, h2 = tf_bloom_sizes(K, fpr)
        assert m == m2 and m >= 1
        assert h == math.ceil((m * 8 / K) * math.log(2)) and h2 == max(1, math.ceil(((m * 8) // K) * math.log(2)))
        assert h2 <= h

"""`compat/` serves the reference's import paths (reference README.md:30-48, pytorch/deepreduce.py:7-8) from this
framework; the usage snippet of the upstream README must run unchanged."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import torch
from grace_dl.dist.helper import grace_from_params
from grace_dl.dist import Compressor
from deepreduce import ValueCompressor, IndexCompressor, DeepReduce, compressor, PolyFit, Bloom, RunLength, QSGD
from deepreduce import GetInputMatrix_Polynomial, LeastSquares, RestoreValues, get_segments, get_BFconfig

params = {'compressor': 'topk', 'memory': 'residual', 'communicator': 'allgather', 'compress_ratio': 0.01,
          'deepreduce': 'index', 'index': 'bloom'}
grc = grace_from_params(params)
deepreduce_wrapper = {'value': ValueCompressor, 'index': IndexCompressor, 'both': DeepReduce}
grc.compressor = deepreduce_wrapper[params['deepreduce']](grc.compressor, params)
assert isinstance(grc.compressor, Compressor)
g = torch.randn(36864)
out = grc.step(g.clone(), 'conv')
assert out.shape == g.shape and 300 < int((out != 0).sum()) <= 368
assert set(['bloom', 'polyfit', 'bloom_cpu', 'polyfit_cpu', 'gzip', 'huffman', 'rle', 'qsgd']) <= set(compressor)

# monomial helpers (reference :308-347) span the same fit as the codec's Gram basis
y = torch.sort(torch.randn(400).abs(), descending=True).values
X = GetInputMatrix_Polynomial(400, 5, 'cpu')
a = LeastSquares(X, y)
fit = RestoreValues(400, a)
from deepreduce_b200.codecs.polyfit import fit_segment_oracle, gram_basis
ref = gram_basis(400, 5) @ fit_segment_oracle(y, 5)
assert torch.allclose(fit, ref, atol=1e-6), float((fit - ref).abs().max())
print("COMPAT_OK")
'''


def test_reference_readme_snippet_runs_against_compat_layer():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-c", SNIPPET], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "COMPAT_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]

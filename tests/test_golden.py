"""The oracle against the committed golden fixtures (CPU); the CUDA path against the same fixtures (GPU)."""
import glob
import os

import numpy as np
import pytest

from oracle import bruteforce as bf

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "cosine_topk_*.npz")))


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    k = int(z["k"])
    s, i = bf.cosine_topk_f64(z["queries"], z["corpus"], k)
    assert (i == z["index"]).all()
    assert np.abs(s - z["score"])[np.isfinite(z["score"])].max() < 1e-14
    chunks = [(lo, z["corpus"][lo:lo + 512]) for lo in range(0, len(z["corpus"]), 512)]
    s2, i2 = bf.cosine_topk_fast(z["queries"], chunks, k)
    assert (i2 == z["index"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=os.path.basename)
def test_engine_reproduces_golden(path):
    import torch
    from qsa_b200.engine import VectorIndex
    z = np.load(path)
    k = int(z["k"])
    c, q = z["corpus"], z["queries"]
    ix = VectorIndex(dim=c.shape[1], capacity=len(c) + 7, max_batch=256, max_k=max(k, 10))
    ix.append_bf16_bits(c)
    for cg in (1, 2):
        ix.set_option("cta_group", cg)
        s, i = ix.search(torch.from_numpy(q.view(np.int16)).view(torch.bfloat16).cuda(), k)
        torch.cuda.synchronize()
        assert (i.cpu().numpy() == z["index"]).all(), f"cta_group {cg}"
        fin = np.isfinite(z["score"])
        assert np.abs(s.cpu().numpy().astype(np.float64) - z["score"])[fin].max() < 1e-6
    ix.close()

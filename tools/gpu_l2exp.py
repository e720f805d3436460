"""L2-reuse experiment: one scan launch per setting, run under `ncu --metrics dram__bytes_read.sum,...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import fill_corpus
from qsa_b200.engine import VectorIndex

rows, dim, B = int(sys.argv[1]), 1536, int(sys.argv[2])
ix = VectorIndex(dim=dim, capacity=rows, max_batch=4096, max_k=10)
fill_corpus(ix, rows, dim, 1234)
q = torch.randn((B, dim), device="cuda").to(torch.bfloat16)
settings = [(2, 0, 0), (2, 1, 0), (2, 2, 0), (2, 4, 0), (2, 0, 1), (2, 2, 1), (1, 0, 0), (1, 2, 0), (1, 0, 1), (1, 2, 1)]
for i, (cg, d, um) in enumerate(settings):
    ix.set_option("cta_group", cg); ix.set_option("max_drift", d); ix.set_option("unit_map", um)
    ix.search(q, 10); torch.cuda.synchronize()
    print(f"launch {i}: cg={cg} drift={d} unit_map={um} algorithmic_GB={rows*dim*2/1e9:.2f}", flush=True)

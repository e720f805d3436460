"""L2-reuse experiment: one scan launch per setting, run under `ncu --metrics dram__bytes_read.sum,...`.
usage: gpu_l2exp.py ROWS BATCH "cg,drift,gain,max;cg,drift,gain,max;..." """
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import fill_corpus
from qsa_b200.engine import VectorIndex

rows, dim, B = int(sys.argv[1]), 1536, int(sys.argv[2])
settings = [tuple(int(x) for x in s.split(",")) for s in sys.argv[3].split(";")]
ix = VectorIndex(dim=dim, capacity=rows, max_batch=4096, max_k=10)
fill_corpus(ix, rows, dim, 1234)
q = torch.randn((B, dim), device="cuda").to(torch.bfloat16)
for i, (cg, d, g, m) in enumerate(settings):
    ix.set_option("cta_group", cg); ix.set_option("max_drift", d); ix.set_option("pace_gain", g); ix.set_option("pace_max", m)
    ix.search(q, 10); torch.cuda.synchronize()
    print(f"launch {i}: cg={cg} drift={d} gain={g} max={m} algorithmic_GB={rows*dim*2/1e9:.2f}", flush=True)

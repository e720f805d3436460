"""Host-side throughput of the search stage (Avro decode -> [search] -> flatten -> Avro encode -> produce) with the GPU
call stubbed out: the ceiling the Python serve loop puts on end-to-end QPS over the file-log transport."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qsa_b200.operator import VectorTable
from qsa_b200.pipeline.serve import Codec, Lab2Pipeline
from qsa_b200.transport.filelog import Producer


class NullIndex:
    dim = 1536
    def __len__(self): return 1000
    def search_host(self, q, k):
        return np.full((len(q), k), 0.5, np.float32), np.tile(np.arange(k, dtype=np.int32), (len(q), 1))
    def append(self, x): return 0
    def reset(self): pass
    def delete_rows(self, r): pass


if __name__ == "__main__":
    d = tempfile.mkdtemp()
    gpu = "--gpu" in sys.argv
    if gpu:
        sys.argv.remove("--gpu")
    t = VectorTable(NullIndex()); t.load_columns([f"d{i}" for i in range(1000)], ["chunk text " * 20] * 1000)
    if gpu:   # the real engine over a 1M x 1536 corpus: QPS_e2e of the search stage over the file-log transport
        import torch
        from tools.gpu_prof import fill as _fill
        fill_corpus = lambda ix, rows, dim, seed: _fill(ix, rows, dim, seed)
        from qsa_b200.engine import VectorIndex
        rows = 1_000_000
        ix = VectorIndex(dim=1536, capacity=rows, max_batch=1024, max_k=3)
        fill_corpus(ix, rows, 1536, 7)
        t = VectorTable(ix); t.load_columns([f"d{i}" for i in range(rows)], ["chunk text " * 20] * rows)
    pipe = Lab2Pipeline(d, t, k=3, max_batch=1024)
    codec = Codec(d); p = Producer({"log.dir": d})
    vec = np.random.randn(1536).astype(np.float32)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    t0 = time.time()
    for i in range(N):
        p.produce("queries_embed", value=codec.encode("queries_embed", {"query": f"question {i}", "embedding": vec}))
    p.flush()
    pipe._query_buffers()          # one-time setup (imports the engine module, page-locks the staging buffers)
    t1 = time.time()
    n = 0
    while (m := pipe.stage_search()):
        n += m
    t2 = time.time()
    print(f"produce queries_embed: {N/(t1-t0):.0f} msg/s   stage_search {'GPU 1M x 1536' if gpu else 'host path only'}: {n/(t2-t1):.0f} msg/s"
          f"   (search_host seconds {pipe.stats['search_seconds']:.3f} of {t2-t1:.3f})")

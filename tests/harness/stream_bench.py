"""Config 5 as a stream (BASELINE.json configs[4]): sensor embeddings arrive in epochs while queries are served.

One GPU holds `--rows` x `--dim` bf16 rows of capacity.  Every `--steps-per-epoch` search steps an epoch of
`--epoch-rows` fp32 rows is appended (device fp32 -> bf16 RNE + norms, `sa_corpus_append_f32`) on a side stream while
the search stream keeps going; each search sees the prefix committed when it was launched.  Prints one JSON line:
steady-state QPS during ingest, ingest throughput, and the final-epoch parity check against the oracle on a sample.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from qsa_b200.engine import VectorIndex

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=6_250_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--k", type=int, default=5)
ap.add_argument("--epoch-rows", type=int, default=1_000_000 // 8)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--steps-per-epoch", type=int, default=4)
a = ap.parse_args()

ix = VectorIndex(dim=a.dim, capacity=a.rows, max_batch=a.batch, max_k=a.k)
g = torch.Generator(device="cuda").manual_seed(5678)
side = torch.cuda.Stream()
q = torch.randn((a.batch, a.dim), generator=g, device="cuda").to(torch.bfloat16)
epoch = torch.randn((a.epoch_rows, a.dim), generator=g, device="cuda")           # one epoch of fp32 embeddings
# start half full so the scan has real work from step 0
half = (a.rows // 2) // a.epoch_rows * a.epoch_rows
for lo in range(0, half, a.epoch_rows):
    ix.append(epoch)
torch.cuda.synchronize()
n0 = len(ix)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
appended = 0
rows_seen = []
t0 = time.perf_counter()
ev0.record()
for step in range(a.steps):
    if step % a.steps_per_epoch == 0 and len(ix) + a.epoch_rows <= a.rows:
        with torch.cuda.stream(side):                                              # ingest overlaps the searches
            ix.lib.sa_corpus_append_f32(ix._h, epoch.data_ptr(), a.epoch_rows, side.cuda_stream)
        appended += a.epoch_rows
        torch.cuda.current_stream().wait_stream(side)                              # next search sees the new epoch
    rows_seen.append(len(ix))
    s, i = ix.search(q, a.k)
ev1.record()
torch.cuda.synchronize()
dt = ev0.elapsed_time(ev1) * 1e-3
# parity on the final state for a few queries
from oracle import bruteforce as bf
nq = 4
qb = q[:nq].view(torch.int16).cpu().numpy().view(np.uint16)
chunks = ((lo, ix.rows[lo:lo + (1 << 18)][: max(0, min(1 << 18, len(ix) - lo))].view(torch.int16).cpu().numpy().view(np.uint16))
          for lo in range(0, len(ix), 1 << 18))
rs, ri = bf.cosine_topk_fast(qb, chunks, a.k)
ok = bool((i[:nq].cpu().numpy() == ri).all())
mean_rows = float(np.mean(rows_seen))
print(json.dumps({"workload": f"stream: {a.rows}x{a.dim} capacity, batch {a.batch}, top-{a.k}, epochs of {a.epoch_rows} rows every {a.steps_per_epoch} steps",
                  "steps": a.steps, "qps_during_ingest": a.batch * a.steps / dt, "ms_per_step": dt / a.steps * 1e3,
                  "rows_start": n0, "rows_end": len(ix), "mean_rows_scanned": mean_rows,
                  "scan_gbs_on_mean_rows": mean_rows * a.dim * 2 / (dt / a.steps) / 1e9,
                  "ingest_rows_per_s": appended / dt, "final_state_parity_exact": ok}))

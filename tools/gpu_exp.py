"""Experiment harness (one process, corpus built once): sweeps engine options and prints scan/total ms, clocks, power."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import ClockSampler, fill_corpus
from qsa_b200.engine import VectorIndex


def run(ix, q, k, iters, label, host_q=None, pinned=None):
    sm = ClockSampler(0); sm.start(); time.sleep(0.05)
    for _ in range(3):
        ix.search(q, k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ix.search(q, k)
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    scan, tot, n = ix.timing_mean(iters)
    c = sm.summary(t0, t1); sm.stop()
    t = ix.last_timing()
    B = q.shape[0]
    msg = (f"{label:28s} step {e0.elapsed_time(e1)/iters:8.3f} ms  scan {scan:8.3f} ms  -> {t.flops/scan/1e9:7.1f} TF "
           f"{t.bytes/scan/1e6:7.1f} GB/s  {B/(e0.elapsed_time(e1)/iters)*1e3:9.0f} QPS  clk {c['sm_mhz']} MHz pw {c.get('power_w_max')} {c['reasons']}")
    if host_q is not None:
        for name, qq, out in (("pageable", host_q, None), ("pinned", pinned[0], (pinned[1], pinned[2]))):
            if qq is None: continue
            for _ in range(2): ix.search_host(qq, k, out=out)
            t2 = time.perf_counter()
            for _ in range(iters): ix.search_host(qq, k, out=out)
            dt = (time.perf_counter() - t2) / iters
            msg += f"  e2e[{name}] {dt*1e3:7.3f} ms"
    print(msg, flush=True)


if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    # settings: "cg,drift,gain,max;..."
    settings = [tuple(int(x) for x in s.split(",")) for s in (sys.argv[4] if len(sys.argv) > 4 else "2,1,0,128").split(";")]
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    ix = VectorIndex(dim=dim, capacity=rows, max_batch=4096, max_k=10)
    fill_corpus(ix, rows, dim, 1234)
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((B, dim), generator=g, device="cuda").to(torch.bfloat16)
    for st in settings:
        cg, d, gain, mx = st[:4]
        qpu2 = st[4] if len(st) > 4 else 1
        ix.set_option("cta_group", cg); ix.set_option("max_drift", d); ix.set_option("pace_gain", gain); ix.set_option("pace_max", mx)
        ix.set_option("qpu2", qpu2)
        run(ix, q, 10, iters, f"B={B} cg={cg} drift={d} gain={gain} max={mx} qpu2={qpu2}")

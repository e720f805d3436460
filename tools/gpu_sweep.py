#!/usr/bin/env python
"""Sweep one engine option over the BASELINE shapes: production scan kernel, CUDA-event time, answers compared with the
first setting (an option may only move work around, never change a result).

    python tools/gpu_sweep.py --opt l2_prefetch=0,4,8,16,24 --shapes cfg5,b128,b1024,cfg2 [--rounds 2]

Settings are interleaved over `--rounds` passes so that thermal drift under the power cap does not favour any of them.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler  # noqa: E402
from tools.gpu_prof import SHAPES, fill  # noqa: E402
from qsa_b200.engine import VectorIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--opt", required=True, help="name=v1,v2,...")
    ap.add_argument("--fixed", default="", help="other options held fixed: a=1,b=2")
    ap.add_argument("--shapes", default="cfg5,b128,b1024,cfg2")
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--preheat", type=float, default=1.5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    name, vals = a.opt.split("=")
    vals = [int(v) for v in vals.split(",")]
    sampler = ClockSampler(0)
    sampler.start()
    res = []
    cache = {}
    for shape in a.shapes.split(","):
        n, dim, B, k = SHAPES[shape]
        if (n, dim) not in cache:
            cache.clear()
            torch.cuda.empty_cache()
            ix = VectorIndex(dim=dim, capacity=n, max_batch=4096, max_k=10)
            fill(ix, n, dim, 1234)
            cache[(n, dim)] = ix
        ix = cache[(n, dim)]
        for kv in filter(None, a.fixed.split(",")):
            o, v = kv.split("=")
            ix.set_option(o, int(v))
        g = torch.Generator(device="cuda").manual_seed(4321)
        q = torch.randn((B, dim), generator=g, device="cuda").to(torch.bfloat16)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < a.preheat:
            for _ in range(4):
                ix.search(q, k)
            torch.cuda.synchronize()
        ref = None
        acc = {v: [] for v in vals}
        clk = {v: [] for v in vals}
        for _ in range(a.rounds):
            for v in vals:
                ix.set_option(name, v)
                for _ in range(4):
                    s, i = ix.search(q, k)
                torch.cuda.synchronize()
                tw0 = time.perf_counter()
                for _ in range(a.iters):
                    s, i = ix.search(q, k)
                torch.cuda.synchronize()
                tw1 = time.perf_counter()
                scan_ms, total_ms, m = ix.timing_mean(min(a.iters, 16))
                acc[v].append((scan_ms, total_ms))
                c = sampler.summary(tw0, tw1)
                clk[v].append(c.get("sm_mhz") or 0)
                if ref is None:
                    ref = i.clone()
                elif not torch.equal(ref, i):
                    raise SystemExit(f"{shape}: {name}={v} changed the answer")
        t = ix.last_timing()
        for v in vals:
            scan = float(np.mean([x[0] for x in acc[v]]))
            tot = float(np.mean([x[1] for x in acc[v]]))
            row = {"shape": shape, name: v, "scan_ms": scan, "search_ms": tot,
                   "gbs": (n * dim * 2 + n * 4) * t.launches / (scan * 1e-3) / 1e9,
                   "tflops": t.flops / (scan * 1e-3) / 1e12, "sm_mhz": float(np.mean(clk[v])),
                   "per_round_scan_ms": [round(x[0], 4) for x in acc[v]]}
            res.append(row)
            print(json.dumps(row), flush=True)
    sampler.stop()
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

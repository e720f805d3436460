"""Development harness: one staged hardware check per process (a device trap poisons the CUDA context, so each
stage runs in its own interpreter under `timeout`).  Usage: python tests/harness/gpu_stage.py <stage> [args]."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from oracle import bruteforce as bf
from qsa_b200.engine import VectorIndex


def bits_to_dev(bits):
    return torch.from_numpy(bits.view(np.int16)).view(torch.bfloat16).cuda()


def stage_dots(cg, n=1024, dim=1536, nq=128):
    c = bf.synth_rows(11, 0, n, dim)
    q = bf.synth_rows(22, 0, nq, dim)
    ix = VectorIndex(dim=dim, capacity=n + 300, max_batch=max(nq, 256), max_k=10)
    ix.append_bf16_bits(c)
    qd = bits_to_dev(q)
    ok = True
    for tile in range((n + 255) // 256):
        got = ix.debug_tile_dots(qd, tile, cg)[:nq].cpu().numpy()
        torch.cuda.synchronize()
        lo, hi = tile * 256, min(n, tile * 256 + 256)
        ref = bf.bf16_bits_to_f32(q).astype(np.float64) @ bf.bf16_bits_to_f32(c[lo:hi]).astype(np.float64).T
        err = np.abs(got[:, : hi - lo] - ref).max()
        scale = np.abs(ref).max()
        print(f"cg={cg} dim={dim} nq={nq} tile={tile}: max|err|={err:.3e} (scale {scale:.3e})", flush=True)
        if not err < 1e-3 * scale:
            ok = False
            bad = np.argwhere(np.abs(got[:, : hi - lo] - ref) > 1e-3 * scale)
            print("  first bad (row,col):", bad[:8].tolist(), "got", got[tuple(bad[0])], "ref", ref[tuple(bad[0])])
    return ok


def stage_search(cg, n=20000, dim=1536, nq=200, k=10, force_fix=0):
    c = bf.synth_rows(1234, 0, n, dim)
    q = bf.synth_queries(4321, nq, dim, c)
    ix = VectorIndex(dim=dim, capacity=n + 1000, max_batch=max(nq, 256), max_k=max(k, 10))
    ix.set_option("cta_group", cg)
    ix.set_option("force_fix", force_fix)      # 1: every (query, lane) also goes through the exact fallback scan
    ix.append_bf16_bits(c)
    s, i = ix.search(bits_to_dev(q), k)
    torch.cuda.synchronize()
    rs, ri = bf.cosine_topk_f64(q, c, k)
    rep = bf.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri, rs)
    print(f"search cg={cg} n={n} dim={dim} nq={nq} k={k}: {rep}", flush=True)
    t = ix.last_timing()
    print("  timing", t, flush=True)
    return rep["strict_order"] == 1.0 and rep["max_abs_dscore"] < 1e-6


def stage_perf(cg, n=1_000_000, dim=1536, nq=256, k=10, iters=5):
    ix = VectorIndex(dim=dim, capacity=n, max_batch=max(nq, 256), max_k=10)
    ix.set_option("cta_group", cg)
    step = 1 << 18
    g = torch.Generator(device="cuda").manual_seed(5)
    for lo in range(0, n, step):
        m = min(step, n - lo)
        ix.rows[lo:lo + m].copy_(torch.randn((m, dim), generator=g, device="cuda", dtype=torch.float32))
    ix.commit(0, n)
    q = torch.randn((nq, dim), generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    res = []
    for it in range(iters + 2):
        ix.search(q, k)
        torch.cuda.synchronize()
        t = ix.last_timing()
        res.append(t)
    t = res[-1]
    ms = np.median([r.scan_ms for r in res[2:]])
    tot = np.median([r.total_ms for r in res[2:]])
    print(f"perf cg={cg} n={n} dim={dim} nq={nq}: scan {ms:.3f} ms total {tot:.3f} ms launches {t.launches} "
          f"-> {t.bytes / ms / 1e6:.1f} GB/s, {t.flops / ms / 1e9:.1f} TFLOP/s, {nq / tot * 1e3:.0f} QPS", flush=True)
    return True


if __name__ == "__main__":
    st = sys.argv[1]
    args = [int(a) for a in sys.argv[2:]]
    t0 = time.time()
    ok = {"dots": stage_dots, "search": stage_search, "perf": stage_perf}[st](*args)
    print(f"STAGE {st} {args} -> {'OK' if ok else 'FAIL'} in {time.time() - t0:.1f}s", flush=True)
    sys.exit(0 if ok else 1)

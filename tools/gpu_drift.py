"""Per-CTA start/end timestamps of one scan launch: how far do the query blocks of a tile lane drift apart?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import fill_corpus
from qsa_b200.engine import VectorIndex

rows, dim, B = int(sys.argv[1]), 1536, int(sys.argv[2])
ix = VectorIndex(dim=dim, capacity=rows, max_batch=4096, max_k=10)
fill_corpus(ix, rows, dim, 1234)
q = torch.randn((B, dim), device="cuda").to(torch.bfloat16)
ix.set_option("record_times", 1)
cudart = ctypes.CDLL("libcudart.so.12") if False else None
for cg, d in ((2, 0), (2, 2), (1, 0), (1, 2)):
    ix.set_option("cta_group", cg); ix.set_option("max_drift", d)
    for _ in range(3):
        ix.search(q, 10)
    torch.cuda.synchronize()
    grid = ix.info("last_grid"); ptr = ix.info("dbg_times_ptr")
    buf = torch.empty(grid * 2, dtype=torch.int64, device="cuda")
    # device-to-device copy of the debug buffer through a ctypes view
    src = (ctypes.c_int64 * (grid * 2)).from_address  # noqa (placeholder to keep flake quiet)
    t = torch.from_dlpack  # noqa
    import torch.cuda as tc
    tmp = torch.empty(grid * 2, dtype=torch.int64)
    libc = ctypes.CDLL(None)
    rt = ctypes.CDLL([l.split()[-1] for l in open("/proc/self/maps") if "libcudart" in l][0])
    rt.cudaMemcpy(ctypes.c_void_p(tmp.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(grid * 16), ctypes.c_int(2))
    tt = tmp.numpy().reshape(grid, 2).astype(np.float64)
    t0 = tt[:, 0].min()
    dur = (tt[:, 1] - tt[:, 0]) / 1e6
    end = (tt[:, 1] - t0) / 1e6
    units = dur.reshape(-1, cg)[:, 0]
    nqb = (B + 128 * cg - 1) // (128 * cg); TL = len(units) // nqb
    per_lane = end.reshape(-1, cg)[:, 0].reshape(TL, nqb)
    print(f"cg={cg} drift={d}: grid {grid} kernel {end.max():.3f} ms; CTA duration min/mean/max {dur.min():.3f}/{dur.mean():.3f}/{dur.max():.3f} ms; "
          f"start skew {((tt[:,0]-t0)/1e3).max():.1f} us; within-lane end spread mean {np.mean(per_lane.max(1)-per_lane.min(1))*1e3:.1f} us "
          f"max {np.max(per_lane.max(1)-per_lane.min(1))*1e3:.1f} us; lane-mean end min/max {per_lane.mean(1).min():.3f}/{per_lane.mean(1).max():.3f}", flush=True)

"""The N>1 path on CPU: two gloo ranks, each holding a row shard, one all-gather of per-shard candidates, merge.
The shard search itself is an oracle-backed double here (tests/doubles.py); the GPU flavour of the same flow is
tests/test_gpu_parity.py::test_merge_shards_equals_global and the multi-rank bench."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bruteforce as bf


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, nq, k, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from doubles import OracleIndex
    from qsa_b200.sharded import ShardedIndex, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = bf.synth_rows(41, 0, n, dim)
    c[n - 5] = c[3]                                   # a tie that straddles the shard boundary
    q = bf.synth_queries(42, nq, dim, c)
    q[0] = c[3]
    lo, hi = shard_bounds(n, world, rank)
    ix = OracleIndex(dim)
    ix.bits = c[lo:hi]
    sh = ShardedIndex(ix, row_offset=lo)
    s, gi = sh.search(torch.from_numpy(bf.bf16_bits_to_f32(q)), k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), s=s.numpy(), i=gi.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_search_equals_unsharded(tmp_path):
    n, dim, nq, k, world = 3001, 64, 9, 10, 2
    mp.spawn(_worker, args=(world, _free_port(), n, dim, nq, k, str(tmp_path)), nprocs=world, join=True)
    c = bf.synth_rows(41, 0, n, dim)
    c[n - 5] = c[3]
    q = bf.synth_queries(42, nq, dim, c)
    q[0] = c[3]
    rs, ri = bf.cosine_topk_f64(q, c, k)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert (r0["i"] == ri).all() and (r1["i"] == ri).all()          # every rank ends with the global answer
    assert ri[0, 0] == 3 and ri[0, 1] == n - 5                      # lower global row wins the cross-shard tie
    assert np.abs(r0["s"].astype(np.float64) - rs).max() < 1e-6


def test_shard_bounds_cover_everything():
    from qsa_b200.sharded import shard_bounds
    for n in (10, 10_000_000, 7):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(x[1] == y[0] for x, y in zip(b, b[1:]))

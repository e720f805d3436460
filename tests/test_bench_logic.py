"""Host logic of bench.py that must hold on every rank count: loops that contain collectives run the same number of
iterations on all ranks (a per-rank clock deadlocked an 8-GPU run once), and the reference arm prints the contract's
JSON line without touching the GPU."""
import json
import os
import socket
import subprocess
import sys
import time

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf = [torch.zeros(4) for _ in range(world)]
    calls = [0]

    def step():                                   # a step with a collective in it; rank 1 is 5x slower than rank 0
        time.sleep(0.002 if rank == 0 else 0.010)
        dist.all_gather(buf, torch.full((4,), float(rank)))
        calls[0] += 1

    def reduce_max(flag):
        t = torch.tensor([flag], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    n = bench.collective_preheat(step, 0.25, world, lambda: None, reduce_max)
    for _ in range(3):                            # "warm-up + timed" steps afterwards must still pair up
        step()
    dist.barrier()
    with open(os.path.join(out_dir, f"r{rank}.json"), "w") as f:
        json.dump({"n": n, "calls": calls[0]}, f)
    dist.destroy_process_group()


def test_preheat_runs_the_same_number_of_collective_steps_on_every_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [json.load(open(tmp_path / f"r{i}.json")) for i in range(world)]
    assert r[0] == r[1] and r[0]["n"] >= 4 and r[0]["calls"] == r[0]["n"] + 3


def test_preheat_single_rank_and_disabled():
    sys.path.insert(0, ROOT)
    import bench
    calls = [0]

    def step():
        calls[0] += 1
        time.sleep(0.001)
    assert bench.collective_preheat(step, 0.0, 1, lambda: None) == 0 and calls[0] == 0
    n = bench.collective_preheat(step, 0.05, 1, lambda: None)
    assert n == calls[0] and n % 4 == 0 and n >= 4


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample-queries", "8", "--cpu-sample-rows", "4096", "--gpus", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("10000000x1536")
    # other ranks of a torchrun launch print nothing and exit 0
    out1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert out1.returncode == 0 and out1.stdout.strip() == ""


def test_harness_and_tool_scripts_compile():
    """The GPU harness can only run on a B200 box; at least keep it syntactically alive here."""
    import glob
    import py_compile
    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "harness", "*.py")) + \
        [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)
    for sh in glob.glob(os.path.join(ROOT, "tools", "*.sh")):
        assert subprocess.run(["bash", "-n", sh]).returncode == 0, sh
        for ref in __import__("re").findall(r"(?:python|bash) ((?:tools|tests)/[\w/.]+)", open(sh).read()):
            assert os.path.exists(os.path.join(ROOT, ref)), f"{sh} refers to missing {ref}"

"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py / the replay driver (sequence
partitioning, barrier, max-over-ranks timing, sum-over-ranks units).  No data-path collective exists."""
import json
import os
import subprocess
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from lidar_visual_inertial_slam_b200 import multi
    r, w, lr = multi.init(backend="gloo")
    assert (r, w) == (rank, world)
    mine = multi.partition_sequences(64, r, w)
    seeds = [multi.sequence_seed(0x5EED0000, s) for s in mine]
    multi.barrier()
    elapsed_ms = 100.0 * (r + 1)                       # rank 1 is the slow one
    max_ms, total = multi.reduce_timing(elapsed_ms, units=len(mine) * 100)
    multi.barrier()
    with open(os.path.join(out_dir, "rank%d.json" % r), "w") as f:
        json.dump(dict(mine=mine, seeds=seeds, max_ms=max_ms, total=total,
                       value=multi.throughput(total, max_ms)), f)


def test_two_rank_partition_and_timing(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / ("rank%d.json" % r))) for r in range(world)]
    all_ids = sorted(res[0]["mine"] + res[1]["mine"])
    assert all_ids == list(range(64))                                   # every sequence exactly once
    assert res[0]["mine"] == list(range(0, 64, 2)) and res[1]["mine"] == list(range(1, 64, 2))
    assert res[0]["seeds"][1] == 0x5EED0000 + 2
    for r in res:
        assert r["max_ms"] == 200.0 and r["total"] == 6400.0            # max over ranks, sum over ranks
        assert abs(r["value"] - 6400.0 / 0.2) < 1e-6


def test_reference_arm_only_rank0_works():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""                 # other ranks exit 0 without work


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "2", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "registrations/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["steps"] == 2

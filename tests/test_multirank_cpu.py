"""CPU (gloo, world_size 2): the multi-GPU path of bench.py — rendezvous, barrier, MAX-over-ranks reduction,
whole-job throughput, rank-0-only JSON — exercised with synthetic per-rank timings (replicas
only, no data-path collective)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, extra):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29653", os.path.join(ROOT, "bench.py"), "--gpus", str(nproc),
           "--steps", "4", "--warmup", "3", "--dry-run"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line (rank 0 only): %r" % out.stdout
    return json.loads(lines[0])


def test_two_rank_reduction_is_max_over_ranks_and_sum_of_frames():
    d = _run(2, ["--streams", "16"])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["dry_run"]
    # rank r reports (10 + r) ms per step: the job takes the slower rank's 11 ms; 2 x 16 streams per step
    assert abs(d["ms_per_step"] - 11.0) < 1e-9
    assert abs(d["value"] - 32 / 11e-3) < 1e-6
    assert abs(d["e2e"]["value"] - 32 * 3 / (21e-3 * 3)) < 1e-6


def test_single_rank_dry_run_matches_formula():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dry-run", "--streams", "8", "--steps", "5"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["n_gpus"] == 1 and abs(d["value"] - 8 / 10e-3) < 1e-6


def test_reference_arm_only_rank0_works_under_torchrun():
    """`--impl reference` under torchrun: rank 0 prints the line, the other ranks exit 0 without work."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29654", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0

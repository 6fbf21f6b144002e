"""bench.py's multi-rank control flow on a machine without a GPU: two ranks over gloo, the C ABI over the simulated CUDA runtime,
torch.cuda replaced by stand-ins (tools/mock_bench.py).  Numbers are meaningless; what is checked is that the whole product arm
— binding, process group, kernel-only arm, end-to-end arm with the per-step consensus gather, reductions, rank 0's JSON line,
the other rank's wait for it — runs through and both ranks exit cleanly.  (This run is what found the ranks handing the
gather blocks of different widths.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_bench_flow_over_the_simulated_runtime():
    from racon_b200 import build
    build.build_simapi()
    env = dict(os.environ, RP_BENCH_NO_BY_REFERENCE="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "RACON_B200_LIB"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mock_bench.py"), "--mock-world", "2", "--windows", "1",
                        "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["gpu_launches"] >= 1
    assert set(d["e2e"]["host_ms_per_step_rank0"]) == {"add", "run", "wait", "fetch", "gather"}
    assert d["config"]["baseline_config"] == 2 and "workload" in d["config"]
    assert d["roofline"]["kernel"] == "rp_poa_kernel" and d["roofline"]["algorithmic_bytes_per_window"] > 0

"""bench.py contract (CPU-checkable part): the reference arm prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-n", "200", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    # the N=8192 solve does not fit a budget extrapolated from this machine's N=200 sample: the arm reports the sample
    # under ITS OWN name (never under the N=8192 label)
    assert line["impl"] == "reference" and line["metric"] == "zhegvdx_n200_gflops" and line["unit"] == "GFLOP/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "ZHEGVDX N=200 " in line["config"]["workload"] and line["config"]["same_workload_as_gpu_arm"] is False
    assert line["cpu_baseline"]["cores"] == os.cpu_count()      # explicit thread count, whatever OMP_NUM_THREADS says


def test_reference_arm_runs_the_real_order_once_when_it_fits():
    env = dict(os.environ, OMP_NUM_THREADS="1")                 # what torchrun exports
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "300", "--ref-n", "100",
                          "--steps", "20", "--warmup", "5"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["metric"] == "zhegvdx_n300_gflops" and line["steps"] == 1 and line["requested_steps"] == 20
    assert "ZHEGVDX N=300 " in line["config"]["workload"] and line["config"]["same_workload_as_gpu_arm"] is True
    assert line["cpu_baseline"]["cores"] == os.cpu_count()


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

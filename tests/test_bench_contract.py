"""CPU suite: the JSON contract of `bench.py --impl reference` (the arm the driver runs beside the product arm).
One bounded step of the reference's CPU implementation on the host cores -- the real reference from baseline/_ref when
tools/install_reference.py staged it (kind "reference"), else the oracle port (kind "port"); no GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_rays_per_sec" and line["unit"] == "rays/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    assert line["value"] > 0 and abs(line["value"] - line["cpu_baseline"]["value"]) < 1e-9
    staged = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "src", "nerf_models"))
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same workload string as the product arm
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"]["workload"] == bench.WORKLOAD % bench.N_RAND

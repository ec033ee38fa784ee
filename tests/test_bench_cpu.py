"""bench.py pieces that need no GPU: the reference arm (the oracle port timed on the host cores) and the helpers that read the
committed evidence."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "640x480 depth frames/sec integrated+tracked"
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and abs(line["ms_per_step"] - 1e3 / line["value"]) < 0.01 * line["ms_per_step"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "frame" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_committed_evidence_is_readable():
    sys.path.insert(0, str(ROOT))
    import bench
    t = bench.ncu_traffic()
    assert t is not None and 1e5 < t < 1e9                      # DRAM bytes per launch of the dominant kernel, from profiles/
    pk = bench.peaks()
    assert pk["bf16_sustained"] > 100 and pk["hbm"] > 1000
    for name in ("r01_bench_final.json", "r02_bench_final.json"):
        final = json.loads((ROOT / "profiles" / name).read_text().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                    "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert key in final, (name, key)
        assert final["gpu_launches"] > 0 and final["roofline"]["frac"] == round(final["roofline"]["achieved"] / final["roofline"]["peak"], 5)
    # round 2 adds the reference's own CUDA path timed on the same GPU and the parity of the DEFAULT engines against it
    assert final["cuda_reference"]["value"] > 0 and final["parity"]["voxel_ids_equal"] is True
    assert final["parity"]["H_rel_on_reference_map"] < 1e-4 and final["parity"]["sdf_max_abs_m_on_reference_map"] < 1e-4
    assert final["e2e"]["h2d_bytes_per_step"] > 0 and final["clocks"]["reasons"] == []

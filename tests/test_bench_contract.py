"""bench.py contract pieces that do not need a GPU: the reference arm (`--impl reference`) prints ONE JSON line
with the agreed keys for every workload, and the workload table matches BASELINE.json's configs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


@pytest.mark.parametrize("workload,extra", [("tc", ["--scale", "10"]), ("clique4", ["--scale", "10"]),
                                            ("diamond", ["--shape-div", "2048"]), ("motif4", ["--shape-div", "32768"])])
def test_reference_arm_json_line(workload, extra):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                        "--steps", "1", "--warmup", "0"] + extra, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert p.returncode == 0, p.stderr[-800:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["unit"] in ("edges/s", "matches/s")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["workload"].startswith({"tc": "tc_rmat", "clique4": "kclique4_rmat", "diamond": "sgl_diamond",
                                               "motif4": "motif4_friendster"}[workload])


def test_workloads_cover_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5

    class A:
        scale = 0; shape_div = 0
    names = {}
    for w in ("tc", "clique4", "diamond", "motif4"):
        A.workload = w
        names[w] = bench.Workload(A, 1)
    # default = the north_star Target size (TC on R-MAT scale 24 at every N); configs[1] is `--scale 22`
    assert names["tc"].scale == 24 and "scale-24" in json.load(open(os.path.join(ROOT, "BASELINE.json")))["north_star"]
    A.workload = "tc"; A.scale = 22
    assert bench.Workload(A, 1).name == "tc_rmat_scale22" and "scale-22" in cfgs[1]
    A.scale = 0
    assert names["clique4"].scale == 23 and "scale-23" in cfgs[2]
    assert bench.LJ_NV == 4_847_571 and "4.8M" in cfgs[3]
    assert bench.FR_NV == 65_608_366 and "65M" in cfgs[4]
    assert all(x.scaling == "strong" for x in names.values())

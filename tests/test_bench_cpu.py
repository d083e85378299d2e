"""bench.py's contract on a box without a GPU: the reference arm (the reference's own CPU path from oracle/_ref) prints
exactly ONE JSON line on stdout -- whatever libraries write to stdout goes to stderr -- with the keys the driver
reads; record helpers keep the two arms' `config` key sets identical."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args):
    env = dict(os.environ, PYTHONUNBUFFERED="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_one_json_line_small_workload():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref (the patched reference build) is not present on this box")
    rec = _run("--impl", "reference", "--workload", "plume128_jacobi28", "--steps", "2", "--warmup", "1")
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "Mcells/s"
    assert rec["steps"] == 2 and rec["warmup"] == 1            # the requested counts, not shortened ones
    assert rec["config"]["workload"] == "plume128_jacobi28" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] == "reference" and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    assert tuple(rec["config"]) == bench.CONFIG_KEYS            # same key set as the GPU arm's records


def test_reference_arm_unavailable_for_3d():
    rec = _run("--impl", "reference", "--workload", "cube256_scalenet_slicewise")
    assert rec["impl"] == "reference" and "unavailable" in rec


def test_workload_table_is_consistent():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.DEFAULT_WORKLOAD in bench.WORKLOADS and all(w in bench.WORKLOADS for w in bench.ALSO_SINGLE)
    d = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert d["res"] == (1, 4096, 4096) and d["jacobi_iters"] == 100     # BASELINE.json configs[3]: the largest 1-GPU config
    for name, wl in bench.WORKLOADS.items():
        assert wl["method"] in ("jacobi", "convnet") and len(wl["res"]) == 3, name
        assert bench.step_bytes_per_cell(wl) == (80 if wl["res"][0] == 1 else 104) + 16 * wl["jacobi_iters"], name
        mconf = bench.workload_mconf(wl)
        assert mconf["simMethod"] == wl["method"] and mconf["pTol"] == 0.0, name

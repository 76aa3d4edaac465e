"""bench.py's JSON contract that can be checked without a GPU: both arms describe the workload with the SAME `config`
dict (the driver compares them), and the reference arm prints the keys the contract names."""
import argparse
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_config_is_a_pure_function_of_the_arguments():
    b = _bench()
    args = argparse.Namespace(streams=2, batches_per_call=20)
    for name, wl in b.WORKLOADS.items():
        c1, c2 = b.workload_config(wl, name, args), b.workload_config(wl, name, args)
        assert c1 == c2 and c1["workload"] == wl["name"] and "l2" in c1
        json.dumps(c1)
    tum = b.workload_config(b.WORKLOADS["tum"], "tum", args)
    assert "222 MB" in tum["l2"]            # 64 pairs x (4C+2) maps x 1.33 levels at 120x160
    vga = b.workload_config(b.WORKLOADS["vga"], "vga", args)
    assert "444 MB" in vga["l2"]


def test_reference_arm_line_has_the_contract_keys():
    from baseline import reference as REF
    if not REF.available():
        pytest.skip("baseline/_ref not installed")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "deepic",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    b = _bench()
    args = argparse.Namespace(streams=2, batches_per_call=20)
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["config"] == b.workload_config(b.WORKLOADS["deepic"], "deepic", args)
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

"""CPU check of bench.py's reference arm (the unmodified reference timed on the host cores) and of the
JSON contract keys; uses BASELINE config 1 (256x256, 4 spp) so it runs in seconds."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "0", "--sqrtspp", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Mray/s" and line["value"] > 0
    assert line["vs_baseline"] is None and line["higher_is_better"] is True
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]

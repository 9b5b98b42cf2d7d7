"""CPU: the C-ABI library is built, loads, and exports every symbol include/mflow_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "mflow_b200.h")
LIB = os.path.join(REPO, "manifold-flow_b200", "lib", "libmflow_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MFLOW_API\s+[\w\s\*]+?\b(mflow_\w+)\s*\(", text)))


def test_header_declares_the_abi():
    syms = declared_symbols()
    assert len(syms) >= 18
    for must in ("mflow_rqs_apply", "mflow_rqs_backward", "mflow_rqs_search", "mflow_chanmix_apply", "mflow_jtj_logdet",
                 "mflow_gauss_logprob", "mflow_version", "mflow_last_error_string"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__

        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/mflow_b200.h but not exported"
    lib.mflow_version.restype = ctypes.c_int
    assert lib.mflow_version() >= 1


def test_python_binding_covers_the_header():
    from manifold_flow.b200 import cabi

    assert sorted(cabi.EXPORTED_SYMBOLS) == declared_symbols()
    cabi.lib()  # loads and types every entry point


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "manifold-flow_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "mflow_oracle" not in text and "import oracle" not in text, f"{f} references the oracle"


def test_cpu_tensors_fail_loudly():
    import torch
    from manifold_flow.b200 import cabi, factory

    model = factory.create_model("cfg1")
    with pytest.raises(cabi.MflowError):
        model(torch.randn(4, 3), mode="projection")


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm: the oracle port on the host cores) prints ONE JSON line with
    the contract's keys; run here on the small vector configuration."""
    import json
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]

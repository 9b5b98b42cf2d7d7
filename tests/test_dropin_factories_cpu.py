"""CPU, build container only (skipped where /root/reference is absent): the reference's OWN model
factories (experiments/architectures/create_model.py) run unchanged against the drop-in
``manifold_flow`` package and produce the models the golden fixtures describe."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types, json
import numpy as np
np.int = int; np.float = float; np.product = np.prod          # numpy-2 shims the reference needs (SURVEY 8(c))
for n in ("matplotlib", "matplotlib.pyplot", "configargparse"):
    sys.modules.setdefault(n, types.ModuleType(n))
sys.path.insert(0, "{repo}/manifold-flow_b200")                  # the drop-in package FIRST
sys.path.insert(0, "{repo}/tests")
sys.path.append("{ref}/experiments")                            # reference factories, unchanged
import torch, golden_util as gu
import manifold_flow
assert "manifold-flow_b200" in manifold_flow.__file__, manifold_flow.__file__
from architectures.create_model import create_model             # reference code
meta = gu.load_json("models.json")
out = {{}}
for name in ("cfg1", "cfg3", "cfg4s", "cfg5s"):
    torch.manual_seed(gu.SEED)
    model = create_model(gu.config_args(name), gu.FakeSimulator(name))
    sd = model.state_dict()
    out[name] = (gu.state_dict_hashes(sd) == meta[name]["hashes"]) and ([k for k in sd] == [k for k, _, _ in meta[name]["keys"]])
print(json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_reference_factories_build_dropin_models():
    res = subprocess.run([sys.executable, "-c", SCRIPT.format(repo=REPO, ref=REF)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out == {"cfg1": True, "cfg3": True, "cfg4s": True, "cfg5s": True}

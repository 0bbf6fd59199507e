"""tools/make_model_golden.py -- tests/golden/minkunet_state_dicts.json from the reference's own model code.

Runs in the build container (needs /root/reference): imports the UNMODIFIED utils/minkunet.py + utils/resnet.py of
the reference on top of this repository's `MinkowskiEngine/` compat package, instantiates every variant the file
defines with working PLANES (utils/minkunet.py:208-245) and records the state-dict keys and shapes in order.
tests/test_oracle_sparse.py::test_model_family_matches_reference_state_dicts compares
canonicalvoting_b200/minkunet.py (which describes the family by a table) against it.  What this pins: the layer
structure built by the reference's `network_initialization` / `_make_layer`.  What it cannot pin: MinkowskiEngine's
own parameter names (`kernel`, `bn.weight`, ...), which come from the compat package [ME-recall].
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

import utils.minkunet as ref  # noqa: E402

VARIANTS = ["14A", "14B", "14C", "14D", "18A", "18B", "18D", "34A", "34B", "34C"]
HEADS = {"joint": (3, 64), "separate": (3, 8)}          # train_joint.py:218, train_separate.py:210

out = {}
for v in VARIANTS:
    for head, (cin, cout) in HEADS.items():
        if head == "separate" and v != "34C":
            continue
        m = getattr(ref, "MinkUNet" + v)(cin, cout)
        out["MinkUNet%s/%s" % (v, head)] = [[k, list(t.shape)] for k, t in m.state_dict().items()]
path = os.path.join(ROOT, "tests", "golden", "minkunet_state_dicts.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print(path, {k: len(v) for k, v in out.items()})

#!/usr/bin/env python
"""Golden vectors for ValueNetwork.attention_weights (rl/policy/sarl.py:36,69-71,130-131), produced by the UNMODIFIED
reference network.

Runs only in the build container (needs /root/reference); writes tests/golden/attention_weights.npz.

    python tests/golden/make_attention_golden.py      (re-executes itself with the reference on PYTHONPATH)

For the steps of three recorded traces that hold the network inputs (EB-CADRL net on the cfg2 shape and on the shipped 24-human scene, SARL baseline
on 5 adults) the reference's rl.policy.sarl.ValueNetwork -- built with the dimensions of the shipped policy configs,
loaded with the golden state_dict -- is run on each of the 81 rotated lookahead states the trace recorded (`vin`, the
very tensors the reference fed its network); after every forward `model.attention_weights` is read.  The reference's
`get_attention_weights()` after a `predict` is the entry of the LAST action (index 80).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

if os.environ.get("EBC_GOLDEN_CHILD") != "1":
    env = dict(os.environ)
    env["EBC_GOLDEN_CHILD"] = "1"
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "oracle", "shims"), REF, os.path.join(REF, "tests")])
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], cwd=REF, env=env))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from rl.policy.sarl import ValueNetwork  # noqa: E402

torch.set_num_threads(1)
CASES = [  # (trace, weights, steps)
    ("trace_cfg2_h10_seed7", "weights_ebcadrl.npz", [0, 12]),
    ("trace_ebcadrl_h24_seed1000000", "weights_ebcadrl.npz", [0, 8]),
    ("trace_cfg1_adults5_seed1002", "weights_sarl_baseline.npz", [0, 10, 40]),
]


def network(w):
    dims = lambda prefix, idx: [int(w["%s.%d.weight" % (prefix, i)].shape[0]) for i in idx]  # noqa: E731
    D = int(w["mlp1.0.weight"].shape[1])
    net = ValueNetwork(D, 6, dims("mlp1", (0, 2)), dims("mlp2", (0, 2)), dims("mlp3", (0, 2, 4, 6)),
                       dims("attention", (0, 2, 4)), int(w["attention.0.weight"].shape[1]) == 2 * int(w["mlp1.2.weight"].shape[0]),
                       None, None)
    net.load_state_dict({k: torch.tensor(v) for k, v in w.items()})
    return net


def main():
    out = {}
    names = []
    for trace, weights, steps in CASES:
        z = np.load(os.path.join(HERE, trace + ".npz"))
        w = np.load(os.path.join(HERE, weights))
        net = network({k: w[k] for k in w.files})
        for t in steps:
            vin = z["s%03d_vin" % t]                       # [A, n, D] float32, n = the state's real rows
            att = np.zeros(vin.shape[:2], np.float32)
            val = np.zeros(vin.shape[0], np.float32)
            for a in range(vin.shape[0]):
                with torch.no_grad():
                    val[a] = net(torch.tensor(vin[a:a + 1])).item()
                att[a] = net.attention_weights
            assert np.allclose(val, z["s%03d_la_value" % t], atol=1e-6), "the rebuilt network is the trace's network"
            key = "%s_s%03d" % (trace, t)
            out[key] = att
            names.append(key)
            print(key, att.shape, "sum", att.sum(1).min(), att.sum(1).max(), "max weight", att.max())
    out["keys"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "attention_weights.npz"), **out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Run the reference's OWN test-suite, unmodified, with oracle/shims on PYTHONPATH.

This is how the oracle's rvo2 restatement is pinned at event level: the
reference's episode-outcome tests must pass with `rvo2` := oracle/ebc_oracle.c.
Only runs in the build container (needs /root/reference); never on the GPU box.

    python oracle/run_reference_tests.py [--with-train]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EBC_REFERENCE", "/root/reference")

SNIPPET = r"""
import sys, time, unittest, json
results = {}
t0 = time.time()
import tests.test_collisions as tc
suite = unittest.defaultTestLoader.loadTestsFromModule(tc)
r = unittest.TextTestRunner(verbosity=0).run(suite)
results["unit_collisions"] = (r.wasSuccessful(), r.testsRun)
from tests.test_collisions_simulation import run_test_collisions
ok, info = run_test_collisions()
results["collision_scenes"] = (ok, [str(x[1][0]) for x in info["info"]])
from test_basic_simulation import run_basic_simulation
for p in ["configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_2.config",
          "configs/test_configs/test_env_configs/env_adults_3_bikes_3.config",
          "configs/test_configs/test_env_configs/env_adults_5.config"]:
    results[p.split("/")[-1]] = run_basic_simulation(p)
from test_scene_simulation import run_scene_simulation
results["scene_static_10"] = run_scene_simulation(
    "configs/test_configs/test_env_configs/env_adults_3_bikes_3_static_10.config",
    "tests/test_scenes/test_scene_adults_3_bikes_3_static_10.json")
from test_save_load_map import run_test_save_load_map
results["save_load_map"] = run_test_save_load_map()
if WITH_TRAIN:
    from test_basic_train import run_basic_train
    results["basic_train"] = run_basic_train()
results["wall_s"] = round(time.time() - t0, 1)
print("REFERENCE_TESTS " + json.dumps(results, default=str))
"""


def main():
    with_train = "--with-train" in sys.argv
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), REF, os.path.join(REF, "tests")])
    code = "WITH_TRAIN = %r\n" % with_train + SNIPPET
    return subprocess.call([sys.executable, "-c", code], cwd=REF, env=env)


if __name__ == "__main__":
    sys.exit(main())

"""ValueNetwork.attention_weights / SARL.get_attention_weights (rl/policy/sarl.py:36,69-71,130-131).

Golden vectors: tests/golden/attention_weights.npz, written by tests/golden/make_attention_golden.py -- the reference's
own ValueNetwork run on the 81 rotated lookahead states that the step traces recorded.  Stated tolerances (absolute, on
weights that sum to 1): oracle 5e-6 (it accumulates every layer in fp64 and rounds once, the reference's sgemm
accumulates in fp32), CUDA kernels 2e-5 (the tensor-core mode carries 22 mantissa bits per operand; the scores feed an
exp()).  Rows past a state's row count are exactly 0.
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "eb-cadrl_b200"))
sys.path.insert(0, HERE)

import oracle_backend as ob  # noqa: E402
from ebc.engine import BatchedSim  # noqa: E402

Z = np.load(os.path.join(ob.GOLDEN, "attention_weights.npz"))
KEYS = [str(k) for k in Z["keys"]]


def _run(key, device, backend, mode=None, fused=False, pad=0):
    name, t = key[:-5], int(key[-3:])
    tr = ob.Trace(name)
    H, S, R = tr.dims()
    sim = BatchedSim(tr.sim_config(), 1, H + pad, S, R, 81, device=device, backend=backend)
    sim.set_actions(tr.z["actions"])
    sim.set_weights(ob.load_weights(ob.TRACE_WEIGHTS[name]))
    if mode is not None:
        sim.set_value_mode(mode)
    tr.load_into(sim, t)
    att = sim.enable_attention()
    att.fill_(7.0)                      # every entry of the batch must be overwritten
    sim.orca()
    if fused:
        sim.lookahead(build_inputs=False)
        sim.value(fused=True)
    else:
        sim.lookahead()
        sim.value()
    rows = int(sim.hum_count[0]) + int(sim.stat_count[0])
    return sim.attention.cpu().numpy(), rows


@pytest.mark.parametrize("key", KEYS)
def test_oracle_attention_weights_match_reference(key):
    got, rows = _run(key, "cpu", ob.OracleBackend())
    gold = Z[key]
    assert gold.shape == (81, rows)
    assert np.abs(got[:, :rows] - gold).max() <= 5e-6
    assert (got[:, rows:] == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
@pytest.mark.parametrize("mode,fused", [("tc_fp16x2", False), ("tc_fp16x2", True), ("fp32", False), ("tc_fp32", False)])
def test_gpu_attention_weights_match_reference(key, mode, fused):
    got, rows = _run(key, "cuda:0", None, mode=mode, fused=fused)
    gold = Z[key]
    assert np.abs(got[:, :rows] - gold).max() <= 2e-5
    assert (got[:, rows:] == 0).all()
    ref, _ = _run(key, "cpu", ob.OracleBackend())
    assert np.abs(got - ref).max() <= 2e-5


@pytest.mark.gpu
def test_gpu_attention_weights_padded_rows_and_policy_surface():
    """Padded human slots (Hmax larger than the episode's count) leave zeros; SARL.get_attention_weights() after
    predict = the last action's weights (the reference's last forward), and env.step records them like env.py:355-356."""
    key = KEYS[0]
    got, rows = _run(key, "cuda:0", None, mode="tc_fp16x2", fused=True, pad=3)
    assert got.shape[1] == rows + 3 and (got[:, rows:] == 0).all()
    assert np.abs(got[:, :rows] - Z[key]).max() <= 2e-5
    from simulator.utils.test_utils import configure_env_policy_robot
    cfg = os.path.join(ob.GOLDEN, "configs")
    env, policy, robot = configure_env_policy_robot(os.path.join(cfg, "env_adults_5.config"), os.path.join(cfg, "policy.config"),
                                                    os.path.join(ob.GOLDEN, "weights_sarl_baseline.npz"))
    assert policy.get_attention_weights() is None                  # sarl.py:36: None before the first forward
    ob_, _ = env.reset("test", test_case=2)                        # = step 0 of trace_cfg1_adults5_seed1002
    action = robot.act(ob_, env=env)
    w = policy.get_attention_weights()
    assert w.shape == (5,) and abs(float(w.sum()) - 1.0) < 1e-5 and (w > 0).all()
    assert np.abs(w - Z["trace_cfg1_adults5_seed1002_s000"][80]).max() <= 2e-5     # the reference's last forward
    # the torch module (the training-side network, same weights) on the last action's rotated state agrees
    sim = env.native
    sim.lookahead()
    model = policy.get_model()
    vin = sim.vin[0, sim.A - 1, :5].unsqueeze(0).to(next(model.parameters()).device)
    with torch.no_grad():
        model(vin)
    assert np.abs(model.attention_weights - w).max() <= 2e-5
    model._attention = torch.as_tensor(w)
    env.step(action)
    assert len(env.attention_weights) == 1 and np.array_equal(env.attention_weights[0], w)   # env.py:355-356

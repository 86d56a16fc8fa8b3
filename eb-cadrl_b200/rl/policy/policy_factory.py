"""Registry of robot policies (rl/policy/policy_factory.py:1-10).  `cadrl`, `lstm_rl` and `sail` are
alternative policies outside the BASELINE configs; their keys resolve to a stub that says so."""
from rl.policy.sarl import SARL
from simulator.policy.policy_factory import policy_factory


def _out_of_scope(name):
    def make():
        raise NotImplementedError("policy %r is outside the B200 hot path (SURVEY §2 #13): use 'sarl'" % name)
    return make


policy_factory["sarl"] = SARL
for _name in ("cadrl", "lstm_rl", "sail"):
    policy_factory[_name] = _out_of_scope(_name)

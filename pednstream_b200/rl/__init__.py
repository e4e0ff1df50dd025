"""Control environments on top of the B200 LTM step.

`PedNetParallelEnv` keeps the reference's PettingZoo-style dict API for one network
(reference: rl/pz_pednet_env.py); `BatchedPedNetEnv` steps R independent replicas per launch with
actions, observations and rewards as device tensors; `GroupedPedNetEnv` runs several batches with differently
perturbed origin / destination nodes side by side.
"""
from .pz_pednet_env import PedNetParallelEnv  # noqa: F401
from .batched_env import BatchedPedNetEnv  # noqa: F401
from .grouped_env import GroupedPedNetEnv  # noqa: F401

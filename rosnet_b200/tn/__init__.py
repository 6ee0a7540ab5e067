"""Tensor-network front-end: equation / circuit generators, a
deterministic greedy contraction-path finder, a slice finder and a contraction driver that
issues ``rosnet_b200.tensordot`` / ``transpose`` on BlockArrays — what opt_einsum / cotengra /
quimb do for the reference's benchmarks (``benchmark/bench_random_tn.py``, ``bench_qft.py``,
``bench_sycamore.py``); none of those packages is available here."""
from .network import TensorNetwork, rand_equation, qft_amplitude_network, sycamore_amplitude_network  # noqa: F401
from .path import greedy_path, path_cost, find_slices, search_path, step_time_model, reconfigure_path  # noqa: F401
from .contract import contract_network, einsum, stage_slices  # noqa: F401

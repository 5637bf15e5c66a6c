"""CPU: randomised interleavings of the auxiliary-stage mbarrier protocol of conv_gemm_kernel
(tools/sim_epilogue_protocol.py).  Pins what the shipped protocol guarantees and reproduces the
phase-aliasing failure of the first two-group epilogue."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import sim_epilogue_protocol as sim  # noqa: E402


def _first_error(stages, bpt, variant, runs, threads_per_group=2, seed=0):
    rng = random.Random(seed)
    for _ in range(runs):
        try:
            sim.simulate(stages, bpt, rng.randint(3, 9), variant, rng, threads_per_group)
        except sim.ProtocolError as e:
            return str(e)
    return None


@pytest.mark.parametrize("stages", [2, 3, 4])
@pytest.mark.parametrize("blocks_per_tile", [1, 2, 4])
def test_shipped_protocol_is_hazard_free_with_two_or_more_stages(stages, blocks_per_tile):
    assert _first_error(stages, blocks_per_tile, "main", runs=120) is None


def test_single_stage_needs_the_release_wait():
    # the configuration that trapped on hardware: one stage shared by both epilogue groups
    assert _first_error(1, 2, "no_prewait", runs=200, threads_per_group=1) is not None
    assert _first_error(1, 2, "main", runs=200, threads_per_group=1) is None
    assert _first_error(1, 4, "main", runs=200, threads_per_group=1) is None


def test_single_stage_is_still_not_safe_for_unsynchronised_sibling_threads():
    # documented limitation: never launch with one auxiliary stage (run_conv's tile-width rule)
    assert _first_error(1, 2, "main", runs=400, threads_per_group=2) is not None


@pytest.mark.parametrize("slots", [2, 4])
@pytest.mark.parametrize("blocks_per_tile", [1, 2, 4])
def test_inplace_residual_epilogue_is_hazard_and_deadlock_free(slots, blocks_per_tile):
    # the lean inference epilogue (LEAN && RES): result written over the consumed residual tile,
    # slot handed back by each thread once its own bulk store has read it
    rng = random.Random(7)
    for _ in range(150):
        sim.simulate_inplace(slots, blocks_per_tile, rng.randint(3, 9), rng, threads_per_group=2)
    for _ in range(50):
        sim.simulate_inplace(slots, blocks_per_tile, rng.randint(3, 9), rng, threads_per_group=4)

"""Rooms with solid blocks inside (tests/cases.py OBSTACLE_CASES: a pillar, a one-node plate, isolated voxels with K = 0, an
L-shaped block): boundary nodes away from the walls with every adjacency pattern a staircase produces.  CUDA path against the
oracle and the reference's golden traces, bit for bit: traces, and whole grids + branch state from noise, every kernel variant.
(Last file of the suite on purpose: these cases were added after the round's GPU budget was spent; the CPU oracle is pinned on them.)"""
import numpy as np
import pytest

from cases import OBSTACLE_CASES, make_sim_data, noise_grids
from oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", (2, 1))
@pytest.mark.parametrize("name", sorted(OBSTACLE_CASES))
def test_obstacle_rooms_bit_exact(name, precision):
    from pffdtd_b200.engine import Engine
    sd = make_sim_data(name, precision)
    ref = Oracle(sd).run_all()
    g1, g0 = noise_grids(sd)
    o = Oracle(sd)
    o.write_grid(1, g1)
    o.write_grid(0, g0)
    o.run_steps(0, 20)
    vo, go = o.read_boundary_state()
    for ak, fuse in (((0, 0), (1, 0), (1, 1)) if sd.fcc_flag == 0 else ((0, 0), (1, 0))):
        with Engine(sd) as e:
            e.set_option("air_kernel", ak)
            e.set_option("fuse", fuse)
            e.run_steps(0, sd.Nt)
            assert np.array_equal(e.read_outputs(), ref), f"{name} p{precision} ak={ak} fuse={fuse}: traces"
        with Engine(sd) as e:
            e.set_option("air_kernel", ak)
            e.set_option("fuse", fuse)
            e.write_grid(1, g1)
            e.write_grid(0, g0)
            e.run_steps(0, 20)
            for which in (1, 0):
                a, b = e.read_grid(which)[1:-1, 1:-1, 1:-1], o.read_grid(which)[1:-1, 1:-1, 1:-1]
                assert np.array_equal(a, b), f"{name} p{precision} ak={ak} fuse={fuse} grid{which}: {np.abs(a - b).max():.3e}"
            v, g = e.read_boundary_state()
            assert np.array_equal(v, vo) and np.array_equal(g, go)

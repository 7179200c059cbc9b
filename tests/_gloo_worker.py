"""Worker of test_two_rank_gloo_shard_and_gather: exercises moveit_b200.sharding over gloo on the CPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_py  # noqa: E402
from common import OracleSetup, oracle_world_points  # noqa: E402
from moveit_b200 import robot_description as rd  # noqa: E402
from moveit_b200 import scenarios  # noqa: E402
from moveit_b200.sharding import broadcast_field_arrays, gather_verdicts, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = rd.panda()
    objs = scenarios.random_scene()
    fa = scenarios.FIELD_A
    # rank 0 "builds" the field (occupancy voxels), every other rank receives it by ONE broadcast
    setup = OracleSetup(oracle_py, desc, fa, (), self_state=np.zeros(desc.nvar))
    dims = setup.env.world.dims
    occ = torch.zeros(dims, dtype=torch.uint8)
    if rank == 0:
        tmp = oracle_py.Field(fa["size"], fa["resolution"], scenarios.field_corner(fa), fa["max_distance"])
        tmp.add_points(oracle_world_points(oracle_py, objs, fa["resolution"]))
        occ = torch.from_numpy((tmp.d2() == 0).astype(np.uint8))
    broadcast_field_arrays([occ], dist, src=0)
    setup.env.world.add_voxels(np.argwhere(occ.numpy()))
    n = 5001
    q = desc.sample_states(n, 77)
    lo, hi = shard_range(n, rank, world)
    local = setup.env.check_batch(q[lo:hi], mode=3)
    full = gather_verdicts(local, n, rank, world, dist)
    if rank == 0:
        single = OracleSetup(oracle_py, desc, fa, objs, self_state=np.zeros(desc.nvar)).env.check_batch(q, mode=3)
        assert np.array_equal(full, single)
        print("GATHER_OK", int(full.sum()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

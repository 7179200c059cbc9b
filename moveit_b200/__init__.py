"""moveit_b200 — B200-native batched state-validity engine for MoveIt's distance-field collision path.

The product is libb200df.so (C ABI in include/b200df.h, CUDA kernels in moveit_b200/csrc/) and the C++ host
mirror of CollisionEnv in moveit_b200/host/.  The Python modules here are plumbing: build recipe, ctypes binding,
robot descriptions and synthetic workloads.
"""
__all__ = ["binding", "engine", "robot_description", "scenarios", "rng"]

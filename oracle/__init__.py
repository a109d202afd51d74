"""CPU oracle for the collective path.  TEST INFRASTRUCTURE ONLY — never imported by ant_ray_b200.

See oracle_reduce.c for what is restated and how the oracle itself is pinned.
"""

"""CPU oracle for the RfD-Net point-cloud hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/pointnet2_oracle.c header.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package rfdnet_b200 must never import it.
"""
from .cpu_ref import (  # noqa: F401
    furthest_point_sampling, gather_points, gather_points_grad, ball_query, group_points,
    group_points_grad, three_nn, three_interpolate, three_interpolate_grad, opt_n_threads,
    num_threads, query_and_group,
)

#!/usr/bin/env python
"""Golden fixture for the reduced-global -> full-local pose post-step, generated from the
UNMODIFIED reference in the build container:  ``python oracle/gen_golden_ik.py``  ->  tests/golden/ik_cases.pt.

It runs exactly what ``PoseNet3._reduced_glb_to_full_local_mat`` / ``_reduced_glb_6d_to_full_local_mat`` run
(net_aagc.py:788-800): scatter by ``joint_set.reduced`` (config.py:29), ``articulate.math.inverse_kinematics_R``
(articulate/math/spatial.py:197-221) and identity on ``joint_set.ignored`` (config.py:30), with the reference's own
functions.  The one thing the reference cannot supply here is the parent table: ``ParametricModel`` reads it from the
SMPL model file (``articulate/model.py:37``, ``kintree_table[0]``), which is not shipped (config.py:23).  The table used
is the published 24-joint SMPL kinematic tree; it is recorded in the fixture.
"""
import os
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")

import articulate as art          # noqa: E402
from config import joint_set      # noqa: E402

SMPL_PARENT = [None, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]


def ref_rot9(glb_reduced_pose):   # net_aagc.py:795-800 with global_to_local_pose = inverse_kinematics_R(., parent)
    g = torch.eye(3).repeat(glb_reduced_pose.shape[0], 24, 1, 1)
    g[:, joint_set.reduced] = glb_reduced_pose
    pose = art.math.inverse_kinematics_R(g, SMPL_PARENT).view(-1, 24, 3, 3)
    pose[:, joint_set.ignored] = torch.eye(3)
    return pose


def ref_rot6(glb_reduced_pose):   # net_aagc.py:788-793
    r = art.math.r6d_to_rotation_matrix(glb_reduced_pose).view(-1, joint_set.n_reduced, 3, 3)
    return ref_rot9(r)


def main():
    g = torch.Generator().manual_seed(2024)
    x9 = torch.randn(37, 15, 3, 3, generator=g)                       # raw network output viewed as 3x3 (not orthonormal)
    q, _ = torch.linalg.qr(torch.randn(11, 15, 3, 3, generator=g))   # proper rotations as well
    x9 = torch.cat((x9, q), 0)
    x6 = torch.randn(29, 15, 6, generator=g)
    out = {"parent": [-1] + SMPL_PARENT[1:], "reduced": list(joint_set.reduced), "ignored": list(joint_set.ignored),
           "x9": x9, "y9": ref_rot9(x9.clone()), "x6": x6, "y6": ref_rot6(x6.clone())}
    torch.save(out, os.path.join(ROOT, "tests", "golden", "ik_cases.pt"))
    print("wrote ik_cases.pt", out["y9"].shape, out["y6"].shape)


if __name__ == "__main__":
    main()

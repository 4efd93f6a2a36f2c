"""Seeded camera cases for the ray-generation step (shared by the golden generator and the tests)."""
import math


def _pose(yaw, pitch, t):
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    R = [[cy, sy * sp, sy * cp], [0.0, cp, -sp], [-sy, cy * sp, cy * cp]]
    return [R[0] + [t[0]], R[1] + [t[1]], R[2] + [t[2]]]


RAY_CASES = {
    # Technicolor-like forward-facing camera, NDC rays (datasets/technicolor.py:355-358)
    "ndc_73x41": dict(H=41, W=73, K=[[60.5, 0.0, 36.2], [0.0, 58.25, 20.1], [0.0, 0.0, 1.0]],
                      pose=_pose(0.07, -0.04, [0.12, -0.05, 0.3]), use_ndc=True, near=0.6, cam_idx=3.0, time=17.0 / 49.0),
    # world-space rays (DoNeRF-like)
    "world_50x37": dict(H=37, W=50, K=[[45.0, 0.0, 25.0], [0.0, 45.0, 18.5], [0.0, 0.0, 1.0]],
                        pose=_pose(-0.9, 0.35, [1.5, 0.4, -2.0]), use_ndc=False, near=0.5, cam_idx=0.0, time=0.0),
}

"""Geometry of the reference's analytic validation cases (validation/common_validation.py:41-121), restated:
axis-aligned rectangles in the xy / yz planes and a triangle-fan disc, each with a selectable normal direction."""
import math

import numpy as np


def _quad(name, corners, outward_positive):
    V = np.asarray(corners, dtype=np.float32)
    F = np.asarray([[0, 1, 2], [0, 2, 3]] if outward_positive else [[0, 2, 1], [0, 3, 2]], dtype=np.int32)
    return name, V, F


def rectangle_xy(name, width, depth, z, *, normal=1, center=(0.0, 0.0)):
    cx, cy = center
    x0, x1, y0, y1 = cx - width / 2.0, cx + width / 2.0, cy - depth / 2.0, cy + depth / 2.0
    return _quad(name, [[x0, y0, z], [x1, y0, z], [x1, y1, z], [x0, y1, z]], normal >= 0)


def rectangle_yz(name, length_y, height_z, x, *, normal=1, y_center=0.0, z_min=0.0):
    y0, y1, z0, z1 = y_center - length_y / 2.0, y_center + length_y / 2.0, z_min, z_min + height_z
    return _quad(name, [[x, y0, z0], [x, y1, z0], [x, y1, z1], [x, y0, z1]], normal >= 0)


def disk_xy(name, radius, z, *, segments=128, normal=1):
    ring = [[radius * math.cos(2.0 * math.pi * i / segments), radius * math.sin(2.0 * math.pi * i / segments), z] for i in range(segments)]
    V = np.asarray([[0.0, 0.0, z]] + ring, dtype=np.float32)
    F = np.asarray([[0, i + 1, 1 + (i + 1) % segments] if normal >= 0 else [0, 1 + (i + 1) % segments, i + 1] for i in range(segments)], np.int32)
    return name, V, F


def analytic_cases():
    """(key in shipped.json, meshes, samples, rays, emitter, receiver, closed form) for validation 01-05."""
    def squares(w):
        x = math.sqrt(1.0 + w * w)
        y = x * math.atan(w / x) - math.atan(w)
        return (math.log(x ** 4 / (1.0 + 2.0 * w * w)) + 4.0 * w * y) / (math.pi * w * w)

    def rectangles(x, y):
        x1, y1 = math.sqrt(1.0 + x * x), math.sqrt(1.0 + y * y)
        return (math.log((x1 * x1 * y1 * y1) / (x1 * x1 + y1 * y1 - 1.0)) + 2.0 * x * (y1 * math.atan(x / y1) - math.atan(x))
                + 2.0 * y * (x1 * math.atan(y / x1) - math.atan(y))) / (math.pi * x * y)

    def perpendicular(h):
        h1 = math.sqrt(1.0 + h * h)
        h2 = h1 ** 4 / (h * h * (2.0 + h * h))
        return 0.25 + (h * math.atan(1.0 / h) - h1 * math.atan(1.0 / h1) - 0.25 * math.log(h2)) / math.pi

    return [
        ("01_parallel_equal_square", [rectangle_xy("plate_1", 1, 1, 0.0, normal=+1), rectangle_xy("plate_2", 1, 1, 1.0, normal=-1)],
         32, 1024, "plate_1", "plate_2", squares(1.0)),
        ("02_parallel_equal_rectangle", [rectangle_xy("plate_1", 2, 1, 0.0, normal=+1), rectangle_xy("plate_2", 2, 1, 1.0, normal=-1)],
         16, 512, "plate_1", "plate_2", rectangles(2.0, 1.0)),
        ("03_equal_coaxial_discs", [disk_xy("disc_1", 1.0, 0.0, segments=256, normal=+1), disk_xy("disc_2", 1.0, 1.0, segments=256, normal=-1)],
         16, 512, "disc_1", "disc_2", 1.0 + (1.0 - math.sqrt(5.0)) / 2.0),
        ("04_patch_to_disc", [rectangle_xy("patch", 0.04, 0.04, 0.0, normal=+1), disk_xy("disc", 1.0, 1.0, segments=256, normal=-1)],
         8, 1024, "patch", "disc", 0.5),
        ("05_perpendicular_square_rectangle", [rectangle_xy("square", 1, 1, 0.0, normal=+1, center=(0.5, 0.0)),
                                                rectangle_yz("adjacent_rectangle", 1, 1, 0.0, normal=+1, y_center=0.0, z_min=0.0)],
         32, 512, "square", "adjacent_rectangle", perpendicular(1.0)),
    ]

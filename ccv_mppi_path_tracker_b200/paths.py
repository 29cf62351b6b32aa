"""Reference-path sources feeding pathCallback (src/diff_drive_mppi.cpp:48-52), ROS-free.

  sin_path     src/reference_path_creator.cpp:37-56   (accumulated `s += resolution` loop: course_length=10 -> 101 points)
  circle_path  src/reference_path_creator.cpp:57-68
  dkan_path    src/dkan_path_creator.cpp:11-52        (L-shaped corridor)
  load_csv / save_csv   rows of `x,y,` as in data/data.csv:1 (trailing comma)
"""
import math

import numpy as np


def sin_path(course_length=10.0, resolution=0.1, A1=0.0, omega1=0.0, delta1=1.57, A2=0.0, omega2=0.0, delta2=1.57,
             A3=0.0, omega3=0.0, delta3=1.57, init_x=0.0, init_y=0.0):
    pts = []
    s = 0.0
    while s < course_length:
        x = init_x + s
        y = (A1 * math.cos(2 * math.pi * omega1 * s + delta1) + A2 * math.cos(2 * math.pi * omega2 * s + delta2)
             + A3 * math.cos(2 * math.pi * omega3 * s + delta3) + init_y)
        y -= A1 + A2 + A3
        pts.append((x, y))
        s += resolution
    return np.asarray(pts, dtype=np.float64).reshape(-1, 2)


def circle_path(R=10.0, resolution=0.1, init_x=0.0, init_y=0.0):
    pts = []
    s = 0.0
    step = resolution / 2 * math.pi * R  # reference_path_creator.cpp:58 (sic)
    while s <= 200 * math.pi:
        pts.append((init_x + R * math.cos(s), init_y + R * math.sin(s) + R))
        s += step
    return np.asarray(pts, dtype=np.float64).reshape(-1, 2)


def dkan_path(resolution=0.1):
    """(0,0)->(17.7,0)->(17.7,8)->(0,8) at `resolution` spacing (src/dkan_path_creator.cpp:11-52)."""
    pts = []
    x = 0.0
    while x < 17.7:
        pts.append((x, 0.0))
        x += resolution
    y = 0.0
    while y < 8.0:
        pts.append((17.7, y))
        y += resolution
    x = 17.7
    while x > 0.0:
        pts.append((x, 8.0))
        x -= resolution
    return np.asarray(pts, dtype=np.float64).reshape(-1, 2)


def load_csv(path):
    rows = []
    with open(path) as f:
        for line in f:
            parts = [q for q in line.strip().split(",") if q != ""]
            if len(parts) >= 2:
                rows.append((float(parts[0]), float(parts[1])))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 2)


def save_csv(path, xy):
    with open(path, "w") as f:
        for x, y in np.asarray(xy, dtype=np.float64).reshape(-1, 2):
            f.write(f"{x:.6g},{y:.6g},\n")

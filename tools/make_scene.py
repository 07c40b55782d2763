#!/usr/bin/env python
"""Writes scene description files for rays1_b200 --scene-file (sphere-count sweeps without a
recompile).  `--preset large` reproduces create_large_scene() (src/latest/rayweek1.cpp:654-719 of the reference) through
the file format; `--grid W H` builds the same recipe on a W x H grid (66 x 62 = the synthetic 4096-sphere scene).

  python tools/make_scene.py --grid 30 16 --out large.r1scene
  ./rays1bench_b200/rays1_b200 --scene-file large.r1scene --spp 64
"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fma32(a, b, c):
    """float32 fma: the product of two float32 values is exact in float64, one rounding at the end"""
    return np.float32(np.float64(np.float32(a)) * np.float64(np.float32(b)) + np.float64(np.float32(c)))


def grid_scene(gw, gh, ior_mod, lookfrom, focus):
    """The reference's large-scene recipe; colours from glibc srand(111)/rand() like the C++ builders."""
    libc = ctypes.CDLL(None)
    libc.srand(111)
    f = np.float32
    spheres = []
    for y in range(gh):
        for x in range(gw):
            px, py, pz = f(x - gw // 2) * f(1.1), f(0), f(y - gh // 2) * f(1.1)
            # the constant expressions as the reference's fast-math build evaluates them (rays1_host.cpp:grid_scene)
            r, g, b = (f(libc.rand() & 0xff) * (f(1.0) / f(255.0)) for _ in range(3))
            i = x + y * gw
            if i % 20 == 0:
                k = i % ior_mod if ior_mod else i
                spheres.append((px, py, pz, f(0.45), 2, 1, 1, 1, fma32(k, 0.05, 1.2)))
            elif i % 10 == 0:
                spheres.append((px, py + f(0.1), pz, f(0.45), 1, r, g, b, fma32(f(0.5) * f(y), f(1.0) / f(gh), 0.01)))
            else:
                spheres.append((px, py, pz, f(0.45), 0, r, g, b, 0))
    spheres.append((0, f(-1000.5), 0, 1000, 0, 0.5, 0.5, 0.5, 0))
    spheres.append((5, 3, 0, 2, 1, 0.5, 0.5, 0.8, 0.65))
    spheres.append((0, 3, 0, 2, 2, 1, 1, 1, 1.5))
    spheres.append((-5, 3, 0, 2, 1, 0.8, 0.2, 0.2, 0.05))
    camera = (*lookfrom, 0, 0, 0, 60, 0.1, focus)
    return camera, spheres


def main():
    import rays1bench_b200 as r1
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", choices=["large", "synth4096"])
    ap.add_argument("--grid", type=int, nargs=2, metavar=("W", "H"))
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    if args.preset == "large" or (not args.preset and not args.grid):
        cam, sph = grid_scene(30, 16, 0, (3, 8, 15), 10.0)
    elif args.preset == "synth4096":
        cam, sph = grid_scene(66, 62, 480, (6, 16, 30), 20.0)
    else:
        gw, gh = args.grid
        scale = max(gw / 30.0, gh / 16.0)
        cam, sph = grid_scene(gw, gh, 480, (3 * scale, 8 * scale, 15 * scale), 10.0 * scale)
    r1.write_scene_file(args.out, cam, sph)
    print("%s: %d spheres" % (args.out, len(sph)))


if __name__ == "__main__":
    main()

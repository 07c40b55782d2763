#!/usr/bin/env python
"""Record the golden fixtures under tests/golden/ from the reference's own compiled code.

Runs only where /root/reference exists (the build container): it needs oracle/_ref/libref_rays1.so, i.e.
oracle/ref_harness.cpp linked against the unmodified /root/reference/src/latest sources (oracle/Makefile).
Everything it writes is small and committed, because /root/reference does not travel to the GPU box.

  tests/golden/rays_<scene>.npz    ray segments recorded from real paths walked with the reference's camera,
                                   Hitable::hit and Material::scatter (+ the random inputs each scatter consumed),
                                   plus hand-made edge rays answered by Hitable::hit, plus (large, synth4096) rays leaving
                                   the ior >= 5 dielectric spheres from inside, answered by hit + scatter
  <scene> = small | medium | large (unmodified src/latest) | synth4096 (BASELINE.json config 5: the same sources with
  MAX_SPHERES = 4096, oracle/_ref/libref_rays1_4096.so)
  tests/golden/replay_<scene>.npz  per-pixel replay: generator states before each of 4096 pixels and the float colour the
                                   reference's own color() returned for it (spp = 1)
  tests/golden/render_<scene>.npz  reference render (its own TileRenderScheduler + render_tile) at 320x180 and
                                   REF_SPP samples per pixel, with the ray count
  tests/golden/ref_stats.json      rays-per-sample of the reference at the default 1280x720x250 workload
                                   (three runs of the unmodified executable) and of the golden renders

usage: python oracle/make_golden.py [--spp 16384] [--skip-render] [--skip-exe]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cpu_checkers import REF_EXE, RefLib  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_SEGMENTS = 3072
N_REPLAY = 4096
RENDER_W, RENDER_H = 320, 180


def unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True).astype(np.float32)).astype(np.float32)


def edge_rays(ref, scene, name):
    """Hand-made rays for the cases real paths rarely produce."""
    rng = np.random.default_rng(20260 + len(name))
    soa = ref.scene_soa(scene)
    cam = ref.scene_camera(scene)
    org, dr = [], []

    def add(o, d):
        org.append(np.asarray(o, np.float32))
        dr.append(unit(d))

    n = len(soa["cx"])
    centers = np.stack([soa["cx"], soa["cy"], soa["cz"]], 1)
    real = np.where(soa["inv_radius"] > 0)[0]
    radius = np.zeros(n, np.float32)
    radius[real] = 1.0 / soa["inv_radius"][real]
    # (a) from inside every dielectric sphere (and the small scene's hollow shell), random directions
    for i in np.where(soa["kind"] == 2)[0][:12]:
        for _ in range(8):
            add(centers[i] + rng.normal(size=3) * 0.1 * abs(radius[i] if radius[i] else 0.4), rng.normal(size=3))
    # (b) aimed straight at a placeholder sphere (radius 0 at 999999999) and at every 7th real centre from the camera
    for i in np.where(soa["inv_radius"] == 0)[0][:4]:
        add(cam[0:3], centers[i] - cam[0:3])
        add([0, 1, 0], [1, 1, 1])
    for i in real[::7][:64]:
        add(cam[0:3], centers[i] - cam[0:3])
    # (c) skimming a row of spheres at centre height (many discriminant-positive spheres per ray)
    for z in np.unique(centers[real, 2])[:24]:
        add([-40.0, float(np.median(centers[real, 1])), float(z)], [1, 0, 0])
        add([40.0, float(np.median(centers[real, 1])) + 0.2, float(z) + 0.1], [-1, 0.001, 0])
    # (d) starting ON a sphere surface (self-hit at t ~ 0 must be rejected by t > t_min = 0.001)
    for i in real[::5][:96]:
        nrm = unit(rng.normal(size=3))
        add(centers[i] + nrm * radius[i], unit(rng.normal(size=3)) + nrm * 0.5)
        add(centers[i] + nrm * radius[i], -nrm)  # straight through the sphere: far root
    # (e) grazing the largest sphere (the ground) at ever smaller angles, and straight down from above
    g = real[np.argmax(radius[real])]
    top = centers[g] + np.array([0, radius[g], 0], np.float32)
    for k in range(24):
        add(top + np.array([-30.0, 0.5, 0.3 * k], np.float32), [1, -0.5 ** (k * 0.5 + 1), 0])
    for k in range(16):
        add([rng.uniform(-10, 10), 50.0, rng.uniform(-10, 10)], [0, -1, 0])
    # (f) pointing away from everything
    for k in range(16):
        add(cam[0:3], [rng.normal(), abs(rng.normal()) + 2, rng.normal()])
    return np.stack(org).astype(np.float32), np.stack(dr).astype(np.float32)


def dielectric_exit_rays(ref, scene, min_ior=5.0, per_sphere=16):
    """Rays that start INSIDE the high-index dielectric spheres and leave them: Dielectric::scatter's refraction branch with
    ni_over_nt = ior, where 1 - k^2 (1 - dt^2) amplifies float32 rounding by k^2 (ior up to 24.2 in the large scene,
    rayweek1.cpp:692).  Real camera paths reach these rarely (6 of 3072 recorded segments)."""
    rng = np.random.default_rng(4242)
    soa = ref.scene_soa(scene)
    sel = np.where((soa["kind"] == 2) & (soa["param"] >= min_ior) & (soa["inv_radius"] > 0))[0]
    org, dr = [], []
    for i in sel:
        c = np.array([soa["cx"][i], soa["cy"][i], soa["cz"][i]], np.float32)
        rad = 1.0 / soa["inv_radius"][i]
        for k in range(per_sphere):
            # half of the rays close to the axis (refraction happens: sin(theta) < 1 / ior), half anywhere inside (mostly
            # total internal reflection, the discriminant <= 0 branch)
            off = unit(rng.normal(size=3)) * rad * (rng.uniform(0, 0.9 / soa["param"][i]) if k % 2 == 0 else rng.uniform(0, 0.95))
            org.append(c + off.astype(np.float32))
            dr.append(unit(rng.normal(size=3)))
    return np.stack(org).astype(np.float32), np.stack(dr).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=16384)
    ap.add_argument("--skip-render", action="store_true")
    ap.add_argument("--skip-exe", action="store_true")
    ap.add_argument("--scenes", default="small,medium,large,synth4096")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    ref1024, ref4096 = RefLib(), RefLib(4096)
    stats_path = os.path.join(GOLD, "ref_stats.json")
    stats = json.load(open(stats_path)) if os.path.exists(stats_path) else {}
    stats["source"] = "montib/rays1bench src/latest, g++ 13.3 -pthread -ffast-math -O3 -march=x86-64-v3 (oracle/Makefile)"

    stats["source_synth4096"] = ("same sources and flags with MAX_SPHERES = 4096 (rayweek1.cpp:174 patched in a temporary copy, "
                                "oracle/Makefile) + the 4096-sphere builder of oracle/ref_harness.cpp on the reference's classes")
    for name in args.scenes.split(","):
        ref = ref4096 if name == "synth4096" else ref1024
        s = ref.scene_create(name)
        rec = ref.record_paths(s, N_SEGMENTS, seed=3)
        eo, ed = edge_rays(ref, s, name)
        ei, et, ep, en = ref.hit(s, eo, ed)
        soa = ref.scene_soa(s)
        out = {("seg_" + k): v for k, v in rec.items()}
        out.update(edge_org=eo, edge_dir=ed, edge_index=ei, edge_t=et, edge_p=ep, edge_normal=en,
                   camera=ref.scene_camera(s), **{("soa_" + k): v for k, v in soa.items()})
        if name in ("large", "synth4096"):
            do, dd = dielectric_exit_rays(ref, s)
            rec_d = ref.hit_scatter(s, do, dd)
            out.update({("diel_" + k): v for k, v in rec_d.items()})
            ex = (rec_d["index"] >= 0) & (soa["kind"][np.maximum(rec_d["index"], 0)] == 2) & ((rec_d["dir"] * rec_d["normal"]).sum(1) > 0)
            print("  dielectric exits: %d rays, %d leave an ior >= 5 sphere, %d of them refract" %
                  (len(do), int(ex.sum()), int((ex & (np.abs((rec_d["scat_dir"] * rec_d["normal"]).sum(1)) > 0) &
                                                 ((rec_d["scat_dir"] * rec_d["normal"]).sum(1) > 0)).sum())))
        np.savez_compressed(os.path.join(GOLD, "rays_%s.npz" % name), **out)
        print(name, "segments", len(rec["t"]), "hits", int((rec["index"] >= 0).sum()), "edge rays", len(et),
              "edge hits", int((ei >= 0).sum()))
        # per-pixel replay fixtures: generator states before each pixel + the reference's own colour sums (spp = 1: every
        # pixel is one independent sample, so a flipped float decision cannot leak into other pixels)
        prng = np.random.default_rng(77 + len(name))
        xy = np.stack([prng.integers(0, 1280, N_REPLAY), prng.integers(0, 720, N_REPLAY)], 1).astype(np.int32)
        st, st4, col, rays_px = ref.replay_pixels(s, xy, 1280, 720, 1)
        np.savez_compressed(os.path.join(GOLD, "replay_%s.npz" % name), xy=xy, state=st, state4=st4, color=col, rays=rays_px)
        print("  replay: %d pixels, %.3f rays/sample" % (N_REPLAY, rays_px.mean()))
        if not args.skip_render:
            rgb, rays, el = ref.render(s, RENDER_W, RENDER_H, args.spp)
            np.savez_compressed(os.path.join(GOLD, "render_%s.npz" % name), rgb=rgb, spp=args.spp, num_rays=rays)
            stats.setdefault("render", {})[name] = dict(w=RENDER_W, h=RENDER_H, spp=args.spp, num_rays=rays,
                                                        rays_per_sample=rays / (RENDER_W * RENDER_H * args.spp))
            print("  render %dx%dx%d: %d rays, %.3f rays/sample, %.1f s" % (RENDER_W, RENDER_H, args.spp, rays,
                                                                          stats["render"][name]["rays_per_sample"], el))
        ref.scene_destroy(s)

    if "synth4096" in args.scenes.split(","):
        # no executable holds the 4096-sphere scene: rays per sample at 1280x720 from the patched library (16 spp = 14.7 M samples)
        s = ref4096.scene_create("synth4096")
        _, rays, el = ref4096.render(s, 1280, 720, 16)
        ref4096.scene_destroy(s)
        stats.setdefault("default_workload", {})["synth4096"] = dict(w=1280, h=720, spp=16, num_rays=[rays], rays_per_sample=rays / (1280 * 720 * 16),
                                                                     note="patched library render at 16 spp (rays per sample does not depend on spp)")
        print("synth4096 1280x720x16: %.4f rays/sample, %.2f Mrays/s here" % (rays / (1280 * 720 * 16), rays / el / 1e6))
    if not args.skip_exe:
        # the unmodified executable at its compiled-in workload (1280x720x250, common.h:19-25), three runs
        with tempfile.TemporaryDirectory() as tmp:
            txt = subprocess.run([REF_EXE, "-n", "3"], cwd=tmp, check=True, capture_output=True, text=True).stdout
        runs = {}
        cur = None
        for line in txt.splitlines():
            if line.strip() in ("small", "medium", "large"):
                cur = line.strip()
            m = re.match(r"total rays:\s+(\d+)", line)
            if m and cur:
                runs.setdefault(cur, []).append(int(m.group(1)))
        stats.setdefault("default_workload", {}).update({k: dict(w=1280, h=720, spp=250, num_rays=v,
                                                                 rays_per_sample=float(np.mean(v)) / (1280 * 720 * 250))
                                                         for k, v in runs.items()})
        print(stats["default_workload"])
    json.dump(stats, open(stats_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

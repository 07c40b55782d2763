#!/usr/bin/env python
"""The reference's bench.py driver (bench.py:62-211 of montib/rays1bench) for the B200 build: builds the executable with
nvcc, runs it with the reference's flags and leaves out_<scene>.txt / out_<scene>.tga where update_readme.py looks for
them ("one more step directory").

  python tools/bench_steps.py --latest [--num N] [--save] [--quick] [--gpus G] [--compile-only] [--outdir DIR]

--quick mirrors -DQUICKBENCH (common.h:3-16: 80 x 60, 100 spp); --num / --save are the reference's -n / -w.
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latest", action="store_true", help="accepted for compatibility: there is one B200 build")
    ap.add_argument("--num", type=int, default=1)
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--compile-only", action="store_true")
    ap.add_argument("--outdir", default=os.path.join(ROOT, "rays1bench_b200"))
    args = ap.parse_args()
    # rays1bench_b200/build.py is loaded by path: importing the package needs the library this step builds
    import importlib.util
    spec = importlib.util.spec_from_file_location("rays1bench_b200_build", os.path.join(ROOT, "rays1bench_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib, exe = b.build()
    print("compiled", exe)
    if args.compile_only:
        return 0
    cmd = [exe, "--gpus", str(args.gpus)]
    if args.save:
        cmd.append("-w")
    if args.num > 1:
        cmd += ["-n", str(args.num)]
    if args.quick:
        cmd += ["--width", "80", "--height", "60", "--spp", "100"]
    os.makedirs(args.outdir, exist_ok=True)
    print("RUN " + " ".join(cmd) + "\n")
    rc = subprocess.call(cmd, cwd=args.outdir)
    for scene in ("small", "medium", "large"):
        path = os.path.join(args.outdir, "out_%s.txt" % scene)
        if os.path.exists(path):
            print(scene, open(path).read())
    return rc


if __name__ == "__main__":
    sys.exit(main())

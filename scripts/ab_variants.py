"""A/B of library variants: runs a quick script once per library in variants/ (and the default).
usage: python scripts/ab_variants.py scripts/bench_c4.py [reps]"""
import glob, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
script = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
libs = [None] + sorted(glob.glob(os.path.join(root, "variants", "lib_*.so")))
for rep in range(reps):
    for lib in libs:
        env = dict(os.environ)
        if lib: env["CYRK_B200_LIB"] = lib
        print("==", os.path.basename(lib) if lib else "default", flush=True)
        subprocess.run([sys.executable, os.path.join(root, script)], env=env, check=False)

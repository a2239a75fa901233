"""A/B probe: C2 stage timings for every libfluidmarch.so under build_variants/ (each in its own process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
for name in sorted(os.listdir(os.path.join(ROOT, "build_variants"))):
    lib = os.path.join(ROOT, "build_variants", name, "libfluidmarch.so")
    if not os.path.exists(lib):
        continue
    env = dict(os.environ, FLUIDMARCH_LIB=lib, FLUIDMARCH_AB="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_step.py"), cfg, "6"], env=env, capture_output=True, text=True)
    print(name, out.stdout.strip().split("{'pixels'")[0][:400], out.stderr[-400:], flush=True)

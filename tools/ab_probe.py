"""A/B probe: C2 stage timings for every libfluidmarch.so under build_variants/ (each in its own process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
for name in sorted(os.listdir(os.path.join(ROOT, "build_variants"))):
    lib = os.path.join(ROOT, "build_variants", name, "libfluidmarch.so")
    if not os.path.exists(lib):
        continue
    env = dict(os.environ, FLUIDMARCH_LIB=lib, FLUIDMARCH_AB="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_step.py"), cfg, "12"], env=env, capture_output=True, text=True)
    txt = out.stdout.strip()
    tail = txt.split("'first_candidates'")[-1][:120] if "'first_candidates'" in txt else ""
    print(name, txt.split("{'pixels'")[0][:400], tail, out.stderr[-400:], flush=True)

"""Golden particle files, written and read back BY THE REFERENCE'S VENDORED PARTIO (oracle/_ref/libfluidref.so).

Run in the build container (needs /root/reference at build time of oracle/_ref):
    python tests/golden/make_golden_bgeo.py
Outputs (committed, a few KB each) in tests/golden/bgeo/:
  plain.bgeo            positions only                      (point block at byte 41: offset % 4 == 1)
  vel_id_dens.bgeo      + velocity (vector), id (int), density (float)   (offset % 4 == 2)
  vel_id.bgeo           + velocity, id                                    (offset % 4 == 3)
  all.bgeo              + an indexed-string attribute                     (offset % 4 == 0)
  all_gz.bgeo           the same, gzip'd by partio
  extreme.bgeo          16 particles with huge / infinite / NaN coordinates (decoders only)
  ParticleData_Fluid_{1,2,3}.bgeo   a three-frame sequence named like the reference's datasets (main.cpp:48-53)
  expected.npz          positions partio reads from each file (bit-exact float32), and the partio sample files'
                        known answers (vendor/partio/misc/data: particle count and first position)
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib  # noqa: E402

scenes = importlib.import_module("bachelor-thesis_b200.scenes")
OUT = os.path.join(HERE, "bgeo")


def main():
    ref = oracle_lib.Ref()
    os.makedirs(OUT, exist_ok=True)
    xyz = scenes.dam_break(320, t=0.4)
    xyz[0] = (0.0, -0.0, 1e-38)                       # zero, negative zero, a denormal: bit patterns must survive
    files = {"plain": (0, False), "vel_id_dens": (7, False), "vel_id": (3, False), "all": (15, False), "all_gz": (15, True)}
    expected = {}
    for name, (extras, comp) in files.items():
        path = os.path.join(OUT, name + ".bgeo")
        ref.partio_write(path, xyz, extras, comp)
        expected[name] = ref.partio_read(path)
        assert np.array_equal(expected[name].view(np.uint32), xyz.view(np.uint32))
    # values no frame can be built from, for the decoders alone
    ext = xyz[:16].copy()
    ext[1] = (-123456.789, 3.4e38, -1.17549435e-38)
    ext[2] = (np.inf, -np.inf, np.nan)
    path = os.path.join(OUT, "extreme.bgeo")
    ref.partio_write(path, ext, 2, False)
    expected["extreme"] = ref.partio_read(path)
    assert np.array_equal(expected["extreme"].view(np.uint32), ext.view(np.uint32))
    for i in (1, 2, 3):
        f = scenes.dam_break(900 + 60 * i, t=0.3 + 0.1 * i)
        path = os.path.join(OUT, f"ParticleData_Fluid_{i}.bgeo")
        ref.partio_write(path, f, 1 if i == 2 else 0, i == 3)
        expected[f"seq{i}"] = ref.partio_read(path)
    sample_dir = "/root/reference/vendor/partio/misc/data"
    for s in ("test", "base", "scatter", "reindeer"):
        p = ref.partio_read(os.path.join(sample_dir, s + ".bgeo"))
        expected["sample_" + s + "_n"] = np.array([len(p)], np.int64)
        expected["sample_" + s + "_p0"] = p[0]
    np.savez(os.path.join(OUT, "expected.npz"), **expected)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()

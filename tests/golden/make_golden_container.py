"""Generates tests/golden/golden_container.npz from the .astc fixtures the reference's own tests use
(Test/Data/*.astc, Test/Data/Tiles/*.astc; astc_test_functional.py:2195-2262 runs the negative_* files through -dl and
expects a refusal). The files are 31-80 bytes of test DATA; they are stored as byte arrays keyed by file name.
    python tests/golden/make_golden_container.py
"""
import glob, os
import numpy as np

REF = "/root/reference/Test/Data"
out = {}
for f in sorted(glob.glob(REF + "/*.astc") + glob.glob(REF + "/Tiles/*.astc")):
    out[os.path.basename(f)] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
    print(os.path.basename(f), len(out[os.path.basename(f)]))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_container.npz"), **out)

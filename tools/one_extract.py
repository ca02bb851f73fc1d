import sys, importlib
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from synth import synth_image
pkg = importlib.import_module("self_commit_orb-slam2_b200")
img = synth_image(640, 480, 1)
ex = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
kps, desc = ex(img)
print("ok", len(kps))

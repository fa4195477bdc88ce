"""Small end-to-end invocation for compute-sanitizer (racecheck / initcheck / synccheck are slow: one image, few features)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import multicol_slam_b200.api as api
from multicol_slam_b200 import synth
import oracle_api as oa
cams = synth.lafida_cams()
cam = synth.scaled_cam(cams[0], 333, 211)
img, mask = synth.frame(cam, 5), synth.mirror_mask(cam)
ex = api.mdBRIEFextractorOct(nfeatures=300, nlevels=4, do_dBrief=True, learnMasks=True)
k, d, m = ex(img, mask, cam)
ok, od, om = oa.OracleExtractor(nfeatures=300, nlevels=4, do_dbrief=True, learn_masks=True).extract(img, mask, cam)
assert k.tobytes() == ok.tobytes() and np.array_equal(d, od) and np.array_equal(m, om)
idx, dist = api.hamming_topk(d, d[::-1].copy(), 2, m, m[::-1].copy())
F = api.Frame.from_cameras([(k, d, m)], [(333, 211)], [ex.info.scale_factor[l] for l in range(4)])
prev = np.stack([k["x"], k["y"]], 1).astype(np.float64)
n, m12 = api.cORBmatcher(0.9, False, 32, True).SearchForInitialization(F, F, prev, 30)
print("sanitize_small ok", len(k), n)

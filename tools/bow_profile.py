"""Small driver for profiling the bag-of-words kernels (ncu -k regex:"bow_descend|group_distance"): transform the
descriptors of two synthetic 3-camera frames and run the feature-vector guided SearchByBoW between them.  Prints timings."""
import pathlib, sys, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import multicol_slam_b200.api as api
from multicol_slam_b200 import synth

cams = synth.lafida_cams()
voc = api.ORBVocabulary(np.load(ROOT / "tests" / "golden" / "voc_small_9_6.npz"))
ex = api.mdBRIEFextractorOct(nfeatures=2000, do_dBrief=True, learnMasks=True)
imgs = [synth.texture_stream(cams[c], 2, seed=4 + c) for c in range(3)]
D, M = [], []
for f in range(2):
    ds, ms = [], []
    for c in range(3):
        _, d, m = ex(imgs[c][f], synth.mirror_mask(cams[c]), cams[c])
        ds.append(d); ms.append(m)
    D.append(np.concatenate(ds)); M.append(np.concatenate(ms))
mt = api.cORBmatcher(0.9, False, 32, True)
for rep in range(3):
    t0 = time.perf_counter(); a = voc.transform(D[0]); t1 = time.perf_counter(); b = voc.transform(D[1])
    t2 = time.perf_counter(); n, _ = mt.SearchByBoWFrame(D[0], a[2], D[1], b[2], M[0], M[1]); t3 = time.perf_counter()
print("descriptors %d + %d   transform %.3f ms   SearchByBoW(KF,F) %.3f ms   matches %d   words %d" %
      (len(D[0]), len(D[1]), (t1 - t0) * 1e3, (t3 - t2) * 1e3, n, len(a[0])))

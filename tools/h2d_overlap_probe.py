"""Does a concurrent pinned H2D copy slow the extraction kernels down?  (e2e pipeline diagnosis)
Times K1/K2/K3 of a 96-image device-resident batch with and without a 139 MB pinned H2D copy in flight on another
stream, and with a device-to-device copy of the same size for comparison."""
import sys, pathlib, time
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import multicol_slam_b200.api as api
from multicol_slam_b200 import synth

dev = torch.device("cuda:0")
cams = synth.lafida_cams()
masks = np.stack([synth.mirror_mask(c) for c in cams])
B = 96
img = np.stack([synth.frame(cams[i % 3], 100 + i) for i in range(6)])
imgs = torch.from_numpy(np.tile(img, (B // 6, 1, 1))).to(dev)
pitched = torch.zeros((B, 480, 768), dtype=torch.uint8, device=dev); pitched[:, :, :754] = imgs
ex = api.mdBRIEFextractorOct(nfeatures=2000, do_dBrief=True, learnMasks=True)
ex.set_profiling(True)
coi = [i % 3 for i in range(B)]
st = torch.cuda.Stream(dev); sc = torch.cuda.Stream(dev)
big_h = torch.empty(139_000_000, dtype=torch.uint8).pin_memory()
big_d = torch.empty(139_000_000, dtype=torch.uint8, device=dev)
big_d2 = torch.empty_like(big_d)
out = None
for mode in ("none", "h2d", "d2h", "d2d", "none", "h2d"):
    res = []
    for it in range(6):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sc):
            if mode == "h2d":
                for _ in range(2): big_d.copy_(big_h, non_blocking=True)
            elif mode == "d2h":
                for _ in range(2): big_h.copy_(big_d, non_blocking=True)
            elif mode == "d2d":
                for _ in range(20): big_d2.copy_(big_d, non_blocking=True)
        with torch.cuda.stream(st):
            e0.record(st)
            out = ex.extract_batch_device(pitched, masks, cams, coi, out=out, stream=st, width=754)
            e1.record(st)
        torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1),) + ex.get_timings())
    r = np.median(np.array(res[1:]), axis=0)
    print("mode %-5s total %.3f ms  K1 %.3f  K2 %.3f  K3 %.3f" % (mode, r[0], r[1], r[2], r[3]), flush=True)

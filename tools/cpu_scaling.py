import sys, time, os
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo") else ".")
import numpy as np, bench
from multicol_slam_b200 import synth
cams = synth.lafida_cams(); masks = np.stack([synth.mirror_mask(c) for c in cams])
images = bench.make_stream(cams, 43, 1000)
bench.cpu_oracle_run(cams, masks, images, os.cpu_count())
for nt in (1, 8, 16, 32, 64, 128):
    n, dt = bench.cpu_oracle_run(cams, masks, images[: max(2, min(43, nt))], nt)
    print(f"threads {nt:3d}: {n} features in {dt:.2f} s -> {n/dt/1e6:.4f} Mfeat/s, {n/dt/1e6/nt*1e3:.3f} kfeat/s/thread")
print(open("/sys/fs/cgroup/cpu.max").read() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max")

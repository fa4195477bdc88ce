"""profiles/k1_traffic.json from an ncu --set full report of one bench step: DRAM read + write bytes of the 8 pyr_fast_kernel
launches, per camera-frame.   python tools/make_k1_traffic.py gpurun_out/r2_step.ncu-rep <images in the step> > profiles/k1_traffic.json"""
import csv, io, json, subprocess, sys
rep, images = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, U = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n = 0.0, 0
for r in rows[2:]:
    if "pyr_fast_kernel" not in r[H.index("Kernel Name")]:
        continue
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = H.index(m)
        tot += float(r[i]) * scale.get(U[i], 1)
    n += 1
print(json.dumps({"kernel": "pyr_fast_kernel", "launches": n, "images": images, "dram_bytes_all_launches": tot,
                  "dram_bytes_per_image": tot / images,
                  "source": f"{rep.split('/')[-1]} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of the {n} level "
                            f"launches of one bench step, {images} images)"}, indent=1))

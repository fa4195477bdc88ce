"""Convert the Lafida interior-orientation YAMLs of the reference (data fixtures,
/root/reference/Examples/Lafida/InteriorOrientationFisheye{0,1,2}.yaml) into
multicol_slam_b200/data/lafida_cams.json.  Run in the build container only."""
import json, re, pathlib
root = pathlib.Path(__file__).resolve().parents[1]
cams = []
for c in range(3):
    txt = pathlib.Path(f"/root/reference/Examples/Lafida/InteriorOrientationFisheye{c}.yaml").read_text(encoding="latin-1")
    kv = {m.group(1): float(m.group(2)) for m in re.finditer(r"^Camera\.(\w+):\s*([-+0-9.eE]+)", txt, re.M)}
    nrpol, nrinv = int(kv["nrpol"]), int(kv["nrinvpol"])
    cams.append(dict(c=kv["c"], d=kv["d"], e=kv["e"], u0=kv["u0"], v0=kv["v0"],
                     pol=[kv[f"a{i}"] for i in range(nrpol)], inv_pol=[kv[f"pol{i}"] for i in range(nrinv)],
                     width=int(kv["Iw"]), height=int(kv["Ih"]), mirror_mask=int(kv["mirrorMask"])))
(root / "multicol_slam_b200/data/lafida_cams.json").write_text(json.dumps(cams, indent=1))
print(cams[0])

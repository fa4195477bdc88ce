"""The C++ mirror of the reference interface (include/mcs_shim.hpp): compiles against the C ABI on CPU; on the GPU
the compiled program must reproduce the oracle's descriptors through mdBRIEFextractorOct::operator()."""
import pathlib
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
EXE = ROOT / "tests" / "cpp" / "shim_smoke"


def build():
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(EXE), str(ROOT / "tests/cpp/shim_smoke.cpp"),
                           "-L" + str(ROOT / "multicol_slam_b200"), "-lmcs_b200", "-Wl,-rpath," + str(ROOT / "multicol_slam_b200")])


def test_shim_compiles_and_links(api):
    build()
    assert EXE.exists()


@pytest.mark.gpu
def test_shim_extract_and_match(api, oa, cams, tmp_path):
    from multicol_slam_b200 import synth
    build()
    cam = cams[0]
    img, mask = synth.frame(cam, 4), synth.mirror_mask(cam)
    (tmp_path / "i.raw").write_bytes(img.tobytes())
    (tmp_path / "m.raw").write_bytes(mask.tobytes())
    out = subprocess.run([str(EXE), str(tmp_path / "i.raw"), str(tmp_path / "m.raw"), "754", "480"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    kv = dict(t.split("=") for t in out.stdout.split()[1:])
    ok, od, om = oa.OracleExtractor(nfeatures=1000, do_dbrief=True, learn_masks=True).extract(img, mask, cam)
    h = 0
    for a, b in zip(od.reshape(-1), om.reshape(-1)):
        h = (h * 1315423911 + int(a) + 7 * int(b)) & 0xFFFFFFFFFFFFFFFF
    assert int(kv["nkp"]) == len(ok) and int(kv["hash"]) == h
    assert kv["same_mask"] == "1" and kv["untouched"] == "1" and kv["levels"] == "8" and kv["ds"] == "32" and kv["d01"] == "1"
    n, m12 = oa.match_bruteforce(od, od ^ np.eye(32, dtype=np.uint8)[np.arange(len(od)) % 32], 32, 0.9, om, om)
    assert int(kv["matches"]) == n and int(kv["self"]) == int((m12 == np.arange(len(od))).sum())

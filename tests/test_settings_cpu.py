"""Configuration loaders (multicol_slam_b200/settings.py) on the values of the reference's Lafida configuration
(tests/golden/lafida_config.json holds the key/value content of Examples/Lafida/*.yaml; the files are re-written here in the
flat cv::FileStorage YAML layout, with a header line and comments, so that the parser is exercised too).  CPU only."""
import json
import pathlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def CFG(tmp_path_factory):
    d = tmp_path_factory.mktemp("lafida_config")
    content = json.loads((pathlib.Path(__file__).resolve().parent / "golden" / "lafida_config.json").read_text())
    for name, kv in content.items():
        lines = ["%YAML:1.0", "", "# written by tests/test_settings_cpu.py"]
        for k, v in kv.items():
            lines.append(f"{k}: {v!r}   # {k.split('.')[-1]}" if isinstance(v, float) else f"{k}: {v}")
        (d / name).write_text("\n".join(lines) + "\n")
    return d


def test_interior_orientation_matches_packaged_cameras(cams, CFG):
    from multicol_slam_b200 import settings as S
    M_c, loaded = S.load_rig(CFG)
    assert loaded == cams                                     # multicol_slam_b200/data/lafida_cams.json came from the same files
    assert M_c.shape == (3, 4, 4)
    R = M_c[:, :3, :3]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-15 and np.allclose(np.linalg.det(R), 1.0)
    assert np.array_equal(R[2], np.eye(3))                    # camera 3 carries zero Cayley parameters
    assert M_c[0, 0, 3] == -0.140202124607334 and M_c[2, 2, 3] == 0.201416323496156
    # Cayley parameters of camera 1 reproduce: c = (R - R^T) entries / (1 + trace)
    c = np.array([R[0][2, 1] - R[0][1, 2], R[0][0, 2] - R[0][2, 0], R[0][1, 0] - R[0][0, 1]]) / (1 + np.trace(R[0]))
    assert np.allclose(c, [-0.0238361786473007, -2.05998171167958, 0.695126790868671], atol=1e-13)


def test_rig_matrices_and_essential(cams, CFG):
    from multicol_slam_b200 import settings as S
    M_c, _ = S.load_rig(CFG)
    rng = np.random.default_rng(0)

    def pose():
        M = np.eye(4); M[:3, :3] = S.cayley2rot(*rng.normal(0, 0.3, 3)); M[:3, 3] = rng.normal(0, 0.5, 3)
        return M
    m1, m1i = S.rig_matrices(pose(), M_c)
    m2, m2i = S.rig_matrices(pose(), M_c)
    assert np.abs(m1 @ m1i - np.eye(4)).max() < 1e-14 and np.array_equal(m1i[1], S.inv_rigid(m1[1]))
    # ComputeE(T1, T2) is the essential matrix of two world->camera transforms: x1^T E x2 = 0 for any scene point
    X = np.array([2.0, -1.0, 5.0, 1.0])
    for i in range(3):
        for j in range(3):
            E = S.compute_E(m1i[i], m2i[j])
            x1, x2 = (m1i[i] @ X)[:3], (m2i[j] @ X)[:3]
            assert abs(x1 @ E @ x2) / (np.linalg.norm(x1) * np.linalg.norm(x2)) < 1e-14
            assert abs(np.linalg.det(E)) < 1e-14 and np.allclose(np.linalg.svd(E, compute_uv=False)[:2], 1.0)
    Es = S.essential_matrices(m1i, m2)                          # the argument pair SearchForTriangulationRaw uses (ref :988-1001)
    assert Es.shape == (3, 3, 3, 3) and np.array_equal(Es[1, 2], S.compute_E(m1i[1], m2[2]))


def test_extractor_settings(CFG):
    from multicol_slam_b200 import settings as S
    track, init = S.extractor_settings(CFG / "Slam_Settings_indoor1.yaml")
    assert track["nfeatures"] == 400 and init["nfeatures"] == 800 and track["fastThreshold"] == 20 and init["fastThreshold"] == 5
    assert track["scaleFactor"] == float(np.float32(1.2)) and track["nlevels"] == 8 and track["descSize"] == 32
    assert track["do_dBrief"] is False and track["learnMasks"] is False and track["useAgast"] is False and track["fastAgastType"] == 2
    assert {k: v for k, v in track.items() if k not in ("nfeatures", "fastThreshold")} == {k: v for k, v in init.items() if k not in ("nfeatures", "fastThreshold")}
    kv = S.read_opencv_yaml(CFG / "Slam_Settings_indoor1.yaml")
    assert kv["Camera.fps"] == 25.0 and kv["Camera.RGB"] == 1 and kv["UseMotionModel"] == 1 and kv["traj.EndFrame"] == 759


def test_extractor_settings_feed_the_oracle(oa, cams, CFG):
    """the keyword sets are accepted by the extractor constructors (oracle here; same ExtractorParams as the GPU class)"""
    from multicol_slam_b200 import settings as S, synth
    track, init = S.extractor_settings(CFG / "Slam_Settings_indoor1.yaml")
    for kw, lo, hi in ((track, 380, 420), (init, 780, 830)):
        e = oa.OracleExtractor(nfeatures=kw["nfeatures"], scale_factor=kw["scaleFactor"], nlevels=kw["nlevels"], fast_threshold=kw["fastThreshold"],
                               do_dbrief=kw["do_dBrief"], learn_masks=kw["learnMasks"], desc_size=kw["descSize"])
        k, d, m = e.extract(synth.frame(cams[0], 2), synth.mirror_mask(cams[0]), cams[0])
        assert lo <= len(k) <= hi and d.shape[1] == 32

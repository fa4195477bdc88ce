"""Deterministic synthetic fisheye frames and calibrations.  numpy only, so the very
same bytes are produced in the build container and on the GPU box."""
import json
import pathlib
import numpy as np

_DATA = pathlib.Path(__file__).resolve().parent / "data"


def lafida_cams():
    """The three Lafida interior orientations (ref Examples/Lafida/InteriorOrientationFisheye*.yaml)."""
    return json.loads((_DATA / "lafida_cams.json").read_text())


def scaled_cam(cam, width, height):
    """A Lafida-like camera rescaled to another sensor size (synthetic 1280x720 / 1920x1080 rigs):
    image-plane quantities scale with s = height/cam.height, polynomials are re-parameterised so that
    f_s(rho) = s * f(rho / s) and rho_s(theta) = s * rho(theta)."""
    s = height / cam["height"]
    out = dict(cam)
    out["width"], out["height"] = width, height
    out["u0"] = cam["u0"] * s + (width - cam["width"] * s) / 2.0
    out["v0"] = cam["v0"] * s
    out["pol"] = [a * s ** (1 - i) for i, a in enumerate(cam["pol"])]
    out["inv_pol"] = [a * s for a in cam["inv_pol"]]
    return out


def mirror_mask(cam):
    """Level-0 mirror mask, same float32 arithmetic as ref src/cam_model_omni.cpp:181-220."""
    h, w = cam["height"], cam["width"]
    if cam.get("mirror_mask", 1) != 1:
        return np.ones((h, w), np.uint8)
    u0 = np.float32(cam["v0"])
    v0 = np.float32(cam["u0"])
    i = np.arange(h, dtype=np.float32)[:, None]
    j = np.arange(w, dtype=np.float32)[None, :]
    a = ((i - u0) * (i - u0)).astype(np.float32) + ((j - v0) * (j - v0)).astype(np.float32)
    return np.where(np.sqrt(a, dtype=np.float32) < np.float32(u0 + np.float32(22.0)), 255, 0).astype(np.uint8)


_G = np.array([0.00443305, 0.05400558, 0.24203623, 0.39905028, 0.24203623, 0.05400558, 0.00443305])


def _gauss(img):
    p = np.pad(img, 3, mode="edge")
    t = sum(_G[k] * p[:, k:k + img.shape[1]] for k in range(7))
    return sum(_G[k] * t[k:k + img.shape[0], :] for k in range(7))


def frame(cam, seed):
    """0.5*(nearest-upsampled quarter-res uniform noise) + 0.5*(gaussian-blurred uniform noise), uint8,
    multiplied by the mirror-mask disc."""
    h, w = cam["height"], cam["width"]
    rng = np.random.default_rng(seed)
    coarse = rng.integers(0, 256, size=((h + 3) // 4, (w + 3) // 4)).astype(np.float64)
    coarse = np.repeat(np.repeat(coarse, 4, axis=0), 4, axis=1)[:h, :w]
    fine = _gauss(rng.integers(0, 256, size=(h, w)).astype(np.float64))
    img = np.clip(np.rint(0.5 * coarse + 0.5 * fine), 0, 255).astype(np.uint8)
    return img * (mirror_mask(cam) > 0).astype(np.uint8)


def texture_stream(cam, n_frames, seed, step=(3, 2)):
    """A stream with real inter-frame motion: crops of one big texture sliding by `step` px/frame."""
    h, w = cam["height"], cam["width"]
    big = dict(cam)
    big["width"], big["height"], big["mirror_mask"] = w + step[0] * n_frames, h + step[1] * n_frames, 0
    tex = frame(big, seed)
    m = (mirror_mask(cam) > 0).astype(np.uint8)
    return np.stack([tex[t * step[1]:t * step[1] + h, t * step[0]:t * step[0] + w] * m for t in range(n_frames)])

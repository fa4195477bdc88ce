"""Generates tests/golden/bow_small_voc.npz from the REFERENCE's own DBoW2 compiled in place (oracle/_ref/libdbow2_ref.so,
`make -C oracle ref`; needs /root/reference) on the reference's vocabulary fixture.  Run in the authoring container:
    python tests/golden/make_bow_golden.py
Inputs are stored with the outputs, so the fixture is self-contained on the GPU box."""
import pathlib, sys
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import ref_dbow2_api as ra  # noqa: E402


def descriptors(voc, n, seed):
    """a third uniform random, a third vocabulary nodes with 0..40 flipped bits (exercises ties and deep descents),
    a third blends of two node descriptors"""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    nodes = voc["desc"][1:]
    a = nodes[rng.integers(0, len(nodes), n)].copy()
    bits = np.unpackbits(a, axis=1)
    for i in range(n):
        k = rng.integers(0, 41)
        bits[i, rng.choice(256, k, replace=False)] ^= 1
    a = np.packbits(bits, axis=1)
    b = nodes[rng.integers(0, len(nodes), n)]
    m = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    c = (a & m) | (b & ~m)
    kind = rng.integers(0, 3, n)
    return np.where(kind[:, None] == 0, d, np.where(kind[:, None] == 1, a, c)).astype(np.uint8)


def main():
    voc = np.load(ROOT / "tests" / "golden" / "voc_small_9_6.npz")
    out = {}
    cases = [(0, 0, 4, 600, 1), (0, 0, 3, 300, 2), (1, 1, 4, 300, 3), (5, 3, 6, 300, 4), (2, 2, 4, 300, 5), (0, 0, 7, 200, 6)]
    for ci, (sc, wg, levelsup, n, seed) in enumerate(cases):
        r = ra.RefVocabulary(scoring=sc, weighting=wg)
        d = descriptors(voc, n, seed)
        bw, bv, fn, fo, ff = r.transform(d, levelsup)
        w, wt = r.words(d)
        half = n // 2
        a = r.transform(d[:half], levelsup); b = r.transform(d[half:], levelsup)
        out.update({f"c{ci}_cfg": np.array([sc, wg, levelsup, n, seed]), f"c{ci}_desc": d, f"c{ci}_bow_words": bw, f"c{ci}_bow_values": bv,
                    f"c{ci}_fv_nodes": fn, f"c{ci}_fv_off": fo, f"c{ci}_fv_feat": ff, f"c{ci}_word": w, f"c{ci}_weight": wt,
                    f"c{ci}_score": np.array([r.score(a[0], a[1], b[0], b[1])])})
    out["n_cases"] = np.array([len(cases)])
    p = ROOT / "tests" / "golden" / "bow_small_voc.npz"
    np.savez_compressed(p, **out)
    print("wrote", p, p.stat().st_size, "bytes")


if __name__ == "__main__":
    main()

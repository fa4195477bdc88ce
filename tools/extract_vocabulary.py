"""Read the reference's ORB vocabulary fixture (Examples/small_orb_omni_voc_9_6.yml, DBoW2 YAML layout written by
TemplatedVocabulary::save, ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1476-1568) and write

  tests/golden/voc_small_9_6.npz   the tree as flat arrays (what mcs_vocabulary_create takes)
  <txt path>                       optional: the same tree in DBoW2's text layout (loadFromTextFile, :1338-1425), which is
                                   how the compiled reference (oracle/_ref/libdbow2_ref.so) loads it without cv::FileStorage

Data only, no code is copied.  Run in the authoring container (needs /root/reference):
    python tools/extract_vocabulary.py [--npz] [--txt oracle/_ref/voc_small_9_6.txt]
"""
import argparse, pathlib, re
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent


def parse_yaml(path):
    txt = pathlib.Path(path).read_text()
    head = {k: int(re.search(r"\b%s:\s*(\d+)" % k, txt).group(1)) for k in ("k", "L", "scoringType", "weightingType")}
    nodes = re.findall(r"nodeId:(\d+),\s*parentId:(\d+),\s*weight:([-+0-9.eE]+),\s*descriptor:\"([^\"]*)\"", txt)
    words = re.findall(r"wordId:(\d+),\s*nodeId:(\d+)", txt)
    return head, nodes, words


def flatten(head, nodes, words):
    n = len(nodes) + 1                                 # + root
    parent = np.full(n, -1, np.int32); weight = np.zeros(n, np.float64); desc = np.zeros((n, 32), np.uint8)
    order = np.zeros(n - 1, np.int32)                   # node ids in file order = the order load() pushes children (:1596-1608)
    for i, (nid, pid, w, d) in enumerate(nodes):
        nid = int(nid); order[i] = nid
        parent[nid] = int(pid); weight[nid] = float(w if not w.endswith(".") else w + "0")
        b = [int(t) for t in d.split()]
        assert len(b) == 32
        desc[nid] = b
    word_node = np.zeros(len(words), np.int32)
    for wid, nid in words:
        word_node[int(wid)] = int(nid)
    return dict(k=np.int32(head["k"]), L=np.int32(head["L"]), scoring=np.int32(head["scoringType"]), weighting=np.int32(head["weightingType"]),
                parent=parent, weight=weight, desc=desc, node_order=order, word_node=word_node)


def depth_of(parent):
    d = np.zeros(len(parent), np.int32)
    for i in range(1, len(parent)):                    # parents precede children in id order (checked below)
        d[i] = d[parent[i]] + 1
    return d


def write_text(v, path):
    """DBoW2 text layout: 'k L scoring weighting' then one line per node in node-id order: 'parent isLeaf d0..d31 weight'.
    No trailing newline: loadFromTextFile's while(!f.eof()) would turn an empty last line into a bogus child of the root."""
    n = len(v["parent"])
    is_leaf = np.ones(n, bool); is_leaf[v["parent"][1:]] = False
    lines = ["%d %d %d %d" % (v["k"], v["L"], v["scoring"], v["weighting"])]
    for i in range(1, n):
        lines.append("%d %d %s %s" % (v["parent"][i], int(is_leaf[i]), " ".join(str(int(b)) for b in v["desc"][i]), repr(float(v["weight"][i]))))
    pathlib.Path(path).parent.mkdir(parents=True, exist_ok=True)
    pathlib.Path(path).write_text("\n".join(lines))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--yml", default="/root/reference/Examples/small_orb_omni_voc_9_6.yml")
    ap.add_argument("--txt", default=None)
    ap.add_argument("--npz", action="store_true", help="(re)write tests/golden/voc_small_9_6.npz; without it only a missing file is written")
    a = ap.parse_args()
    head, nodes, words = parse_yaml(a.yml)
    v = flatten(head, nodes, words)
    n = len(v["parent"])
    assert (v["parent"][1:] < np.arange(1, n)).all(), "a child precedes its parent"
    # the text layout implies: children in ascending id order, word ids in ascending leaf-node order -- check that the YAML agrees
    pos = np.zeros(n, np.int64); pos[v["node_order"]] = np.arange(n - 1)
    for p in range(n):
        ch = np.nonzero(v["parent"] == p)[0]
        assert (np.diff(pos[ch]) > 0).all(), "file order of children differs from id order"
    is_leaf = np.ones(n, bool); is_leaf[v["parent"][1:]] = False; is_leaf[0] = False
    assert (np.nonzero(is_leaf)[0] == v["word_node"]).all(), "word ids are not in leaf-node order"
    d = depth_of(v["parent"])
    print("nodes %d  words %d  k %d  L %d  scoring %d  weighting %d" % (n, len(v["word_node"]), v["k"], v["L"], v["scoring"], v["weighting"]))
    print("leaf depth histogram", np.bincount(d[is_leaf]), " max children", np.bincount(v["parent"][1:]).max())
    out = ROOT / "tests" / "golden" / "voc_small_9_6.npz"
    if a.npz or not out.exists():
        np.savez_compressed(out, **v)
        print("wrote", out, out.stat().st_size, "bytes")
    else:
        old = np.load(out)
        assert all(np.array_equal(old[k], v[k]) for k in v), "committed fixture differs from the reference's vocabulary file"
    if a.txt:
        write_text(v, a.txt); print("wrote", a.txt)


if __name__ == "__main__":
    main()

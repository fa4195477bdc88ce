"""CPU, world_size 2, gloo: the N>1 host logic -- camera-to-rank sharding, slot packing and the single all_gather
reproduce exactly the single-process camera-order concatenation (ref src/cMultiFrame.cpp:168-184).  Features come
from the oracle here (no GPU); on the GPU box bench.py --gpus N exercises the same collective over NCCL."""
import os
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch.distributed as dist
    import oracle_api as oa
    from multicol_slam_b200 import rig, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cams = synth.lafida_cams()
    e = oa.OracleExtractor(nfeatures=300, do_dbrief=True, learn_masks=True)
    cap = e.info.capacity
    local = {c: e.extract(synth.frame(cams[c], 90 + c), synth.mirror_mask(cams[c]), cams[c]) for c in rig.cameras_of_rank(3, world, rank)}
    allc = rig.allgather_rig(local, 3, cap, 32)
    q.put((rank, [(k.tobytes(), d.tobytes(), m.tobytes()) for k, d, m in allc]))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_two_ranks(oa, cams):
    import torch.multiprocessing as mp
    from multicol_slam_b200 import rig, synth
    assert rig.cameras_of_rank(3, 2, 0) == [0, 2] and rig.cameras_of_rank(3, 2, 1) == [1]
    assert rig.cameras_of_rank(8, 8, 5) == [5] and rig.cameras_of_rank(3, 8, 5) == []
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    e = oa.OracleExtractor(nfeatures=300, do_dbrief=True, learn_masks=True)
    ref = [e.extract(synth.frame(cams[c], 90 + c), synth.mirror_mask(cams[c]), cams[c]) for c in range(3)]
    ref = [(k.tobytes(), d.tobytes(), m.tobytes()) for k, d, m in ref]
    assert got[0] == ref and got[1] == ref


def test_slot_roundtrip(api):
    from multicol_slam_b200 import rig
    from multicol_slam_b200.ctypes_defs import KEYPOINT_DTYPE
    rng = np.random.default_rng(0)
    k = np.zeros(17, KEYPOINT_DTYPE)
    k["x"] = rng.random(17)
    d, m = rng.integers(0, 256, (17, 32)).astype(np.uint8), rng.integers(0, 256, (17, 32)).astype(np.uint8)
    buf = rig.pack_slot(k, d, m, 40, 32)
    assert buf.size == api.lib().mcs_slot_bytes(40, 32)
    # the numpy packers follow the library's layout (mcs_packed_layout) for any batch size
    import ctypes as C
    for n, cap, dim in [(1, 40, 32), (3, 2016, 32), (384, 2016, 32), (7, 4016, 64), (2, 416, 16)]:
        off = (C.c_size_t * 4)()
        api.lib().mcs_packed_layout.restype = C.c_size_t
        total = api.lib().mcs_packed_layout(n, cap, dim, off)
        assert (list(off), total) == rig.packed_layout(n, cap, dim)
    k2, d2, m2 = rig.unpack_slot(buf, 40, 32)
    assert k2.tobytes() == k.tobytes() and np.array_equal(d, d2) and np.array_equal(m, m2)
    # empty camera
    k3, d3, m3 = rig.unpack_slot(rig.pack_slot(k[:0], d[:0], m[:0], 40, 32), 40, 32)
    assert len(k3) == 0 and d3.shape == (0, 32)

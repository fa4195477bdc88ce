"""Multi-GPU host logic of the path: the cameras of a rig (or chunks of a frame stream) are sharded
one-camera-per-GPU, every GPU extracts straight into ONE packed feature buffer (mcs_packed_layout: counts | keypoints |
descriptors | masks, the layout K3 writes in place) and ONE all-gather over NCCL / NVLink gives every rank the buffers of all
ranks, i.e. the concatenated mvKeys / mDescriptors in camera order (ref src/cMultiFrame.cpp:168-184).

On the GPUs the collective is the library's own mcs_allgather_features (ncclAllGather on the extractor's stream, see
`Communicator`); torch.distributed only carries the 128-byte rendezvous token.  The numpy packers below follow the same layout
and exist for host-side consumers and for the CPU tests (gloo, world size 2)."""
import ctypes as C

import numpy as np

from .ctypes_defs import KEYPOINT_DTYPE


def cameras_of_rank(n_cams, world, rank):
    """camera c -> GPU c mod G (the reference runs one OpenMP thread per camera, src/cMultiFrame.cpp:128)."""
    return [c for c in range(n_cams) if c % world == rank]


def packed_layout(n_images, capacity, dim):
    """(offsets[4], total bytes) of the packed feature buffer: the same arithmetic as mcs_packed_layout() (checked against the
    library in tests/test_multi_gloo.py)."""
    sizes = [n_images * 4, n_images * capacity * KEYPOINT_DTYPE.itemsize, n_images * capacity * dim, n_images * capacity * dim]
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (s + 255) // 256 * 256
    return offs, o


def slot_bytes(capacity, dim):
    """mcs_slot_bytes(): the packed buffer of one camera."""
    return packed_layout(1, capacity, dim)[1]


def pack(per_image, capacity, dim, out=None):
    """per_image: list of (kps, desc, dmask or None) -> packed buffer (uint8) in the library's layout"""
    n = len(per_image)
    offs, total = packed_layout(n, capacity, dim)
    buf = np.zeros(total, np.uint8) if out is None else out
    buf[:] = 0
    counts = buf[offs[0]:offs[0] + 4 * n].view(np.int32)
    kps = buf[offs[1]:offs[1] + n * capacity * 28].view(KEYPOINT_DTYPE).reshape(n, capacity)
    desc = buf[offs[2]:offs[2] + n * capacity * dim].reshape(n, capacity, dim)
    dmask = buf[offs[3]:offs[3] + n * capacity * dim].reshape(n, capacity, dim)
    for i, (k, d, m) in enumerate(per_image):
        assert len(k) <= capacity
        counts[i] = len(k)
        kps[i, :len(k)] = k
        desc[i, :len(k)] = d
        if m is not None:
            dmask[i, :len(k)] = m
    return buf


def unpack(buf, n_images, capacity, dim):
    offs, _ = packed_layout(n_images, capacity, dim)
    counts = buf[offs[0]:offs[0] + 4 * n_images].view(np.int32)
    kps = buf[offs[1]:offs[1] + n_images * capacity * 28].view(KEYPOINT_DTYPE).reshape(n_images, capacity)
    desc = buf[offs[2]:offs[2] + n_images * capacity * dim].reshape(n_images, capacity, dim)
    dmask = buf[offs[3]:offs[3] + n_images * capacity * dim].reshape(n_images, capacity, dim)
    return [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy(), dmask[i, :counts[i]].copy()) for i in range(n_images)]


def pack_slot(kps, desc, dmask, capacity, dim, out=None):
    return pack([(kps, desc, dmask)], capacity, dim, out=out)


def unpack_slot(buf, capacity, dim):
    return unpack(buf, 1, capacity, dim)[0]


def allgather_rig(per_cam_local, n_cams, capacity, dim, device="cpu"):
    """per_cam_local: {camera index: (kps, desc, dmask)} for the cameras this rank owns.  One all_gather of the rank's packed
    buffer over torch.distributed (host-side consumers / CPU tests); returns [(kps, desc, dmask)] for ALL cameras in camera order."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_cams + world - 1) // world            # images per rank (the last ranks may carry empty ones)
    mine = cameras_of_rank(n_cams, world, rank)
    empty = (np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, dim), np.uint8), None)
    block = pack([per_cam_local[mine[i]] if i < len(mine) else empty for i in range(per_rank)], capacity, dim)
    send = torch.from_numpy(block).to(device)
    recv = torch.empty(world * block.size, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    flat = recv.cpu().numpy()
    out = [None] * n_cams
    for r in range(world):
        got = unpack(flat[r * block.size:(r + 1) * block.size], per_rank, capacity, dim)
        for i, c in enumerate(cameras_of_rank(n_cams, world, r)):
            out[c] = got[i]
    return out


class Communicator:
    """mcs_comm of the rig's GPUs (one process per GPU): the rendezvous token is drawn by rank 0 with mcs_comm_unique_id and
    broadcast over the already initialised torch.distributed group; the collective itself is the library's ncclAllGather."""

    def __init__(self, device):
        import torch
        import torch.distributed as dist
        from . import api
        self.api, self.lib = api, api.lib()
        rank, world = dist.get_rank(), dist.get_world_size()
        tok = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = np.zeros(128, np.uint8)
            api._check(self.lib.mcs_comm_unique_id(t.ctypes.data_as(C.c_void_p)))
            tok = torch.from_numpy(t)
        tok = tok.to(device)
        dist.broadcast(tok, 0)
        t = tok.cpu().numpy()
        self.h = C.c_void_p()
        api._check(self.lib.mcs_comm_create(t.ctypes.data_as(C.c_void_p), rank, world, C.byref(self.h)))
        self.rank, self.world = rank, world

    def allgather(self, packed_t, gathered_t, stream):
        """packed_t / gathered_t: uint8 torch tensors on the GPU (gathered = world x packed); stream: torch.cuda.Stream"""
        self.api._check(self.lib.mcs_allgather_features(self.h, C.c_void_p(packed_t.data_ptr()), C.c_size_t(packed_t.numel()),
                                                        C.c_void_p(gathered_t.data_ptr()), C.c_void_p(stream.cuda_stream)))

    def close(self):
        if self.h:
            self.lib.mcs_comm_destroy(self.h)
            self.h = C.c_void_p()

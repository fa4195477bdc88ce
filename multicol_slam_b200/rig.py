"""Multi-GPU host logic of the path: cameras of a rig (or chunks of a frame stream) are sharded
one-camera-per-GPU, every GPU packs its cameras' features into fixed-size slots and ONE all_gather over
NCCL/NVLink gives every rank the concatenated mvKeys / mDescriptors in camera order (ref src/cMultiFrame.cpp:168-184).
torch.distributed is only the plumbing (nccl on GPUs, gloo in the CPU tests)."""
import numpy as np

from .ctypes_defs import KEYPOINT_DTYPE


def cameras_of_rank(n_cams, world, rank):
    """camera c -> GPU c mod G (the reference runs one OpenMP thread per camera, src/cMultiFrame.cpp:128)."""
    return [c for c in range(n_cams) if c % world == rank]


def slot_bytes(capacity, dim):
    """Same as mcs_slot_bytes(): int32 n, 12 pad bytes, kp[capacity], desc[capacity*dim], dmask[capacity*dim]."""
    return 16 + capacity * (KEYPOINT_DTYPE.itemsize + 2 * dim)


def pack_slot(kps, desc, dmask, capacity, dim, out=None):
    n = len(kps)
    assert n <= capacity
    buf = np.zeros(slot_bytes(capacity, dim), np.uint8) if out is None else out
    buf[:4] = np.array([n], np.int32).view(np.uint8)
    o = 16
    buf[o:o + n * 28] = np.ascontiguousarray(kps, KEYPOINT_DTYPE).view(np.uint8).reshape(-1)
    o += capacity * 28
    buf[o:o + n * dim] = np.ascontiguousarray(desc, np.uint8).reshape(-1)
    o += capacity * dim
    if dmask is not None:
        buf[o:o + n * dim] = np.ascontiguousarray(dmask, np.uint8).reshape(-1)
    return buf


def unpack_slot(buf, capacity, dim):
    n = int(buf[:4].view(np.int32)[0])
    o = 16
    kps = buf[o:o + n * 28].view(KEYPOINT_DTYPE).copy()
    o += capacity * 28
    desc = buf[o:o + n * dim].reshape(n, dim).copy()
    o += capacity * dim
    dmask = buf[o:o + n * dim].reshape(n, dim).copy()
    return kps, desc, dmask


def allgather_rig(per_cam_local, n_cams, capacity, dim, device="cpu"):
    """per_cam_local: {camera index: (kps, desc, dmask)} for the cameras this rank owns.  One all_gather of the
    rank's slot block; returns the list [(kps, desc, dmask)] for ALL cameras in camera order."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = (n_cams + world - 1) // world            # slots per rank (the last ranks may carry empty slots)
    sb = slot_bytes(capacity, dim)
    block = np.zeros(per_rank * sb, np.uint8)
    for i, c in enumerate(cameras_of_rank(n_cams, world, rank)):
        pack_slot(*per_cam_local[c], capacity, dim, out=block[i * sb:(i + 1) * sb])
    send = torch.from_numpy(block).to(device)
    recv = torch.empty(world * block.size, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    flat = recv.cpu().numpy()
    out = [None] * n_cams
    for r in range(world):
        for i, c in enumerate(cameras_of_rank(n_cams, world, r)):
            out[c] = unpack_slot(flat[(r * per_rank + i) * sb:(r * per_rank + i + 1) * sb], capacity, dim)
    return out

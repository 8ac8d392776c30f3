"""Frame-parallel sharding across GPUs (one process per GPU) and the single gather of the per-frame rectangle lists.

The reference is single-device; frames are independent, so a batch of F frames is cut into
contiguous shards, rank r owning frames [r*ceil(F/G), ...).  There is no data-path collective: the
only exchange is one gather of the variable-length rect_t lists (176 B per rectangle) to rank 0, done as an
all_gather of the per-rank sizes followed by an all_gather of size-padded byte tensors - over NCCL (NVLink/NVSwitch)
when the tensors live on the GPU, over gloo in the CPU tests.
"""
import numpy as np

from .api import RECT_DTYPE, RectLists


def shard_range(nframes, world_size, rank):
    """contiguous shard [lo, hi) of rank `rank` for a batch of `nframes` frames"""
    per = -(-nframes // world_size)
    lo = min(rank * per, nframes)
    return lo, min(lo + per, nframes)


def pack_rect_lists(first_frame, rect_lists):
    """per-frame rect arrays (a list, or a RectLists) -> one uint8 blob: int64 header [nframes, first_frame], int64 counts, then the rect_t bytes"""
    head = np.array([len(rect_lists), first_frame], np.int64)
    if isinstance(rect_lists, RectLists):
        counts = rect_lists.counts().astype(np.int64)
        body = np.ascontiguousarray(rect_lists.flat[: int(counts.sum())]).view(np.uint8).reshape(-1)
    else:
        counts = np.array([len(r) for r in rect_lists], np.int64)
        body = np.concatenate([np.ascontiguousarray(r).view(np.uint8).reshape(-1) for r in rect_lists]) if len(rect_lists) and counts.sum() else np.zeros(0, np.uint8)
    return np.concatenate([head.view(np.uint8), counts.view(np.uint8), body])


def unpack_rect_lists(blob):
    """inverse of pack_rect_lists -> (first_frame, RectLists over one copy of the payload)"""
    blob = np.ascontiguousarray(blob, np.uint8)
    nframes, first = (int(v) for v in blob[:16].view(np.int64))
    counts = blob[16: 16 + 8 * nframes].view(np.int64)
    off = 16 + 8 * nframes
    flat = blob[off: off + int(counts.sum()) * RECT_DTYPE.itemsize].copy().view(RECT_DTYPE)
    return first, RectLists(flat, counts)


def gather_rect_lists(first_frame, rect_lists, nframes_total, device=None):
    """gather every rank's per-frame rect lists; returns the full list (index = frame) on rank 0, None elsewhere.
    Works without an initialised process group (single process)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rect_lists
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    blob = torch.from_numpy(pack_rect_lists(first_frame, rect_lists)).to(dev)
    size = torch.tensor([blob.numel()], dtype=torch.int64, device=dev)
    if dist.get_backend() == "nccl":
        # one tensor per collective and one read-back each (a list of per-rank tensors costs a device round trip per rank)
        allsz = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allsz, size)
        sizes = [int(v) for v in allsz.cpu().tolist()]
        mx = max(sizes)
        padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
        padded[: blob.numel()] = blob
        allparts = torch.empty(world * mx, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allparts, padded)
        if rank != 0:
            return None
        host = allparts.cpu().numpy()
        parts = [host[r * mx: r * mx + sizes[r]] for r in range(world)]
    else:
        szl = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(szl, size)
        sizes = [int(s.item()) for s in szl]
        mx = max(sizes)
        padded = torch.zeros(mx, dtype=torch.uint8, device=dev)
        padded[: blob.numel()] = blob
        pl = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(pl, padded)
        if rank != 0:
            return None
        parts = [pl[r][: sizes[r]].cpu().numpy() for r in range(world)]
    # ranks own contiguous shards in rank order (shard_range), so the full list is the concatenation of the parts
    firsts, flats, counts = [], [], []
    for r in range(world):
        first, lists = unpack_rect_lists(parts[r])
        firsts.append(first)
        flats.append(lists.flat)
        counts.append(lists.counts())
    order = np.argsort(firsts, kind="stable")
    pos = 0
    for r in order:
        if len(counts[r]) and firsts[r] != pos:
            raise ValueError("rank shards are not contiguous: shard starting at frame %d arrives at position %d" % (firsts[r], pos))
        pos += len(counts[r])
    if pos != nframes_total:
        raise ValueError("gathered %d frames, expected %d" % (pos, nframes_total))
    return RectLists(np.concatenate([flats[r] for r in order]), np.concatenate([counts[r] for r in order]))

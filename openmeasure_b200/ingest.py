"""Streaming ingest of on-disk snapshot matrices.

The reference's users load `X = np.load('X_train.npy')` (README.md:50-53) -- a C-order FP64 (n, m)
matrix, feature-major rows -- and hand it to ROM/SPR.  Here a rank reads ONLY its own cells of
every feature straight from the file into HBM: the byte ranges of the shard are read with
os.preadv into a ring of pinned staging buffers by a reader thread while the previous buffer is
in flight to the device on a copy stream, so disk, PCIe and the first kernels overlap and the
matrix never exists in pageable host memory.

    Xd = load_npy_shard(path, n_features, rank=r, world=G)     # (F * n_c_loc, m) on the GPU
    spr = SPR.from_device(Xd, n_features, xyz_loc)
"""
import os
import threading

import numpy as np
import torch

CHUNK_BYTES = 64 << 20


def npy_header(path):
    """(shape, data_offset) of a C-order little-endian float64 .npy file; raises ValueError otherwise."""
    with open(path, "rb") as f:
        version = np.lib.format.read_magic(f)
        if version == (1, 0):
            shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
        elif version in ((2, 0), (3, 0)):
            shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
        else:
            raise ValueError(f"unsupported .npy version {version}")
        offset = f.tell()
    if fortran or len(shape) != 2:
        raise ValueError("the snapshot matrix must be a C-order 2-D array")
    if np.dtype(dtype) != np.dtype("<f8"):
        raise ValueError(f"the snapshot matrix must be little-endian float64, not {dtype}")
    return tuple(int(v) for v in shape), offset


def shard_cells(n_c, rank, world):
    """Cells [c0, c0 + n_c_loc) of every feature owned by `rank` (contiguous, balanced)."""
    base, extra = divmod(int(n_c), int(world))
    c0 = rank * base + min(rank, extra)
    return c0, base + (1 if rank < extra else 0)


def shard_ranges(shape, offset, n_features, rank=0, world=1):
    """Byte ranges (file_offset, nbytes, dst_row) of this rank's rows, one per feature block."""
    n, m = shape
    if n % n_features != 0:
        raise Exception('The number of rows of X is not a multiple of n_features')
    n_c = n // n_features
    c0, ncl = shard_cells(n_c, rank, world)
    row_bytes = 8 * m
    return [(offset + (f * n_c + c0) * row_bytes, ncl * row_bytes, f * ncl) for f in range(n_features)], ncl


def load_npy_shard(path, n_features, rank=0, world=1, device=None, chunk_bytes=CHUNK_BYTES):
    """This rank's (n_features * n_c_loc, m) float64 shard of the .npy snapshot matrix, on the GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("load_npy_shard needs a CUDA device (no CPU fallback)")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    shape, offset = npy_header(path)
    ranges, ncl = shard_ranges(shape, offset, n_features, rank, world)
    m = shape[1]
    out = torch.empty(n_features * ncl, m, dtype=torch.float64, device=device)
    flat = out.view(-1).view(torch.uint8)
    chunk_bytes = max(8 * m, (int(chunk_bytes) // (8 * m)) * 8 * m)          # whole rows per chunk
    nbuf = 3
    stage = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(nbuf)]
    free = [threading.Semaphore(1) for _ in range(nbuf)]
    ready = [threading.Semaphore(0) for _ in range(nbuf)]
    plan = []                                           # (file_offset, nbytes, dst_byte)
    for off, nb, dst_row in ranges:
        done = 0
        while done < nb:
            take = min(chunk_bytes, nb - done)
            plan.append((off + done, take, dst_row * 8 * m + done))
            done += take
    err = []

    def reader():
        try:
            fd = os.open(path, os.O_RDONLY)
            try:
                for i, (off, nb, _) in enumerate(plan):
                    b = i % nbuf
                    free[b].acquire()
                    view = memoryview(stage[b].numpy())[:nb]
                    got = 0
                    while got < nb:
                        k = os.preadv(fd, [view[got:]], off + got)
                        if k <= 0:
                            raise IOError("short read from " + path)
                        got += k
                    ready[b].release()
            finally:
                os.close(fd)
        except Exception as e:                          # surfaced on the main thread
            err.append(e)
            for s in ready:
                s.release()

    th = threading.Thread(target=reader, daemon=True)
    th.start()
    copy_stream = torch.cuda.Stream(device=device)
    # `out` was allocated on the current stream: a block the caching allocator recycled may still have kernels of its
    # previous owner pending there, so the copies wait for that stream, and the block is tied to the copy stream
    copy_stream.wait_stream(torch.cuda.current_stream(device))
    out.record_stream(copy_stream)
    events = [None] * nbuf
    with torch.cuda.stream(copy_stream):
        for i, (_, nb, dst) in enumerate(plan):
            b = i % nbuf
            ready[b].acquire()
            if err:
                raise err[0]
            flat[dst:dst + nb].copy_(stage[b][:nb], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            events[b] = ev
            # the staging buffer may be refilled once its copy has left the host
            threading.Thread(target=lambda e=ev, s=free[b]: (e.synchronize(), s.release()), daemon=True).start()
    th.join()
    torch.cuda.current_stream(device).wait_stream(copy_stream)
    return out

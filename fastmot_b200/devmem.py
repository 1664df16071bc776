"""Device-memory plumbing: torch tensors are used only as containers (allocation, streams, pinned memory).

`Uplink` packs many small host arrays into one pinned block and ships them with a single async H2D copy;
`ptr()` turns tensors into raw addresses for the C-ABI.
"""
import ctypes as C

import numpy as np
import torch


def ptr(t):
    """Raw device (or pinned-host) address of a torch tensor, or None."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


class Uplink:
    """Ring of pinned staging blocks -> one device block each; one cudaMemcpyAsync per `flush`."""

    ALIGN = 256

    def __init__(self, nbytes=1 << 20, depth=4, device="cuda"):
        self.nbytes = nbytes
        self.host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_np = [h.numpy() for h in self.host]
        self.dev = [torch.empty(nbytes, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.events = [None] * depth
        self.cur = 0
        self.off = 0
        self._begin()

    def _begin(self):
        ev = self.events[self.cur]
        if ev is not None:
            ev.synchronize()
        self.off = 0

    def put(self, arr):
        """Stage a host ndarray; returns the device address (c_void_p) it will have after flush()."""
        arr = np.ascontiguousarray(arr)
        n = arr.nbytes
        off = self.off
        if off + n > self.nbytes:
            raise MemoryError("Uplink block overflow")
        if n:
            self.host_np[self.cur][off:off + n] = arr.view(np.uint8).reshape(-1)
        self.off = (off + n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        return C.c_void_p(self.dev[self.cur].data_ptr() + off)

    def flush(self):
        if self.off:
            self.dev[self.cur][:self.off].copy_(self.host[self.cur][:self.off], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self.events[self.cur] = ev
        self.cur = (self.cur + 1) % len(self.host)
        self._begin()


class Downlink:
    """Device scratch block mirrored by a pinned host block; kernels write results at `alloc`ed offsets,
    `fetch()` brings the used prefix back with one D2H copy + stream synchronize."""

    ALIGN = 256

    def __init__(self, nbytes=1 << 20, device="cuda"):
        self.nbytes = nbytes
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.host_np = self.host.numpy()
        self.off = 0
        self.items = []

    def reset(self):
        self.off = 0
        self.items = []

    def alloc(self, shape, dtype):
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        off = self.off
        if off + n > self.nbytes:
            raise MemoryError("Downlink block overflow")
        self.off = (off + n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.items.append((off, n, tuple(shape), dtype))
        return C.c_void_p(self.dev.data_ptr() + off), len(self.items) - 1

    def fetch(self):
        """Returns list of ndarrays (views into the pinned block — copy if kept across fetches)."""
        if self.off:
            self.host[:self.off].copy_(self.dev[:self.off], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = []
        for off, n, shape, dtype in self.items:
            out.append(self.host_np[off:off + n].view(dtype).reshape(shape))
        return out


class FrameUploader:
    """HxWx3 u8 host frame -> device tensor with one async copy.  Page-locked sources (detected with
    cudaPointerGetAttributes) are copied directly; pageable ones are staged through an internal pinned ring."""

    def __init__(self, size, depth=2, device="cuda"):
        from . import _lib
        self._lib = _lib.load()
        w, h = size
        self.shape = (h, w, 3)
        self.nbytes = h * w * 3
        self.dev = [torch.empty(self.shape, dtype=torch.uint8, device=device) for _ in range(depth)]
        self.host = [torch.empty(self.shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_np = [t.numpy() for t in self.host]
        self.events = [None] * depth
        self.cur = 0
        # read-ahead (GPU-resident ingest, role of the reference's VideoIO frame queue, fastmot/videoio.py:125-142):
        # prefetch(frame) starts the H2D copy of a FUTURE frame on a dedicated upload stream while the current step
        # computes; upload() of that same frame then only waits for the copy's event.
        self._up_stream = torch.cuda.Stream(device=device)
        self._prefetched = None            # (key, slot, event)

    @staticmethod
    def _key(frame):
        return (frame.ctypes.data, frame.shape)

    def _copy(self, frame, k, stream):
        src = frame.ctypes.data
        if not self._lib.fm_host_is_pinned(C.c_void_p(src)):
            ev = self.events[k]
            if ev is not None:
                ev.synchronize()
            np.copyto(self.host_np[k], frame)
            src = self.host[k].data_ptr()
        self._lib.fm_memcpy_async(C.c_void_p(self.dev[k].data_ptr()), C.c_void_p(src), self.nbytes,
                                  C.c_void_p(stream.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(stream)
        self.events[k] = ev
        return ev

    def prefetch(self, frame):
        """Starts the upload of a frame that a later upload() call will ask for (the same ndarray)."""
        frame = np.ascontiguousarray(frame)
        if frame.shape != self.shape or frame.dtype != np.uint8:
            raise ValueError(f"frame must be uint8 {self.shape}, got {frame.dtype} {frame.shape}")
        k = self.cur
        self.cur = (k + 1) % len(self.dev)
        # the slot's previous contents may still be read by kernels of the main stream
        self._up_stream.wait_stream(torch.cuda.current_stream())
        ev = self._copy(frame, k, self._up_stream)
        self._prefetched = (self._key(frame), k, ev)

    def upload(self, frame):
        frame = np.ascontiguousarray(frame)
        if self._prefetched is not None and self._prefetched[0] == self._key(frame):
            _, k, ev = self._prefetched
            self._prefetched = None
            torch.cuda.current_stream().wait_event(ev)
            return self.dev[k]
        if frame.shape != self.shape or frame.dtype != np.uint8:
            raise ValueError(f"frame must be uint8 {self.shape}, got {frame.dtype} {frame.shape}")
        k = self.cur
        self.cur = (k + 1) % len(self.dev)
        self._copy(frame, k, torch.cuda.current_stream())
        return self.dev[k]

"""Driver-side I/O off the critical path: ``DepthMapWriter`` -- the save path of the reference's eval
drivers without their stalls -- and ``WindowIO`` -- double-buffered host<->device traffic around ``forward``.


The drivers save each map as ``np.save(path, np.float16(outputs[key].squeeze(1).cpu().numpy()))`` (eval_hybrid.py:260-264,
282-286, 276-277, 306-307): a blocking device->host copy of fp32 data, a host-side cast and a synchronous file write per map,
with the GPU idle meanwhile.  ``DepthMapWriter.save`` produces byte-identical files but casts to fp16 ON THE DEVICE (half the
bytes over PCIe; torch and numpy both round to nearest even), copies into pinned memory on a side stream behind the work that
produced the map, and leaves the wait for that copy and the file write to a worker thread.  The caller's stream is never
synchronised.  Host tensors take the same route minus the copy (that is what the CPU test pins against the reference's formula).
"""
import queue
import threading

import numpy as np
import torch


class DepthMapWriter(object):
    def __init__(self, max_pending=16, check=None):
        """``check``: optional callable run by ``close()`` after the last map is on disk, e.g. ``model.check`` -- the blocking
        fp16-range check of DepthNetHybrid, so that a sequence whose last windows overflowed is reported where its maps
        are finalised (the per-call check inside ``forward`` is deliberately late by one or two calls)."""
        self._check = check
        self._q = queue.Queue(maxsize=max_pending)       # back-pressure: at most max_pending maps in flight
        self._error = None
        self._streams = {}
        self._thread = threading.Thread(target=self._drain, name="estd-depth-writer", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ producer side
    def save(self, tensor, path, squeeze_channel=True):
        """tensor [B,1,H,W] (or any shape) fp32, CUDA or host -> ``path`` (.npy, float16).  ``squeeze_channel`` drops dim 1 as
        the drivers do for depth maps (``.squeeze(1)``); pass False and a pre-squeezed tensor for the probability maps
        (``.squeeze()``)."""
        if self._error is not None:
            self.close()
        t = tensor.detach()
        if squeeze_channel and t.dim() >= 2 and t.shape[1] == 1:
            t = t.squeeze(1)
        if t.is_cuda:
            main = torch.cuda.current_stream(t.device)
            side = self._streams.get(t.device)
            if side is None:
                side = self._streams[t.device] = torch.cuda.Stream(device=t.device)
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                half = t.to(torch.float16)
                host = torch.empty(half.shape, dtype=torch.float16, pin_memory=True)
                host.copy_(half, non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
            t.record_stream(side)                        # the producer may free / reuse the map as soon as the cast has read it
        else:
            host, done = t.to(torch.float16), None
        self._q.put((host, done, str(path)))

    def save_outputs(self, outputs, paths):
        """``paths``: {output key: file path}, e.g. {("depth", 0, 2): ".../init_depth/frame-000010.color.npy", ...}."""
        for key, path in paths.items():
            self.save(outputs[key], path)

    # ------------------------------------------------------------------ consumer side
    def _drain(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            host, done, path = item
            try:
                if self._error is None:
                    if done is not None:
                        done.synchronize()               # waits for this copy only, on this thread only
                    np.save(path, host.numpy())
            except BaseException as e:                   # surfaced to the producer by close() / the next save()
                self._error = e

    def close(self):
        """Waits until every submitted map is on disk; re-raises the first write error."""
        if self._thread.is_alive():
            self._q.put(None)
            self._thread.join()
        if self._error is not None:
            err, self._error = self._error, None
            raise err
        if self._check is not None:
            self._check()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class WindowIO(object):
    """Host<->device traffic of a stream of windows, off the critical path.

    The reference's drivers move every window to the GPU right before the forward (``tocuda(sample)``, eval_hybrid.py:231-243) and
    read the maps back right after it (``.cpu()``, :259-286): the GPU idles during both copies.  With this helper the NEXT window's
    images travel host->device on a copy stream while the current window computes, and a window's maps travel device->host while
    the next one computes; the host only ever waits for maps that were requested a window ago.

        io = WindowIO(device)
        nxt = io.upload(first_window_images)                    # pinned host tensor -> device, asynchronous
        for k, (poses, K) in enumerate(cameras):
            cur, nxt = nxt, (io.upload(images[k + 1]) if k + 1 < n else None)
            outputs, state, pose_state = model(io.ready(cur), poses, K, sample, state, pose_state, mode="val")
            pending = io.download([outputs[key] for key in keys])     # asynchronous, into pinned buffers
            if previous is not None:
                maps = previous.result()                        # the PREVIOUS window's maps: long since on the host
            previous = pending

    Nothing here changes what is computed; it is plain stream / event plumbing over pinned buffers."""

    class _Upload(object):
        __slots__ = ("tensor", "done")

    class _Download(object):
        __slots__ = ("buffers", "done")

        def result(self):
            """The maps as pinned host tensors (waits for this download only)."""
            self.done.synchronize()
            return self.buffers

    def __init__(self, device):
        self.device = torch.device(device)
        self.copy_in = torch.cuda.Stream(device=self.device)
        self.copy_out = torch.cuda.Stream(device=self.device)
        self._pool = {}            # (shape, dtype) -> free pinned buffers for downloads

    def upload(self, host_tensor):
        """Pinned host tensor -> device on the upload stream; returns a handle for ``ready``."""
        if not host_tensor.is_pinned():
            host_tensor = host_tensor.pin_memory()
        up = WindowIO._Upload()
        with torch.cuda.stream(self.copy_in):
            up.tensor = host_tensor.to(self.device, non_blocking=True)
            up.done = torch.cuda.Event()
            up.done.record(self.copy_in)
        return up

    def ready(self, upload):
        """The uploaded tensor, usable on the CURRENT stream (which is made to wait for the copy; the host does not)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(upload.done)
        upload.tensor.record_stream(cur)
        return upload.tensor

    def download(self, tensors):
        """Device tensors -> pinned host buffers on the download stream, behind the work already enqueued on the current
        stream.  Returns a handle whose ``result()`` gives the host tensors; buffers are recycled after ``release``."""
        cur = torch.cuda.current_stream(self.device)
        produced = torch.cuda.Event()
        produced.record(cur)
        self.copy_out.wait_event(produced)
        dl = WindowIO._Download()
        dl.buffers = []
        with torch.cuda.stream(self.copy_out):
            for t in tensors:
                key = (tuple(t.shape), t.dtype)
                free = self._pool.get(key)
                buf = free.pop() if free else torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf.copy_(t, non_blocking=True)
                t.record_stream(self.copy_out)
                dl.buffers.append(buf)
            dl.done = torch.cuda.Event()
            dl.done.record(self.copy_out)
        return dl

    def release(self, download):
        """Hands a finished download's pinned buffers back for reuse."""
        for buf in download.buffers:
            self._pool.setdefault((tuple(buf.shape), buf.dtype), []).append(buf)
        download.buffers = []

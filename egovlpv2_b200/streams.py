"""Two-stream execution of one training step: the RoBERTa tower (small, launch-latency-bound kernels on B*S = 256 rows)
runs on a side CUDA stream next to the video tower (large kernels) instead of in front of it.

The reference runs everything on the default stream ("Threading"); the two towers are independent
until the fused levels, where text layer i and video block i both read the PREVIOUS level's outputs (model.py:263-271),
so level i's pair is independent as well.  torch's autograd replays each backward node on the stream of its forward
and orders producers / consumers with events, so the backward overlaps the same way; everything stays capturable in
one CUDA graph (the side stream forks from and joins the capturing stream).

Off unless `enable(True)` is called (PretrainStep does; env EGV_TEXT_STREAM=0 vetoes): a bare `FrozenInTime` behaves
like the reference, single-stream."""
import contextlib
import os

import torch

_enabled = False
_side = {}


def enable(on=True):
    global _enabled
    _enabled = bool(on) and os.environ.get("EGV_TEXT_STREAM", "1") != "0" and torch.cuda.is_available()
    return _enabled


def enabled():
    return _enabled


def _side_stream():
    dev = torch.cuda.current_device()
    s = _side.get(dev)
    if s is None:
        s = _side[dev] = torch.cuda.Stream(device=dev)
    return s


@contextlib.contextmanager
def side():
    """Run the body on the side stream, ordered after everything already enqueued on the current stream."""
    if not _enabled:
        yield
        return
    main = torch.cuda.current_stream()
    s = _side_stream()
    s.wait_stream(main)
    with torch.cuda.stream(s):
        yield


def exchange():
    """Both streams wait for each other's enqueued work (start of a fused level)."""
    if _enabled:
        main, s = torch.cuda.current_stream(), _side_stream()
        s.wait_stream(main)
        main.wait_stream(s)


def join(*tensors):
    """The current stream waits for the side stream; `tensors` were produced there and are consumed here."""
    if _enabled:
        main = torch.cuda.current_stream()
        main.wait_stream(_side_stream())
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(main)


def to_side(*tensors):
    """Tensors produced on the current stream that the side stream is about to read (allocator bookkeeping)."""
    if _enabled:
        s = _side_stream()
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(s)


def to_main(*tensors):
    """Tensors produced on the side stream that the current stream is about to read."""
    if _enabled:
        main = torch.cuda.current_stream()
        for t in tensors:
            if t is not None and t.is_cuda:
                t.record_stream(main)

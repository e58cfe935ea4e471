"""Sub-batch pipelining of the env step (host plumbing: CUDA streams and events only).

Two launches of the step kernel over the SAME envs depend on each other, and programmatic dependent launch only half
hides what that costs: the next grid's prologue overlaps the previous grid's tail, but its first global access waits
until the previous grid has drained, and the stores of the new grid only start flowing one load latency plus the
dynamics later.  At 64K envs of the C1 chip that bubble is ~2.5 us of a 13 us step during which HBM is idle.  Envs are
independent, so the batch can be cut into K sub-batches whose step chains run on K streams: while sub-batch A waits for
its own previous step, the kernels of B, C, D keep the memory system busy.  Measured on a B200 (64K envs, C1): 13.0 ->
10.5 (K=2) -> 10.2 us (K=4) per step of the whole batch, i.e. the step reaches the measured HBM peak.

The sub-batches only pipeline ACROSS steps if nothing joins them between two steps: `step(..., join=False)` leaves the
side streams running and `join()` (called implicitly by every other method) makes the caller's stream wait for them.
With the default `join=True` every step is complete on the caller's stream when `step` returns, as before."""
import ctypes as C

import torch


def sub_batch_ranges(n_envs, k, unit=64):
    """[lo, hi) env ranges of at most k sub-batches that cover [0, n_envs); every boundary but the last is a multiple of
    `unit` envs, which keeps each sub-batch's observation rows 16-byte aligned (the TMA bulk store needs that) and the
    tiles of the step kernel whole."""
    units = (n_envs + unit - 1) // unit
    k = max(1, min(int(k), units))
    b = [min(n_envs, (units * c // k) * unit) for c in range(k + 1)]
    b[k] = n_envs
    return [(b[c], b[c + 1]) for c in range(k) if b[c + 1] > b[c]]


class SubBatches:
    def __init__(self, device, n_envs, k, unit=64):
        self.ranges = sub_batch_ranges(n_envs, k, unit)
        self.k = len(self.ranges)
        self.device = device
        self.streams = [torch.cuda.Stream(device=device) for _ in self.ranges]
        self.pending = False

    def handles(self):
        return [C.c_void_p(s.cuda_stream) for s in self.streams]

    def fork(self, keep_alive=()):
        """The side streams wait for everything enqueued so far on the caller's stream (the step's inputs)."""
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        capturing = torch.cuda.is_current_stream_capturing()
        for s in self.streams:
            s.wait_event(ev)
            if not capturing:                      # inside a graph capture the tensors belong to the graph's pool
                for t in keep_alive:
                    if t is not None:
                        t.record_stream(s)
        self.pending = True

    def join(self):
        """The caller's stream waits for the side streams."""
        if not self.pending:
            return
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)
        self.pending = False


def offset_ptr(t, lo):
    """Device pointer of row `lo` of a tensor whose first dimension runs over the envs (None stays None)."""
    if t is None:
        return None
    return t.data_ptr() + lo * t.stride(0) * t.element_size()

"""Overlapped device->host read-back of SR frames.

The engine reuses one output buffer per plan (`net.reuse_output_buffers`), and a batch of 32 ACDC x4 sequences is 209 MB
of fp32 frames: copied on the compute stream that is ~4 ms of PCIe time per step during which the GPU idles.
`HostFrameRing` moves the copy to its own stream: `submit()` enqueues the D2H of a step's frames behind an event of the
compute stream, `before_launch()` makes the compute stream wait for the copies still reading the buffer it is about to
overwrite (they finished long ago in steady state: the last head kernel runs at the END of a forward), `drain()` joins.
Every step still pays its H2D and D2H; they just overlap the next step's compute.
(The reference predictor reads results back with one `.item()` per frame: acdc_vsr_refinenet_predictor.py:74-75,170-174.)
"""
import torch


def _flat_view(frames):
    """One tensor over a list of equally shaped frames that are consecutive views of a single buffer (what
    RefineNet returns with reuse_output_buffers); None if they are not."""
    base = frames[0]
    step = base.numel() * base.element_size()
    if not base.is_contiguous():
        return None
    for t, f in enumerate(frames):
        if f.shape != base.shape or f.dtype != base.dtype or not f.is_contiguous() or \
                f.data_ptr() != base.data_ptr() + t * step:
            return None
    return torch.as_strided(base, (len(frames),) + tuple(base.shape), (base.numel(),) + tuple(base.stride()))


class HostFrameRing:
    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.slots = slots
        self.host = [None] * slots
        self.done = [None] * slots        # event: copy into slot finished
        self.src = [None] * slots         # device byte range [begin, end) each outstanding copy reads from
        self.next = 0
        self.bytes_per_step = 0

    def before_launch(self, dst_ptr=None, dst_bytes=None):
        """Call before enqueueing a forward that overwrites a buffer earlier submits read from: the compute stream
        waits for the copies still reading any byte of [dst_ptr, dst_ptr + dst_bytes)
        (RefineNetEngine.next_output_ptr / next_output_bytes) - or for all of them when the address is not given.
        A submitted tensor may be a sub-view of the output buffer (e.g. the last list of an all-heads forward), so
        the test is a range overlap, not pointer equality; without `dst_bytes` any copy whose source starts at or
        after `dst_ptr` is waited for.  With engine.output_slots = 2 the copy of step i never reads what step i + 1
        writes, so the forward starts immediately and the copy overlaps it."""
        cur = torch.cuda.current_stream(self.device)
        for ev, src in zip(self.done, self.src):
            if ev is None:
                continue
            if dst_ptr is not None:
                s0, s1 = src
                d1 = dst_ptr + dst_bytes if dst_bytes is not None else None
                if s1 <= dst_ptr or (d1 is not None and s0 >= d1):
                    continue
            cur.wait_event(ev)

    def submit(self, frames):
        """Enqueues the D2H of `frames` (list of tensors or one tensor); returns the slot index."""
        src = frames if isinstance(frames, torch.Tensor) else _flat_view(frames)
        if src is None:
            src = torch.stack(list(frames), dim=0)
        s = self.next
        self.next = (s + 1) % self.slots
        if self.host[s] is None or self.host[s].shape != src.shape or self.host[s].dtype != src.dtype:
            self.host[s] = torch.empty(src.shape, dtype=src.dtype).pin_memory()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        if self.done[s] is not None:
            self.done[s].synchronize()          # the host slot is about to be overwritten: its previous copy must be out
        with torch.cuda.stream(self.stream):
            self.host[s].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[s] = ev
        self.src[s] = (src.data_ptr(), src.data_ptr() + src.numel() * src.element_size())
        self.bytes_per_step = src.numel() * src.element_size()
        return s

    def result(self, slot):
        """Host tensor of a submitted step (blocks until its copy has landed)."""
        self.done[slot].synchronize()
        return self.host[slot]

    def drain(self):
        """Makes the current stream (and the host) wait for every outstanding copy."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)
                ev.synchronize()

"""Asynchronous epoch log: the reference reads every train_result scalar back with `.item()` /
`torch_ext.mean_list(...).item()` once per epoch and writes TensorBoard scalars on the spot (learning/common_agent.py:109-152,551-564,
amp_agent.py:244-262), which stalls the host on the device every epoch.  Here the per-minibatch train_result series of an epoch (a
[n_minibatches, TR_COUNT] device tensor the learner fills without any host sync) is copied into one slot of a small ring of PINNED
host buffers with a non-blocking D2H copy, together with the epoch's CUDA events; `poll()` hands back, without ever blocking, the
epochs whose copy has completed.  The training loop therefore never waits for the device; statistics arrive an epoch or two late.
Plumbing only (torch tensors, events): no arithmetic lives here."""
import torch


class AsyncEpochLog:
    def __init__(self, names, depth=4):
        self.names = list(names)
        self.depth = int(depth)
        self._slots = [None] * self.depth          # pinned [n_minibatches, len(names)] buffers, allocated on first use
        self._pending = []                          # (slot, epoch, frames, done_event, (ev_start, ev_play, ev_end))
        self._head = 0

    def _slot_buffer(self, i, like):
        buf = self._slots[i]
        if buf is None or tuple(buf.shape) != tuple(like.shape):
            buf = torch.empty(like.shape, dtype=like.dtype, device='cpu')
            if like.is_cuda:
                buf = buf.pin_memory()
            self._slots[i] = buf
        return buf

    def push(self, epoch, series, frames=0, events=None):
        """series: [n_minibatches, len(names)] tensor on the training device (the learner's per-minibatch train_result rows).
        events: optional (start, after_rollout, end) CUDA events of the epoch.  Never blocks unless the ring is full, in which
        case the oldest epoch is drained first (that is the only place a wait can happen, depth epochs behind the device)."""
        out = []
        if len(self._pending) == self.depth:
            out = self.flush(max_epochs=1)
        i = self._head
        self._head = (self._head + 1) % self.depth
        buf = self._slot_buffer(i, series)
        buf.copy_(series, non_blocking=True)
        done = None
        if series.is_cuda:
            done = torch.cuda.Event()
            done.record()
        self._pending.append((i, epoch, frames, done, events))
        return out

    def _finish(self, rec):
        i, epoch, frames, done, events = rec
        buf = self._slots[i]
        scalars = {n: float(buf[:, j].mean()) for j, n in enumerate(self.names)}      # torch_ext.mean_list(...).item() of the reference
        r = {'epoch': epoch, 'frames': frames, 'scalars': scalars, 'series': buf.clone()}
        if events is not None:
            r['play_time'] = events[0].elapsed_time(events[1]) / 1e3
            r['update_time'] = events[1].elapsed_time(events[2]) / 1e3
        return r

    def poll(self):
        """Completed epochs, oldest first; returns immediately (possibly empty)."""
        out = []
        while self._pending and (self._pending[0][3] is None or self._pending[0][3].query()):
            out.append(self._finish(self._pending.pop(0)))
        return out

    def flush(self, max_epochs=None):
        """Wait for (up to max_epochs of) the pending epochs -- end of training, or a full ring."""
        out = []
        while self._pending and (max_epochs is None or len(out) < max_epochs):
            rec = self._pending.pop(0)
            if rec[3] is not None:
                rec[3].synchronize()
            out.append(self._finish(rec))
        return out

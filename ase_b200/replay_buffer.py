"""Ring buffer of AMP observations with permutation-based sampling -- host-side mirror of
learning/replay_buffer.py:3-84 (device tensors, torch indexing only: plumbing, no arithmetic)."""
import torch


class ReplayBuffer:
    def __init__(self, buffer_size, device):
        self._head = 0
        self._total_count = 0
        self._buffer_size = int(buffer_size)
        self._device = device
        self._data_buf = None
        self._sample_idx = torch.randperm(self._buffer_size)
        self._sample_head = 0

    def get_buffer_size(self):
        return self._buffer_size

    def get_total_count(self):
        return self._total_count

    def store(self, data_dict):
        if self._data_buf is None:
            self._data_buf = {k: torch.zeros((self._buffer_size,) + tuple(v.shape[1:]), device=self._device, dtype=v.dtype)
                              for k, v in data_dict.items()}
        n = next(iter(data_dict.values())).shape[0]
        assert n <= self._buffer_size
        for key, buf in self._data_buf.items():
            v = data_dict[key]
            store_n = min(n, self._buffer_size - self._head)
            buf[self._head:self._head + store_n] = v[:store_n]
            if n - store_n > 0:
                buf[0:n - store_n] = v[store_n:]
        self._head = (self._head + n) % self._buffer_size
        self._total_count += n

    def sample_indices(self, n):
        idx = torch.arange(self._sample_head, self._sample_head + n) % self._buffer_size
        rand_idx = self._sample_idx[idx]
        if self._total_count < self._buffer_size:
            rand_idx = rand_idx % self._head
        self._sample_head += n
        if self._sample_head >= self._buffer_size:
            self._sample_idx[:] = torch.randperm(self._buffer_size)
            self._sample_head = 0
        return rand_idx

    def sample(self, n):
        rand_idx = self.sample_indices(n).to(self._device)
        return {k: v[rand_idx] for k, v in self._data_buf.items()}

    def rows(self, key, idx):
        """Gather only the rows `idx` (the learner consumes amp_minibatch_size rows per minibatch, ase_agent.py:172-181)."""
        return self._data_buf[key][idx]

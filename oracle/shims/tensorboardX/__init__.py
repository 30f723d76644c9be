"""oracle shim: no-op SummaryWriter (common_agent.py:23, amp_agent.py:19)."""
class SummaryWriter:
    def __init__(self, *a, **k): pass
    def add_scalar(self, *a, **k): pass
    def add_text(self, *a, **k): pass
    def flush(self): pass
    def close(self): pass

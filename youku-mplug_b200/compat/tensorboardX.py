"""utils.py:176 calls SummaryWriter(logdir=...) (tensorboardX spelling)."""


class SummaryWriter:
    def __init__(self, logdir=None, log_dir=None, **kwargs):
        self.logdir = logdir or log_dir
        self._impl = None
        try:
            from torch.utils.tensorboard import SummaryWriter as _W
            self._impl = _W(log_dir=self.logdir)
        except Exception:
            self._impl = None

    def add_scalar(self, *a, **k):
        if self._impl is not None:
            self._impl.add_scalar(*a, **k)

    def flush(self):
        if self._impl is not None:
            self._impl.flush()

    def close(self):
        if self._impl is not None:
            self._impl.close()

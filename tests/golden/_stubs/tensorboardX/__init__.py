"""Import stub for `tensorboardX` (not installed). TEST INFRASTRUCTURE ONLY."""


class SummaryWriter:
    def __init__(self, *a, **kw):
        pass

    def add_scalar(self, *a, **kw):
        pass

    def add_scalars(self, *a, **kw):
        pass

    def add_histogram(self, *a, **kw):
        pass

    def flush(self):
        pass

    def close(self):
        pass

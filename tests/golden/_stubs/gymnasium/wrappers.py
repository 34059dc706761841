class _W:
    def __init__(self, *a, **kw):
        raise RuntimeError('gymnasium stub')


class AtariPreprocessing(_W):
    pass


class FrameStackObservation(_W):
    pass


class FlattenObservation(_W):
    pass


class TimeLimit(_W):
    pass


class RecordVideo(_W):
    pass

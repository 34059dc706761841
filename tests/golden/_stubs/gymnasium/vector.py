class SyncVectorEnv:
    def __init__(self, *a, **kw):
        raise RuntimeError('gymnasium stub')


class AsyncVectorEnv(SyncVectorEnv):
    pass


class AutoresetMode:
    NEXT_STEP = 'next_step'
    SAME_STEP = 'same_step'
    DISABLED = 'disabled'

def get_checkpoint_state(*a, **k):
    return None


def latest_checkpoint(*a, **k):
    return None


class Checkpoint:
    def __init__(self, **kw):
        pass

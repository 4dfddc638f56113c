class SummaryWriter:  # no-op writer so trainRGB.py / trainmask.py import
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass

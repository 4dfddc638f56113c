class BufferedRansEncoder:  # names only: compress()/decompress() are out of scope
    def __init__(self, *a, **k):
        raise NotImplementedError("rANS coder stand-in")


class RansDecoder:
    def __init__(self, *a, **k):
        raise NotImplementedError("rANS coder stand-in")

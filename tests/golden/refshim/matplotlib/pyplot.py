def close(*a, **k):
    return None


def figure(*a, **k):
    raise RuntimeError("matplotlib stub: plotting is out of scope")

class Line2D:
    def __init__(self, *a, **kw):
        pass

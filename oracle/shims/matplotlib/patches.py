class Patch:
    def __init__(self, *a, **kw):
        pass

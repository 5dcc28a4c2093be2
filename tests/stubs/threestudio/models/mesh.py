class Mesh:
    def __init__(self, *a, **k):
        pass

class PlyData:
    pass


class PlyElement:
    pass

"""Minimal functional stand-in for the ``plyfile`` package (dranjan/python-plyfile; not installed in this image, the
reference leaves it unpinned) -- just the surface the reference's geometry/gaussian_io.py:52-172 touches:
``PlyElement.describe(structured_array, "vertex")``, ``PlyData([el]).write(path)`` (binary little endian, plyfile's
default for ``text=False``), ``PlyData.read(path)``, ``plydata.elements[0][name]`` and ``.properties[i].name``.
Written independently of the product's b200splat/ply.py so that each can check the other."""
import numpy as np

_NAMES = {"f4": "float", "f8": "double", "i4": "int", "u1": "uchar", "i2": "short", "u2": "ushort", "u4": "uint",
          "i1": "char"}
_TYPES = {v: k for k, v in _NAMES.items()}
_TYPES.update({"float32": "f4", "float64": "f8", "int32": "i4", "uint8": "u1"})


class PlyProperty:
    def __init__(self, name, dtype):
        self.name, self.dtype = name, dtype


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = tuple(PlyProperty(n, data.dtype[n].str[1:]) for n in data.dtype.names)

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    def __getitem__(self, key):
        return self.data[key]


class PlyData:
    def __init__(self, elements=()):
        self.elements = list(elements)

    def write(self, path):
        with open(path, "wb") as f:
            lines = ["ply", "format binary_little_endian 1.0"]
            for el in self.elements:
                lines.append(f"element {el.name} {len(el.data)}")
                lines += [f"property {_NAMES[p.dtype]} {p.name}" for p in el.properties]
            lines.append("end_header")
            f.write(("\n".join(lines) + "\n").encode("ascii"))
            for el in self.elements:
                le = el.data.astype(el.data.dtype.newbyteorder("<"))
                f.write(le.tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            assert f.readline().strip() == b"ply"
            elements, fmt = [], None
            while True:
                tok = f.readline().decode("ascii").split()
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    elements.append([tok[1], int(tok[2]), []])
                elif tok[0] == "property":
                    elements[-1][2].append((tok[2], _TYPES[tok[1]]))
                elif tok[0] == "end_header":
                    break
            assert fmt == "binary_little_endian"
            out = []
            for name, count, props in elements:
                dt = np.dtype([(n, "<" + t) for n, t in props])
                out.append(PlyElement(name, np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)))
            return PlyData(out)

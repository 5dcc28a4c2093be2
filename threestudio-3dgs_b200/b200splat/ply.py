"""INRIA-layout Gaussian PLY import / export, the on-disk format next to the hot path.

Same file the reference's ``GaussianIO.save_ply`` / ``load_ply`` write and read through ``plyfile``
(geometry/gaussian_io.py:36-172): one ``vertex`` element, all properties ``float`` (f4), in the order

    x y z | nx ny nz (zeros) | f_dc_0..2 | f_rest_0..3(M-1)-1 | opacity | scale_0..2 | rot_0..3

holding the RAW parameters (log scales, pre-sigmoid opacity, un-normalised quaternion).  ``f_dc`` / ``f_rest`` are stored
CHANNEL-major: the (P, K, 3) tensors are transposed to (P, 3, K) and flattened (gaussian_io.py:54-69), and read back by
reshaping to (P, 3, K) and transposing (:112-114, :147-156).  Host-side numpy code (file IO is not GPU work); the
writer emits ``binary_little_endian`` like plyfile's default, the reader also accepts ``ascii`` and
``binary_big_endian`` and ignores unknown extra properties.  plyfile itself is not installed in this image; the PLY
container format is restated from its public specification.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
              "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
              "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}


def attribute_names(n_dc: int, n_rest: int, n_scale: int = 3, n_rot: int = 4):
    """construct_list_of_attributes (geometry/gaussian_io.py:37-50)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation) -> None:
    """Raw parameters as the reference keeps them: xyz (P,3), features_dc (P,1,3), features_rest (P,M-1,3),
    opacity (P,1), scaling (P,3), rotation (P,4)."""
    np32 = lambda t: t.detach().float().cpu().numpy()
    xyz_, P = np32(xyz), xyz.shape[0]
    f_dc = np32(features_dc.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    f_rest = np32(features_rest.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    cols = np.concatenate((xyz_, np.zeros_like(xyz_), f_dc, f_rest, np32(opacity).reshape(P, -1), np32(scaling),
                           np32(rotation)), axis=1).astype("<f4")
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scaling.shape[1], rotation.shape[1])
    assert cols.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(cols).tobytes())


def read_vertex_table(path) -> Dict[str, np.ndarray]:
    """{property name: (P,) array} of the file's ``vertex`` element (scalar properties only)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, current = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                current = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(current)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported")
                current["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: the first element must be 'vertex'")
        el = elements[0]
        if fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=el["count"], ndmin=2)
            return {n: rows[:, i].astype(t) for i, (n, t) in enumerate(el["props"])}
        order = {"binary_little_endian": "<", "binary_big_endian": ">"}.get(fmt)
        if order is None:
            raise ValueError(f"{path}: unknown PLY format {fmt!r}")
        dt = np.dtype([(n, order + t) for n, t in el["props"]])
        data = np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"])
        return {n: np.asarray(data[n]) for n, _ in el["props"]}


def load_ply(path, max_sh_degree: int, device="cpu") -> Dict[str, torch.Tensor]:
    """Returns the raw parameter tensors in the reference's shapes (geometry/gaussian_io.py:86-172):
    xyz (P,3), features_dc (P,1,3), features_rest (P,(D+1)^2-1,3), opacity (P,1), scaling (P,3), rotation (P,4)."""
    tab = read_vertex_table(path)
    col = lambda n: np.asarray(tab[n], dtype=np.float64)
    xyz = np.stack((col("x"), col("y"), col("z")), axis=1)
    P = xyz.shape[0]
    opacities = col("opacity")[..., np.newaxis]
    features_dc = np.zeros((P, 3, 1))
    for c in range(3):
        features_dc[:, c, 0] = col(f"f_dc_{c}")
    by_index = lambda prefix: sorted((n for n in tab if n.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    K = (max_sh_degree + 1) ** 2 - 1
    if max_sh_degree > 0:
        names = by_index("f_rest_")
        if len(names) != 3 * K:
            raise ValueError(f"{path}: {len(names)} f_rest properties, expected {3 * K} for SH degree {max_sh_degree}")
        features_extra = np.stack([col(n) for n in names], axis=1).reshape(P, 3, K)
    else:
        features_extra = np.zeros((P, 3, 0))
    scales = np.stack([col(n) for n in by_index("scale_")], axis=1)
    rots = np.stack([col(n) for n in by_index("rot")], axis=1)
    t = lambda a: torch.tensor(a, dtype=torch.float, device=device)
    return dict(xyz=t(xyz), features_dc=t(features_dc).transpose(1, 2).contiguous(),
                features_rest=t(features_extra).transpose(1, 2).contiguous(), opacity=t(opacities), scaling=t(scales),
                rotation=t(rots))

"""CPU: b200splat.ply against the reference's own ``GaussianIO.save_ply`` / ``load_ply`` code (imported unchanged from
/root/reference, running on tests/stubs/plyfile), plus format round trips that need no reference."""
import warnings

import numpy as np
import pytest
import torch

import ref_harness as H
from b200splat import ply

NAMES = ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")


def _params(P, deg, seed):
    g = torch.Generator().manual_seed(seed)
    K = (deg + 1) ** 2 - 1
    return dict(xyz=torch.randn(P, 3, generator=g), features_dc=torch.randn(P, 1, 3, generator=g),
                features_rest=torch.randn(P, K, 3, generator=g), opacity=torch.randn(P, 1, generator=g),
                scaling=torch.randn(P, 3, generator=g), rotation=torch.randn(P, 4, generator=g))


@pytest.mark.parametrize("deg", [0, 1, 3])
def test_round_trip_and_layout(tmp_path, deg):
    p = _params(37, deg, 5)
    path = tmp_path / "a.ply"
    ply.save_ply(path, **p)
    head = path.read_bytes().split(b"end_header\n")[0].decode().splitlines()
    assert head[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 37"]
    props = [l.split()[-1] for l in head[3:]]
    assert props == ply.attribute_names(3, 3 * ((deg + 1) ** 2 - 1))
    assert all(l.split()[1] == "float" for l in head[3:])
    back = ply.load_ply(path, deg)
    for k in NAMES:
        assert torch.equal(back[k], p[k]), k
    tab = ply.read_vertex_table(path)
    assert np.all(tab["nx"] == 0)
    if deg:   # channel-major storage: f_rest_0..K-1 are the red channel of coefficients 1..K
        K = (deg + 1) ** 2 - 1
        assert np.array_equal(tab["f_rest_1"], p["features_rest"][:, 1, 0].numpy())
        assert np.array_equal(tab[f"f_rest_{K}"], p["features_rest"][:, 0, 1].numpy())


def test_ascii_and_big_endian_files_are_read(tmp_path):
    p = _params(5, 1, 6)
    path = tmp_path / "b.ply"
    ply.save_ply(path, **p)
    tab = ply.read_vertex_table(path)
    names = list(tab)
    rows = np.stack([tab[n] for n in names], 1)
    txt = tmp_path / "t.ply"
    txt.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 5\n" +
                   "".join(f"property float {n}\n" for n in names) + "end_header\n" +
                   "\n".join(" ".join(repr(float(x)) for x in r) for r in rows) + "\n")
    big = tmp_path / "g.ply"
    big.write_bytes(("ply\nformat binary_big_endian 1.0\nelement vertex 5\n" +
                     "".join(f"property float {n}\n" for n in names) + "end_header\n").encode() +
                    rows.astype(">f4").tobytes())
    for f in (txt, big):
        back = ply.load_ply(f, 1)
        for k in NAMES:
            assert torch.allclose(back[k], p[k], rtol=0, atol=0), (f.name, k)


@pytest.mark.skipif(not H.available(), reason="/root/reference not present")
@pytest.mark.parametrize("deg", [0, 2])
def test_interchange_with_the_reference_gaussian_io(tmp_path, deg):
    import importlib
    warnings.filterwarnings("ignore")
    H.load("oracle")
    io_mod = importlib.import_module("ref3dgs.geometry.gaussian_io")
    p = _params(64, deg, 7)

    class Holder(io_mod.GaussianIO):
        pass

    # reference writes -> we read
    h = Holder()
    h._xyz, h._features_dc, h._features_rest = p["xyz"], p["features_dc"], p["features_rest"]
    h._opacity, h._scaling, h._rotation = p["opacity"], p["scaling"], p["rotation"]
    ref_file = tmp_path / "ref.ply"
    h.save_ply(str(ref_file))
    back = ply.load_ply(ref_file, deg)
    for k in NAMES:
        assert torch.equal(back[k], p[k]), k
    # we write -> reference reads; and both writers produce the same bytes
    ours = tmp_path / "ours.ply"
    ply.save_ply(ours, **p)
    assert ours.read_bytes() == ref_file.read_bytes()
    h2 = Holder()
    h2.max_sh_degree = deg
    with H.CudaToCpu():
        h2.load_ply(str(ours))
    got = dict(xyz=h2._xyz, features_dc=h2._features_dc, features_rest=h2._features_rest, opacity=h2._opacity,
               scaling=h2._scaling, rotation=h2._rotation)
    for k in NAMES:
        assert torch.equal(got[k].detach(), p[k]), k

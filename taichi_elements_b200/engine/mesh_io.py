"""PLY helpers with the reference's on-disk layout (engine/mesh_io.py:5-44)."""
import numpy as np


def load_mesh(fn, scale=1, offset=(0, 0, 0)):
    """PLY triangle mesh -> (num_tris, 9) float32, vertices scaled then shifted
    (reference engine/mesh_io.py:5-24).  Needs the optional `plyfile` package."""
    try:
        from plyfile import PlyData
    except ImportError as e:  # pragma: no cover - plyfile is not in this image
        raise ImportError('load_mesh needs the `plyfile` package') from e
    if isinstance(scale, (int, float)):
        scale = (scale, scale, scale)
    print(f'loading {fn}')
    ply = PlyData.read(fn)
    vert = ply['vertex']
    xyz = np.stack([np.asarray(vert[k], np.float64) for k in 'xyz'], axis=1)
    xyz = xyz * np.asarray(scale, np.float64) + np.asarray(offset, np.float64)
    faces = ply['face']['vertex_indices']
    tris = np.zeros((len(faces), 9), dtype=np.float32)
    for i, face in enumerate(faces):
        assert len(face) == 3
        tris[i] = xyz[np.asarray(face)].reshape(9)
    return tris


_PLY_HEADER = """ply
format binary_little_endian 1.0
comment Created by taichi
element vertex {n}
property float x
property float y
property float z
property uchar red
property uchar green
property uchar blue
property uchar placeholder
end_header
"""


def write_point_cloud(fn, pos_and_color):
    """Binary PLY of (n, 4) float32 rows: x, y, z and a packed 0x00BBGGRR word
    viewed as float (reference engine/mesh_io.py:27-44)."""
    with open(fn, 'wb') as f:
        f.write(_PLY_HEADER.format(n=len(pos_and_color)).encode())
        f.write(np.ascontiguousarray(pos_and_color).tobytes())

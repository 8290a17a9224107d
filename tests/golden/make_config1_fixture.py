"""Generate tests/golden/config1_cloud.npz — BASELINE config 1 (the reference's own CPU-runnable case) as a small
committed fixture, so that the GPU box (which has no /root/reference) can run the rigid_icp.cpp recipe on REAL scan data.

Source: the reference's bundled scan examples/test_clouds/test.ply (573 663 vertices with normals and colours),
read here with a few lines of numpy (binary little endian: float x y z, uchar r g b, float nx ny nz, float radius),
voxel-downsampled with the ORACLE's restatement of PointCloud::gridDownsample at 12 mm (the example uses 5 mm; a
coarser grid keeps the fixture under 1 MB). Stored: points, normals (float32), plus the oracle's 5 mm bin count as a
known answer for the CPU test. Run in the build container:  python tests/golden/make_config1_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PLY = "/root/reference/examples/test_clouds/test.ply"


def read_test_ply(path=PLY):
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            header += f.readline()
        n = int([ln for ln in header.decode().splitlines() if ln.startswith("element vertex")][0].split()[-1])
        dt = np.dtype([("p", "<f4", 3), ("c", "u1", 3), ("n", "<f4", 3), ("radius", "<f4")])
        v = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return (np.ascontiguousarray(v["p"]), np.ascontiguousarray(v["n"]),
            (np.float32(1.0 / 255.0) * v["c"].astype(np.float32)).astype(np.float32))


if __name__ == "__main__":
    import oracle

    oracle.build()
    pts, nrm, col = read_test_ply()
    p5, n5, _ = oracle.grid_downsample(pts, 0.005, normals=nrm, colors=col)
    p12, n12, _ = oracle.grid_downsample(pts, 0.012, normals=nrm)
    np.savez_compressed(os.path.join(HERE, "config1_cloud.npz"), points=p12, normals=n12,
                        n_source=np.int64(pts.shape[0]), n_bins_5mm=np.int64(p5.shape[0]))
    print(f"{pts.shape[0]} vertices -> {p5.shape[0]} bins at 5 mm, {p12.shape[0]} at 12 mm (fixture)")

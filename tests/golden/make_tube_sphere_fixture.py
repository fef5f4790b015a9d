"""Test infrastructure: the one geometry the reference ships (/root/reference/input/geometry/tube_sphere.dat, 12-byte
header (60, 60, 80) int32 + 288 000 int8 voxels, x fastest) as a compressed fixture, so that the GPU box - where
/root/reference does not exist - can run BASELINE configs[0].

    python tests/golden/make_tube_sphere_fixture.py      # run where /root/reference exists
"""
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/input/geometry/tube_sphere.dat")
DST = Path(__file__).resolve().parent / "tube_sphere_60_60_80.npz"

if __name__ == "__main__":
    raw = SRC.read_bytes()
    nx, ny, nz = (int(v) for v in np.frombuffer(raw, np.int32, 3))
    vox = np.frombuffer(raw, np.int8, nx * ny * nz, 12).reshape(nz, ny, nx)
    assert (nx, ny, nz) == (60, 60, 80) and len(raw) == 12 + nx * ny * nz
    np.savez_compressed(DST, solid=vox)
    print(DST, vox.shape, int(vox.sum()), "solid voxels")

"""write_obj / write_ply of the native library (csrc/meshio.cu, host code: no GPU needed).  The OBJ bytes must equal
what the reference's Python loop writes (src/isoext/utils.py:42-63 of the reference): `repr(float(x))` per coordinate."""
import struct

import numpy as np
import torch

from isoext_b200 import utils as U


def reference_write_obj(obj_path, v, f):
    """Restatement of src/isoext/utils.py:42-63 (the oracle for the bytes)."""
    with open(obj_path, "w") as obj_file:
        if v is None or f is None or v.numel() == 0 or f.numel() == 0:
            return
        lines = []
        for v0, v1, v2 in v.tolist():
            lines.append(f"v {v0} {v1} {v2}\n")
        for f0, f1, f2 in (f + 1).tolist():
            lines.append(f"f {f0} {f1} {f2}\n")
        obj_file.writelines(lines)


def _special_vertices():
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 1e-4, 9.9999e-5, 1e-5, 123456.789, 1e15, 1e16, 9.999999e15, 1.5e16, 3.4028235e38, 1e-38,
            1.4e-45, 0.5, 2.0 / 3.0, 16777216.0, 1e7, 0.001, 12345678.0, float("inf"), float("-inf"), float("nan"), 1e22, 5e-324]
    a = torch.tensor(vals, dtype=torch.float32)
    n = (len(a) + 2) // 3 * 3
    return torch.cat([a, torch.zeros(n - len(a))]).view(-1, 3)


def test_write_obj_bytes_equal_the_reference_loop(tmp_path):
    gen = torch.Generator().manual_seed(0)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (60000, 3), generator=gen, dtype=torch.int64).to(torch.int32)
    rnd = bits.view(torch.float32)                                   # every exponent, both signs, denormals, NaNs
    grid = (torch.arange(3000, dtype=torch.float32) / 2999 * 2 - 1).repeat(3, 1).t().contiguous()
    v = torch.cat([_special_vertices(), rnd, grid, torch.randn((50000, 3), generator=gen)])
    f = torch.randint(0, len(v), (70001, 3), generator=gen, dtype=torch.int64).to(torch.int32)
    a, b = tmp_path / "ours.obj", tmp_path / "ref.obj"
    U.write_obj(str(a), v, f)
    reference_write_obj(str(b), v, f)
    assert a.read_bytes() == b.read_bytes()
    # small mesh (single-thread path), int64 faces, empty and None meshes
    U.write_obj(a, v[:5], f[:2].long() % 5)
    reference_write_obj(b, v[:5], f[:2].long() % 5)
    assert a.read_bytes() == b.read_bytes()
    U.write_obj(a, None, None)
    assert a.read_bytes() == b"" 
    U.write_obj(a, torch.zeros((0, 3)), torch.zeros((0, 3), dtype=torch.int32))
    assert a.read_bytes() == b""


def test_write_ply_round_trip(tmp_path):
    gen = torch.Generator().manual_seed(1)
    v = torch.randn((1000, 3), generator=gen)
    f = torch.randint(0, 1000, (1999, 3), generator=gen, dtype=torch.int64).to(torch.int32)
    p = tmp_path / "m.ply"
    U.write_ply(p, v, f)
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex 1000" in head and b"element face 1999" in head and b"binary_little_endian" in head
    vv = np.frombuffer(body[:12000], dtype="<f4").reshape(-1, 3)
    assert np.array_equal(vv.view(np.uint32), v.numpy().view(np.uint32))
    rec = np.frombuffer(body[12000:], dtype=np.uint8).reshape(-1, 13)
    assert (rec[:, 0] == 3).all()
    ff = np.frombuffer(rec[:, 1:].tobytes(), dtype="<i4").reshape(-1, 3)
    assert np.array_equal(ff, f.numpy())
    assert struct.calcsize("<B3i") == 13

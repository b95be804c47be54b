"""The pure-Python HDF5 writer against the pure-Python reader (both replace h5py / libhdf5,
which are absent where the engine runs).  The reader itself is pinned to real HDF5 files by
tests/golden/make_golden.py, which parses the reference's fixtures with it."""
import numpy as np

from hyperion_b200.io import h5min, h5write


def test_roundtrip_groups_datasets_attrs_links(tmp_path):
    fn = str(tmp_path / "t.h5")
    f = h5write.File()
    f.attrs["date_started"] = "2026-01-01"
    f.attrs["converged"] = b"no"
    f.attrs["iterations"] = np.int32(5)
    f.attrs["cpu_time"] = 1.25
    f.attrs["vec"] = np.arange(3, dtype=np.float64)
    g = f.create_group("iteration_00001")
    se = np.random.default_rng(0).random((2, 3, 5, 7))
    d = g.create_dataset("specific_energy", se)
    d.attrs["geometry"] = "abcdef0123456789abcdef0123456789"
    g.attrs["killed_photons_geo"] = np.int64(0)
    f["Input"] = h5write.ExternalLink("model.rtin", "/")
    f["alias"] = h5write.SoftLink("/iteration_00001")
    tab = np.zeros(4, dtype=[("nu", "<f8"), ("P1", "<f8", (3,)), ("n", "<i4")])
    tab["nu"] = [1., 2., 3., 4.]
    tab["P1"] = np.arange(12).reshape(4, 3)
    tab["n"] = [7, 8, 9, 10]
    f.create_dataset("Dust/dust_001/optical_properties", tab)
    f.create_dataset("f4", np.arange(6, dtype=np.float32).reshape(2, 3))
    f.create_dataset("empty", np.zeros((0,), dtype=np.float64))
    f.write(fn)

    r = h5min.File(fn)
    assert bytes(r.attrs["converged"]).strip(b"\x00") == b"no"
    assert int(r.attrs["iterations"]) == 5 and float(r.attrs["cpu_time"]) == 1.25
    assert np.array_equal(r.attrs["vec"], [0., 1., 2.])
    assert sorted(r.keys()) == ["Dust", "Input", "alias", "empty", "f4", "iteration_00001"]
    got = r["iteration_00001/specific_energy"]
    assert got.shape == (2, 3, 5, 7) and np.array_equal(got[...], se)
    assert bytes(got.attrs["geometry"]).decode() == "abcdef0123456789abcdef0123456789"
    assert int(r["iteration_00001"].attrs["killed_photons_geo"]) == 0
    link = r.get_link("Input")
    assert isinstance(link, h5min.ExternalLink) and link.filename == "model.rtin"
    assert np.array_equal(r["alias/specific_energy"][...], se)
    t2 = r["Dust/dust_001/optical_properties"][...]
    assert t2.dtype.names == ("nu", "P1", "n")
    assert np.array_equal(t2["P1"], tab["P1"]) and np.array_equal(t2["n"], tab["n"])
    assert r["f4"][...].dtype == np.float32 and np.array_equal(r["f4"][...], np.arange(6).reshape(2, 3))
    assert r["empty"][...].shape == (0,)


def test_superblock_and_header_layout(tmp_path):
    """Byte-level checks of the fixed parts of the format (HDF5 file format spec, version 0
    superblock / version 1 object header), so that libhdf5 can open what we write."""
    fn = str(tmp_path / "t.h5")
    f = h5write.File()
    f.create_dataset("x", np.arange(4, dtype=np.int64))
    f.write(fn)
    b = open(fn, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8] == 0 and b[13] == 8 and b[14] == 8           # superblock v0, 8-byte offsets/lengths
    eof = int.from_bytes(b[40:48], "little")
    assert eof == len(b)
    root = int.from_bytes(b[64:72], "little")
    assert root % 8 == 0 and b[root] == 1                    # object header version 1, aligned
    nmsg = int.from_bytes(b[root + 2:root + 4], "little")
    size = int.from_bytes(b[root + 8:root + 12], "little")
    p, end, types = root + 16, root + 16 + size, []
    while p < end:
        t, n = int.from_bytes(b[p:p + 2], "little"), int.from_bytes(b[p + 2:p + 4], "little")
        assert n % 8 == 0
        types.append(t)
        p += 8 + n
    assert p == end and len(types) == nmsg
    assert types[:2] == [0x0002, 0x000A] and 0x0006 in types  # link info, group info, link

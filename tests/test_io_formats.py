"""Phantom dump container of libshamb200 (csrc/io_formats.cu) against the oracle restatement (oracle/io_formats.py)
on the CPU: the reference's own test of its reader / writer (src/tests/phantom_read_test.cpp:20-34 — read a dump,
write it again, `cmp`) on synthetic dumps that use all eight element types and two blocks; header queries
(PhantomDump.hpp read_header_float / read_header_int / has_header_entry), gen_config_from_phantom_dump
(Model.cpp:1203-1222) and compare_phantom_dumps (PhantomDump.cpp:389-467).  No device needed."""
import struct

import numpy as np
import pytest

from oracle import io_formats as O
from shamrock_b200 import _capi


def synthetic_dump(seed=0, n0=257, n1=3, ieos=2, periodic=True):
    rng = np.random.default_rng(seed)
    ph = O.PhantomDump()
    ph.fileid = "FT:Phantom synthetic dump (tests)".ljust(100)
    ph.add("fort_int", "nparttot", n0)
    ph.add("fort_int", "isink", 0)
    ph.add("i8", "tiny", -3)
    ph.add("i16", "small", -1234)
    ph.add("i32", "ieos", ieos)
    ph.add("i32", "isink", 0)
    ph.add("i64", "nparttot", n0)
    for tag, v in (("gamma", 1.4), ("RK2", 1.5 * 0.25), ("polyk2", 0.0), ("qfacdisc", 0.35), ("qfacdisc2", 0.75),
                   ("time", 0.375), ("C_cour", 0.3), ("C_force", 0.25), ("alphau", 0.9), ("massoftype", 1e-5),
                   ("massoftype", 0.0), ("hfact", 1.2)):
        ph.add("fort_real", tag, v)
    if periodic:
        for tag, v in (("xmin", -1.0), ("xmax", 1.0), ("ymin", -0.5), ("ymax", 0.5), ("zmin", -0.25), ("zmax", 0.25)):
            ph.add("fort_real", tag, v)
    ph.add("f32", "single", 0.5)
    for tag, v in (("udist", 1.0), ("umass", 1.0), ("utime", 1.0), ("umagfd", 3.54491)):
        ph.add("f64", tag, v)
    b0 = {"tot_count": n0, "arrays": {
        "fort_int": [("itype", rng.integers(0, 8, n0))],
        "i8": [("iphase", rng.integers(-5, 5, n0))],
        "i16": [("flags", rng.integers(-300, 300, n0))],
        "i64": [("ids", rng.integers(0, 2**40, n0))],
        "fort_real": [(t, rng.uniform(-0.2, 0.2, n0)) for t in ("x", "y", "z", "vx", "vy", "vz", "u")],
        "f32": [("h", rng.uniform(0.01, 0.02, n0)), ("alpha", rng.uniform(0, 1, n0))],
        "f64": [("extra", rng.normal(size=n0))]}}
    b1 = {"tot_count": n1, "arrays": {"fort_real": [(t, rng.normal(size=n1)) for t in
                                                    ("x", "y", "z", "m", "h", "vx", "vy", "vz")]}}
    ph.blocks = [b0, b1]
    return ph


def test_oracle_round_trip():
    ph = synthetic_dump()
    raw = ph.gen_file()
    again = O.PhantomDump.from_bytes(raw)
    assert again.gen_file() == raw
    assert again.header("ieos") == ("i32", 2) and again.header("tiny") == ("i8", -3)
    assert np.array_equal(again.array(0, "x"), ph.blocks[0]["arrays"]["fort_real"][0][1])
    # record framing: the first record is (i1, r1, i2, iversion, i3) = 24 bytes between two byte counts
    assert struct.unpack_from("<i", raw, 0)[0] == 24 and struct.unpack_from("<i", raw, 28)[0] == 24
    assert struct.unpack_from("<idiii", raw, 4) == (60769, 60878.0, 60878, 1, 690706)


@pytest.mark.parametrize("seed,n0,n1", [(0, 257, 3), (1, 1, 0), (2, 0, 0), (3, 4099, 17)])
def test_read_write_is_byte_identical(tmp_path, seed, n0, n1):
    """phantom_read_test.cpp: from_file -> gen_file -> cmp"""
    fin, fout = tmp_path / "in.phdump", tmp_path / "out.phdump"
    raw = synthetic_dump(seed, n0, n1).gen_file()
    fin.write_bytes(raw)
    _capi.phantom_copy(fin, fout)
    assert fout.read_bytes() == raw


def test_header_queries(tmp_path):
    f = tmp_path / "d.phdump"
    f.write_bytes(synthetic_dump().gen_file())
    assert _capi.phantom_header_float(f, "gamma") == 1.4
    assert _capi.phantom_header_float(f, "single") == 0.5  # f32 table
    assert _capi.phantom_header_float(f, "udist") == 1.0  # f64 table
    assert _capi.phantom_header_float(f, "massoftype") == 0.0  # fetch: the last entry with the tag
    assert _capi.phantom_header_int(f, "ieos") == 2
    assert _capi.phantom_header_int(f, "tiny") == -3 and _capi.phantom_header_int(f, "small") == -1234
    assert _capi.phantom_header_int(f, "nparttot") == 257  # fort_int before i64
    assert _capi.phantom_header_float(f, "nothing") is None
    with pytest.raises(_capi.ShamB200Error):  # an integer entry is not a float entry (read_header_float)
        _capi.phantom_header_float(f, "ieos")


@pytest.mark.parametrize("ieos", [1, 2, 3])
@pytest.mark.parametrize("periodic", [True, False])
def test_gen_config(tmp_path, ieos, periodic):
    f = tmp_path / "d.phdump"
    f.write_bytes(synthetic_dump(ieos=ieos, periodic=periodic).gen_file())
    cfg = _capi.phantom_gen_config(f)
    assert cfg.gpart_mass == 1e-5 and cfg.cfl_cour == 0.3 and cfg.cfl_force == 0.25
    assert cfg.bc == (1 if periodic else 0)
    assert (cfg.av, cfg.alpha_min, cfg.alpha_max, cfg.sigma_decay, cfg.alpha_u, cfg.beta_AV) == (3, 0.0, 1.0, 0.1, 0.9, 2.0)
    polyk = 2.0 / 3.0 * (1.5 * 0.25)
    if ieos == 1:
        assert cfg.eos == 1 and cfg.cs0 == np.sqrt(polyk)
    elif ieos == 2:
        assert cfg.eos == 0 and cfg.gamma == 1.4
    else:
        assert cfg.eos == 2 and cfg.cs0 == np.sqrt(polyk) and cfg.eos_q == 0.35 and cfg.eos_r0 == 1.0


def test_gen_config_unknown_eos(tmp_path):
    f = tmp_path / "d.phdump"
    f.write_bytes(synthetic_dump(ieos=15).gen_file())
    with pytest.raises(_capi.ShamB200Error, match="ieos=15"):
        _capi.phantom_gen_config(f)
    cfg = _capi.phantom_gen_config(f, bypass_error=True)  # warning only; the rest of the configuration is read
    assert cfg.gpart_mass == 1e-5


def test_compare(tmp_path):
    a, b, c = tmp_path / "a", tmp_path / "b", tmp_path / "c"
    a.write_bytes(synthetic_dump(seed=0).gen_file())
    b.write_bytes(synthetic_dump(seed=5, n0=11).gen_file())  # other particles, one other header value (nparttot)
    c.write_bytes(synthetic_dump(seed=0, periodic=False).gen_file())  # six entries missing
    assert _capi.phantom_compare(a, a) == 0
    assert _capi.phantom_compare(a, b) == 1
    assert _capi.phantom_compare(a, c) == 6 and _capi.phantom_compare(c, a) == 6


@pytest.mark.parametrize("damage", ["magic", "marker", "truncated"])
def test_damaged_files_are_refused(tmp_path, damage):
    raw = bytearray(synthetic_dump().gen_file())
    if damage == "magic":
        raw[4:8] = struct.pack("<i", 1234)
    elif damage == "marker":
        raw[28:32] = struct.pack("<i", 25)  # closing byte count of the first record
    else:
        raw = raw[: len(raw) // 2]
    f = tmp_path / "bad.phdump"
    f.write_bytes(bytes(raw))
    with pytest.raises(_capi.ShamB200Error):
        _capi.phantom_copy(f, tmp_path / "out")


# ---- the record layer against the reference's own FortranIOFile (oracle/_ref/fortran_io_ref) --------------------
def _ref(args):
    import subprocess

    exe = O.ref_binary()
    if exe is None:
        pytest.skip("oracle/_ref/fortran_io_ref is not built (needs /root/reference)")
    r = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
    return r.returncode, r.stderr


def test_reference_io_writes_what_the_oracle_writes(tmp_path):
    """a dump written record by record with the reference's FortranIOFile == gen_file() of the restatement"""
    f = tmp_path / "ref.phdump"
    assert _ref(["write", f]) == (0, "")
    raw = f.read_bytes()
    assert raw == O.ref_synthetic_dump().gen_file()
    # the library reads the reference-written file and writes it back unchanged
    _capi.phantom_copy(f, tmp_path / "lib.phdump")
    assert (tmp_path / "lib.phdump").read_bytes() == raw
    assert _capi.phantom_header_int(f, "tag_1_1") == 6 and _capi.phantom_header_float(f, "tag_6_2") == 12.0


@pytest.mark.parametrize("seed,n0,n1", [(0, 257, 3), (2, 0, 0), (3, 4099, 17)])
def test_reference_io_reads_what_oracle_and_library_write(tmp_path, seed, n0, n1):
    """phantom_read_test.cpp with the reference's record IO: read every record with FortranIOFile, write it again"""
    a, b, c = tmp_path / "oracle.phdump", tmp_path / "lib.phdump", tmp_path / "ref.phdump"
    raw = synthetic_dump(seed, n0, n1).gen_file()
    a.write_bytes(raw)
    _capi.phantom_copy(a, b)  # written by the library
    assert _ref(["copy", b, c]) == (0, "")
    assert c.read_bytes() == raw


def test_reference_io_refuses_what_the_library_refuses(tmp_path):
    raw = bytearray(synthetic_dump().gen_file())
    raw[28:32] = struct.pack("<i", 25)  # closing byte count of the first record
    f = tmp_path / "bad.phdump"
    f.write_bytes(bytes(raw))
    rc, err = _ref(["copy", f, tmp_path / "out"])
    assert rc == 2 and "invalid" in err
    with pytest.raises(_capi.ShamB200Error):
        _capi.phantom_copy(f, tmp_path / "out2")


def test_model_shaped_dump_through_library_and_reference_io(tmp_path):
    """what Model::make_phantom_dump writes (oracle restatement of Model.cpp:1491-1638; the CUDA model's file equals
    these bytes: tests/test_gpu_io_formats.py) is read by the library's configuration path and by the reference's
    record IO class"""
    rng = np.random.default_rng(1)
    n = 1000
    fields = dict(xyz=rng.normal(size=(n, 3)), vxyz=rng.normal(size=(n, 3)), hpart=rng.uniform(0.01, 0.02, n),
                  uint=rng.uniform(1, 2, n), alpha_AV=rng.uniform(0, 1, n), divv=rng.normal(size=n))
    cfg = dict(eos="lp07", gamma=5 / 3, cs0=0.05, q=0.25, r0=2.0, av_has_alpha=True, time=0.125, dt=1e-3, hfact=1.2,
               cfl_cour=0.3, cfl_force=0.25, gpart_mass=1e-6, periodic=True, bmin=(-1, -2, -3), bmax=(1, 2, 3))
    f = tmp_path / "model.phdump"
    raw = O.make_phantom_dump(fields, cfg).gen_file()
    f.write_bytes(raw)
    assert _capi.phantom_header_int(f, "nparttot") == n and _capi.phantom_header_float(f, "hfact") == 1.2
    assert _capi.phantom_header_float(f, "ymax") == 1.0  # Phantom2Shamrock.cpp:203-209: bmax.x()
    c = _capi.phantom_gen_config(f)
    assert (c.eos, c.eos_q, c.eos_r0, c.bc, c.gpart_mass) == (2, 0.25, 1.0, 1, 1e-6)
    assert c.cs0 == np.sqrt(2.0 / 3.0 * (1.5 * (0.05 * 0.05 / 4.0)))  # polyk = cs0² / r0², RK2 = 1.5 polyk
    assert c.alpha_u == 1.0  # the writer's "alphau" entry
    rc, err = _ref(["copy", f, tmp_path / "copy"])
    assert (rc, err) == (0, "") and (tmp_path / "copy").read_bytes() == raw


def test_error_behaviour_of_the_file_level_calls(tmp_path):
    missing = tmp_path / "nothing.phdump"
    for call in (lambda: _capi.phantom_copy(missing, tmp_path / "o"), lambda: _capi.phantom_gen_config(missing),
                 lambda: _capi.phantom_header_int(missing, "ieos"), lambda: _capi.phantom_compare(missing, missing)):
        with pytest.raises(_capi.ShamB200Error, match="cannot open"):
            call()
    f = tmp_path / "d.phdump"
    ph = synthetic_dump()
    ph.tables["fort_real"] = [e for e in ph.tables["fort_real"] if e[0].strip() != "C_cour"]
    f.write_bytes(ph.gen_file())
    with pytest.raises(_capi.ShamB200Error, match="C_cour"):  # read_header_float: the entry cannot be found
        _capi.phantom_gen_config(f)
    L = _capi.lib()
    assert L.shamb200_phantom_copy(None, None) == -1 and b"null" in L.shamb200_last_error()

"""Oracle (TEST INFRASTRUCTURE, never imported by shamrock_b200): the reference's Phantom-dump container, its
SPH model's Phantom header / particle block, and its legacy VTK writer, restated with numpy / struct.

ref (paths relative to /root/reference/src):
  shambase/include/shambase/fortran_io.hpp:100-270       FortranIOFile: [i32 n][payload][i32 n] records
  shammodels/sph/include/shammodels/sph/io/PhantomDump.hpp  tables, blocks, fetch (last entry wins), fill_vec,
                                                            magic numbers 60769 / 60878 / 690706
  shammodels/sph/src/io/PhantomDump.cpp:34-107,152-218    table / block array read + write
  shammodels/sph/src/io/PhantomDump.cpp:276-375           gen_file / from_file
  shammodels/sph/src/Model.cpp:1432-1488                  add_pdat_to_phantom_block
  shammodels/sph/src/Model.cpp:1491-1638                  make_phantom_dump
  shammodels/sph/src/io/PhantomDumpEOSUtils.cpp:43-60,170-247   EOS header entries
  shammodels/sph/src/io/Phantom2Shamrock.cpp:147-233      units (SI default) and boundary entries
  shamrock/include/shamrock/io/LegacyVtkWriter.hpp:160-420, shammodels/common/.../io/VTKDumpUtils.hpp:42-160,
  shammodels/sph/src/modules/io/VTKDump.cpp:36-178        legacy VTK: header text, big-endian f32 / i32 sections

Parity pin: (1) the record layer — byte counts, argument order, fixed strings, string arrays, value arrays of all
eight element types — against the REFERENCE'S OWN CODE run here: oracle/_ref/fortran_io_ref is a driver compiled
against shambase/fortran_io.hpp where it lies (ref_fortran_io.cpp, `make -C oracle ref`); a dump it writes with
FortranIOFile equals the bytes gen_file() below builds for the same content, it reads the dumps written by this
module and by the library, and its copy of them is byte-identical (tests/test_io_formats.py).  (2) the container on
the reference's own test of it (src/tests/phantom_read_test.cpp: read -> gen_file -> `cmp`) through the library's
reader / writer.  (3) the model's header tables and particle block, and the VTK file, on the reference's source
only: PhantomDump.cpp / Model.cpp / VTKDump.cpp need SYCL and cannot be compiled here, and no Phantom or VTK file
written by the reference exists offline — for the CONTENT of a model dump the comparison with a reference-written
file stays "parity unpinned".  Where the reference writes uninitialised memory (isink, polyk2 and, for the adiabatic EOS,
polyk of EOSPhConfig) 0 is used."""
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref", "fortran_io_ref")


def ref_binary(build=True):
    """oracle/_ref/fortran_io_ref: the driver around the REFERENCE's FortranIOFile (ref_fortran_io.cpp), built from
    /root/reference where that exists (this container); elsewhere the prebuilt file that travelled, or None"""
    src = os.path.join(_HERE, "ref_fortran_io.cpp")
    hdr = "/root/reference/src/shambase/include/shambase/fortran_io.hpp"
    if build and os.path.exists(hdr) and (
            not os.path.exists(REF_BIN) or os.path.getmtime(REF_BIN) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return REF_BIN if os.path.exists(REF_BIN) else None


def ref_bmi(values):
    """expand_bits<u32,2>, expand_bits<u64,2>, contract_bits<u32,2>, contract_bits<u64,2> of the reference
    (shammath/sfc/bmi.hpp through oracle/_ref/bmi_ref, ref_bmi.cpp) for every value; None if the driver is absent"""
    ref_binary()
    exe = os.path.join(_HERE, "_ref", "bmi_ref")
    if not os.path.exists(exe):
        return None
    out = subprocess.run([exe], input="\n".join(str(int(v)) for v in values), capture_output=True, text=True,
                         check=True).stdout
    return np.array([[int(t) for t in line.split()] for line in out.strip().splitlines()], dtype=np.uint64)


def ref_units(unit_time, unit_length, unit_mass):
    """(G, year, au, sol_mass) of the reference's shamunits::Constants in these code units (oracle/_ref/units_ref,
    ref_units.cpp); None if the driver is absent"""
    ref_binary()
    exe = os.path.join(_HERE, "_ref", "units_ref")
    if not os.path.exists(exe):
        return None
    out = subprocess.run([exe, repr(float(unit_time)), repr(float(unit_length)), repr(float(unit_mass))],
                         capture_output=True, text=True, check=True).stdout
    return tuple(float(t) for t in out.split())


def ref_synthetic_dump():
    """the content `fortran_io_ref write` puts into its file (ref_fortran_io.cpp), as a PhantomDump of this module"""
    ph = PhantomDump()
    ph.fileid = "FT:Phantom reference-IO pin".ljust(100)
    for t, name in enumerate(TYPES):
        fl = name in ("fort_real", "f32", "f64")
        for k in range(t + 1):
            ph.add(name, f"tag_{t}_{k}", (k + 1) * (t + 2) * (0.5 if fl else 1))
    for b, (tot, counts) in enumerate(((37, (1, 1, 1, 1, 1, 2, 2, 1)), (3, (0, 0, 0, 0, 0, 3, 0, 0)))):
        arrays = {}
        i = np.arange(tot)
        for t, name in enumerate(TYPES):
            fl = name in ("fort_real", "f32", "f64")
            for j in range(counts[t]):
                v = ((7 * i + 3 * j + t) % 101 - 50) * (0.25 if fl else 1)
                arrays.setdefault(name, []).append((f"arr_{b}_{t}_{j}", v))
        ph.blocks.append({"tot_count": tot, "arrays": arrays})
    return ph

TYPES = ("fort_int", "i8", "i16", "i32", "i64", "fort_real", "f32", "f64")
DTYPE = {"fort_int": "<i4", "i8": "i1", "i16": "<i2", "i32": "<i4", "i64": "<i8", "fort_real": "<f8", "f32": "<f4",
         "f64": "<f8"}


def pad16(s):
    return s if len(s) >= 16 else s + " " * (16 - len(s))


def _rec(payload):
    n = struct.pack("<i", len(payload))
    return n + payload + n


class PhantomDump:
    def __init__(self):
        self.i1, self.i2, self.iversion, self.i3, self.r1 = 60769, 60878, 1, 690706, 60878.0
        self.fileid = ""
        self.tables = {t: [] for t in TYPES}  # type -> [(tag16, value)]
        self.blocks = []  # [{"tot_count": n, "arrays": {type: [(tag16, np.ndarray)]}}]

    def add(self, t, tag, val):
        self.tables[t].append((pad16(tag), val))

    # -- PhantomDump::gen_file
    def gen_file(self):
        out = [_rec(struct.pack("<idiii", self.i1, self.r1, self.i2, self.iversion, self.i3)),
               _rec(self.fileid.ljust(100).encode()[:100])]
        for t in TYPES:
            ent = self.tables[t]
            out.append(_rec(struct.pack("<i", len(ent))))
            if not ent:
                continue
            out.append(_rec("".join(tag[:16] for tag, _ in ent).encode()))
            out.append(_rec(np.array([v for _, v in ent]).astype(DTYPE[t]).tobytes()))
        out.append(_rec(struct.pack("<i", len(self.blocks))))
        for b in self.blocks:
            counts = [len(b["arrays"].get(t, [])) for t in TYPES]
            out.append(_rec(struct.pack("<q8i", b["tot_count"], *counts)))
        for b in self.blocks:
            for t in TYPES:
                for tag, vals in b["arrays"].get(t, []):
                    out.append(_rec(pad16(tag)[:16].encode()))
                    out.append(_rec(np.asarray(vals).astype(DTYPE[t])[: b["tot_count"]].tobytes()))
        return b"".join(out)

    # -- PhantomDump::from_file
    @staticmethod
    def from_bytes(data):
        pos = 0

        def rec(expect=None):
            nonlocal pos
            (n,) = struct.unpack_from("<i", data, pos)
            if expect is not None and n != expect:
                raise ValueError("the byte count is not correct")
            payload = data[pos + 4: pos + 4 + n]
            (m,) = struct.unpack_from("<i", data, pos + 4 + n)
            if m != n:
                raise ValueError("fortran 4 bytes invalid")
            pos += 8 + n
            return payload

        ph = PhantomDump()
        ph.i1, ph.r1, ph.i2, ph.iversion, ph.i3 = struct.unpack("<idiii", rec(24))
        assert (ph.i1, ph.i2, ph.i3) == (60769, 60878, 690706) and ph.r1 == ph.i2
        ph.fileid = rec(100).decode()
        for t in TYPES:
            (nv,) = struct.unpack("<i", rec(4))
            if nv == 0:
                continue
            tags = rec(16 * nv).decode()
            vals = np.frombuffer(rec(nv * np.dtype(DTYPE[t]).itemsize), dtype=DTYPE[t])
            ph.tables[t] = [(tags[16 * k: 16 * k + 16], vals[k].item()) for k in range(nv)]
        (nb,) = struct.unpack("<i", rec(4))
        heads = [struct.unpack("<q8i", rec(40)) for _ in range(nb)]
        for h in heads:
            b = {"tot_count": h[0], "arrays": {}}
            for t, cnt in zip(TYPES, h[1:]):
                for _ in range(cnt):
                    tag = rec(16).decode()
                    vals = np.frombuffer(rec(h[0] * np.dtype(DTYPE[t]).itemsize), dtype=DTYPE[t])
                    b["arrays"].setdefault(t, []).append((tag, vals))
            ph.blocks.append(b)
        assert pos == len(data), "some data was not read"
        return ph

    def header(self, key):
        """(type, value) of the LAST entry with this tag in the first table that has it, None if absent"""
        k = pad16(key)
        for t in TYPES:
            hit = [v for tag, v in self.tables[t] if tag == k]
            if hit:
                return t, hit[-1]
        return None

    def array(self, iblock, name):
        k = pad16(name)
        out = [np.asarray(v, dtype=np.float64) for t in TYPES for tag, v in self.blocks[iblock]["arrays"].get(t, [])
               if pad16(tag) == k]
        return np.concatenate(out) if out else np.zeros(0)


def make_phantom_dump(fields, cfg):
    """Model::make_phantom_dump.  fields: name -> array in dump order (xyz, vxyz [n,3]; hpart, uint, alpha_AV, divv
    [n]); cfg: dict(eos = "adiabatic" | "isothermal" | "lp07", gamma, cs0, q, r0, av_has_alpha, time, dt, hfact,
    cfl_cour, cfl_force, gpart_mass, periodic, bmin, bmax)"""
    ph = PhantomDump()
    ph.fileid = "FT:Phantom Shamrock writer".ljust(100)
    n = len(fields["xyz"])
    for t in ("fort_int", "i64"):
        ph.add(t, "nparttot", n)
        ph.add(t, "ntypes", 8)
        ph.add(t, "npartoftype", n)
        for _ in range(7):
            ph.add(t, "npartoftype", 0)
    for tag, v in (("nblocks", 1), ("nptmass", 0), ("ndustlarge", 0), ("ndustsmall", 0), ("idust", 7), ("idtmax_n", 1),
                   ("idtmax_frac", 0), ("idumpfile", 0), ("majorv", 2023), ("minorv", 0), ("microv", 0), ("isink", 0)):
        ph.add("fort_int", tag, v)
    ph.add("i32", "iexternalforce", 0)
    # write_shamrock_eos_in_phantom_dump -> eosN_write -> write_headeropts_eos
    gamma, polyk, qfac = 1.0, 0.0, 0.75
    if cfg["eos"] == "isothermal":
        ieos, polyk = 1, cfg["cs0"] * cfg["cs0"]
    elif cfg["eos"] == "adiabatic":
        ieos, gamma = 2, cfg["gamma"]
    elif cfg["eos"] == "lp07":
        ieos, polyk, qfac = 3, cfg["cs0"] * cfg["cs0"] / (cfg["r0"] * cfg["r0"]), cfg["q"]
    else:
        raise ValueError("The current shamrock EOS is not implemented in phantom dump conversion")
    ph.add("i32", "ieos", ieos)
    ph.add("i32", "isink", 0)
    for tag, v in (("gamma", gamma), ("RK2", 1.5 * polyk), ("polyk2", 0.0), ("qfacdisc", qfac), ("qfacdisc2", 0.75)):
        ph.add("fort_real", tag, v)
    for tag, v in (("time", cfg["time"]), ("dtmax", cfg["dt"]), ("rhozero", 0.0), ("hfact", cfg["hfact"]),
                   ("tolh", 0.0001), ("C_cour", cfg["cfl_cour"]), ("C_force", cfg["cfl_force"]), ("alpha", 0.0),
                   ("alphau", 1.0), ("alphaB", 1.0), ("massoftype", cfg["gpart_mass"])):
        ph.add("fort_real", tag, v)
    for _ in range(7):
        ph.add("fort_real", "massoftype", 0.0)
    for tag in ("Bextx", "Bexty", "Bextz", "dum"):
        ph.add("fort_real", tag, 0.0)
    if cfg["periodic"]:  # Phantom2Shamrock.cpp:203-209 writes bmax.x() for ymax and zmax
        bmin, bmax = cfg["bmin"], cfg["bmax"]
        for tag, v in (("xmin", bmin[0]), ("xmax", bmax[0]), ("ymin", bmin[1]), ("ymax", bmax[0]), ("zmin", bmin[2]),
                       ("zmax", bmax[0])):
            ph.add("fort_real", tag, v)
    for tag, v in (("get_conserv", -1.0), ("etot_in", 0.59762), ("angtot_in", 0.0189694), ("totmom_in", 0.0306284)):
        ph.add("fort_real", tag, v)
    for tag, v in (("udist", 1.0), ("umass", 1.0), ("utime", 1.0), ("umagfd", 3.54491)):  # no unit system: SI
        ph.add("f64", tag, v)
    xyz, v = np.asarray(fields["xyz"]).reshape(n, 3), np.asarray(fields["vxyz"]).reshape(n, 3)
    real = [("x", xyz[:, 0]), ("y", xyz[:, 1]), ("z", xyz[:, 2]), ("vx", v[:, 0]), ("vy", v[:, 1]), ("vz", v[:, 2]),
            ("u", np.asarray(fields["uint"]))]
    f32 = [("h", np.asarray(fields["hpart"]))]
    if cfg["av_has_alpha"]:
        f32 += [("alpha", np.asarray(fields["alpha_AV"])), ("divv", np.asarray(fields["divv"]))]
    ph.blocks.append({"tot_count": n, "arrays": {"fort_real": real, "f32": f32}})
    return ph


def vtk_dump_bytes(fields, cfg, add_patch_world_id, patch_ids=None, world_ranks=None):
    """modules::VTKDump::do_dump: the whole file.  fields as above + axyz, dtdivv, curlv, soundspeed; cfg: av in
    ("none", "constant", "mm97", "cd10", "disc"), eos, gpart_mass, hfact"""
    n = len(fields["hpart"])

    def f32be(a):
        return np.asarray(a, dtype=np.float64).astype(">f4").tobytes()

    def i32be(a):
        return np.asarray(a).astype(">i4").tobytes()

    has_alpha = cfg["av"] in ("mm97", "cd10")
    has_cd10 = cfg["av"] == "cd10"
    has_cs = has_alpha or cfg["eos"] == "lp07"
    out = [b"# vtk DataFile Version 4.2\nvtk output\nBINARY\nDATASET UNSTRUCTURED_GRID",
           f"\n\nPOINTS {n} float\n".encode(), f32be(fields["xyz"]), f"\n\nPOINT_DATA {n}".encode()]
    fnum = 5 + (2 if add_patch_world_id else 0) + (2 if has_alpha else 0) + (2 if has_cd10 else 0) + (1 if has_cs else 0)
    out.append(f"\nFIELD FieldData {fnum}".encode())

    def field(name, nvar, typ, payload):
        out.append(f"\n{name} {nvar} {n} {typ}\n".encode())
        out.append(payload)

    if add_patch_world_id:
        field("patchid", 1, "int", i32be(patch_ids))
        field("world_rank", 1, "int", i32be(world_ranks))
    field("h", 1, "float", f32be(fields["hpart"]))
    field("u", 1, "float", f32be(fields["uint"]))
    field("v", 3, "float", f32be(fields["vxyz"]))
    field("a", 3, "float", f32be(fields["axyz"]))
    if has_alpha:
        field("alpha_AV", 1, "float", f32be(fields["alpha_AV"]))
        field("divv", 1, "float", f32be(fields["divv"]))
    if has_cd10:
        field("dtdivv", 1, "float", f32be(fields["dtdivv"]))
        field("curlv", 3, "float", f32be(fields["curlv"]))
    if has_cs:
        field("soundspeed", 1, "float", f32be(fields["soundspeed"]))
    h = np.asarray(fields["hpart"], dtype=np.float64)
    hf = cfg["hfact"] / h
    field("rho", 1, "float", f32be(cfg["gpart_mass"] * hf * hf * hf))  # rho_h: m (hfact/h)(hfact/h)(hfact/h)
    return b"".join(out)

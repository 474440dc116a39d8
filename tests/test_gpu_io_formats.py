"""Phantom dumps and legacy VTK files of a model on the device (csrc/io_formats.cu) against the bytes the oracle
restatement builds from the same state (oracle/io_formats.py; Model.cpp:1432-1638, VTKDump.cpp:36-178), and the way
back: gen_config_from_phantom_dump + init_from_phantom_dump (Model.cpp:1203-1429, the protocol of
examples/sph/test_ph_dump_writer.py: load a dump, start a model from it, dump again, compare)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import io_formats as O  # noqa: E402
from shamrock_b200 import _capi  # noqa: E402
from tests import scenarios as S  # noqa: E402

NAMES = ("xyz", "vxyz", "axyz", "hpart", "uint", "alpha_AV", "divv", "dtdivv", "curlv", "soundspeed")
AV = {1: "constant", 2: "mm97", 3: "cd10"}
HFACT = {"M4": 1.2, "M6": 1.0}


def all_fields(m):
    """the model's fields in dump order: patch by patch in list order"""
    out = {}
    for nm in NAMES:
        parts = [m.get(ip, nm) for ip in range(m.patch_count) if m.patch_size(ip)]
        nv = 3 if nm in ("xyz", "vxyz", "axyz", "curlv") else 1
        out[nm] = np.concatenate(parts).reshape(-1, nv) if nv == 3 else np.concatenate(parts)
    return out


def oracle_cfg(sc, m):
    st, c = m.state(), sc["cfg"]
    return dict(eos={0: "adiabatic", 1: "isothermal", 2: "lp07"}[c.get("eos", 0)], gamma=c.get("gamma", 5 / 3),
                cs0=c.get("cs0", 0.0), q=c.get("eos_q", 0.0), r0=c.get("eos_r0", 1.0),
                av=AV.get(c["av"], "none"), av_has_alpha=c["av"] in (2, 3), time=st["time"], dt=st["dt"],
                hfact=HFACT[sc["kernel"]], cfl_cour=c["cfl_cour"], cfl_force=c["cfl_force"],
                gpart_mass=c["gpart_mass"], periodic=c["bc"] == 1, bmin=sc["bmin"], bmax=sc["bmax"])


@pytest.mark.parametrize("av,grid,kernel", [("cd10", (1, 1, 1), "M4"), ("mm97", (2, 1, 1), "M4"),
                                            ("constant", (2, 2, 1), "M6")])
def test_phantom_dump_bytes(tmp_path, av, grid, kernel):
    sc = S.periodic_box(3000, kernel, av, jitter=0.2, grid=grid)
    m = S.make_cuda(sc, keep_step_data=False)
    for _ in range(2):
        m.evolve_once()
    f = tmp_path / "dump_0001"
    m.phantom_dump(f)
    expect = O.make_phantom_dump(all_fields(m), oracle_cfg(sc, m)).gen_file()
    got = f.read_bytes()
    assert len(got) == len(expect)
    assert got == expect
    # and the reference's own reader test on it: read, write, compare
    _capi.phantom_copy(f, tmp_path / "copy")
    assert (tmp_path / "copy").read_bytes() == got
    # the same through the reference's own record IO class (oracle/_ref/fortran_io_ref, prebuilt: ref_fortran_io.cpp)
    exe = O.ref_binary()
    if exe is not None:
        import subprocess

        try:
            r = subprocess.run([exe, "copy", str(f), str(tmp_path / "refcopy")], capture_output=True, text=True)
        except OSError:
            r = None  # the prebuilt driver does not run on this machine: the CPU tests cover it where it was built
        if r is not None and r.returncode in (0, 2, 3):  # 2 / 3: the reference's reader refused the file
            assert r.returncode == 0, r.stderr
            assert (tmp_path / "refcopy").read_bytes() == got


@pytest.mark.parametrize("av,grid,ids", [("cd10", (2, 1, 1), True), ("constant", (1, 1, 1), False),
                                         ("mm97", (1, 2, 2), True)])
def test_vtk_dump_bytes(tmp_path, av, grid, ids):
    sc = S.periodic_box(3000, "M4", av, jitter=0.2, grid=grid)
    m = S.make_cuda(sc, keep_step_data=False)
    for _ in range(2):
        m.evolve_once()
    f = tmp_path / "out_0001.vtk"
    m.vtk_dump(f, ids)
    pid = np.concatenate([np.full(m.patch_size(ip), m.patch_info(ip)["id"]) for ip in range(m.patch_count)])
    expect = O.vtk_dump_bytes(all_fields(m), oracle_cfg(sc, m), ids, patch_ids=pid, world_ranks=np.zeros_like(pid))
    got = f.read_bytes()
    assert len(got) == len(expect)
    assert got == expect
    with pytest.raises(_capi.ShamB200Error, match="vtk"):
        m.vtk_dump(tmp_path / "out.bin", ids)


def test_model_from_phantom_dump(tmp_path):
    """test_ph_dump_writer.py: dump -> gen_config + init_from_phantom_dump -> dump again -> same header, same particles"""
    sc = S.periodic_box(4000, "M4", "cd10", jitter=0.2, grid=(2, 1, 1))
    a = S.make_cuda(sc, keep_step_data=False)
    for _ in range(2):
        a.evolve_once()
    f1, f2 = tmp_path / "a.phdump", tmp_path / "b.phdump"
    a.phantom_dump(f1)
    cfg = _capi.phantom_gen_config(f1)
    cfg.kernel = 0
    b = _capi.Model(_capi.Context(0), cfg)
    b.set_box((0, 0, 0), (1, 1, 1), (2, 1, 1))  # any box: init_from_phantom_dump resizes it, the patch grid stays
    kept = b.init_from_phantom_dump(f1)
    fa = all_fields(a)
    # Phantom2Shamrock.cpp:203-209 writes bmax.x() as ymax and zmax: the box comes back as the reference would read
    # it, and a particle outside of it is not inserted (Model.cpp:1329)
    bmin, bmax = np.array(sc["bmin"]), np.array([sc["bmax"][0]] * 3)
    keep = np.all((fa["xyz"] >= bmin) & (fa["xyz"] < bmax), axis=1)
    assert keep.sum() > 0.9 * len(keep)
    for nm in fa:
        fa[nm] = fa[nm][keep]
    assert kept == int(keep.sum()) == b.total_part_count()
    assert b.state()["time"] == a.state()["time"]
    info = [b.patch_info(ip) for ip in range(b.patch_count)]
    assert info[0]["lo"] == tuple(bmin) and np.allclose(info[-1]["hi"], bmax, rtol=1e-14)
    b.phantom_dump(f2)
    # dtmax (the next dt is not part of what a Phantom dump restores); nparttot / npartoftype if particles were cut
    assert 1 <= _capi.phantom_compare(f1, f2) <= 3
    p1, p2 = O.PhantomDump.from_bytes(f1.read_bytes()), O.PhantomDump.from_bytes(f2.read_bytes())
    # particles may change patch (the box changed): compare as sets, sorted by position
    def table(ph):
        cols = [ph.array(0, t) for t in ("x", "y", "z", "vx", "vy", "vz", "u", "h", "alpha")]
        t = np.stack(cols, axis=1)
        return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]
    t1, t2 = table(p1), table(p2)
    t1 = t1[np.all((t1[:, :3] >= bmin) & (t1[:, :3] < bmax), axis=1)]
    assert np.array_equal(t1, t2)  # h and alpha went through f32 once: a second f32 rounding changes nothing
    # what the model holds: f64 fields exact, h and alpha rounded to f32 by the file
    fb = all_fields(b)
    ob, oa = np.lexsort(fb["xyz"].T[::-1]), np.lexsort(fa["xyz"].T[::-1])
    assert np.array_equal(fb["xyz"][ob], fa["xyz"][oa]) and np.array_equal(fb["vxyz"][ob], fa["vxyz"][oa])
    assert np.array_equal(fb["uint"][ob], fa["uint"][oa])
    assert np.array_equal(fb["hpart"][ob], fa["hpart"][oa].astype(np.float32).astype(np.float64))
    assert np.array_equal(fb["alpha_AV"][ob], fa["alpha_AV"][oa].astype(np.float32).astype(np.float64))


def test_phantom_free_box_and_dead_particles(tmp_path):
    """no xmin..zmax in the header: the box is the positions' bounding box grown by 1.2 about its centre
    (Model.cpp:1253-1273); particles with h < 0 are left out (:1329)"""
    from tests.test_io_formats import synthetic_dump

    ph = synthetic_dump(seed=3, n0=500, n1=0, periodic=False)
    h = np.array(ph.blocks[0]["arrays"]["f32"][0][1])
    h[::7] = -1.0
    ph.blocks[0]["arrays"]["f32"][0] = ("h", h)
    f = tmp_path / "free.phdump"
    f.write_bytes(ph.gen_file())
    cfg = _capi.phantom_gen_config(f)
    assert cfg.bc == 0
    m = _capi.Model(_capi.Context(0), cfg)
    kept = m.init_from_phantom_dump(f, hpart_fact_load=1.5)
    assert kept == int((h >= 0).sum()) == m.total_part_count()
    x = np.stack([ph.array(0, t) for t in "xyz"], axis=1)
    lo, hi = x.min(axis=0), x.max(axis=0)
    c, d = (lo + hi) * 0.5, (hi - lo) * 0.5 * 1.2
    info = m.patch_info(0)
    assert np.array_equal(np.array(info["lo"]), c - d)
    got_h = np.sort(m.get(0, "hpart"))
    assert np.array_equal(got_h, np.sort(h[h >= 0].astype(np.float32).astype(np.float64) * 1.5))
    assert m.state()["time"] == 0.375


def test_dump_writer_script_through_the_python_surface(tmp_path):
    """examples/sph/test_ph_dump_writer.py with `from shamrock_b200 import pyshamrock as shamrock`"""
    from shamrock_b200 import pyshamrock as shamrock

    def new_model():
        ctx = shamrock.Context()
        ctx.pdata_layout_new()
        return ctx, shamrock.get_Model_SPH(context=ctx, vector_type="f64_3", sph_kernel="M4")

    ctx, model = new_model()
    cfg = model.gen_default_config()
    cfg.set_artif_viscosity_VaryingCD10(alpha_min=0.0, alpha_max=1, sigma_decay=0.1, alpha_u=1, beta_AV=2)
    cfg.set_boundary_free()
    cfg.set_eos_adiabatic(1.4)
    model.set_solver_config(cfg)
    model.init_scheduler(int(1e8), 1)
    model.resize_simulation_box((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
    model.add_cube_hcp_3d(0.07, ((-0.5, -0.5, -0.5), (0.5, 0.5, 0.5)))
    model.set_value_in_a_box("uint", "f64", 2.5, (-1, -1, -1), (1, 1, 1))
    model.set_particle_mass(1e-4)
    model.set_cfl_cour(0.3)
    model.set_cfl_force(0.25)
    n = model.get_total_part_count()
    fname = tmp_path / "ref_00000"
    model.make_phantom_dump().save_dump(str(fname))
    model.do_vtk_dump(str(tmp_path / "ref_00000.vtk"), True)
    assert (tmp_path / "ref_00000.vtk").read_bytes().startswith(b"# vtk DataFile Version 4.2\nvtk output\nBINARY\n")

    dump_ref = shamrock.load_phantom_dump(str(fname))
    assert dump_ref.read_header_int("nparttot") == n and dump_ref.read_header_float("gamma") == 1.4
    ctx2, model2 = new_model()
    cfg2 = model2.gen_config_from_phantom_dump(dump_ref)
    model2.set_solver_config(cfg2)
    model2.init_scheduler(int(1e8), 1)
    model2.init_from_phantom_dump(dump_ref)
    assert model2.get_total_part_count() == n
    dump_2 = model2.make_phantom_dump()
    assert shamrock.compare_phantom_dumps(dump_ref, dump_2)
    d1, d2 = ctx.collect_data(), ctx2.collect_data()
    o1, o2 = np.lexsort(d1["xyz"].T[::-1]), np.lexsort(d2["xyz"].T[::-1])
    assert np.array_equal(d1["xyz"][o1], d2["xyz"][o2]) and np.array_equal(d1["uint"][o1], d2["uint"][o2])

"""GPU parity, module by module: the stage-level SPH entry points of the C ABI (shamb200_compute_eos,
shamb200_update_divv_curlv, shamb200_update_dtdivv, shamb200_update_viscosity, shamb200_update_derivs,
shamb200_vsig_cfl, shamb200_leapfrog_predict / _correct) against the CPU oracle.

The oracle runs two steps of a scenario; the MERGED arrays and the ObjectCache of its second step (what the
reference's solver graph hands to ComputeEos, DiffOperators, DiffOperatorDtDivv, UpdateViscosity, UpdateDerivs
and the v_sig / CFL kernels) are uploaded and every module is run on its own through the C ABI.  Contract:
bit-identical float64 outputs (rtol = 0); the LP07 equation of state (`pow`) 1e-12 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from shamrock_b200 import _capi  # noqa: E402
from tests import scenarios as S  # noqa: E402

KNAME = {0: "M4", 1: "M6"}
EOSNAME = {0: "adiabatic", 1: "isothermal", 2: "locally_isothermal_lp07"}
AVNAME = {1: "constant", 2: "varying_mm97", 3: "varying_cd10", 4: "constant_disc"}

SCENARIOS = {
    "periodic_M4_cd10": lambda: S.periodic_box(3000, "M4", "cd10", jitter=0.15),
    "periodic_M6_mm97": lambda: S.periodic_box(2500, "M6", "mm97", jitter=0.15),
    "periodic_M4_constant": lambda: S.periodic_box(2500, "M4", "constant", jitter=0.15),
    "disc_M4": lambda: disc_without_removal(),
    "periodic_M6_cd10_combined": lambda: combined_cd10(),
}


def combined_cd10():
    """CD10 with combined_dtdiv_divcurlv_compute: div v, curl v and d(div v)/dt from one velocity-gradient matrix"""
    sc = S.periodic_box(2500, "M6", "cd10", jitter=0.15)
    sc["cfg"] = dict(sc["cfg"], combined_dtdiv_divcurlv_compute=1)
    return sc


def disc_without_removal():
    """the disc scenario (free boundaries, point mass, LP07, ConstantDisc AV) without the kill sphere and with
    an accretion radius inside the inner edge: the particle ids of the second step are those of the first"""
    sc = S.disc(2500, "M4")
    sc["kill"] = []
    sc["cfg"]["pm_racc"] = 0.5
    return sc


def dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def same(a, b, what, rtol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if rtol == 0:
        bad = np.argwhere(a != b)
        assert len(bad) == 0, f"{what}: {len(bad)} of {a.size} differ (bit-exact contract), max |d| {np.abs(a - b).max():.3e}"
    else:
        err = np.abs(a - b) / np.maximum(np.abs(b), np.abs(b).mean())
        assert err.max() <= rtol, f"{what}: max rel err {err.max():.3e}"


@pytest.fixture(scope="module")
def ctx():
    c = _capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_modules_against_oracle(ctx, name):
    sc = SCENARIOS[name]()
    cfg = sc["cfg"]
    kernel, eos, av = KNAME[cfg["kernel"]], EOSNAME[cfg.get("eos", 0)], AVNAME[cfg["av"]]
    pmass = cfg["gpart_mass"]
    o = S.make_oracle(sc)
    o.evolve_once()
    assert o.patch_count == 1
    ip = 0
    # what the second step starts from
    before = {k: o.get(ip, k).copy() for k in ("xyz", "vxyz", "axyz", "uint", "duint")}
    vary = cfg["av"] in (2, 3)
    if vary:
        alpha_old, cs_old = o.get(ip, "alpha_AV").copy(), o.get(ip, "soundspeed").copy()
    dt, mult = o.state()["dt"], o.state()["cfl_multiplier"]
    o.evolve_once()
    st = o.state()
    n = o.patch_size(ip)
    if n != len(before["uint"]):
        pytest.skip("particles left the patch during the step (accretion / kill sphere): ids moved")
    g = {k: o.get(ip, "step." + k) for k in ("mxyz", "g_h", "g_v", "g_u", "g_omega", "pressure", "soundspeed",
                                             "vsig", "cfl_dt")}
    m_cnt = len(g["g_h"])
    cache = {k: o.get(ip, "cache." + k) for k in ("cnt_neigh", "scanned_cnt", "index_neigh_map")}
    assert len(cache["cnt_neigh"]) == n and m_cnt >= n

    t = {k: dev(v) for k, v in g.items()}
    cnt, scanned, lst = (dev(cache[k].astype(np.int64), torch.int32) for k in ("cnt_neigh", "scanned_cnt",
                                                                               "index_neigh_map"))
    cv = _capi.CsrView()
    cv.obj_cnt, cv.sum_neigh_cnt = n, int(len(cache["index_neigh_map"]))
    cv.d_cnt_neigh, cv.d_scanned_cnt, cv.d_index_neigh_map = cnt.data_ptr(), scanned.data_ptr(), lst.data_ptr()
    g_alpha = dev(o.get(ip, "step.g_alpha")) if vary else None
    g_a = dev(o.get(ip, "step.g_a")) if cfg["av"] == 3 else None

    def fields(**kw):
        return ctx.merged_fields(m_cnt, n, t["mxyz"], t["g_h"], **kw)

    # ---- ComputeEos over the merged range
    P, cs = torch.empty(m_cnt, dtype=torch.float64, device="cuda"), torch.empty(m_cnt, dtype=torch.float64, device="cuda")
    ctx.compute_eos(kernel, eos, fields(uint=t["g_u"]), pmass, cfg.get("gamma", 5.0 / 3.0), cfg.get("cs0", 0.0),
                    cfg.get("eos_q", 0.0), cfg.get("eos_r0", 1.0), P, cs)
    ctx.synchronize()
    rt = 1e-12 if cfg.get("eos", 0) == 2 else 0.0
    same(P.cpu().numpy(), g["pressure"], "pressure", rt)
    same(cs.cpu().numpy(), g["soundspeed"], "soundspeed", rt)

    # ---- DiffOperators / DiffOperatorDtDivv / UpdateViscosity (MM97, CD10)
    combined = bool(cfg.get("combined_dtdiv_divcurlv_compute"))
    if vary and combined:
        divv = torch.empty(n, dtype=torch.float64, device="cuda")
        curlv = torch.empty(3 * n, dtype=torch.float64, device="cuda")
        dtdivv = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.update_dtdivv(kernel, cv, fields(vxyz=t["g_v"], axyz=g_a), pmass, dtdivv, also_divv_curlv=True, divv_t=divv,
                          curlv_t=curlv)
        ctx.synchronize()
        same(divv.cpu().numpy(), o.get(ip, "divv"), "divv (combined)")
        same(curlv.cpu().numpy().reshape(-1, 3), o.get(ip, "curlv"), "curlv (combined)")
        same(dtdivv.cpu().numpy(), o.get(ip, "dtdivv"), "dtdivv (combined)")
    elif vary:
        divv = torch.empty(n, dtype=torch.float64, device="cuda")
        curlv = torch.empty(3 * n, dtype=torch.float64, device="cuda") if cfg["av"] == 3 else None
        ctx.update_divv_curlv(kernel, cv, fields(vxyz=t["g_v"], omega=t["g_omega"]), pmass, divv, curlv)
        ctx.synchronize()
        same(divv.cpu().numpy(), o.get(ip, "divv"), "divv")
        dtdivv = None
        if cfg["av"] == 3:
            same(curlv.cpu().numpy().reshape(-1, 3), o.get(ip, "curlv"), "curlv")
            dtdivv = torch.empty(n, dtype=torch.float64, device="cuda")
            ctx.update_dtdivv(kernel, cv, fields(vxyz=t["g_v"], axyz=g_a), pmass, dtdivv)
            ctx.synchronize()
            same(dtdivv.cpu().numpy(), o.get(ip, "dtdivv"), "dtdivv")
    if vary:
        alpha_new = torch.empty(n, dtype=torch.float64, device="cuda")
        ctx.update_viscosity(av, n, dt, cfg["sigma_decay"], cfg["alpha_min"], cfg["alpha_max"], divv, curlv, dtdivv,
                             dev(cs_old), t["g_h"], dev(alpha_old), alpha_new)
        ctx.synchronize()
        same(alpha_new.cpu().numpy(), o.get(ip, "step.alpha_updated"), "alpha_updated")

    # ---- UpdateDerivs
    axyz = torch.empty(3 * n, dtype=torch.float64, device="cuda")
    duint = torch.empty(n, dtype=torch.float64, device="cuda")
    ext = dev(o.get(ip, "axyz_ext"))
    ctx.update_derivs(kernel, av, cv,
                      fields(vxyz=t["g_v"], uint=t["g_u"], omega=t["g_omega"], pressure=t["pressure"],
                             soundspeed=t["soundspeed"], alpha_AV=g_alpha),
                      pmass, cfg["alpha_u"], cfg["alpha_AV"], cfg["beta_AV"], ext, axyz, duint)
    ctx.synchronize()
    same(axyz.cpu().numpy().reshape(-1, 3), o.get(ip, "axyz"), "axyz")
    same(duint.cpu().numpy(), o.get(ip, "duint"), "duint")

    # ---- v_sig + CFL (the multiplier the step started with: it is only halved when the corrector is repeated
    # and relaxed towards 1 after the time step has been taken)
    vsig = torch.empty(n, dtype=torch.float64, device="cuda")
    cfl = torch.empty(n, dtype=torch.float64, device="cuda")
    mult = mult / 2 ** (int(st["corrector_iter"]) - 1)
    dt_min = ctx.vsig_cfl(kernel, cv, fields(vxyz=t["g_v"], soundspeed=t["soundspeed"]), axyz,
                          cfg["cfl_cour"] * mult, cfg["cfl_force"] * mult, vsig, cfl)
    same(vsig.cpu().numpy(), g["vsig"], "vsig")
    same(cfl.cpu().numpy(), g["cfl_dt"], "cfl_dt")
    assert dt_min == g["cfl_dt"].min()

    # ---- leapfrog: the predictor of the second step and its (single) corrector pass
    if st["corrector_iter"] == 1 and not cfg.get("has_point_mass"):
        x, v, u = dev(before["xyz"]), dev(before["vxyz"]), dev(before["uint"])
        a_old, du_old = dev(before["axyz"]), dev(before["duint"])
        ctx.leapfrog_predict(n, dt, x, v, a_old, u, du_old)
        max_dv2, sum_v2 = ctx.leapfrog_correct(n, dt / 2, v, axyz, a_old, u, duint, du_old)
        same(v.cpu().numpy().reshape(-1, 3), o.get(ip, "vxyz"), "vxyz after the corrector")
        same(u.cpu().numpy(), o.get(ip, "uint"), "uint after the corrector")
        if cfg.get("bc", 0) == 0:  # periodic runs wrap the drifted positions afterwards
            same(x.cpu().numpy().reshape(-1, 3), o.get(ip, "xyz"), "xyz after the predictor")
        vn = o.get(ip, "vxyz")
        assert abs(sum_v2 - (vn * vn).sum()) <= 1e-12 * (vn * vn).sum() + 1e-300
        eps_v = np.sqrt(max_dv2) / np.sqrt(sum_v2 / n) if sum_v2 > 0 else 0.0
        assert abs(eps_v - st["eps_v"]) <= 1e-12 * max(st["eps_v"], 1e-300)


def test_modules_on_empty_patch(ctx):
    """a patch without objects: every module returns without touching anything"""
    z = torch.zeros(4, dtype=torch.float64, device="cuda")
    zi = torch.zeros(4, dtype=torch.int32, device="cuda")
    f = ctx.merged_fields(0, 0, z, z, vxyz=z, uint=z, axyz=z, omega=z, pressure=z, soundspeed=z, alpha_AV=z)
    cv = _capi.CsrView()
    cv.obj_cnt, cv.sum_neigh_cnt = 0, 0
    cv.d_cnt_neigh, cv.d_scanned_cnt, cv.d_index_neigh_map = zi.data_ptr(), zi.data_ptr(), zi.data_ptr()
    out = torch.full((4,), 7.0, dtype=torch.float64, device="cuda")
    ctx.compute_eos("M4", "adiabatic", f, 1.0, 1.4, 0.0, 0.0, 1.0, out, out)
    ctx.update_divv_curlv("M4", cv, f, 1.0, out, out)
    ctx.update_dtdivv("M6", cv, f, 1.0, out, also_divv_curlv=True, divv_t=out, curlv_t=out)
    ctx.update_viscosity("varying_cd10", 0, 0.1, 0.1, 0.0, 1.0, out, out, out, z, z, z, out)
    ctx.update_derivs("M4", "varying_cd10", cv, f, 1.0, 1.0, 1.0, 2.0, None, out, out)
    assert ctx.vsig_cfl("M4", cv, f, z, 0.3, 0.25, out, out) == float("inf")
    ctx.leapfrog_predict(0, 0.1, out, out, z, out, z)
    assert ctx.leapfrog_correct(0, 0.05, out, z, z, out, z, z) == (0.0, 0.0)
    ctx.synchronize()
    assert torch.all(out == 7.0)


def test_modules_reject_bad_arguments(ctx):
    x = torch.zeros(12, dtype=torch.float64, device="cuda")
    h = torch.ones(4, dtype=torch.float64, device="cuda")
    f = ctx.merged_fields(4, 4, x, h)
    out = torch.empty(4, dtype=torch.float64, device="cuda")
    with pytest.raises(_capi.ShamB200Error):  # the adiabatic EOS reads uint
        ctx.compute_eos("M4", "adiabatic", f, 1.0, 1.4, 0.0, 0.0, 1.0, out, out)
    cv = _capi.CsrView()
    cv.obj_cnt = 3  # disagrees with real_cnt
    with pytest.raises(_capi.ShamB200Error):
        ctx.update_divv_curlv("M4", cv, ctx.merged_fields(4, 4, x, h, vxyz=x, omega=h), 1.0, out)
    with pytest.raises(_capi.ShamB200Error):  # the switch exists for MM97 / CD10 only
        ctx.update_viscosity("constant", 4, 0.1, 0.1, 0.0, 1.0, out, None, None, h, h, h, out)

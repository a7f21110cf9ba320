"""The fp32 screen's lower bound (ndtpso_pso_sliced.cuh, "fp32 screen"), restated in numpy float32 from the derivation in that
header and checked against the fp64 cost of the oracle on the golden tables — on the CPU, independently of the kernel's code
(which tests/test_gpu_parity.py and tests/test_gpu_screen_soak.py check pose by pose on the GPU).  The restatement rounds every
operation to fp32 separately (no fused multiply-add), i.e. it makes MORE roundings than the kernel: the derivation allows three per
term, so the bound must hold for it as well.
"""
import numpy as np
import pytest

from tests.problems import Golden

f32 = np.float32
U24 = 2.0 ** -24
L2E = 1.4426950408889634


def screen_tables(flat):
    """Per built cell: the record {l00, l11, c0, c1, l10} in fp32, and the table's kappa2, du, beta (host: screen_params,
    prologue: screen_record / screen_kappa2)."""
    cs, W, gw = flat["cell_side"], flat["width_m"], flat["w_cells"]
    pts = flat["points"]
    pmax = max(float(np.abs(pts).max()) if len(pts) else 0.0, 1.0)
    du = float(f32(U24 * (12.0 * pmax + 3.01 * W) / cs + 2.0 * U24) * f32(1.000001))
    du = float(np.nextafter(f32(du), f32(np.inf)))  # the host rounds to fp32; take the next value up to be on the safe side
    beta = 2.0 * (U24 * (12.0 * pmax + 3.01 * W) / cs + 2.0 * U24)
    idx, mean, ic = flat["cell_index"], flat["mean"], flat["inv_cov"]
    H00, H01, H11 = ic[:, 0] / 2 * cs * cs, ic[:, 1] / 2 * cs * cs, ic[:, 3] / 2 * cs * cs
    ox = (mean[:, 0] + W / 2) / cs - ((idx % gw) + 0.5)
    oy = (mean[:, 1] + W / 2) / cs - ((idx // gw) + 0.5)
    dl0 = du + U24 * (2 * np.abs(ox) + 0.51)
    dl1 = du + U24 * (2 * np.abs(oy) + 0.51)
    dl = np.maximum(dl0, dl1)
    dm0 = 0.5 + np.abs(ox) + dl0
    dm1 = 0.5 + np.abs(oy) + dl1
    sh = 1.0 - 1e-12
    l00 = np.where(H00 > 0, np.sqrt(np.maximum(H00, 0)) * sh, 0.0)
    l10 = np.where(l00 > 0, H01 / np.where(l00 > 0, l00, 1.0) * sh, 0.0)
    l11 = np.sqrt(np.maximum(H11 - l10 * l10 - 4e-15 * H11, 0.0)) * sh
    ez0 = 4 * U24 * (l00 * dm0 + np.abs(l10) * dm1)
    ez1 = 4 * U24 * l11 * dm1
    hs = H00 + 2 * np.abs(H01) + H11
    k0 = ez0 ** 2 + ez1 ** 2 + hs * dl * dl
    k = max(min(L2E * np.sqrt(k0.mean() if len(k0) else 0.0), 0.25), 2.0 ** -20)
    kappa2 = float(np.nextafter(f32(k * 1.000001), f32(np.inf)))
    kd = kappa2 * (1.0 - 2.0 ** -22)
    t = 1.000001 * L2E * k0 / kd
    scale = np.where(t <= 0.5, np.sqrt((1 - t) ** 2 * (1 - 2.0 ** -22) * L2E) * sh, 0.0)
    rec = np.stack([l00 * scale, l11 * scale, -(l00 * ox + l10 * oy) * scale, -(l11 * oy) * scale, l10 * scale], 1).astype(f32)
    return rec, f32(kappa2), f32(0.5 - beta)


def screen_bound(flat, poses):
    """Lower bound of cost_function for every pose: -(sum of the per-point upper bounds)(1 + 2^-14) - 1e-6."""
    rec, kappa2, beta_c = screen_tables(flat)
    cs, W, gw = flat["cell_side"], flat["width_m"], flat["w_cells"]
    slot = np.full(gw * flat["h_cells"], -1, dtype=np.int64)
    slot[flat["cell_index"]] = np.arange(len(flat["cell_index"]))
    px, py = flat["points"][:, 0].astype(f32), flat["points"][:, 1].astype(f32)
    out = []
    for x, y, th in poses:
        c, s = np.cos(th), np.sin(th)
        ck, sk = f32(c / cs), f32(s / cs)
        tu, tv = f32(x / cs + (W / 2) / cs - 0.5), f32(y / cs + (W / 2) / cs - 0.5)
        u = (px * ck + ((-py) * sk + tu)).astype(f32)  # every operation rounded to fp32
        v = (py * ck + (px * sk + tv)).astype(f32)
        nu, nv = np.rint(u).astype(f32), np.rint(v).astype(f32)
        dfu, dfv = (u - nu).astype(f32), (v - nv).astype(f32)
        ix, iy = nu.astype(np.int64), nv.astype(np.int64)
        inside = (ix >= 0) & (ix < gw) & (iy >= 0) & (iy < flat["h_cells"])
        r = np.where(inside, slot[np.clip(ix + gw * iy, 0, len(slot) - 1)], -1)
        hit = r >= 0
        R = rec[np.where(hit, r, 0)]
        z1 = (R[:, 1] * dfv + R[:, 3]).astype(f32)
        z0 = (R[:, 4] * dfv + (R[:, 0] * dfu + R[:, 2]).astype(f32)).astype(f32)
        xe = (-(z0 * z0).astype(f32) + (-(z1 * z1).astype(f32) + kappa2).astype(f32)).astype(f32)
        e = np.where(hit, np.exp2(xe.astype(np.float64)) * (1 + 2.0 ** -21), 0.0)  # ex2.approx: 2 ulp
        edge = np.maximum(np.abs(dfu), np.abs(dfv)) > beta_c
        e = np.where(edge, 1.0 + e, e)  # a point in the band counts as the worst case (the kernel adds the lane's point count)
        out.append(-(e.sum() * (1 + 2.0 ** -14)) - 1e-6)
    return np.array(out)


@pytest.mark.parametrize("case", ["cfg1", "cfg2", "cfg5_0.25", "cfg5_2.0"])
def test_bound_below_fp64_cost_on_golden_tables(golden, oracle, case):
    flat = golden.flat(case, sparse=True)
    dense = golden.flat(case)
    best = golden.case(case)["pose"][0]
    rng = np.random.default_rng(11)
    poses = np.concatenate([best + rng.normal(size=(40, 3)) * np.array(sig)
                            for sig in ((0.002, 0.002, 0.0005), (0.1, 0.1, 0.01), (1.0, 1.0, 0.3), (10.0, 10.0, 3.0))])
    lower = screen_bound(flat, poses)
    cost = oracle.cost_many(dense, poses)
    assert (lower <= cost).all(), (case, np.argwhere(lower > cost)[:3], lower[lower > cost][:3], cost[lower > cost][:3])
    # and it is a useful bound: within a few per cent of the cost on the converged poses (0.6 % at cfg2's optimum, 3 % on the
    # 0.25 m table, whose cells are needles on the scale of the coordinate error: 2^kappa2 on every term plus the t A of the
    # exponents; this restatement also counts every point in the edge band on top of its term)
    conv = slice(0, 40)
    assert np.median((cost[conv] - lower[conv]) / np.abs(cost[conv])) < 0.05


@pytest.mark.parametrize("geom", [(50.0, 0.5), (20.0, 0.25), (50.0, 2.0), (100.0, 1.0)])
def test_bound_on_random_tables_with_needles(oracle, geom):
    """The soak test's random tables (inverse covariances over 7.6 decades in one map, correlations up to 0.9999, means on cell
    corners, points on cell edges and on the frame's border): some records need more than the table's kappa2 gives and take the
    trivial bound; the restated bound must still never exceed the oracle's cost."""
    from tests.test_gpu_screen_soak import pose_sets, random_problem
    rng = np.random.default_rng(5)
    for _ in range(3):
        flat, centre = random_problem(rng, geom[0], geom[1], 500)
        n = flat["w_cells"] * flat["h_cells"]
        dense = dict(flat)
        dense["mean"] = np.zeros((n, 2))
        dense["inv_cov"] = np.zeros((n, 4))
        dense["built"] = np.zeros(n, dtype=np.uint8)
        dense["mean"][flat["cell_index"]] = flat["mean"]
        dense["inv_cov"][flat["cell_index"]] = flat["inv_cov"]
        dense["built"][flat["cell_index"]] = 1
        poses = pose_sets(rng, centre, flat["width_m"], 64)
        poses = poses[np.abs(poses[:, :2]).max(axis=1) < 1e4]  # the kernel's host side admits (pmax + W)/cs <= 2^22 only
        lower = screen_bound(flat, poses)
        cost = oracle.cost_many(dense, poses)
        assert (lower <= cost).all(), (geom, np.argwhere(lower > cost)[:3], lower[lower > cost][:3], cost[lower > cost][:3])

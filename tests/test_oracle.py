"""Pin the CPU oracle (oracle/) against golden vectors produced by the REFERENCE's own modules
(tests/golden/make_golden.py, run in the build container) and against the reference's sparse matrices."""
import os

import numpy as np
import pytest

from oracle.mesh import DropletOracle, PMA2Oracle
from oracle.sh import SHOracle, apply_L_roll, apply_lap_roll, seeded_state, sh13_coefficients

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relmax(a, b):
    return np.abs(np.ravel(a) - np.ravel(b)).max() / max(np.abs(np.ravel(b)).max(), 1e-300)


def test_sh_operator_structure():
    """the 13-point coefficients and storage facts SURVEY.md section 8 (a1) states for the reference's matrices"""
    o = SHOracle(N=64)
    assert o.L.format == "csr" and o.L.nnz == 13 * 64 * 64 and o.Lap.nnz == 5 * 64 * 64
    c0, c1, c2, c3 = sh13_coefficients(o.h, o.r)
    row = o.L.getrow(5 * 64 + 7).toarray().ravel()
    assert np.isclose(row[5 * 64 + 7], c0) and np.isclose(row[5 * 64 + 8], c1)
    assert np.isclose(row[6 * 64 + 8], c2) and np.isclose(row[7 * 64 + 7], c3)
    np.testing.assert_allclose([c0, c1, c2, c3], [-111.582, 47.3088, -13.1072, -6.5536], rtol=1e-12)
    u = seeded_state(64, 3)
    assert relmax(apply_L_roll(u, 64, o.h, o.r), o.L @ u) < 1e-14
    assert relmax(apply_lap_roll(u, 64, o.h), o.Lap @ u) < 1e-14


def test_sh_oracle_is_deterministic():
    o = SHOracle(N=32, d=20.0)
    U0 = seeded_state(32)
    a = o.run(U0, 3)
    b = SHOracle(N=32, d=20.0).run(U0, 3)
    assert np.array_equal(a, b)


def test_pma2_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "pma2_n51.npz"))
    o = PMA2Oracle(N=51)
    o.set_mesh(g["op_Q"])
    for key, ref in (("d2ksi", "op_d2ksi"), ("d2eta", "op_d2eta"), ("dksideta", "op_dksideta"), ("J", "op_J")):
        assert np.array_equal(o.met[key], g[ref]), key
    vxx, vyy = o.ops.laplace_of(g["op_u"], o.met)
    assert np.array_equal(vxx, g["op_vxx"]) and np.array_equal(vyy, g["op_vyy"])
    o.set_prev(g["op_Uval"])
    assert np.array_equal(o.CN, g["op_CN"])
    assert np.array_equal(o.residual(g["op_u"]), g["op_residual"])
    # the script's own run: three passes of the time loop including the DCT mesh update
    o2 = PMA2Oracle(N=51)
    U = np.zeros(51 * 51)
    for s in range(3):
        U = o2.step(U)
        assert np.array_equal(U, g[f"run_U{s}"]), s
        assert np.array_equal(o2.Q, g[f"run_Q{s}"]), s


def test_droplet_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "droplet_91x61.npz"))
    o = DropletOracle()
    o.set_mesh(g["state_Q"])
    for key, ref in (("d2ksi", "op_d2ksi"), ("d2eta", "op_d2eta"), ("dksideta", "op_dksideta"), ("J", "op_J")):
        assert np.array_equal(o.met[key], g[ref]), key
    o.set_prev(g["state_U"], 1e-4)
    assert np.array_equal(o.Uxx, g["op_Uxx"]) and np.array_equal(o.Uyy, g["op_Uyy"])
    assert np.array_equal(o.PI(g["state_U"]), g["op_PI"])
    assert np.array_equal(o.F, g["op_F"])
    vxx, vyy = o.ops.laplace_of(g["op_u"], o.met)
    assert np.array_equal(vxx, g["op_vxx"]) and np.array_equal(vyy, g["op_vyy"])
    assert np.array_equal(o.residual(g["op_u"]), g["op_residual"])
    # two passes of evolve_with_PDE's loop body incl. loop_pma(3e-9, 400)
    o2 = DropletOracle()
    o2.Q = g["state_Q"].copy()
    U = g["state_U"].copy()
    for s in range(2):
        hist = []
        U = o2.step(U, 1e-4, history=hist)
        assert np.array_equal(U, g[f"run_U{s}"]), s
        assert np.array_equal(o2.Q, g[f"run_Q{s}"]), s
        assert np.allclose([h[0] for h in hist[0]["iters"]], g[f"run_hist{s}"], rtol=0, atol=0)


def test_droplet_initialisers_oracle_matches_reference_golden():
    """compute_U2 / initialise_coalescing_droplets / initialise_droplet /
    evolve_R_explicit of droplet.py, restated in oracle/mesh.py, bit for bit against the reference's own functions."""
    g = np.load(os.path.join(GOLD, "droplet_init_91x61.npz"))
    o = DropletOracle()
    ksi, eta = o.ops.ksiksi.reshape(-1), o.ops.etaeta.reshape(-1)
    Q0 = 0.5 * ksi ** 2 + 0.5 * eta ** 2
    o.Q = Q0.copy()
    U = o.initialise_coalescing_droplets(4, g["coal_info"].tolist(), 5e-9, 20)
    assert np.array_equal(U, g["coal_U"]) and np.array_equal(o.Q, g["coal_Q"])
    o = DropletOracle()
    o.Q = Q0.copy()
    U = o.initialise_coalescing_droplets(3, [[0.0, 0.0, 1.0, 1.0]], 5e-9, 20)
    assert np.array_equal(U, g["rect_U"]) and np.array_equal(o.Q, g["rect_Q"])
    pmaloops, Rfinal, tol, dtR, dtmesh = g["evolveR_args"]
    U, R, _ = o.evolve_R_explicit(U, int(pmaloops), Rfinal, tol, R=1.0, V=1.0, dtR=dtR, dtmesh=dtmesh)
    assert R == float(g["evolveR_R"])
    assert np.array_equal(U, g["evolveR_U"]) and np.array_equal(o.Q, g["evolveR_Q"])


def test_sh_oracle_matches_reference_script_golden():
    """tests/golden/sh_n64.npz holds the first three time steps of sh_scipy_nk.py executed UNMODIFIED (legacy RNG
    seeded, its newton_krylov wrapped to record and stop) and of sh_linearised.main(): the oracle's restatement
    reproduces fields and per-iteration Newton norms bit for bit."""
    from oracle.sh import SHLinearisedOracle

    g = np.load(os.path.join(GOLD, "sh_n64.npz"))
    o = SHOracle(N=64)
    U, hist = g["nk_U0"], []
    for s in (1, 2, 3):
        U = o.step(U, history=hist)
        assert np.array_equal(U, g[f"nk_U{s}"]), s
        assert np.array_equal(np.array([it[0] for it in hist[-1]["iters"]]), g[f"nk_hist{s}"]), s
    ol = SHLinearisedOracle(N=64)
    U = g["lin_U0"]
    Uo = U.copy()
    for s in (1, 2, 3):
        U, Uo = ol.step(U, Uo)
        assert np.array_equal(U, g[f"lin_U{s}"]), s


def test_block_row_roll_operator_is_the_full_roll_operator():
    """apply_L_roll_rows / residual_roll_rows (used by bench.py to check the engine at 16384^2 block by block) are the
    rows of apply_L_roll / SHOracle.residual -- bit for bit for L, rounding-close to the CSR residual."""
    from oracle.sh import apply_L_roll_rows, residual_roll_rows

    N = 48
    o = SHOracle(N=N, d=30.0)
    u, uo = seeded_state(N, 3), seeded_state(N, 4)
    full = apply_L_roll(u, N, o.h, o.r).reshape(N, N)
    for a, b in ((0, 5), (3, 17), (40, 48), (0, 48)):
        assert np.array_equal(apply_L_roll_rows(u.reshape(N, N), a, b, o.h, o.r), full[a:b])
    o.set_prev(uo)
    ref = o.residual(u).reshape(N, N)
    got = residual_roll_rows(u.reshape(N, N), uo.reshape(N, N), 0, N, o.h, o.r, o.g, o.k)
    assert relmax(got, ref) < 1e-13

"""The oracle (oracle/ndtpso_oracle.c) against the reference's own outputs (tests/golden).

The reference ships no tests or golden vectors; tests/golden/ref_vectors.npz holds outputs of the
unmodified reference (see tests/golden/make_golden.py).  Bar: bit-exact pose and cost.
"""
import numpy as np
import pytest

from tests.problems import SOLVED, empty_points


@pytest.mark.parametrize("case,inputs", SOLVED)
def test_pso_bit_exact(golden, oracle, case, inputs):
    c = golden.case(case, inputs)
    flat = golden.flat(inputs)
    if case == "edge_empty_scan":
        flat = empty_points(flat)
    for i, seed in enumerate(c["seeds"]):
        pose, cost, st = oracle.pso(flat, c["guess"], c["deviation"], c["P"], c["I"], seed=seed,
                                    w=c["w"], c1=c["c1"], c2=c["c2"], w_dumping=c["w_dumping"])
        assert np.array_equal(pose, c["pose"][i]), (case, seed, pose, c["pose"][i])
        assert cost == c["cost"][i], (case, seed)
        assert st["rand_draws"] == 3 + 3 * c["P"] + 6 * c["P"] * c["I"]


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_cost_function_bit_exact(golden, oracle, name):
    flat = golden.flat(name)
    poses = golden.z[f"{name}/cost_poses"]
    want = golden.z[f"{name}/cost_values"]
    got = oracle.cost_many(flat, poses)
    assert np.array_equal(got, want)
    assert (want[:48] < 0).all()          # near poses hit the map
    assert (want[48:] == 0).sum() > 0     # some far poses are entirely out of bounds


@pytest.mark.parametrize("seed", [1, 42, 123456789, 0, 4294967295])
def test_glibc_rand_stream(golden, oracle, seed):
    want = golden.z[f"rand/{seed}"]
    got = oracle.rand_stream(seed, want.shape[0])
    assert np.array_equal(got, want)


def test_host_stream_equals_seed(golden, oracle):
    """orc_pso fed the rand() outputs explicitly == orc_pso seeding its own generator."""
    c = golden.case("cfg1")
    flat = golden.flat("cfg1")
    n = 3 + 3 * c["P"] + 6 * c["P"] * c["I"]
    for i, seed in enumerate(c["seeds"][:4]):
        stream = oracle.rand_stream(seed, n)
        pose, cost, _ = oracle.pso(flat, c["guess"], c["deviation"], c["P"], c["I"], stream=stream)
        assert np.array_equal(pose, c["pose"][i]) and cost == c["cost"][i]


def test_align_chain(golden, oracle):
    """orc_align (deviation rule + s_* bookkeeping of NDTFrame::align) on the continuing rand() stream."""
    import ctypes as C
    from oracle.binding import OrcAlignState
    want = golden.z["align_chain/pose"]
    flat = golden.flat("align_default")
    prob, _keep = oracle.problem(flat)
    per_call = 3 + 3 * 30 + 6 * 30 * 50
    stream = oracle.rand_stream(int(golden.z["align_chain/srand"][0]), per_call * len(want))
    st = OrcAlignState()
    guess = np.array(golden.case("align_default")["guess"], dtype=np.float64)
    for k in range(len(want)):
        chunk = np.ascontiguousarray(stream[k * per_call:(k + 1) * per_call])
        pose = np.empty(3)
        rc = oracle.lib.orc_align(C.byref(st), C.byref(prob), guess.ctypes.data_as(C.POINTER(C.c_double)), 0,
                                  chunk.ctypes.data_as(C.POINTER(C.c_int32)), pose.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        assert np.array_equal(pose, want[k]), (k, pose, want[k])
        guess = pose.copy()


def test_batch_golden_inputs_and_oracle(oracle, batch_golden, traj_batch):
    """tests/golden/batch_vectors.npz: the drop-in NDTFrame rebuilds the reference's tables and scans of the trajectory problems
    bit for bit (CRC of the reference's own arrays), and the oracle reproduces the reference's results on a sample of them."""
    from tests.problems import oracle_pso_many, table_crc
    flats = traj_batch(256)
    for b, f in enumerate(flats):
        assert table_crc(f) == int(batch_golden.table_crc[b]), b
    sample = [0, 1, 37, 100, 128, 200, 255]
    po, co = oracle_pso_many(oracle, [flats[b] for b in sample], batch_golden.P, batch_golden.I)
    assert np.array_equal(po, batch_golden.pose[sample])
    assert np.array_equal(co, batch_golden.cost[sample])


@pytest.mark.parametrize("cs", [0.25, 0.5, 1.0, 2.0])
def test_cfg5_more_seeds_bit_exact(golden, oracle, batch_golden, cs):
    """configs[4] (200 x 100): the oracle against the unmodified reference on four more seeds per cell size."""
    from tests.problems import oracle_pso_many
    c = golden.case(f"cfg5_{cs}")
    base = golden.flat(f"cfg5_{cs}")
    seeds, want_pose, want_cost = batch_golden.cfg5_more(cs)
    flats = []
    for s in seeds:
        f = dict(base)
        f.update(guess=c["guess"], deviation=c["deviation"], seed=s)
        flats.append(f)
    po, co = oracle_pso_many(oracle, flats, c["P"], c["I"])
    assert np.array_equal(po, want_pose) and np.array_equal(co, want_cost)


def test_ulp_level_table_differences_do_not_move_the_pose(golden, oracle):
    """oracle/_ref is the reference's sources compiled against third_party/eigen_standin, whose 2x2 eigenvalues come from the
    closed form; real Eigen's EigenSolver may differ in the last place of lambda_max, i.e. in the last place of Sigma^-1 of the
    cells that take the eigenvalue-ratio floor (ndtcell.cpp:93-111).  A perturbation of that size (+-2 ulp on every built cell)
    leaves every golden cfg2 pose where it is and moves the scores by ~1e-16 relative: the parity vectors do not hinge on which
    2x2 eigen solver built the table."""
    from tests.problems import oracle_pso_many
    c, flats = golden.problems("cfg2")
    rng = np.random.default_rng(1)
    for _ in range(2):
        pert = []
        for f in flats:
            h = dict(f)
            h["inv_cov"] = f["inv_cov"] * (1 + rng.choice([-2, -1, 0, 1, 2], size=f["inv_cov"].shape[0])[:, None] * 2.2e-16)
            pert.append(h)
        po, co = oracle_pso_many(oracle, pert, c["P"], c["I"])
        assert np.abs(po - c["pose"]).max() <= 1e-9
        assert (np.abs(co - c["cost"]) <= 1e-10 * np.abs(c["cost"])).all()

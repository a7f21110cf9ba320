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

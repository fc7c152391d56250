"""Live cross-check of the CPU oracle against the UNMODIFIED reference cells, at the full size SURVEY section 8(d) asks
for: 2**20 teacher-forced states per variant (uniform box + gate crossings + ground / bounds / rate thresholds).
Runs wherever the reference is reachable (``/root/reference`` in the build container, or the staged copy
``oracle/_ref/reference``) and sympy is importable; skipped otherwise.  The same set is frozen as a digest
(``tests/golden/*_teacher_2p20_digest.npz``) that ``test_oracle_golden.py`` / ``test_gpu_parity.py`` check where the
reference cannot run."""
import numpy as np
import pytest

from conftest import assert_close, golden, scaled_err

R = pytest.importorskip("oracle.reference_exec")
pytest.importorskip("sympy")
if not R.reference_available():
    pytest.skip("reference not reachable (neither /root/reference nor oracle/_ref/reference)", allow_module_level=True)

from oracle import c_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402


@pytest.mark.parametrize("variant", ["e2e", "indi"])
def test_c_oracle_equals_reference_on_the_2p20_teacher_set(variant):
    n = 1 << 20
    track = R.zigzag_track() if variant == "e2e" else R.rectangle_track()
    ws, tg, sc, act, dist = MG.teacher_set_inputs(variant, track, MG.TEACHER_SEED[variant], n)
    ref = MG.teacher_reference_step(variant, ws, tg, sc, act, dist, track)

    gp, gy, sp = track
    ora = O.OracleEnv(variant, n, gp, gy, sp, gates_ahead=1, pause_if_collision=True)
    if variant == "e2e":
        ora.disturbance_ranges = R.training_disturbance_ranges()
    ora.force(ws, tg, sc, dist)
    assert_close(ora.states, ref["obs0"], "obs before the step")
    obs, rew, done, _ = ora.step(act)
    # discrete quantities: bit-exact, every one of the 2**20
    np.testing.assert_array_equal(done, ref["done"])
    np.testing.assert_array_equal(ora.target_gates, ref["tg"])
    np.testing.assert_array_equal(ora.step_counts, ref["sc"])
    # positions are two rounded float32 operations: bit-exact too
    np.testing.assert_array_equal(ora.world_states[:, 0:3], ref["ws"][:, 0:3])
    assert_close(ora.world_states, ref["ws"], "world_states")
    assert_close(obs, ref["obs"], "obs")
    assert_close(rew, ref["rew"], "reward")
    assert done.sum() > n // 20 and (ref["tg"] != tg).sum() > n // 50  # the set exercises the flag logic
    print(variant, "max scaled err oracle vs live reference:", float(scaled_err(ora.world_states, ref["ws"]).max()))

    # ... and the committed digest is exactly this run
    g = golden(f"{variant}_teacher_2p20_digest")
    assert str(g["sha_done"]) == MG.sha(ref["done"].astype(np.uint8))
    assert str(g["sha_tg"]) == MG.sha(ref["tg"].astype(np.int64))
    assert str(g["sha_sc"]) == MG.sha(ref["sc"].astype(np.int64))
    k = int(g["stride"])
    np.testing.assert_array_equal(g["ws"], ref["ws"][::k])
    np.testing.assert_array_equal(g["obs"], ref["obs"][::k])
    np.testing.assert_array_equal(g["rew"], ref["rew"][::k])

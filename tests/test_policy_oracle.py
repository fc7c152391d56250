"""CPU: the policy-forward oracle (row f1) pinned against the reference's own generated C network.
  * golden fixture tests/golden/policy_k4.npz: weights parsed from c_code/neural_network.c, inputs + outputs of the
    reference's compiled nn_forward (oracle/make_policy_fixture.py);
  * where oracle/_ref/libnn_policy_ref.so is present, the live library too."""
import numpy as np
import pytest

from conftest import golden


def net(z):
    n = len(z["dims"]) - 1
    return [z[f"W{l}"] for l in range(n)], [z[f"b{l}"] for l in range(n)]


def test_fixture_shapes_and_std():
    z = golden("policy_k4")
    assert list(z["dims"]) == [24, 120, 120, 120, 4]
    w, b = net(z)
    assert [x.shape for x in w] == [(120, 24), (120, 120), (120, 120), (4, 120)]
    # action std of the shipped controller (`c_code/nn_controller.c:7-12`)
    np.testing.assert_allclose(z["std"], [0.8907372, 0.848785, 0.8973739, 0.8708917], rtol=1e-6)


def test_f32_restatement_matches_reference_c_outputs():
    from oracle import c_oracle as O
    z = golden("policy_k4")
    w, b = net(z)
    y = O.policy_forward(w, b, z["kat_in"])
    # same float32 accumulation order as nn_linear (`c_code/neural_network.c:397-405`): bit-identical
    np.testing.assert_array_equal(y, z["kat_out"])


def test_live_reference_library_if_present():
    from oracle import c_oracle as O
    lib = O.ref_policy_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libnn_policy_ref.so not built (reference not mounted)")
    z = golden("policy_k4")
    x = z["kat_in"][:64]
    y = np.zeros((64, 4), np.float32)
    for i in range(64):
        lib.nn_forward(x[i].ctypes.data_as(O._fp), y[i].ctypes.data_as(O._fp))
    np.testing.assert_array_equal(y, z["kat_out"][:64])


def test_bf16_model_error_budget():
    """The tensor-core path's rounding model stays within 3e-2 absolute of the float32 network on outputs of
    magnitude ~1 (the exploration noise std is 0.85-0.90): the tolerance the GPU tests use against float32."""
    from oracle import c_oracle as O
    z = golden("policy_k4")
    w, b = net(z)
    y16 = O.policy_forward(w, b, z["kat_in"], bf16=True)
    err = np.abs(y16 - z["kat_out"])
    print("bf16-model vs f32: max abs err", err.max(), "mean", err.mean())
    assert err.max() < 3e-2 and err.mean() < 4e-3

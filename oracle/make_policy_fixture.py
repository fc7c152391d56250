"""TEST INFRASTRUCTURE -- freezes the reference's trained controller into fixtures.

Run in the build container (needs ``/root/reference``):

    python oracle/make_policy_fixture.py

* parses the weight / bias arrays of ``c_code/neural_network.c:5-395`` (24 -> 120 -> 120 -> 120 -> 4) and the action
  std of ``c_code/nn_controller.c:7-12`` -> ``tests/golden/policy_k4.npz`` and the packaged
  ``optimal_quad_control_rl_b200/data/policy_k4.npz``;
* evaluates the reference's OWN compiled C (``oracle/_ref/libnn_policy_ref.so``: ``nn_forward``) on seeded inputs
  and stores inputs + outputs in the same fixture, so the GPU box can check parity without the reference.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("QUADSIM_REFERENCE_ROOT", "/root/reference")


def c_array(text, name):
    m = re.search(r"const\s+float\s+" + name + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, flags=re.S)
    return np.array([float(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()], dtype=np.float32)


def main():
    net = open(os.path.join(REF, "c_code", "neural_network.c")).read()
    ctl = open(os.path.join(REF, "c_code", "nn_controller.c")).read()
    dims = [24, 120, 120, 120, 4]
    out = {"dims": np.array(dims, np.int32), "std": c_array(ctl, "output_std")}
    for l in range(4):
        out[f"W{l}"] = c_array(net, f"weights_fc{l + 1}").reshape(dims[l + 1], dims[l])
        out[f"b{l}"] = c_array(net, f"biases_fc{l + 1}")
        assert out[f"b{l}"].shape == (dims[l + 1],)
    subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(HERE, "_ref", "libnn_policy_ref.so"))
    fp = C.POINTER(C.c_float)
    lib.nn_forward.argtypes = [fp, fp]
    rng = np.random.default_rng(4)
    n = 1024
    # observation-like inputs: gate-frame position / velocity / attitude / rates / motor speeds / next gate / disturbances
    scale = np.array([3, 3, 1.5, 6, 6, 3, 1, 1, 3.1, 4, 4, 3, 1, 1, 1, 1, 3, 3, 1, 4.8, 1, 1, 1, 1], np.float32)
    x = (rng.uniform(-1, 1, (n, 24)).astype(np.float32) * scale).astype(np.float32)
    y = np.zeros((n, 4), np.float32)
    for i in range(n):
        lib.nn_forward(x[i].ctypes.data_as(fp), y[i].ctypes.data_as(fp))
    out["kat_in"], out["kat_out"] = x, y
    for path in (os.path.join(ROOT, "tests", "golden", "policy_k4.npz"),
                 os.path.join(ROOT, "optimal_quad_control_rl_b200", "data", "policy_k4.npz")):
        keep = out if "golden" in path else {k: v for k, v in out.items() if not k.startswith("kat_")}
        np.savez_compressed(path, **keep)
        print("wrote", path, os.path.getsize(path), "bytes")
    print("nn_forward output range", y.min(0), y.max(0))


if __name__ == "__main__":
    sys.exit(main())

"""Flight-controller C export: a trained policy plus the track it was trained on become
the dependency-free C files the reference flashes onto the Bebop through Paparazzi.

What it replaces: the notebook cells that write ``c_code/`` -- the dense network (`3D quad race.ipynb:4026-4132`),
the controller with gate switching, gate-frame observation, exploration noise, clip and actuator scaling
(`:4141-4377`; INDI `3D quad race INDI inner loop.ipynb:862-1062`), the residual models (`:4544-4647`), and the
``gcc -fPIC`` / ``ctypes`` round trip (`:4393-4421`).  The generated sources export the reference's symbols with the
reference's signatures and arithmetic (float accumulation ``bias + sum_j x[j]*w[i][j]`` in ascending ``j``;
``cosf/sinf`` gate frame; double-precision ``M_PI`` yaw wrap; ``rand()`` Box-Muller), so firmware written against
``c_code/*.h`` links unchanged:

    void nn_forward(const float *input, float *output);
    void nn_reset(void);
    void nn_control(const float world_state[16], const float disturbances[4], float rpms[4]);   /* E2E            */
    void nn_control(const float world_state[16], float rpms[4]);                                /* E2E, no dist.  */
    void nn_control(const float world_state[13], float indi_cmd[4]);                            /* INDI           */
    void nn_thrust_forward(const float *input, float *output);  void nn_moment_forward(...);    /* residual MLPs  */
    bool deterministic;  uint8_t target_gate_index;  gate_pos / gate_yaw / start_pos / output_std tables

The network is emitted table-driven (one ``layers[]`` descriptor walked by one loop) instead of one unrolled call
per layer, and the headers declare the globals ``extern`` (the reference's headers rely on tentative definitions,
which only link with ``-fcommon``).  Nothing here touches the GPU: inputs are NumPy arrays or any object with the env's
public attributes (``Quadcopter3DGates``, the oracle env, or a plain namespace).  ``tests/test_codegen.py`` compiles
the output and checks it bit for bit against the reference's own shipped ``c_code`` for the same weights and track."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from types import SimpleNamespace

import numpy as np

# actuator ranges of the two models (`3D quad race.ipynb:93`; INDI `:70-74`)
W_MIN, W_MAX = 3000.0, 11000.0
INDI_RANGES = {"p": (-3.0, 3.0), "q": (-3.0, 3.0), "r": (-2.0, 2.0), "T": (0.0, 16.0)}


def _lit(x):
    """A float32 value as a C literal that parses back to the same float32 (shortest repr of its double)."""
    return repr(float(np.float32(x)))


def _rows(a, indent="    "):
    a = np.asarray(a, np.float32)
    a = a.reshape(1, -1) if a.ndim == 1 else a
    return ",\n".join(indent + ", ".join(_lit(v) for v in row) for row in a)


# ------------------------------------------------------------------------------------------------ dense network
def network_sources(weights, biases, prefix="nn", activation="relu", header_name=None):
    """C source + header of ``<prefix>_forward`` for an MLP given as torch-layout ``(out, in)`` float32 matrices.
    ``activation`` in {"relu", "tanh"} is applied after every layer but the last (SB3 ``activation_fn``)."""
    if activation not in ("relu", "tanh"):
        raise ValueError("activation must be 'relu' or 'tanh'")
    weights = [np.ascontiguousarray(w, np.float32) for w in weights]
    biases = [np.ascontiguousarray(b, np.float32).ravel() for b in biases]
    if len(weights) != len(biases) or not weights:
        raise ValueError("need one bias vector per weight matrix")
    for l, (w, b) in enumerate(zip(weights, biases)):
        if w.ndim != 2 or b.shape != (w.shape[0],) or (l and w.shape[1] != weights[l - 1].shape[0]):
            raise ValueError(f"layer {l}: inconsistent shapes W{w.shape} b{b.shape}")
    header_name = header_name or ("neural_network.h" if prefix == "nn" else f"{prefix}.h")
    guard = header_name.upper().replace(".", "_")
    wname = (lambda i: f"weights_fc{i}") if prefix == "nn" else (lambda i: f"{prefix}_weights_fc{i}")
    bname = (lambda i: f"biases_fc{i}") if prefix == "nn" else (lambda i: f"{prefix}_biases_fc{i}")
    widest = max(w.shape[0] for w in weights[:-1]) if len(weights) > 1 else 1
    out = [f'#include "{header_name}"', "#include <math.h>", ""]
    for i, (w, b) in enumerate(zip(weights, biases), 1):
        out += [f"const float {wname(i)}[] = {{", _rows(w), "};", "",
                f"const float {bname(i)}[] = {{", _rows(b), "};", ""]
    out += [
        f"void {prefix}_linear(const float* weights, const float* biases, const float* input, int in_features, "
        "int out_features, float* output) {",
        "    for (int i = 0; i < out_features; ++i) {",
        "        const float* w = weights + i * in_features;",
        "        float acc = biases[i];",
        "        for (int j = 0; j < in_features; ++j) acc += input[j] * w[j];",
        "        output[i] = acc;",
        "    }",
        "}", "",
        f"void {prefix}_relu(float* x, int size) {{",
        "    for (int i = 0; i < size; ++i) x[i] = fmaxf(0, x[i]);",
        "}", "",
        f"void {prefix}_tanh(float* x, int size) {{",
        "    for (int i = 0; i < size; ++i) x[i] = tanh(x[i]);",
        "}", "",
        f"#define {prefix.upper()}_NUM_LAYERS {len(weights)}",
        "static const struct { const float* w; const float* b; int in, out; } " + f"{prefix}_layers[] = {{",
    ]
    out += [f"    {{{wname(i)}, {bname(i)}, {w.shape[1]}, {w.shape[0]}}}," for i, w in enumerate(weights, 1)]
    out += [
        "};", "",
        f"void {prefix}_forward(const float* input, float* output) {{",
        f"    float ping[{widest}], pong[{widest}];",
        "    const float* x = input;",
        f"    for (int l = 0; l < {prefix.upper()}_NUM_LAYERS; ++l) {{",
        f"        const int last = l == {prefix.upper()}_NUM_LAYERS - 1;",
        "        float* y = last ? output : (x == ping ? pong : ping);",
        f"        {prefix}_linear({prefix}_layers[l].w, {prefix}_layers[l].b, x, {prefix}_layers[l].in, "
        f"{prefix}_layers[l].out, y);",
        f"        if (!last) {prefix}_{activation}(y, {prefix}_layers[l].out);",
        "        x = y;",
        "    }",
        "}", "",
    ]
    hdr = [f"#ifndef {guard}", f"#define {guard}", "",
           f"void {prefix}_forward(const float* input, float* output);", "", f"#endif // {guard}", ""]
    return "\n".join(out), "\n".join(hdr)


# ------------------------------------------------------------------------------------------------ controller
def _variant_of(env):
    v = getattr(env, "_VARIANT", None) or getattr(env, "variant", None)
    if isinstance(v, str):
        return v
    if v is not None:  # the oracle env stores the C enum: 0 = E2E, 1 = INDI
        return "e2e" if int(v) == 0 else "indi"
    ns = getattr(env, "ns", None) or getattr(env, "_ns", None)
    if ns in (13, 16):
        return "e2e" if ns == 16 else "indi"
    raise ValueError("cannot tell the model variant of this env: pass variant='e2e'|'indi'")


def controller_sources(env, std, variant=None, disturbance_input=None):
    """C source + header of ``nn_controller``: gate switching on the gate plane, gate-frame observation
    (the same transform as ``update_states_gate``, `3D quad race.ipynb:365-450`), Gaussian exploration noise, clip, and
    the mapping of the network's [-1,1] outputs to rpm (E2E) or rate/thrust commands (INDI)."""
    variant = variant or _variant_of(env)
    if variant not in ("e2e", "indi"):
        raise ValueError(variant)
    e2e = variant == "e2e"
    if disturbance_input is None:
        disturbance_input = e2e
    if disturbance_input and not e2e:
        raise ValueError("the INDI model has no disturbance inputs")
    ng, ga = int(env.num_gates), int(env.gates_ahead)
    if ng > 255:
        raise ValueError("target_gate_index is a uint8_t: at most 255 gates")
    ns, base = (16, 16) if e2e else (13, 13)
    n_in = base + 4 * ga + (4 if disturbance_input else 0)
    std = np.asarray(std, np.float32).ravel()
    if std.size != 4:
        raise ValueError("std must have 4 entries")
    if e2e and disturbance_input:
        sig = f"void nn_control(const float world_state[{ns}], const float disturbances[4], float rpms[4])"
    elif e2e:
        sig = f"void nn_control(const float world_state[{ns}], float rpms[4])"
    else:
        sig = f"void nn_control(const float world_state[{ns}], float indi_cmd[4])"

    hdr = ["#ifndef NN_CONTROLLER_H", "#define NN_CONTROLLER_H", "", "#include <stdint.h>", "#include <stdbool.h>", "",
           f"#define GATES_AHEAD {ga}", f"#define NUM_GATES {ng}", f"#define NN_INPUT_SIZE {n_in}", "",
           '#include "neural_network.h"', "",
           "extern const float gate_pos[NUM_GATES][3];", "extern const float gate_yaw[NUM_GATES];",
           "extern const float start_pos[3];", "extern const float output_std[4];",
           "extern uint8_t target_gate_index;", "extern bool deterministic;", "",
           "void nn_reset(void);", sig + ";", "", "#endif", ""]

    table = lambda name, a, dims: [f"const float {name}{dims} = {{", *[
        "    {" + ", ".join(_lit(v) for v in row) + "}," for row in np.asarray(a, np.float32)], "};", ""]
    vec = lambda name, a, dims: [f"const float {name}{dims} = {{", *[f"    {_lit(v)}," for v in np.asarray(a, np.float32)],
                                 "};", ""]
    src = ['#include "nn_controller.h"', "#include <math.h>", "#include <stdlib.h>", "",
           "bool deterministic = false;", ""]
    src += vec("output_std", std, "[4]")
    src += table("gate_pos", env.gate_pos, "[NUM_GATES][3]")
    src += vec("gate_yaw", env.gate_yaw, "[NUM_GATES]")
    src += ["const float start_pos[3] = {", "    " + ", ".join(_lit(v) for v in np.asarray(env.start_pos, np.float32)),
            "};", ""]
    src += table("gate_pos_rel", env.gate_pos_rel, "[NUM_GATES][3]")
    src += vec("gate_yaw_rel", env.gate_yaw_rel, "[NUM_GATES]")
    src += ["uint8_t target_gate_index = 0;", "", "void nn_reset(void) {", "    target_gate_index = 0;", "}", "",
            "// signed distance past the plane of gate g along its normal",
            "static float gate_plane(int g, float x, float y) {",
            "    return cosf(gate_yaw[g]) * (x - gate_pos[g][0]) + sinf(gate_yaw[g]) * (y - gate_pos[g][1]);",
            "}", "", sig + " {",
            "    const float x = world_state[0], y = world_state[1], z = world_state[2];",
            "    const float vx = world_state[3], vy = world_state[4], vz = world_state[5];",
            "    // through the plane of the target gate: aim at the next one (looped track)",
            "    if (gate_plane(target_gate_index, x, y) > 0) {",
            "        target_gate_index = (uint8_t)((target_gate_index + 1) % NUM_GATES);",
            "    }",
            "    const int g = target_gate_index;",
            "    const float gyaw = gate_yaw[g];",
            "    const float dx = x - gate_pos[g][0], dy = y - gate_pos[g][1];",
            "    float nn_input[NN_INPUT_SIZE];",
            "    // position and velocity in the gate frame",
            "    nn_input[0] = cosf(gyaw) * dx + sinf(gyaw) * dy;",
            "    nn_input[1] = -sinf(gyaw) * dx + cosf(gyaw) * dy;",
            "    nn_input[2] = z - gate_pos[g][2];",
            "    nn_input[3] = cosf(gyaw) * vx + sinf(gyaw) * vy;",
            "    nn_input[4] = -sinf(gyaw) * vx + cosf(gyaw) * vy;",
            "    nn_input[5] = vz;",
            "    // attitude: roll, pitch, heading relative to the gate wrapped to [-pi, pi]",
            "    float yaw_rel = world_state[8] - gyaw;",
            "    while (yaw_rel > M_PI) yaw_rel -= 2*M_PI;",
            "    while (yaw_rel < -M_PI) yaw_rel += 2*M_PI;",
            "    nn_input[6] = world_state[6];",
            "    nn_input[7] = world_state[7];",
            "    nn_input[8] = yaw_rel;",
            "    for (int i = 9; i < 12; i++) nn_input[i] = world_state[i];  // body rates"]
    if e2e:
        src += [f"    const float w_min = {_lit(W_MIN)}, w_max = {_lit(W_MAX)};",
                "    for (int i = 12; i < 16; i++) nn_input[i] = (world_state[i] - w_min) * 2 / (w_max - w_min) - 1;"
                "  // rpm -> [-1,1]"]
    else:
        src += [f"    const float T_min = {_lit(INDI_RANGES['T'][0])}, T_max = {_lit(INDI_RANGES['T'][1])};",
                "    nn_input[12] = (world_state[12] - T_min) / (T_max - T_min) * 2 - 1;  // thrust -> [-1,1]"]
    src += ["    // the gates after the target, each in the frame of its predecessor",
            "    for (int i = 0; i < GATES_AHEAD; i++) {",
            "        const int k = (g + i + 1) % NUM_GATES;",
            f"        float* o = nn_input + {base} + 4*i;",
            "        o[0] = gate_pos_rel[k][0]; o[1] = gate_pos_rel[k][1]; o[2] = gate_pos_rel[k][2]; o[3] = gate_yaw_rel[k];",
            "    }"]
    if disturbance_input:
        r = np.asarray(env.disturbance_ranges, np.float64)
        lo, hi = r[[0, 1, 2, 5], 0].copy(), r[[0, 1, 2, 5], 1].copy()
        same = lo == hi  # the env widens an empty range by -+1 (`3D quad race.ipynb:425-440`)
        lo[same] -= 1
        hi[same] += 1
        src += ["    // disturbance estimates Mx, My, Mz, Fz normalised over their training ranges",
                "    static const float d_min[4] = {" + ", ".join(repr(float(v)) for v in lo) + "};",
                "    static const float d_max[4] = {" + ", ".join(repr(float(v)) for v in hi) + "};",
                "    for (int i = 0; i < 4; i++) {",
                f"        nn_input[{base} + 4*GATES_AHEAD + i] = (disturbances[i] - d_min[i]) * 2 / (d_max[i] - d_min[i]) - 1;",
                "    }"]
    src += ["    float nn_output[4];",
            "    nn_forward(nn_input, nn_output);",
            "    if (!deterministic) {  // exploration noise: Box-Muller on libc rand()",
            "        for (int i = 0; i < 4; i++) {",
            "            float u1 = (float)rand() / RAND_MAX;",
            "            float u2 = (float)rand() / RAND_MAX;",
            "            float rand_std = sqrtf(-2 * logf(u1)) * cosf(2 * M_PI * u2);",
            "            nn_output[i] += output_std[i] * rand_std;",
            "        }",
            "    }",
            "    for (int i = 0; i < 4; i++) {",
            "        if (nn_output[i] > 1) nn_output[i] = 1;",
            "        if (nn_output[i] < -1) nn_output[i] = -1;",
            "    }"]
    if e2e:
        src += ["    for (int i = 0; i < 4; i++) rpms[i] = (w_max - w_min) * (nn_output[i] + 1) / 2 + w_min;"]
    else:
        src += ["    static const float c_min[4] = {" + ", ".join(_lit(INDI_RANGES[k][0]) for k in "pqrT") + "};",
                "    static const float c_max[4] = {" + ", ".join(_lit(INDI_RANGES[k][1]) for k in "pqrT") + "};",
                "    for (int i = 0; i < 4; i++) indi_cmd[i] = (nn_output[i] + 1) / 2 * (c_max[i] - c_min[i]) + c_min[i];"]
    src += ["}", ""]
    return "\n".join(src), "\n".join(hdr)


# ------------------------------------------------------------------------------------------------ files + build
def _policy_arrays(policy):
    """(weights, biases, std) from an ``MlpPolicy`` / ``PPO`` of this package, an SB3 model, or a tuple."""
    if isinstance(policy, (tuple, list)) and len(policy) == 3:
        return policy
    if getattr(policy, "actor", None) is not None:  # our PPO with a device actor: publish the float32 master weights first
        policy._publish()
        policy = policy.actor
    if hasattr(policy, "weights") and hasattr(policy, "biases"):
        return policy.weights, policy.biases, policy.std
    if hasattr(policy, "policy") and hasattr(policy.policy, "mlp_extractor"):  # stable_baselines3.PPO
        pol = policy.policy
        lin = [m for m in pol.mlp_extractor.policy_net if hasattr(m, "weight")] + [pol.action_net]
        return ([m.weight.detach().cpu().numpy() for m in lin], [m.bias.detach().cpu().numpy() for m in lin],
                pol.log_std.detach().exp().cpu().numpy())
    raise TypeError("policy must be MlpPolicy, PPO, an SB3 model or (weights, biases, std)")


def export_controller(policy, env, out_dir="c_code", variant=None, disturbance_input=None, residual_weights=None,
                      activation="relu"):
    """Write ``neural_network.{c,h}`` and ``nn_controller.{c,h}`` (and, for the E2E model with ``residual_weights=
    (thrust[289], moment[451])`` or ``True`` for the packaged ones, ``nn_thrust.{c,h}`` / ``nn_moment.{c,h}``) into
    ``out_dir``.  Returns the list of files written."""
    w, b, std = _policy_arrays(policy)
    variant = variant or _variant_of(env)
    os.makedirs(out_dir, exist_ok=True)
    files = {}
    files["neural_network.c"], files["neural_network.h"] = network_sources(w, b, activation=activation)
    files["nn_controller.c"], files["nn_controller.h"] = controller_sources(env, std, variant, disturbance_input)
    ga = int(env.gates_ahead)
    n_in = (16 if variant == "e2e" else 13) + 4 * ga + (4 if (disturbance_input if disturbance_input is not None
                                                              else variant == "e2e") else 0)
    if np.asarray(w[0]).shape[1] != n_in:
        raise ValueError(f"the policy takes {np.asarray(w[0]).shape[1]} inputs but this controller builds {n_in}")
    if residual_weights is not None and residual_weights is not False:
        if residual_weights is True:
            z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "residual_mlp.npz"))
            parts = {n: ([z[f"{n}_w1"], z[f"{n}_w2"]], [z[f"{n}_b1"], z[f"{n}_b2"]]) for n in ("thrust", "moment")}
        else:
            t, m = (np.asarray(a, np.float32).ravel() for a in residual_weights)
            split = lambda a, i, o: ([a[:32 * i].reshape(32, i), a[32 * i + 32:32 * i + 32 + o * 32].reshape(o, 32)],
                                     [a[32 * i:32 * i + 32], a[32 * i + 32 + o * 32:]])
            parts = {"thrust": split(t, 7, 1), "moment": split(m, 10, 3)}
        for n, (ws_, bs_) in parts.items():
            files[f"nn_{n}.c"], files[f"nn_{n}.h"] = network_sources(ws_, bs_, prefix=f"nn_{n}")
    paths = []
    for name, text in files.items():
        p = os.path.join(out_dir, name)
        with open(p, "w") as f:
            f.write(text)
        paths.append(p)
    return paths


def build_controller(c_dir="c_code", lib_name="libtools.so", cc=None, flags=("-O2",)):
    """``gcc -fPIC -shared`` over every ``.c`` of ``c_dir`` (the reference's round trip, `:4399-4407`)."""
    cc = cc or shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("no C compiler found")
    srcs = sorted(os.path.join(c_dir, f) for f in os.listdir(c_dir) if f.endswith(".c"))
    out = os.path.join(c_dir, lib_name)
    r = subprocess.run([cc, "-fPIC", "-shared", *flags, "-o", out, *srcs, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("controller build failed:\n" + r.stderr)
    return out


class CController:
    """ctypes view of a built controller library with the helper conversions of the notebook's test cells
    (`:4426-4478`): ``forward(obs)`` = clipped ``nn_forward``; ``control(world_state[, disturbances])`` takes the env's
    normalised state, feeds the controller physical units and maps its command back to [-1,1]."""

    def __init__(self, lib_path, variant="e2e", disturbance_input=None):
        self.lib = C.CDLL(os.path.abspath(lib_path))
        self.variant = variant
        self.disturbance_input = (variant == "e2e") if disturbance_input is None else bool(disturbance_input)
        fp = C.POINTER(C.c_float)
        self.lib.nn_forward.argtypes = [fp, fp]
        self.lib.nn_forward.restype = None
        self.lib.nn_control.argtypes = [fp, fp, fp] if self.disturbance_input else [fp, fp]
        self.lib.nn_control.restype = None
        self.lib.nn_reset.restype = None
        self._fp = fp
        self.n_in = None

    deterministic = property(lambda self: bool(C.c_bool.in_dll(self.lib, "deterministic").value),
                             lambda self, v: setattr(C.c_bool.in_dll(self.lib, "deterministic"), "value", bool(v)))
    target_gate_index = property(lambda self: int(C.c_uint8.in_dll(self.lib, "target_gate_index").value))

    def reset(self):
        self.lib.nn_reset()

    def forward(self, obs):
        x = np.ascontiguousarray(obs, np.float32)
        y = np.zeros(4, np.float32)
        self.lib.nn_forward(x.ctypes.data_as(self._fp), y.ctypes.data_as(self._fp))
        return np.clip(y, -1, 1)

    def control(self, world_state, disturbances=None):
        ws = np.array(world_state, dtype=np.float32)
        out = np.zeros(4, np.float32)
        if self.variant == "e2e":
            ws[12:16] = (ws[12:16] + 1) / 2 * (W_MAX - W_MIN) + W_MIN
        else:
            lo, hi = INDI_RANGES["T"]
            ws[12] = (ws[12] + 1) / 2 * (hi - lo) + lo
        args = [ws.ctypes.data_as(self._fp)]
        if self.disturbance_input:
            d = np.ascontiguousarray(disturbances, np.float32)
            args.append(d.ctypes.data_as(self._fp))
        self.lib.nn_control(*args, out.ctypes.data_as(self._fp))
        if self.variant == "e2e":
            return (out - W_MIN) / (W_MAX - W_MIN) * 2 - 1
        lo = np.array([INDI_RANGES[k][0] for k in "pqrT"], np.float32)
        hi = np.array([INDI_RANGES[k][1] for k in "pqrT"], np.float32)
        return (out - lo) / (hi - lo) * 2 - 1


def track_spec(gate_pos, gate_yaw, start_pos, gates_ahead, variant, disturbance_ranges=None):
    """The env attributes the generator reads, computed on the host like ``Quadcopter3DGates.__init__``
    (`3D quad race.ipynb:298-319`) -- for exporting a controller without constructing a GPU env."""
    gp = np.ascontiguousarray(np.asarray(gate_pos).astype(np.float32))
    gy = np.ascontiguousarray(np.asarray(gate_yaw).astype(np.float32))
    ng = gp.shape[0]
    rel, yrel = np.zeros((ng, 3), np.float32), np.zeros(ng, np.float32)
    for i in range(ng):
        d = gp[i] - gp[i - 1]
        c, s = np.cos(gy[i - 1]), np.sin(gy[i - 1])
        rel[i, 0:2] = np.array([[c, s], [-s, c]]) @ d[0:2]
        rel[i, 2] = d[2]
        yrel[i] = gy[i] - gy[i - 1]
    dr = np.zeros((6, 2), np.float32) if disturbance_ranges is None else np.asarray(disturbance_ranges)
    return SimpleNamespace(num_gates=ng, gates_ahead=int(gates_ahead), gate_pos=gp, gate_yaw=gy,
                           start_pos=np.asarray(start_pos).astype(np.float32), gate_pos_rel=rel, gate_yaw_rel=yrel,
                           disturbance_ranges=dr, variant=variant)

"""TEST INFRASTRUCTURE — runs the *unmodified* reference notebook cells as the ground truth.

Reference root, in order: ``$QUADSIM_REFERENCE_ROOT``; ``/root/reference`` (the build container); the staged copy
``oracle/_ref/reference`` (written by ``stage_reference()`` from ``__graft_entry__.build()`` -- git-ignored, it
travels to the GPU box with the snapshot like the ``.so`` files next to it, so that ``bench.py`` can time the
reference's own NumPy ``step()`` on the box's host cores).  Used by ``oracle/make_golden.py`` (golden vectors),
``tests/test_oracle_vs_reference_live.py`` (live cross-check of the C oracle) and ``bench.py``'s CPU legs.  The
product package never imports it.

Recipe:
  * E2E  : ``3D quad race.ipynb`` code cells 2 (sympy EoM -> f_func), 4 (residual MLPs), 6 (Quadcopter3DGates)
  * INDI : ``3D quad race INDI inner loop.ipynb`` code cells 2 (EoM) and 5 (Quadcopter3DGates)
are ``exec``'d verbatim into one namespace after registering stub modules for the two imports that are not
installed here (``gymnasium``/``gym`` ``spaces.Box`` and ``stable_baselines3.common.vec_env.VecEnv``).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types

E2E_NOTEBOOK = "3D quad race.ipynb"
INDI_NOTEBOOK = "3D quad race INDI inner loop.ipynb"
_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(_HERE, "_ref", "reference")
_STAGED_FILES = (E2E_NOTEBOOK, INDI_NOTEBOOK, os.path.join("NNDroneModel", "thrust_model.pt"),
                 os.path.join("NNDroneModel", "moment_model.pt"))


def _find_root():
    cands = [os.environ.get("QUADSIM_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, E2E_NOTEBOOK)):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, E2E_NOTEBOOK))


def stage_reference(src="/root/reference") -> bool:
    """Copy the files the two env notebooks need (the notebooks themselves + NNDroneModel/*.pt) from the mounted
    reference into oracle/_ref/reference, byte for byte.  oracle/_ref is git-ignored: nothing of the reference enters
    the repository history.  Returns False (and changes nothing) where the reference is not mounted."""
    import shutil

    if not os.path.isfile(os.path.join(src, E2E_NOTEBOOK)):
        return False
    for rel in _STAGED_FILES:
        dst = os.path.join(STAGED_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        os.chmod(dst, 0o644)
    return True


class _Box:
    """Minimal stand-in for gymnasium.spaces.Box: stores what the reference passes."""

    def __init__(self, low, high, shape=None, dtype=None):
        import numpy as np

        self.low, self.high = low, high
        self.shape = tuple(shape) if shape is not None else np.asarray(low).shape
        self.dtype = dtype


class _VecEnv:
    """Minimal stand-in for stable_baselines3's VecEnv base class (SB3 is not installed here)."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    for root in ("gymnasium", "gym"):
        sp = mod(root + ".spaces", Box=_Box)
        mod(root, spaces=sp)
    ve = mod("stable_baselines3.common.vec_env", VecEnv=_VecEnv)
    co = mod("stable_baselines3.common", vec_env=ve)
    mod("stable_baselines3", common=co, __version__="stub")


def _cells(nb_name):
    with open(os.path.join(REFERENCE_ROOT, nb_name)) as f:
        nb = json.load(f)
    return ["".join(c["source"]) for c in nb["cells"]]


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_CACHE: dict = {}


def load_reference(variant: str) -> dict:
    """Return the namespace produced by exec'ing the reference cells for ``variant`` in {"e2e","indi"}.

    Keys of interest: ``Quadcopter3DGates``, ``f_func``, and for e2e ``get_body_velocity``,
    ``thrust_moment_model_world_states``, ``thrust_model``, ``moment_model``.
    """
    if variant in _CACHE:
        return _CACHE[variant]
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    import numpy as np
    import torch

    _install_stubs()
    G: dict = {"np": np, "torch": torch, "__name__": f"reference_{variant}"}
    if variant == "e2e":
        cells, wanted = _cells(E2E_NOTEBOOK), (2, 4, 6)
    elif variant == "indi":
        cells, wanted = _cells(INDI_NOTEBOOK), (2, 5)
    else:
        raise ValueError(variant)

    real_load = torch.load

    def load_full_pickle(path, *a, **kw):  # the .pt files are whole-module pickles
        kw.setdefault("weights_only", False)
        kw.setdefault("map_location", "cpu")
        return real_load(path, *a, **kw)

    rng_state = np.random.get_state()  # cell 4 draws from the global stream; do not disturb the caller
    torch.load = load_full_pickle
    try:
        with _cwd(REFERENCE_ROOT), contextlib.redirect_stdout(io.StringIO()):
            for i in wanted:
                exec(compile(cells[i], f"{variant}:cell{i}", "exec"), G)
    finally:
        torch.load = real_load
        np.random.set_state(rng_state)
    _CACHE[variant] = G
    return G


# ---- the tracks the notebooks define (E2E cell 8, INDI cell 7), restated as data ----
def zigzag_track():
    import numpy as np

    gate_pos = np.array([[-3.0, 0, -1.5], [-1, 0, -1.5], [1, 0, -1.5], [3, 0, -1.5],
                         [1, 0, -1.5], [-1, 0, -1.5], [-3, 0, -1.5]])
    gate_yaw = np.array([np.pi / 2, -np.pi / 2] * 3 + [np.pi / 2])
    start_pos = gate_pos[0] + np.array([0, -1.0, 0])
    return gate_pos, gate_yaw, start_pos


def rectangle_track():
    import numpy as np

    gate_pos = np.array([[2, -1.5, -1.5], [2, 1.5, -1.5], [-2, 1.5, -1.5], [-2, -1.5, -1.5]] * 2, dtype=float)
    gate_yaw = np.array([np.pi / 4, 3 * np.pi / 4, 5 * np.pi / 4, 7 * np.pi / 4] * 2)
    start_pos = gate_pos[3]
    return gate_pos, gate_yaw, start_pos


def training_disturbance_ranges():
    """E2E training ranges, float64 exactly as cell 10 builds them (`3D quad race.ipynb:772-779`)."""
    import numpy as np

    return np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])


def make_reference_env(variant, num_envs, gates_ahead=1, pause_if_collision=False, track=None,
                       disturbance_ranges="training"):
    G = load_reference(variant)
    if track is None:
        track = zigzag_track() if variant == "e2e" else rectangle_track()
    gp, gy, sp = track
    env = G["Quadcopter3DGates"](num_envs=num_envs, gates_pos=gp, gate_yaw=gy, start_pos=sp,
                                 gates_ahead=gates_ahead, pause_if_collision=pause_if_collision)
    if variant == "e2e" and disturbance_ranges is not None:
        env.disturbance_ranges = (training_disturbance_ranges() if isinstance(disturbance_ranges, str)
                                  else disturbance_ranges)
    return env

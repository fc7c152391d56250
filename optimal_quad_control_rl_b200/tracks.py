"""The two race tracks the reference notebooks define, as data (float64, like the notebook literals)."""
import numpy as np


def zigzag_track():
    """7-gate zigzag of the end-to-end notebook (`3D quad race.ipynb:640-661`): gates_pos, gate_yaw, start_pos."""
    gate_pos = np.array([[x, 0.0, -1.5] for x in (-3, -1, 1, 3, 1, -1, -3)], dtype=np.float64)
    gate_yaw = np.array([np.pi / 2 * (-1) ** i for i in range(7)])
    return gate_pos, gate_yaw, gate_pos[0] + np.array([0.0, -1.0, 0.0])


def rectangle_track():
    """8-gate (2 laps of 4) rectangle of the INDI notebook (`3D quad race INDI inner loop.ipynb:438-460`)."""
    gate_pos = np.array([[2, -1.5, -1.5], [2, 1.5, -1.5], [-2, 1.5, -1.5], [-2, -1.5, -1.5]] * 2, dtype=np.float64)
    gate_yaw = np.array([np.pi / 4, 3 * np.pi / 4, 5 * np.pi / 4, 7 * np.pi / 4] * 2)
    return gate_pos, gate_yaw, gate_pos[3].copy()


def training_disturbance_ranges():
    """Ranges the E2E training cell assigns after construction (`3D quad race.ipynb:772-781`), float64."""
    return np.array([[-0.03, 0.03], [-0.03, 0.03], [-0.01, 0.01], [0, 0], [0, 0], [-0.5, 0.5]])

"""Constants of the experiment: mirror of reference src/params.py (DroneData :10-70, ExperimentParameters :113-122).
Only what the MPC path reads is kept; the URDF parser is out of scope (its MASS is overridden at params.py:42 and
nothing else it parses reaches the dynamics)."""
import numpy as np


class DroneData:
    def __init__(self):
        self.GRAVITY_ACC = 9.81                       # params.py:37
        self.MASS = 0.03277                           # params.py:42
        self.GRAVITY = self.GRAVITY_ACC * self.MASS   # params.py:45
        self.max_F = 1.3 * self.GRAVITY               # params.py:46
        self.min_F = -0.2 * self.GRAVITY              # params.py:47
        self.min_p_x, self.max_p_x = -1.2, 1.2        # params.py:48-51
        self.min_p_z, self.max_p_z = -1.2, 1.2
        self.min_v_x, self.max_v_x = -1, 1            # params.py:52-55
        self.min_v_z, self.max_v_z = -1, 1
        self.min_a_x, self.max_a_x = -5, 5            # params.py:56-59
        self.min_a_z, self.max_a_z = -5 + self.GRAVITY_ACC, 5 + self.GRAVITY_ACC
        self.min_jerk, self.max_jerk = -5, 5          # params.py:60-61
        self.RAD2DEG = 180 / np.pi
        self.DEG2RAD = np.pi / 180


class ExperimentParameters:
    def __init__(self):
        self.T = 10                                   # params.py:115
        self.dt = 1 / 50                              # params.py:116
        self.dt_conv = 1 / 500                        # params.py:117
        self.ctrls_per_sample = int(self.dt / self.dt_conv)   # params.py:118
        self.N = int(self.T / self.dt)                # params.py:119
        self.N_conv = int(self.T / self.dt_conv)      # params.py:120
        self.N_horizon = 30                           # params.py:121
        self.noise = 0.01                             # params.py:122

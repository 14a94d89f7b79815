"""3-D attitude-and-total-thrust controller path (north-star extension, not in the reference): same module shape as
force_model / jerk_model (OCP, Converter, follow_trajectory) over libbnmpc's BNMPC_MODEL_ATT."""
from .ocp import OCP, Converter, follow_trajectory, follow_trajectory_batched, gen_helix_traj, helix_table  # noqa: F401

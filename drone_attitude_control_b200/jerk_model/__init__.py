"""Jerk-model controller path: mirror of reference src/jerk_model/{dynamics,ocp,controller}.py over libbnmpc."""
from .ocp import OCP, Converter  # noqa: F401
from .ocp import follow_trajectory  # noqa: F401

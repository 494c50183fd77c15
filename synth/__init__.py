"""Synthetic bearing-only worlds (benchmark / test input generator).  Own shared library, no CUDA, no product code."""
from .synth import build, synth_world  # noqa: F401

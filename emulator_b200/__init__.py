"""
emulator_b200: B200-native (sm_100a CUDA) Gaussian-process hot path behind swiftemulator's
emulators.  See DESIGN.md.  No CPU fallback: the CUDA library must be built and a GPU visible.
"""

__version__ = "0.1.0"

"""Dense-stage subset of the reference's `core` package (reference: src/core/__init__.py:6-36).

Same importable names for the hot path: Camera, CameraPose, load_calibration,
DenseStereoReconstructor, create_combined_dense_cloud; PatchMatchMVS is imported lazily from
core.mvs_patchmatch exactly like the reference does (run_reconstruction.py:130)."""
from .camera import Camera, CameraPose, load_calibration
from .dense_stereo import DenseStereoReconstructor, create_combined_dense_cloud

__all__ = ['Camera', 'CameraPose', 'load_calibration', 'DenseStereoReconstructor', 'create_combined_dense_cloud']

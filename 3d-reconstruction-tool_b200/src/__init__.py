"""Drop-in mirror of the reference's `src` package for the dense stage only
(reference: src/__init__.py).  Sparse SfM, GUI, calibration and the COLMAP wrapper are out of scope."""

"""modflow6_b200 -- B200-native IMS linear solve + GWF assembly hot path for MODFLOW 6.

The package holds only what the hot path needs: `csrc/` (CUDA kernels + the
C ABI, built in-tree into libmf6gpu.so) and the host-side mirror of the
reference's LinearSolverBase / MatrixBase / VectorBase / NumericalSolution seam.
"""
__version__ = "0.1.0"

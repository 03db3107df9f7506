"""sga_b200 -- Blackwell-native Stochastic Gumbel Annealing for the mean-scale hyperprior codec.

Host-side mirror of the reference's sga.py (compress CLI, Adam/temperature loop, npz outputs) over a C-ABI
CUDA library (libsga_b200.so, include/sga_b200.h).  Import name: ``sga_b200`` (see sga_b200.py at the repo root;
the directory name carries the upstream project name and is not a valid identifier)."""
from . import schedule, synth, weights  # noqa: F401
from ._lib import OPERAND_BF16, OPERAND_F32, SgaError  # noqa: F401

__all__ = ["schedule", "synth", "weights", "OPERAND_BF16", "OPERAND_F32", "SgaError"]

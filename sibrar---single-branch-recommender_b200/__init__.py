"""sibrar_b200 -- B200-native (sm_100a) SingleBranchNet training step + full-catalog top-k evaluation.

Drop-in for the reference's algorithm API (``algorithms/sgd_alg.py:2009-2144``); all arithmetic runs in the
hand-written CUDA kernels of ``csrc/`` behind the C-ABI declared in ``include/sibrar_b200.h``.
"""
__version__ = "0.1.0"

"""paraiso_b200 — a B200-native stencil backend for the Paraiso Orthotope Machine.

Layers: om/ (graph IR + Builder EDSL), annotation.py, optimization.py
(analysis passes), generator/ (Plan + B200 emitter), examples/ (Life, Hydro, ...),
runtime.py (host side over the generated C ABI), csrc/ (CUDA runtime pieces).
"""
__version__ = "0.1.0"

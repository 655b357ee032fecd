"""srfrd_b200 -- B200 (sm_100a) implementation of the SRFRD data-parallel hot path.

Drop-in model classes: ``srfrd_b200.SRFR_model`` (SRFR, SRFRN, SRFU_B/F/R, SASRec) and
``srfrd_b200.model`` (legacy numpy-input SASRec); trainer: ``srfrd_b200.trainer``;
full-catalogue evaluation: ``srfrd_b200.evaluation``.  The CUDA kernels live in
``srfrd_b200/csrc`` behind the C ABI of ``include/srfrd_b200.h`` and are loaded from
``srfrd_b200/lib/libsrfrd_b200.so``; nothing here falls back to the CPU.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"

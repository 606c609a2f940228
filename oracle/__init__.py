"""CPU oracle for the OFDM link chain -- TEST INFRASTRUCTURE ONLY.

This package is a float64 NumPy/SciPy *restatement* of the MATLAB functions of
ladnlav/OFDM-course (canonical copies: ``Task 5/*.m``, ``Task 4/fine_sync.m``,
``Task 1/OFDM_map_carriers.m``).  Every function cites the reference file:line it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
(``cpu_baseline`` / ``--impl reference``) may import it; the product package
``ofdm-course_b200`` never does and fails loudly without its CUDA library.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path, and neither MATLAB nor GNU Octave exists in the build container, so
the reference itself cannot be executed here.  The oracle is pinned only against
known-answer tests *derived by hand from the reference source* (SURVEY.md section 4,
KAT 1-9; ``tests/test_oracle_kats.py``) and against closed-form identities
(loop-back BER = 0, FFT of the channel taps, spline = not-a-knot).
"""
from .functions import *  # noqa: F401,F403
from . import chains  # noqa: F401

"""contact_b200 -- B200-native hot path of CONTACT (influence-coefficient FFT product, NormCG/NORM, ...).

Host-side mirror of the reference's ``python_intfc`` (same function names, argument meaning and error behaviour)
over the C-ABI of ``lib/libcontact_addon_b200.so``.  The CUDA library is mandatory: importing works without a
GPU (the driver's build check), but any computing call fails loudly when the library or the device is missing.
There is no CPU fallback and nothing here imports ``oracle/``.
"""
from .lib import load_library, library_path, LibraryMissing      # noqa: F401
from .api import *                                               # noqa: F401,F403
from . import lowlevel                                           # noqa: F401

"""ngsF-HMM EM hot path on B200: hand-written sm_100a kernels behind a C ABI.

Python here is host plumbing only (ctypes over ``include/ngsfhmm_b200.h``,
synthetic inputs, multi-rank exchange through torch.distributed).  There is no
CPU fallback: without the compiled library and a CUDA device the calls raise.
"""
from .api import Context, NfhError, library_path, load_library, build_library  # noqa: F401
from .em import EmRank, Group, load_host_library, run_em  # noqa: F401,E402

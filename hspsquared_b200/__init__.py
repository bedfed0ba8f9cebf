"""hspsquared_b200 -- the land-segment hot path of HSP2 (PERLND / IMPLND time loops) on B200 (sm_100a) behind HSP2's own
module contract.  See DESIGN.md; the C ABI is include/hsp2g.h."""
import os as _os

# Ensembles with several option words run one kernel per word on as many streams (engine.GroupedLandBatch); the default of 8
# hardware queues would alias those streams into false dependencies.  Only effective when set before the process's first CUDA
# call, hence here, at import; an explicit setting of the user's wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

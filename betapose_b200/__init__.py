"""betapose_b200: B200-native engine for the per-frame evaluate path of sjtuytc/betapose."""
import os as _os

# Mixed-object batches run up to 13 objects' networks on 13 CUDA streams at once (engine.py: concurrent_slots); with the
# driver's default of 8 hardware work queues streams would share queues and serialise falsely.  Only takes effect when set
# before the CUDA context is created; an explicit setting by the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

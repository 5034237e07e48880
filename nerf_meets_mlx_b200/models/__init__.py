from . import embedding  # noqa: F401
from . import NeRF as _nerf_module  # noqa: F401
from .NeRF import NeRF, create_NeRF, run_model, inference_wrapper_batch  # noqa: F401

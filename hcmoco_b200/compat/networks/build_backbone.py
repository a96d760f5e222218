"""Shadow of pycontrast/networks/build_backbone.py (only what main_contrast.py imports)."""
from hcmoco_b200.api import HCMoCoModel as CMC3HRNetSGCNSingleHead, build_model  # noqa: F401

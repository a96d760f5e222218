"""Shadow of pycontrast/memory/build_memory.py."""
from hcmoco_b200.api import HCMoCoMem as CMCMem3, build_mem  # noqa: F401

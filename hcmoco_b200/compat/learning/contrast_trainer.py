"""Shadow of pycontrast/learning/contrast_trainer.py."""
from hcmoco_b200.api import ContrastTrainer, build_contrast  # noqa: F401

"""Shadow of pycontrast/options/train_options.py."""
from hcmoco_b200.options import TrainOptions  # noqa: F401
